// C-ABI glue: error string, launch counter, GEMM-mode switch and the entry points that dispatch between the
// exact fp32 CUDA-core kernels and the tcgen05 tensor-core kernels.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace mic {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_gemm_mode{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// gemm_simt.cu
int simt_linear_fwd(const float*, int, const float*, int, int, const float*, float*, int, int, int, int, int, float*, int,
                    const float*, int, const float*, int, int, cudaStream_t);
int simt_linear_bwd_data(const float*, int, const float*, int, int, float*, int, int, int, int, const float*, int,
                         const float*, int, int, cudaStream_t);
int simt_linear_bwd_weight(const float*, int, const float*, int, float*, int, int, float*, int, int, int, const float*, int,
                           cudaStream_t);
// gemm_tc.cu (tcgen05); return MIC_ERR_UNSUPPORTED when the shape is not taken
int tc_linear_fwd(const float*, int, const float*, int, int, const float*, float*, int, int, int, int, int, float*, int,
                  const float*, int, const float*, int, int, int, cudaStream_t);
int tc_linear_bwd_data(const float*, int, const float*, int, int, float*, int, int, int, int, const float*, int,
                       const float*, int, int, int, cudaStream_t);
int tc_linear_bwd_weight(const float*, int, const float*, int, float*, int, int, float*, int, int, int, const float*, int, int,
                         cudaStream_t);
bool tc_view_pending();
void tc_view_set(int ch, int dc, int hc, int wc);
// window_attn_tc.cu (tcgen05; large windows, head_dim 32)
int tc_window_attn_fwd(const float*, int, const float*, const float*, int, float*, int, float*, int, int, int, int, int, int,
                       int, int, int, float, cudaStream_t);
// window_attn_tc_bwd.cu (tcgen05; large windows, head_dim 32)
int tc_window_attn_bwd(const float*, int, const float*, const float*, int, const float*, const float*, int, const float*,
                       float*, int, float*, float*, int, int, int, int, int, int, int, int, int, int, float, cudaStream_t);
// window_attn.cu
int simt_window_attn_fwd(const float*, int, const float*, const float*, int, float*, int, float*, int, int, int, int, int,
                         int, int, int, int, float, cudaStream_t);
int simt_window_attn_bwd(const float*, int, const float*, const float*, int, const float*, const float*, int, const float*,
                         float*, int, float*, float*, int, int, int, int, int, int, int, int, int, int, float,
                         cudaStream_t);

}  // namespace mic

using namespace mic;

namespace mic {
// Tensor-core mode: a shape the tcgen05 kernel declines runs the exact fp32 CUDA-core kernel instead.  That is correct but
// slow, so it is reported once per entry point on stderr (MICFORMER_WARN_FALLBACK=0 silences it).
void warn_fallback(const char* op, int a, int b, int c) {
    static const bool on = []() { const char* v = getenv("MICFORMER_WARN_FALLBACK"); return !(v && v[0] == '0'); }();
    if (!on) return;
    static std::atomic<unsigned> seen{0};
    unsigned h = 0;
    for (const char* p = op; *p; ++p) h = h * 31u + (unsigned)*p;
    const unsigned bit = 1u << (h % 32u);
    if (seen.fetch_or(bit) & bit) return;
    fprintf(stderr, "[micformer_b200] %s: shape (%d, %d, %d) is not taken by the tcgen05 kernel; using the fp32 CUDA-core "
                    "kernel (reported once per entry point)\n", op, a, b, c);
}
bool pdl_enabled() {
    static const bool on = []() { const char* v = getenv("MICFORMER_PDL"); return !(v && v[0] == '0'); }();
    return on;
}
}  // namespace mic

extern "C" int mic_version(void) { return 100; }
extern "C" const char* mic_last_error_string(void) { return g_err; }
extern "C" int64_t mic_launch_count(void) { return g_launches.load(); }
extern "C" void mic_reset_launch_count(void) { g_launches.store(0); }
extern "C" int mic_set_gemm_mode(int mode) {
    if (mode < 0 || mode > 2) return fail(MIC_ERR_INVALID, "gemm mode %d not in {0,1,2}", mode);
    g_gemm_mode.store(mode);
    return MIC_OK;
}
extern "C" int mic_get_gemm_mode(void) { return g_gemm_mode.load(); }

extern "C" int mic_linear_fwd(const float* X, int ldx, const float* W, int ldw, int w_is_kn, const float* bias, float* Y,
                              int ldy, int M, int N, int K, int act, float* pre, int ldpre, const float* res, int ldres,
                              const float* rowscale, int rows_per_sample, int accumulate, void* stream) {
    MIC_REQUIRE(X && W && Y && M > 0 && N > 0 && K > 0, "linear_fwd: bad arguments (M=%d N=%d K=%d)", M, N, K);
    MIC_REQUIRE(!(accumulate && (act || res)), "linear_fwd: accumulate cannot be combined with act/res");
    const bool view_was_pending = tc_view_pending();
    const int mode = g_gemm_mode.load();
    if (mode == 1) {
        int rc = tc_linear_fwd(X, ldx, W, ldw, w_is_kn, bias, Y, ldy, M, N, K, act, pre, ldpre, res, ldres, rowscale,
                               rows_per_sample, accumulate, mode, (cudaStream_t)stream);
        if (rc != MIC_ERR_UNSUPPORTED) return rc;
        warn_fallback("mic_linear_fwd", M, N, K);
    }
    if (view_was_pending) { tc_view_set(0, 0, 0, 0); return fail(MIC_ERR_UNSUPPORTED, "mic_linear_fwd: the unpatch view needs the tensor-core path (gemm mode 1) and a 32-cell-wide grid"); }
    return simt_linear_fwd(X, ldx, W, ldw, w_is_kn, bias, Y, ldy, M, N, K, act, pre, ldpre, res, ldres, rowscale,
                           rows_per_sample, accumulate, (cudaStream_t)stream);
}

extern "C" int mic_linear_bwd_data(const float* dY, int lddy, const float* W, int ldw, int w_is_kn, float* dX, int lddx,
                                   int M, int N, int K, const float* gelu_pre, int ldpre, const float* rowscale,
                                   int rows_per_sample, int accumulate, void* stream) {
    MIC_REQUIRE(dY && W && dX && M > 0 && N > 0 && K > 0, "linear_bwd_data: bad arguments");
    const bool view_was_pending = tc_view_pending();
    const int mode = g_gemm_mode.load();
    if (mode == 1) {
        int rc = tc_linear_bwd_data(dY, lddy, W, ldw, w_is_kn, dX, lddx, M, N, K, gelu_pre, ldpre, rowscale, rows_per_sample,
                                    accumulate, mode, (cudaStream_t)stream);
        if (rc != MIC_ERR_UNSUPPORTED) return rc;
        warn_fallback("mic_linear_bwd_data", M, N, K);
    }
    if (view_was_pending) { tc_view_set(0, 0, 0, 0); return fail(MIC_ERR_UNSUPPORTED, "mic_linear_bwd_data: the unpatch view needs the tensor-core path (gemm mode 1) and a 32-cell-wide grid"); }
    return simt_linear_bwd_data(dY, lddy, W, ldw, w_is_kn, dX, lddx, M, N, K, gelu_pre, ldpre, rowscale, rows_per_sample,
                                accumulate, (cudaStream_t)stream);
}

extern "C" int mic_linear_bwd_weight(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, int w_is_kn,
                                     float* db, int M, int N, int K, const float* rowscale, int rows_per_sample,
                                     void* stream) {
    MIC_REQUIRE(dY && X && dW && M > 0 && N > 0 && K > 0, "linear_bwd_weight: bad arguments");
    const bool view_was_pending = tc_view_pending();
    const int mode = g_gemm_mode.load();
    if (mode == 1) {
        int rc = tc_linear_bwd_weight(dY, lddy, X, ldx, dW, lddw, w_is_kn, db, M, N, K, rowscale, rows_per_sample, mode,
                                      (cudaStream_t)stream);
        if (rc != MIC_ERR_UNSUPPORTED) return rc;
        warn_fallback("mic_linear_bwd_weight", M, N, K);
    }
    if (view_was_pending) { tc_view_set(0, 0, 0, 0); return fail(MIC_ERR_UNSUPPORTED, "mic_linear_bwd_weight: the unpatch view needs the tensor-core path (gemm mode 1) and a 32-cell-wide grid"); }
    return simt_linear_bwd_weight(dY, lddy, X, ldx, dW, lddw, w_is_kn, db, M, N, K, rowscale, rows_per_sample,
                                  (cudaStream_t)stream);
}

extern "C" int mic_linear_unpatch_view(int ch, int dc, int hc, int wc) {
    MIC_REQUIRE(ch >= 0 && dc >= 0 && hc >= 0 && wc >= 0, "linear_unpatch_view: negative geometry");
    tc_view_set(ch, dc, hc, wc);
    return MIC_OK;
}

extern "C" int mic_window_attn_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo,
                                   float* lse, int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh, int ww,
                                   float scale, void* stream) {
    MIC_REQUIRE(q && k && v && out && lse, "window_attn_fwd: null pointer");
    if (g_gemm_mode.load() == 1) {
        int rc = tc_window_attn_fwd(q, ldq, k, v, ldkv, out, ldo, lse, B, Dp, Hp, Wp, heads, hd, wd, wh, ww, scale,
                                    (cudaStream_t)stream);
        if (rc != MIC_ERR_UNSUPPORTED) return rc;
    }
    return simt_window_attn_fwd(q, ldq, k, v, ldkv, out, ldo, lse, B, Dp, Hp, Wp, heads, hd, wd, wh, ww, scale,
                                (cudaStream_t)stream);
}

extern "C" int mic_window_attn_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* out,
                                   const float* dout, int ldo, const float* lse, float* dq, int lddq, float* dk,
                                   float* dv, int lddkv, int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh,
                                   int ww, float scale, void* stream) {
    MIC_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv, "window_attn_bwd: null pointer");
    // MICFORMER_ATTN_BWD_SIMT=1 keeps the exact CUDA-core backward for large windows too
    static const bool force_simt = []() { const char* e = getenv("MICFORMER_ATTN_BWD_SIMT"); return e && e[0] == '1'; }();
    if (g_gemm_mode.load() == 1 && !force_simt) {
        int rc = tc_window_attn_bwd(q, ldq, k, v, ldkv, out, dout, ldo, lse, dq, lddq, dk, dv, lddkv, B, Dp, Hp, Wp, heads, hd,
                                    wd, wh, ww, scale, (cudaStream_t)stream);
        if (rc != MIC_ERR_UNSUPPORTED) return rc;
    }
    return simt_window_attn_bwd(q, ldq, k, v, ldkv, out, dout, ldo, lse, dq, lddq, dk, dv, lddkv, B, Dp, Hp, Wp, heads, hd,
                                wd, wh, ww, scale, (cudaStream_t)stream);
}
