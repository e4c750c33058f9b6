// Shared definitions of the fused attention half-block kernels (block_attn.cu forward, block_attn_bwd.cu backward).
#pragma once
#include "common.cuh"
#include "tc5.cuh"

namespace mic {

struct AttnFwdArgs {
    const float* x; const float* kvsrc; float* y;            // kvsrc == nullptr: self block (k/v from LN(x))
    const float* gamma; const float* beta; const float* bq; const float* bkv; const float* bp;
    const uint8_t* wq_hi; const uint8_t* wq_lo;              // q    image: N = C  (n_pad CP) x K = C
    const uint8_t* wkv_hi; const uint8_t* wkv_lo;            // kv   image: N = 2C (n_pad 2C) x K = C
    const uint8_t* wp_hi; const uint8_t* wp_lo;              // proj image: N = C  (n_pad CP) x K = C
    const float* rowscale; int rps;
    int D, H, W;
    int64_t nwin_total;
    int ntiles;
    float scale, eps;
};

struct AttnBwdArgs {
    const float* x; const float* kvsrc; const float* dy;     // dy = gradient w.r.t. the half-block output x1
    float* dx;                                               // dy + LN'(dxn): gradient w.r.t. x through residual + q (+ k/v, self)
    float* dkvsrc;                                           // cross: gradient w.r.t. the k/v source (T, C), written
    const float* gamma; const float* beta; const float* bq; const float* bkv;
    const uint8_t* wq_hi; const uint8_t* wq_lo; const uint8_t* wkv_hi; const uint8_t* wkv_lo;        // forward images (recompute)
    const uint8_t* wpT_hi; const uint8_t* wpT_lo;            // proj transposed: N = C (n_pad CP) x K = C
    const uint8_t* wqT_hi; const uint8_t* wqT_lo;            // q transposed:    N = C (n_pad CP) x K = C
    const uint8_t* wkvT_hi; const uint8_t* wkvT_lo;          // kv transposed:   N = C (n_pad CP) x K = 2C (two panels)
    const float* rowscale; int rps;
    float* dgamma; float* dbeta; float* dWq; float* dbq; float* dWkv; float* dbkv; float* dWp; float* dbp;   // accumulated
    int D, H, W;
    int64_t nwin_total;
    int ntiles;
    float scale, eps;
};

template <int C, int HD>
struct AttnCfg {
    static constexpr int CP = (C + 15) / 16 * 16;
    static constexpr int HEADS = C / HD;
    static constexpr int ROW_WARPS = 4 * HEADS;
    static constexpr int TILE = 128 * 128;                   // one 64-feature panel of 128 rows
    static constexpr int WC_BYTES = CP * 128;                // C x C image (one panel), per hi / lo
    static constexpr int WKV_BYTES = 2 * C * 128;            // 2C x C image (one panel), per hi / lo
    static_assert(C % HD == 0 && HD % 8 == 0 && C <= 64 && HEADS >= 2 && HEADS <= 3, "fused attention: C <= 64, 2..3 heads, head_dim % 8 == 0");
    // ---- forward
    static constexpr int THREADS = 64 + 32 * ROW_WARPS;
    static constexpr int F_XN = 0, F_SP = 2 * TILE, F_WQ = 4 * TILE;
    static constexpr int F_WKV = F_WQ + 2 * WC_BYTES, F_WP = F_WKV + 2 * WKV_BYTES, F_PAR = F_WP + 2 * WC_BYTES;
    static constexpr int F_BAR = F_PAR + 4 * 6 * C;
    static constexpr int F_SMEM = F_BAR + 128 + 1024;
    static constexpr int T_Q = 0, T_K = CP, T_V = CP + C;
    static constexpr int F_TCOLS = 256;
    static_assert(CP + 2 * C <= 256, "TMEM");
};

// 2x2x2 window geometry: global window index -> token row of the (B, D, H, W) grid
struct WinGeom {
    int D, H, W, nx, ny, nz;
    __device__ WinGeom(int d, int h, int w) : D(d), H(h), W(w), nx(w >> 1), ny(h >> 1), nz(d >> 1) {}
    __device__ __forceinline__ int64_t row_of(int64_t gw, int tok, int64_t nwin_total) const {
        if (gw >= nwin_total) return -1;
        int64_t t = gw;
        int wx, wy, wz;
        divmod(t, nx, wx); divmod(t, ny, wy); divmod(t, nz, wz);
        const int z = 2 * wz + (tok >> 2), y = 2 * wy + ((tok >> 1) & 1), x = 2 * wx + (tok & 1);
        return ((t * D + z) * H + y) * (int64_t)W + x;
    }
};

template <int C>
__device__ __forceinline__ void load_row(const float* __restrict__ base, int64_t grow, bool ok, float (&r)[C]) {
    if (ok) {
        const float4* p = reinterpret_cast<const float4*>(base + grow * C);
#pragma unroll
        for (int i = 0; i < C / 4; ++i) {
            const float4 v = __ldg(p + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < C; ++i) r[i] = 0.f;
    }
}

template <int C>
__device__ __forceinline__ void ln_stats(const float (&r)[C], float eps, float& mean, float& rstd) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) m += r[i];
    m *= (1.f / C);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) { const float d = r[i] - m; var = fmaf(d, d, var); }
    mean = m;
    rstd = rsqrtf(var * (1.f / C) + eps);
}

// a row of C features -> the split-bf16 chunks of its tile row (columns C .. CP-1 zero)
template <int C>
__device__ __forceinline__ void store_row_tile(uint8_t* hi, uint8_t* lo, int row, const float (&r)[C]) {
    constexpr int CP = (C + 15) / 16 * 16;
#pragma unroll
    for (int c = 0; c < CP / 8; ++c) {
        float v8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v8[e] = (c * 8 + e) < C ? r[(c * 8 + e) < C ? c * 8 + e : 0] : 0.f;
        t5::store_chunk(hi, lo, row, c, v8);
    }
}

// N consecutive TMEM columns of this thread's lane (N % 8 == 0); the caller issues ld_wait()
template <int N>
__device__ __forceinline__ void ld_cols(uint32_t taddr, float (&v)[N]) {
    static_assert(N % 8 == 0, "ld_cols");
    int c = 0;
#pragma unroll
    for (; c + 16 <= N; c += 16) t5::ld16(taddr + c, &v[c]);
    if (c < N) t5::ld8(taddr + c, &v[c]);
}

}  // namespace mic
