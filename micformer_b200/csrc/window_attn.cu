// Windowed multi-head attention core on CUDA cores (exact fp32): softmax(scale * q k^T) v per (window, head),
// reading q/k/v straight from the token grid (window_partition / window_reverse are address math only) and
// never materialising the score matrix.  One thread per query row, K/V of the (window, head) staged in shared
// memory, online softmax.  Small windows (8 tokens in the train config) are packed G per CTA.
// The tcgen05 kernel in window_attn_tc.cu takes the large-window shapes (343 tokens, hd=32) when enabled.
#include "common.cuh"

namespace mic {

struct WinGeom {
    int B, Dp, Hp, Wp, heads, wd, wh, ww;
    int nwd, nwh, nww;   // windows per axis
    int N;               // tokens per window
    int G;               // (window, head) groups per CTA
    int64_t ngroups;     // B * nW * heads
};

__device__ __forceinline__ int64_t token_row(const WinGeom& g, int64_t win64, int tok) {
    // win = ((b*nwd + wz)*nwh + wy)*nww + wx ; tok = (iz*wh + iy)*ww + ix.  32-bit index math (the host checks that
    // the window count fits): 64-bit div/mod is a ~100-instruction software routine and this runs per staged float4
    uint32_t win = (uint32_t)win64;
    const uint32_t wx = win % (uint32_t)g.nww; win /= (uint32_t)g.nww;
    const uint32_t wy = win % (uint32_t)g.nwh; win /= (uint32_t)g.nwh;
    const uint32_t wz = win % (uint32_t)g.nwd; win /= (uint32_t)g.nwd;
    uint32_t t = (uint32_t)tok;
    const uint32_t ix = t % (uint32_t)g.ww; t /= (uint32_t)g.ww;
    const uint32_t iy = t % (uint32_t)g.wh; t /= (uint32_t)g.wh;
    const uint32_t z = wz * g.wd + t, y = wy * g.wh + iy, x = wx * g.ww + ix;
    return (((int64_t)win * g.Dp + z) * g.Hp + y) * (int64_t)g.Wp + x;
}

template <int HD>
__global__ void __launch_bounds__(HD > 32 ? 128 : 384) window_attn_fwd_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                       const float* __restrict__ v, int ldkv, float* __restrict__ out, int ldo,
                                       float* __restrict__ lse, WinGeom g, float scale) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                              // [G][N][HD]
    float* Vs = smem + (size_t)g.G * g.N * HD;     // [G][N][HD]
    const int N = g.N;
    const int64_t group0 = (int64_t)blockIdx.x * g.G;
    constexpr int V4 = HD / 4;
    // cooperative K/V staging
    for (int idx = threadIdx.x; idx < g.G * N * V4; idx += blockDim.x) {
        const int part = idx % V4;
        const int tok = (idx / V4) % N;
        const int gg = idx / (V4 * N);
        const int64_t grp = group0 + gg;
        if (grp >= g.ngroups) continue;
        const int head = (int)((uint32_t)grp % (uint32_t)g.heads);
        const int64_t row = token_row(g, (uint32_t)grp / (uint32_t)g.heads, tok);
        const float4 kk = *reinterpret_cast<const float4*>(k + row * ldkv + head * HD + part * 4);
        const float4 vv = *reinterpret_cast<const float4*>(v + row * ldkv + head * HD + part * 4);
        *reinterpret_cast<float4*>(Ks + ((size_t)gg * N + tok) * HD + part * 4) = kk;
        *reinterpret_cast<float4*>(Vs + ((size_t)gg * N + tok) * HD + part * 4) = vv;
    }
    __syncthreads();
    const int gg = threadIdx.x / N, i = threadIdx.x % N;
    const int64_t grp = group0 + gg;
    if (gg >= g.G || grp >= g.ngroups) return;
    const int head = (int)((uint32_t)grp % (uint32_t)g.heads);
    const int64_t row = token_row(g, (uint32_t)grp / (uint32_t)g.heads, i);
    float qr[HD], acc[HD];
#pragma unroll
    for (int c = 0; c < V4; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(q + row * ldq + head * HD + c * 4);
        qr[c * 4 + 0] = t.x * scale; qr[c * 4 + 1] = t.y * scale; qr[c * 4 + 2] = t.z * scale; qr[c * 4 + 3] = t.w * scale;
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    const float* kg = Ks + (size_t)gg * N * HD;
    const float* vg = Vs + (size_t)gg * N * HD;
    for (int j = 0; j < N; ++j) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < V4; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(kg + (size_t)j * HD + c * 4);
            s = fmaf(qr[c * 4 + 0], t.x, s); s = fmaf(qr[c * 4 + 1], t.y, s);
            s = fmaf(qr[c * 4 + 2], t.z, s); s = fmaf(qr[c * 4 + 3], t.w, s);
        }
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn);
        const float pj = expf(s - mn);
        l = l * corr + pj;
#pragma unroll
        for (int c = 0; c < V4; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(vg + (size_t)j * HD + c * 4);
            acc[c * 4 + 0] = fmaf(pj, t.x, acc[c * 4 + 0] * corr);
            acc[c * 4 + 1] = fmaf(pj, t.y, acc[c * 4 + 1] * corr);
            acc[c * 4 + 2] = fmaf(pj, t.z, acc[c * 4 + 2] * corr);
            acc[c * 4 + 3] = fmaf(pj, t.w, acc[c * 4 + 3] * corr);
        }
        m = mn;
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < V4; ++c) {
        float4 t = make_float4(acc[c * 4] * inv, acc[c * 4 + 1] * inv, acc[c * 4 + 2] * inv, acc[c * 4 + 3] * inv);
        *reinterpret_cast<float4*>(out + row * ldo + head * HD + c * 4) = t;
    }
    lse[row * g.heads + head] = m + logf(l);
}

// Backward: pass A (thread = query) computes dq; pass B (thread = key) computes dk, dv.  Scores are recomputed
// from q, k and the saved log-sum-exp.
template <int HD>
__global__ void __launch_bounds__(HD > 32 ? 128 : 384) window_attn_bwd_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                       const float* __restrict__ v, int ldkv, const float* __restrict__ out,
                                       const float* __restrict__ dout, int ldo, const float* __restrict__ lse,
                                       float* __restrict__ dq, int lddq, float* __restrict__ dk,
                                       float* __restrict__ dv, int lddkv, WinGeom g, float scale) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    const int N = g.N;
    const size_t GN = (size_t)g.G * N;
    float* Qs = smem;                  // [G][N][HD]
    float* Ks = Qs + GN * HD;
    float* Vs = Ks + GN * HD;
    float* dOs = Vs + GN * HD;
    float* Ls = dOs + GN * HD;         // [G][N] lse
    float* Ds = Ls + GN;               // [G][N] D_i = dO_i . O_i
    const int64_t group0 = (int64_t)blockIdx.x * g.G;
    constexpr int V4 = HD / 4;
    for (int idx = threadIdx.x; idx < (int)GN * V4; idx += blockDim.x) {
        const int part = idx % V4;
        const int tok = (idx / V4) % N;
        const int gg = idx / (V4 * N);
        const int64_t grp = group0 + gg;
        if (grp >= g.ngroups) continue;
        const int head = (int)((uint32_t)grp % (uint32_t)g.heads);
        const int64_t row = token_row(g, (uint32_t)grp / (uint32_t)g.heads, tok);
        const size_t so = ((size_t)gg * N + tok) * HD + part * 4;
        *reinterpret_cast<float4*>(Qs + so) = *reinterpret_cast<const float4*>(q + row * ldq + head * HD + part * 4);
        *reinterpret_cast<float4*>(Ks + so) = *reinterpret_cast<const float4*>(k + row * ldkv + head * HD + part * 4);
        *reinterpret_cast<float4*>(Vs + so) = *reinterpret_cast<const float4*>(v + row * ldkv + head * HD + part * 4);
        *reinterpret_cast<float4*>(dOs + so) = *reinterpret_cast<const float4*>(dout + row * ldo + head * HD + part * 4);
    }
    const int gg = threadIdx.x / N, i = threadIdx.x % N;
    const int64_t grp = group0 + gg;
    const bool active = gg < g.G && grp < g.ngroups;
    int head = 0;
    int64_t row = 0;
    if (active) {
        head = (int)((uint32_t)grp % (uint32_t)g.heads);
        row = token_row(g, (uint32_t)grp / (uint32_t)g.heads, i);
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < V4; ++c) {
            const float4 o = *reinterpret_cast<const float4*>(out + row * ldo + head * HD + c * 4);
            const float4 go = *reinterpret_cast<const float4*>(dout + row * ldo + head * HD + c * 4);
            d += o.x * go.x + o.y * go.y + o.z * go.z + o.w * go.w;
        }
        Ls[(size_t)gg * N + i] = lse[row * g.heads + head];
        Ds[(size_t)gg * N + i] = d;
    }
    __syncthreads();
    if (!active) return;
    const float* qg = Qs + (size_t)gg * N * HD;
    const float* kg = Ks + (size_t)gg * N * HD;
    const float* vg = Vs + (size_t)gg * N * HD;
    const float* dog = dOs + (size_t)gg * N * HD;
    const float* lg = Ls + (size_t)gg * N;
    const float* dg = Ds + (size_t)gg * N;
    float a[HD], b[HD], r1[HD], r2[HD];
    // ---- pass A: thread = query i -> dq_i = scale * sum_j p_ij (dp_ij - D_i) k_j ----
    {
#pragma unroll
        for (int c = 0; c < HD; ++c) { a[c] = qg[(size_t)i * HD + c]; b[c] = dog[(size_t)i * HD + c]; r1[c] = 0.f; }
        const float li = lg[i], di = dg[i];
        for (int j = 0; j < N; ++j) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < V4; ++c) {
                const float4 kk = *reinterpret_cast<const float4*>(kg + (size_t)j * HD + c * 4);
                const float4 vv = *reinterpret_cast<const float4*>(vg + (size_t)j * HD + c * 4);
                s = fmaf(a[c * 4], kk.x, s); s = fmaf(a[c * 4 + 1], kk.y, s);
                s = fmaf(a[c * 4 + 2], kk.z, s); s = fmaf(a[c * 4 + 3], kk.w, s);
                dp = fmaf(b[c * 4], vv.x, dp); dp = fmaf(b[c * 4 + 1], vv.y, dp);
                dp = fmaf(b[c * 4 + 2], vv.z, dp); dp = fmaf(b[c * 4 + 3], vv.w, dp);
            }
            const float pij = expf(s * scale - li);
            const float ds = pij * (dp - di);
#pragma unroll
            for (int c = 0; c < V4; ++c) {
                const float4 kk = *reinterpret_cast<const float4*>(kg + (size_t)j * HD + c * 4);
                r1[c * 4] = fmaf(ds, kk.x, r1[c * 4]); r1[c * 4 + 1] = fmaf(ds, kk.y, r1[c * 4 + 1]);
                r1[c * 4 + 2] = fmaf(ds, kk.z, r1[c * 4 + 2]); r1[c * 4 + 3] = fmaf(ds, kk.w, r1[c * 4 + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < V4; ++c)
            *reinterpret_cast<float4*>(dq + row * lddq + head * HD + c * 4) =
                make_float4(r1[c * 4] * scale, r1[c * 4 + 1] * scale, r1[c * 4 + 2] * scale, r1[c * 4 + 3] * scale);
    }
    // ---- pass B: thread = key j (= i) -> dv_j = sum_i p_ij dO_i ; dk_j = scale * sum_i ds_ij q_i ----
    {
        const int j = i;
#pragma unroll
        for (int c = 0; c < HD; ++c) { a[c] = kg[(size_t)j * HD + c]; b[c] = vg[(size_t)j * HD + c]; r1[c] = 0.f; r2[c] = 0.f; }
        for (int ii = 0; ii < N; ++ii) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < V4; ++c) {
                const float4 qq = *reinterpret_cast<const float4*>(qg + (size_t)ii * HD + c * 4);
                const float4 go = *reinterpret_cast<const float4*>(dog + (size_t)ii * HD + c * 4);
                s = fmaf(a[c * 4], qq.x, s); s = fmaf(a[c * 4 + 1], qq.y, s);
                s = fmaf(a[c * 4 + 2], qq.z, s); s = fmaf(a[c * 4 + 3], qq.w, s);
                dp = fmaf(b[c * 4], go.x, dp); dp = fmaf(b[c * 4 + 1], go.y, dp);
                dp = fmaf(b[c * 4 + 2], go.z, dp); dp = fmaf(b[c * 4 + 3], go.w, dp);
            }
            const float pij = expf(s * scale - lg[ii]);
            const float ds = pij * (dp - dg[ii]);
#pragma unroll
            for (int c = 0; c < V4; ++c) {
                const float4 qq = *reinterpret_cast<const float4*>(qg + (size_t)ii * HD + c * 4);
                const float4 go = *reinterpret_cast<const float4*>(dog + (size_t)ii * HD + c * 4);
                r1[c * 4] = fmaf(ds, qq.x, r1[c * 4]); r1[c * 4 + 1] = fmaf(ds, qq.y, r1[c * 4 + 1]);
                r1[c * 4 + 2] = fmaf(ds, qq.z, r1[c * 4 + 2]); r1[c * 4 + 3] = fmaf(ds, qq.w, r1[c * 4 + 3]);
                r2[c * 4] = fmaf(pij, go.x, r2[c * 4]); r2[c * 4 + 1] = fmaf(pij, go.y, r2[c * 4 + 1]);
                r2[c * 4 + 2] = fmaf(pij, go.z, r2[c * 4 + 2]); r2[c * 4 + 3] = fmaf(pij, go.w, r2[c * 4 + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < V4; ++c) {
            *reinterpret_cast<float4*>(dk + row * lddkv + head * HD + c * 4) =
                make_float4(r1[c * 4] * scale, r1[c * 4 + 1] * scale, r1[c * 4 + 2] * scale, r1[c * 4 + 3] * scale);
            *reinterpret_cast<float4*>(dv + row * lddkv + head * HD + c * 4) =
                make_float4(r2[c * 4], r2[c * 4 + 1], r2[c * 4 + 2], r2[c * 4 + 3]);
        }
    }
}

static int make_geom(WinGeom& g, int B, int Dp, int Hp, int Wp, int heads, int wd, int wh, int ww) {
    if (wd <= 0 || wh <= 0 || ww <= 0 || Dp % wd || Hp % wh || Wp % ww)
        return fail(MIC_ERR_INVALID, "window_attn: grid (%d,%d,%d) not a multiple of window (%d,%d,%d)", Dp, Hp, Wp, wd,
                    wh, ww);
    g.B = B; g.Dp = Dp; g.Hp = Hp; g.Wp = Wp; g.heads = heads; g.wd = wd; g.wh = wh; g.ww = ww;
    g.nwd = Dp / wd; g.nwh = Hp / wh; g.nww = Wp / ww;
    g.N = wd * wh * ww;
    if (g.N > 384) return fail(MIC_ERR_UNSUPPORTED, "window_attn: %d tokens per window > 384", g.N);
    g.G = g.N >= 128 ? 1 : 128 / g.N;
    g.ngroups = (int64_t)B * g.nwd * g.nwh * g.nww * heads;
    if (g.ngroups >= (int64_t)1 << 31) return fail(MIC_ERR_UNSUPPORTED, "window_attn: more than 2^31 (window, head) groups");
    return MIC_OK;
}

template <int HD>
static int launch_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo, float* lse,
                      const WinGeom& g, float scale, cudaStream_t st) {
    const size_t smem = 2 * (size_t)g.G * g.N * HD * sizeof(float);
    auto kern = window_attn_fwd_kernel<HD>;
    if (smem > 48 * 1024) {
        if (smem > 227 * 1024) return fail(MIC_ERR_UNSUPPORTED, "window_attn_fwd: window too large for shared memory");
        static size_t granted = 0;      // per instantiation; raise the opt-in limit only when it grows
        if (smem > granted) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            granted = smem;
        }
    }
    const int threads = ceil_div(g.G * g.N, 32) * 32;
    const int64_t blocks = ceil_div64(g.ngroups, g.G);
    mic::launch(kern, dim3((unsigned)blocks), dim3(threads), smem, st, q, ldq, k, v, ldkv, out, ldo, lse, g, scale);
    return check_launch("window_attn_fwd_kernel");
}

template <int HD>
static int launch_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* out,
                      const float* dout, int ldo, const float* lse, float* dq, int lddq, float* dk, float* dv, int lddkv,
                      const WinGeom& g, float scale, cudaStream_t st) {
    const size_t smem = ((size_t)4 * g.G * g.N * HD + 2 * (size_t)g.G * g.N) * sizeof(float);
    auto kern = window_attn_bwd_kernel<HD>;
    if (smem > 48 * 1024) {
        if (smem > 227 * 1024) return fail(MIC_ERR_UNSUPPORTED, "window_attn_bwd: window too large for shared memory");
        static size_t granted = 0;
        if (smem > granted) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            granted = smem;
        }
    }
    const int threads = ceil_div(g.G * g.N, 32) * 32;
    const int64_t blocks = ceil_div64(g.ngroups, g.G);
    mic::launch(kern, dim3((unsigned)blocks), dim3(threads), smem, st, q, ldq, k, v, ldkv, out, dout, ldo, lse, dq, lddq, dk, dv, lddkv, g,
                                                  scale);
    return check_launch("window_attn_bwd_kernel");
}

int simt_window_attn_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo,
                         float* lse, int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh, int ww,
                         float scale, cudaStream_t st) {
    WinGeom g;
    int rc = make_geom(g, B, Dp, Hp, Wp, heads, wd, wh, ww);
    if (rc) return rc;
    if ((ldq | ldkv | ldo) & 3) return fail(MIC_ERR_INVALID, "window_attn: row strides must be multiples of 4 floats");
    switch (hd) {
        case 8: return launch_fwd<8>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st);
        case 12: return launch_fwd<12>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st);
        case 16: return launch_fwd<16>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st);
        case 24: return launch_fwd<24>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st);
        case 32: return launch_fwd<32>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st);
        case 48: if (g.N <= 128) return launch_fwd<48>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st); break;
        case 64: if (g.N <= 128) return launch_fwd<64>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st); break;
        case 96: if (g.N <= 128) return launch_fwd<96>(q, ldq, k, v, ldkv, out, ldo, lse, g, scale, st); break;
        default: break;
    }
    return fail(MIC_ERR_UNSUPPORTED, "window_attn: head_dim %d (window %d tokens) not built: {8,12,16,24,32} any window, "
                "{48,64,96} windows <= 128 tokens", hd, g.N);
}

int simt_window_attn_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* out,
                         const float* dout, int ldo, const float* lse, float* dq, int lddq, float* dk, float* dv,
                         int lddkv, int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh, int ww,
                         float scale, cudaStream_t st) {
    WinGeom g;
    int rc = make_geom(g, B, Dp, Hp, Wp, heads, wd, wh, ww);
    if (rc) return rc;
    if ((ldq | ldkv | ldo | lddq | lddkv) & 3)
        return fail(MIC_ERR_INVALID, "window_attn: row strides must be multiples of 4 floats");
#define BWD(HD) launch_bwd<HD>(q, ldq, k, v, ldkv, out, dout, ldo, lse, dq, lddq, dk, dv, lddkv, g, scale, st)
    switch (hd) {
        case 8: return BWD(8);
        case 12: return BWD(12);
        case 16: return BWD(16);
        case 24: return BWD(24);
        case 32: return BWD(32);
        case 48: if (g.N <= 128) return BWD(48); break;
        case 64: if (g.N <= 128) return BWD(64); break;
        case 96: if (g.N <= 128) return BWD(96); break;
        default: break;
    }
#undef BWD
    return fail(MIC_ERR_UNSUPPORTED, "window_attn: head_dim %d (window %d tokens) not built: {8,12,16,24,32} any window, "
                "{48,64,96} windows <= 128 tokens", hd, g.N);
}

}  // namespace mic
