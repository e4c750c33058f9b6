// LayerNorm over channels for channels-last token grids, with (a) two-source channel concatenation and
// (b) trailing zero-pad of the spatial grid fused in.  One warp per row; rows are short (C = 16..1536) so the
// row stays in L1 across the three passes.  HBM-bound: algorithmic bytes = 2*rows*C*4 (+ stats).
#include "common.cuh"

namespace mic {

struct LnGeom {
    int B, D, H, W, Dp, Hp, Wp;
};

__device__ __forceinline__ bool padded_to_row(const LnGeom& g, int64_t prow, int64_t& row) {
    int64_t t = prow;
    const int x = (int)(t % g.Wp); t /= g.Wp;
    const int y = (int)(t % g.Hp); t /= g.Hp;
    const int z = (int)(t % g.Dp); t /= g.Dp;
    if (x >= g.W || y >= g.H || z >= g.D) return false;
    row = ((t * g.D + z) * g.H + y) * (int64_t)g.W + x;
    return true;
}

__device__ __forceinline__ float ld_cat(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                                        int64_t row, int c) {
    return c < C0 ? x0[row * C0 + c] : x1[row * C1 + (c - C0)];
}

__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1,
                                                     int C1, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, LnGeom g,
                                                     int64_t prows, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int C = C0 + C1;
    for (int64_t prow = warp; prow < prows; prow += nwarps) {
        int64_t row;
        float* yo = y + prow * C;
        if (!padded_to_row(g, prow, row)) {
            for (int c = lane; c < C; c += 32) yo[c] = 0.f;
            continue;
        }
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += ld_cat(x0, C0, x1, C1, row, c);
        const float mu = warp_sum(s) / C;
        float v = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float d = ld_cat(x0, C0, x1, C1, row, c) - mu;
            v += d * d;
        }
        const float rs = rsqrtf(warp_sum(v) / C + eps);
        for (int c = lane; c < C; c += 32)
            yo[c] = (ld_cat(x0, C0, x1, C1, row, c) - mu) * rs * gamma[c] + beta[c];
        if (lane == 0) {
            mean[row] = mu;
            rstd[row] = rs;
        }
    }
}

// NPL = channels per lane (C <= 32*NPL).  Each warp walks rows with stride; per-lane partial dgamma/dbeta stay in
// registers and are flushed once per CTA (shared-memory reduce over warps, then one atomic per channel).
template <int NPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x0, int C0,
                                                     const float* __restrict__ x1, int C1,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ dres0,
                                                     const float* __restrict__ dres1, float* __restrict__ dx0,
                                                     float* __restrict__ dx1, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, LnGeom g, int64_t prows) {
    extern __shared__ float red[];  // [2][C]
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const int C = C0 + C1;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) red[c] = 0.f;
    __syncthreads();
    float pg[NPL], pb[NPL], gam[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        pg[i] = 0.f; pb[i] = 0.f;
        const int c = lane + 32 * i;
        gam[i] = c < C ? gamma[c] : 0.f;
    }
    const int64_t warp = (int64_t)blockIdx.x * nw + wid;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    for (int64_t prow = warp; prow < prows; prow += nwarps) {
        int64_t row;
        if (!padded_to_row(g, prow, row)) continue;
        const float mu = mean[row], rs = rstd[row];
        const float* dyr = dy + prow * C;
        float xh[NPL], gg[NPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float d = dyr[c];
                xh[i] = (ld_cat(x0, C0, x1, C1, row, c) - mu) * rs;
                gg[i] = d * gam[i];
                pg[i] += d * xh[i];
                pb[i] += d;
                s1 += gg[i];
                s2 += gg[i] * xh[i];
            } else {
                xh[i] = 0.f; gg[i] = 0.f;
            }
        }
        s1 = warp_sum(s1) / C;
        s2 = warp_sum(s2) / C;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                float v = rs * (gg[i] - s1 - xh[i] * s2);
                if (c < C0) {
                    if (dres0) v += dres0[row * C0 + c];
                    dx0[row * C0 + c] = v;
                } else {
                    if (dres1) v += dres1[row * C1 + (c - C0)];
                    dx1[row * C1 + (c - C0)] = v;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
            atomicAdd(&red[c], pg[i]);
            atomicAdd(&red[C + c], pb[i]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(&dgamma[c], red[c]);
        atomicAdd(&dbeta[c], red[C + c]);
    }
}

}  // namespace mic

using namespace mic;

extern "C" int mic_layernorm_fwd(const float* x0, int C0, const float* x1, int C1, const float* gamma,
                                 const float* beta, float* y, float* mean, float* rstd, int B, int D, int H, int W,
                                 int Dp, int Hp, int Wp, float eps, void* stream) {
    MIC_REQUIRE(x0 && gamma && beta && y && mean && rstd, "layernorm_fwd: null pointer");
    MIC_REQUIRE(C0 > 0 && C1 >= 0 && (C1 == 0 || x1), "layernorm_fwd: bad channel split %d|%d", C0, C1);
    MIC_REQUIRE(Dp >= D && Hp >= H && Wp >= W && B > 0 && D > 0 && H > 0 && W > 0, "layernorm_fwd: bad geometry");
    const int64_t prows = (int64_t)B * Dp * Hp * Wp;
    LnGeom g{B, D, H, W, Dp, Hp, Wp};
    const int wpb = 8;
    int64_t blocks = ceil_div64(prows, wpb);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    ln_fwd_kernel<<<(unsigned)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(x0, C0, x1, C1, gamma, beta, y, mean, rstd, g,
                                                                          prows, eps);
    return check_launch("ln_fwd_kernel");
}

extern "C" int mic_layernorm_bwd(const float* dy, const float* x0, int C0, const float* x1, int C1, const float* gamma,
                                 const float* mean, const float* rstd, const float* dres0, const float* dres1,
                                 float* dx0, float* dx1, float* dgamma, float* dbeta, int B, int D, int H, int W,
                                 int Dp, int Hp, int Wp, void* stream) {
    MIC_REQUIRE(dy && x0 && gamma && mean && rstd && dx0 && dgamma && dbeta, "layernorm_bwd: null pointer");
    MIC_REQUIRE(C0 > 0 && C1 >= 0 && (C1 == 0 || (x1 && dx1)), "layernorm_bwd: bad channel split %d|%d", C0, C1);
    const int C = C0 + C1;
    MIC_REQUIRE(C <= 32 * 48, "layernorm_bwd: C=%d > 1536 unsupported", C);
    const int64_t prows = (int64_t)B * Dp * Hp * Wp;
    LnGeom g{B, D, H, W, Dp, Hp, Wp};
    const int wpb = 8;
    int64_t blocks = ceil_div64(prows, (int64_t)wpb * 2);
    const int64_t cap = (int64_t)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t smem = 2 * (size_t)C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define LN_BWD(NPL)                                                                                              \
    ln_bwd_kernel<NPL><<<(unsigned)blocks, wpb * 32, smem, st>>>(dy, x0, C0, x1, C1, gamma, mean, rstd, dres0,   \
                                                                 dres1, dx0, dx1, dgamma, dbeta, g, prows)
    if (C <= 64) LN_BWD(2);
    else if (C <= 128) LN_BWD(4);
    else if (C <= 256) LN_BWD(8);
    else if (C <= 512) LN_BWD(16);
    else if (C <= 768) LN_BWD(24);
    else LN_BWD(48);
#undef LN_BWD
    return check_launch("ln_bwd_kernel");
}
