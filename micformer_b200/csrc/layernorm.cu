// LayerNorm over channels for channels-last token grids, with (a) two-source channel concatenation and
// (b) trailing zero-pad of the spatial grid fused in.  One warp per row; rows are short (C = 16..1536) so the
// row stays in L1 across the three passes.  HBM-bound: algorithmic bytes = 2*rows*C*4 (+ stats).
#include "common.cuh"

namespace mic {

struct LnGeom {
    int B, D, H, W, Dp, Hp, Wp;
};

__device__ __forceinline__ bool padded_to_row(const LnGeom& g, int64_t prow, int64_t& row) {
    if (g.Dp == g.D && g.Hp == g.H && g.Wp == g.W) {     // no trailing pad (every window-2 layer): identity
        row = prow;
        return true;
    }
    int64_t t = prow;
    int x; divmod(t, g.Wp, x);
    int y; divmod(t, g.Hp, y);
    int z; divmod(t, g.Dp, z);
    if (x >= g.W || y >= g.H || z >= g.D) return false;
    row = ((t * g.D + z) * g.H + y) * (int64_t)g.W + x;
    return true;
}

__device__ __forceinline__ float ld_cat(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                                        int64_t row, int c) {
    return c < C0 ? x0[row * C0 + c] : x1[row * C1 + (c - C0)];
}

// Vectorised kernels: a row (C = C0 + C1 channels, both multiples of 4) is held in registers as float4s by a group of
// LPR lanes (8, 16 or 32; 32 / LPR rows per warp pass), VPL float4 per lane; one global read of the row, group-wide
// shuffle reductions, one float4 write.
__device__ __forceinline__ float group_sum(float v, int lpr) {
    for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float4 ld_cat4v(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                                           int64_t row, int c) {
    return c < C0 ? __ldg(reinterpret_cast<const float4*>(x0 + row * C0 + c))
                  : __ldg(reinterpret_cast<const float4*>(x1 + row * C1 + (c - C0)));
}

template <int VPL>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1,
                                                     int C1, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, LnGeom g,
                                                     int64_t prows, float eps, int lpr) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int gl = lane % lpr, gi = lane / lpr, rpw = 32 / lpr;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int C = C0 + C1;
    const float invC = 1.f / C;
    float4 gam[VPL], bet[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = 4 * (gl + lpr * i);
        gam[i] = c < C ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        bet[i] = c < C ? __ldg(reinterpret_cast<const float4*>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t base = warp * rpw; base < prows; base += nwarps * rpw) {
        const int64_t prow = base + gi;
        int64_t row = 0;
        const bool in = prow < prows;
        const bool live = in && padded_to_row(g, prow, row);
        float4 v[VPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = 4 * (gl + lpr * i);
            v[i] = (live && c < C) ? ld_cat4v(x0, C0, x1, C1, row, c) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mu = group_sum(s, lpr) * invC;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = 4 * (gl + lpr * i);
            if (c < C) {
                const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
                q += (a * a + b * b) + (cc * cc + d * d);
            }
        }
        const float rs = rsqrtf(group_sum(q, lpr) * invC + eps);
        if (!in) continue;
        float* yo = y + prow * C;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = 4 * (gl + lpr * i);
            if (c < C) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);          // rows of the trailing zero pad
                if (live)
                    o = make_float4((v[i].x - mu) * rs * gam[i].x + bet[i].x, (v[i].y - mu) * rs * gam[i].y + bet[i].y,
                                    (v[i].z - mu) * rs * gam[i].z + bet[i].z, (v[i].w - mu) * rs * gam[i].w + bet[i].w);
                *reinterpret_cast<float4*>(yo + c) = o;
            }
        }
        if (live && gl == 0) {
            mean[row] = mu;
            rstd[row] = rs;
        }
    }
}

// Per-lane partial dgamma/dbeta stay in registers across all rows of the warp and are flushed once per CTA
// (lane groups folded by shuffles, shared-memory reduce over warps, then one atomic per channel).
template <int VPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x0, int C0,
                                                     const float* __restrict__ x1, int C1,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ dres0,
                                                     const float* __restrict__ dres1, float* __restrict__ dx0,
                                                     float* __restrict__ dx1, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, LnGeom g, int64_t prows, int lpr) {
    pdl_sync();
    extern __shared__ float red[];  // [2][C]
    const int lane = threadIdx.x & 31;
    const int gl = lane % lpr, gi = lane / lpr, rpw = 32 / lpr;
    const int wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const int C = C0 + C1;
    const float invC = 1.f / C;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) red[c] = 0.f;
    __syncthreads();
    float4 pg[VPL], pb[VPL], gam[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        pg[i] = make_float4(0.f, 0.f, 0.f, 0.f); pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = 4 * (gl + lpr * i);
        gam[i] = c < C ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t warp = (int64_t)blockIdx.x * nw + wid;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    for (int64_t base = warp * rpw; base < prows; base += nwarps * rpw) {
        const int64_t prow = base + gi;
        int64_t row = 0;
        const bool live = prow < prows && padded_to_row(g, prow, row);
        const float mu = live ? mean[row] : 0.f, rs = live ? rstd[row] : 0.f;
        float4 xh[VPL], gg[VPL], rr[VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = 4 * (gl + lpr * i);
            xh[i] = make_float4(0.f, 0.f, 0.f, 0.f); gg[i] = make_float4(0.f, 0.f, 0.f, 0.f); rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && c < C) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(dy + prow * C + c));
                const float4 xv = ld_cat4v(x0, C0, x1, C1, row, c);
                // the residual gradient is only needed after the row reductions: fetch it with the other operands
                const float* dres = c < C0 ? dres0 : dres1;
                if (dres) rr[i] = __ldg(reinterpret_cast<const float4*>(c < C0 ? dres + row * C0 + c : dres + row * C1 + (c - C0)));
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                gg[i] = make_float4(d.x * gam[i].x, d.y * gam[i].y, d.z * gam[i].z, d.w * gam[i].w);
                pg[i].x += d.x * xh[i].x; pg[i].y += d.y * xh[i].y; pg[i].z += d.z * xh[i].z; pg[i].w += d.w * xh[i].w;
                pb[i].x += d.x; pb[i].y += d.y; pb[i].z += d.z; pb[i].w += d.w;
                s1 += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
                s2 += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
            }
        }
        s1 = group_sum(s1, lpr) * invC;
        s2 = group_sum(s2, lpr) * invC;
        if (!live) continue;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = 4 * (gl + lpr * i);
            if (c < C) {
                float4 v = make_float4(rs * (gg[i].x - s1 - xh[i].x * s2), rs * (gg[i].y - s1 - xh[i].y * s2),
                                       rs * (gg[i].z - s1 - xh[i].z * s2), rs * (gg[i].w - s1 - xh[i].w * s2));
                float* dx = c < C0 ? dx0 + row * C0 + c : dx1 + row * C1 + (c - C0);
                v.x += rr[i].x; v.y += rr[i].y; v.z += rr[i].z; v.w += rr[i].w;
                *reinterpret_cast<float4*>(dx) = v;
            }
        }
    }
    // fold the lane groups of the warp (same channels) by shuffles, then the warps take turns adding their partial sums to the
    // shared accumulators (float atomicAdd on shared memory is a compare-and-swap loop: 8 contending warps cost microseconds)
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        for (int o = lpr; o < 32; o <<= 1) {
            pg[i].x += __shfl_xor_sync(0xffffffffu, pg[i].x, o); pg[i].y += __shfl_xor_sync(0xffffffffu, pg[i].y, o);
            pg[i].z += __shfl_xor_sync(0xffffffffu, pg[i].z, o); pg[i].w += __shfl_xor_sync(0xffffffffu, pg[i].w, o);
            pb[i].x += __shfl_xor_sync(0xffffffffu, pb[i].x, o); pb[i].y += __shfl_xor_sync(0xffffffffu, pb[i].y, o);
            pb[i].z += __shfl_xor_sync(0xffffffffu, pb[i].z, o); pb[i].w += __shfl_xor_sync(0xffffffffu, pb[i].w, o);
        }
    }
    for (int w = 0; w < nw; ++w) {
        if (wid == w && gi == 0) {
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = 4 * (gl + lpr * i);
                if (c < C) {
                    float4* rg = reinterpret_cast<float4*>(red + c);
                    float4* rb = reinterpret_cast<float4*>(red + C + c);
                    float4 a = *rg, b = *rb;
                    a.x += pg[i].x; a.y += pg[i].y; a.z += pg[i].z; a.w += pg[i].w;
                    b.x += pb[i].x; b.y += pb[i].y; b.z += pb[i].z; b.w += pb[i].w;
                    *rg = a; *rb = b;
                }
            }
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(&dgamma[c], red[c]);
        atomicAdd(&dbeta[c], red[C + c]);
    }
}

// lanes per row / float4 per lane for C channels
static void ln_shape(int C, int& lpr, int& vpl) {
    const int nv = C / 4;
    if (nv <= 8) { lpr = 8; vpl = 1; }
    else if (nv <= 16) { lpr = 16; vpl = 1; }
    else { lpr = 32; vpl = (nv + 31) / 32; }
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
    MIC_REQUIRE((C0 & 3) == 0 && (C1 & 3) == 0 && C0 + C1 <= 1536, "layernorm_fwd: channel counts must be multiples of 4, <= 1536 (%d|%d)", C0, C1);
    MIC_REQUIRE(((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(y) |
                  reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "layernorm_fwd: pointers must be 16-byte aligned");
    int lpr, vpl;
    ln_shape(C0 + C1, lpr, vpl);
    const int wpb = 8;
    int64_t blocks = ceil_div64(prows, (int64_t)wpb * (32 / lpr));
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define LN_FWD(V) mic::launch((ln_fwd_kernel<V>), dim3((unsigned)blocks), dim3(wpb * 32), 0, st, x0, C0, x1, C1, gamma, beta, y, mean, rstd, g, prows, eps, lpr)
    switch (vpl) {
        case 1: LN_FWD(1); break;
        case 2: LN_FWD(2); break;
        case 3: LN_FWD(3); break;
        case 4: LN_FWD(4); break;
        case 5: case 6: LN_FWD(6); break;
        case 7: case 8: LN_FWD(8); break;
        default: LN_FWD(12); break;
    }
#undef LN_FWD
    return check_launch("ln_fwd_kernel");
}

extern "C" int mic_layernorm_bwd(const float* dy, const float* x0, int C0, const float* x1, int C1, const float* gamma,
                                 const float* mean, const float* rstd, const float* dres0, const float* dres1,
                                 float* dx0, float* dx1, float* dgamma, float* dbeta, int B, int D, int H, int W,
                                 int Dp, int Hp, int Wp, void* stream) {
    MIC_REQUIRE(dy && x0 && gamma && mean && rstd && dx0 && dgamma && dbeta, "layernorm_bwd: null pointer");
    MIC_REQUIRE(C0 > 0 && C1 >= 0 && (C1 == 0 || (x1 && dx1)), "layernorm_bwd: bad channel split %d|%d", C0, C1);
    const int C = C0 + C1;
    MIC_REQUIRE(C <= 1536 && (C0 & 3) == 0 && (C1 & 3) == 0, "layernorm_bwd: channel counts must be multiples of 4, <= 1536 (%d|%d)", C0, C1);
    MIC_REQUIRE(((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(dy) |
                  reinterpret_cast<uintptr_t>(dx0) | reinterpret_cast<uintptr_t>(dx1) | reinterpret_cast<uintptr_t>(dres0) |
                  reinterpret_cast<uintptr_t>(dres1) | reinterpret_cast<uintptr_t>(gamma)) & 15) == 0,
                "layernorm_bwd: pointers must be 16-byte aligned");
    const int64_t prows = (int64_t)B * Dp * Hp * Wp;
    LnGeom g{B, D, H, W, Dp, Hp, Wp};
    int lpr, vpl;
    ln_shape(C, lpr, vpl);
    const int wpb = 8;
    int64_t blocks = ceil_div64(prows, (int64_t)wpb * (32 / lpr) * 2);
    const int64_t cap = (int64_t)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const size_t smem = 2 * (size_t)C * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define LN_BWD(V)                                                                                              \
    mic::launch((ln_bwd_kernel<V>), dim3((unsigned)blocks), dim3(wpb * 32), smem, st, dy, x0, C0, x1, C1, gamma, mean, rstd, dres0,   \
                                                               dres1, dx0, dx1, dgamma, dbeta, g, prows, lpr)
    switch (vpl) {
        case 1: LN_BWD(1); break;
        case 2: LN_BWD(2); break;
        case 3: LN_BWD(3); break;
        case 4: LN_BWD(4); break;
        case 5: case 6: LN_BWD(6); break;
        case 7: case 8: LN_BWD(8); break;
        default: LN_BWD(12); break;
    }
#undef LN_BWD
    return check_launch("ln_bwd_kernel");
}
