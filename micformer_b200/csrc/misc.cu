// Block <-> row permutation for the stride==kernel convolutions, the fused sigmoid-Dice + BCE loss, and the
// crop + residual helper used when windows need zero padding.  All HBM-bound streaming kernels.
#include "common.cuh"

namespace mic {

// grid (B, k*Dq, k*Hq, k*Wq, C) channels-last  <->  rows (B*Dq*Hq*Wq, k^3*C), columns ordered (kz,ky,kx,c).
// A run of L = k*C floats (kx, c) is contiguous in both layouts; one thread moves VEC floats of a run.
template <int VEC>
__global__ void __launch_bounds__(256) block_permute_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            int Dq, int Hq, int Wq, int k, int C,
                                                            int64_t grid_batch_stride, int to_rows, int64_t total) {
    pdl_sync();
    const int L = k * C;
    const int LV = L / VEC;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx;
        int lq; divmod(t, LV, lq);
        const int l = lq * VEC;
        int ky; divmod(t, k, ky);
        int kz; divmod(t, k, kz);
        const int64_t r = t;                      // row index
        int wx; divmod(t, Wq, wx);
        int hy; divmod(t, Hq, hy);
        int dz; divmod(t, Dq, dz);
        const int64_t b = t;
        const int64_t goff = b * grid_batch_stride +
                             ((((int64_t)dz * k + kz) * (Hq * k) + (hy * k + ky)) * (int64_t)(Wq * k) + (int64_t)wx * k) * C + l;
        const int64_t roff = (r * k * k + (kz * k + ky)) * L + l;
        const float* s = to_rows ? src + goff : src + roff;
        float* d = to_rows ? dst + roff : dst + goff;
        if (VEC == 4) *reinterpret_cast<float4*>(d) = *reinterpret_cast<const float4*>(s);
        else *d = *s;
    }
}

// ------------------------------------------------------------------------------------------------ loss
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// TT = float (one-hot float labels, what train_mmwhs_noPad.py:181 hands over after .float()) or uint8_t (the bool / uint8
// one-hot the dataset produces, dataset/MMWHS.py:392,414-425, read without the 4x wider float copy)
template <typename TT>
__global__ void __launch_bounds__(256) dice_partial_kernel(const float* __restrict__ logits,
                                                           const TT* __restrict__ target,
                                                           double* __restrict__ sums, int C, int64_t S) {
    pdl_sync();
    // grid: (chunks, B*C); one (b,c) slab per blockIdx.y
    const int64_t slab = blockIdx.y;
    const int c = (int)(slab % C);
    const float* lg = logits + slab * S;
    const TT* tg = target + slab * S;
    float s_pt = 0.f, s_pp = 0.f, s_tt = 0.f, s_ce = 0.f;
    const int64_t S4 = S >> 2;
    const bool vec = aligned16(lg) && (reinterpret_cast<uintptr_t>(tg) & (4 * sizeof(TT) - 1)) == 0;
    if (vec) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S4; i += (int64_t)gridDim.x * blockDim.x) {
            const float4 x = reinterpret_cast<const float4*>(lg)[i];
            float ts[4];
            if (sizeof(TT) == 4) {
                const float4 t = reinterpret_cast<const float4*>(tg)[i];
                ts[0] = t.x; ts[1] = t.y; ts[2] = t.z; ts[3] = t.w;
            } else {
                const uchar4 t = reinterpret_cast<const uchar4*>(tg)[i];
                ts[0] = t.x ? 1.f : 0.f; ts[1] = t.y ? 1.f : 0.f; ts[2] = t.z ? 1.f : 0.f; ts[3] = t.w ? 1.f : 0.f;
            }
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float p = sigmoidf_(xs[k]);
                s_pt = fmaf(p, ts[k], s_pt); s_pp = fmaf(p, p, s_pp); s_tt = fmaf(ts[k], ts[k], s_tt);
                s_ce += (ts[k] - 1.f) * fmaxf(log1pf(-p), -100.f) - ts[k] * fmaxf(logf(p), -100.f);
            }
        }
    }
    for (int64_t i = (vec ? S4 * 4 : 0) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float p = sigmoidf_(lg[i]);
        const float t = sizeof(TT) == 4 ? (float)tg[i] : (tg[i] ? 1.f : 0.f);
        s_pt = fmaf(p, t, s_pt); s_pp = fmaf(p, p, s_pp); s_tt = fmaf(t, t, s_tt);
        s_ce += (t - 1.f) * fmaxf(log1pf(-p), -100.f) - t * fmaxf(logf(p), -100.f);
    }
    __shared__ double red[4][8];
    double v[4] = {(double)s_pt, (double)s_pp, (double)s_tt, (double)s_ce};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) red[k][wid] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
        atomicAdd(&sums[c * 4 + threadIdx.x], t);
    }
}

// loss = (w_dice * sum_c dice_c + w_bce * sum_c bce_c) / C: (0.7, 0.3) = MDiceLoss (loss/dice.py:158-166), (1, 0) = MDiceLoss_Val
// (loss/dice.py:216-221)
__global__ void dice_finalize_kernel(const double* __restrict__ sums, float* __restrict__ loss, float* __restrict__ coef,
                                     int C, double n, double w_dice, double w_bce) {
    pdl_sync();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double dice = 0.0, ce = 0.0;
    for (int c = 0; c < C; ++c) {
        const double I = sums[c * 4], P2 = sums[c * 4 + 1], T2 = sums[c * 4 + 2], B = sums[c * 4 + 3];
        const double den = P2 + T2 + 1.0;
        dice += 1.0 - (2.0 * I + 1.0) / den;
        ce += B / n;
        // d loss / d p = a*t + b*p + e*(p-t)/max(p(1-p),1e-12)
        coef[c * 3 + 0] = (float)(-2.0 * w_dice / (C * den));
        coef[c * 3 + 1] = (float)(w_dice * 2.0 * (2.0 * I + 1.0) / (C * den * den));
        coef[c * 3 + 2] = (float)(w_bce / (C * n));
    }
    *loss = (float)((w_dice * dice + w_bce * ce) / C);
}

template <typename TT>
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ logits, const TT* __restrict__ target,
                                                       const float* __restrict__ coef, const float* __restrict__ dloss,
                                                       float* __restrict__ dlogits, int C, int64_t S) {
    pdl_sync();
    const int64_t slab = blockIdx.y;
    const int c = (int)(slab % C);
    const float g = dloss ? *dloss : 1.f;
    const float a = coef[c * 3] * g, b = coef[c * 3 + 1] * g, e = coef[c * 3 + 2] * g;
    const float* lg = logits + slab * S;
    const TT* tg = target + slab * S;
    float* dl = dlogits + slab * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
        const float p = sigmoidf_(lg[i]);
        const float t = sizeof(TT) == 4 ? (float)tg[i] : (tg[i] ? 1.f : 0.f);
        const float q = p * (1.f - p);
        const float dp = a * t + b * p + e * (p - t) / fmaxf(q, 1e-12f);
        dl[i] = dp * q;
    }
}

// ------------------------------------------------------------------------------------------- crop + residual
__global__ void __launch_bounds__(256) crop_residual_kernel(const float* __restrict__ res, const float* __restrict__ br,
                                                            const float* __restrict__ rowscale, float* __restrict__ y,
                                                            int D, int H, int W, int Dp, int Hp, int Wp, int C4,
                                                            int64_t total) {
    pdl_sync();
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx;
        int c; divmod(t, C4, c);
        const int64_t row = t;
        int x; divmod(t, W, x);
        int yy; divmod(t, H, yy);
        int z; divmod(t, D, z);
        const int64_t prow = ((t * Dp + z) * Hp + yy) * (int64_t)Wp + x;
        const float s = rowscale ? rowscale[t] : 1.f;
        const float4 r = reinterpret_cast<const float4*>(res)[row * C4 + c];
        const float4 v = reinterpret_cast<const float4*>(br)[prow * C4 + c];
        reinterpret_cast<float4*>(y)[row * C4 + c] = make_float4(r.x + s * v.x, r.y + s * v.y, r.z + s * v.z, r.w + s * v.w);
    }
}

__global__ void __launch_bounds__(256) crop_residual_bwd_kernel(const float* __restrict__ dy,
                                                                const float* __restrict__ rowscale,
                                                                float* __restrict__ dbr, int D, int H, int W, int Dp,
                                                                int Hp, int Wp, int C4, int64_t total) {
    pdl_sync();
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx;
        int c; divmod(t, C4, c);
        const int64_t prow = t;
        int x; divmod(t, Wp, x);
        int yy; divmod(t, Hp, yy);
        int z; divmod(t, Dp, z);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x < W && yy < H && z < D) {
            const int64_t row = ((t * D + z) * H + yy) * (int64_t)W + x;
            const float s = rowscale ? rowscale[t] : 1.f;
            const float4 d = reinterpret_cast<const float4*>(dy)[row * C4 + c];
            v = make_float4(s * d.x, s * d.y, s * d.z, s * d.w);
        }
        reinterpret_cast<float4*>(dbr)[prow * C4 + c] = v;
    }
}

static unsigned grid_for(int64_t total, int threads) {
    int64_t b = ceil_div64(total, threads);
    const int64_t cap = (int64_t)num_sms() * 32;
    return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace mic

using namespace mic;

extern "C" int mic_block_permute(const float* src, float* dst, int B, int Dq, int Hq, int Wq, int k, int C,
                                 int64_t grid_batch_stride, int to_rows, void* stream) {
    MIC_REQUIRE(src && dst && B > 0 && Dq > 0 && Hq > 0 && Wq > 0 && k > 0 && C > 0, "block_permute: bad arguments");
    const int L = k * C;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (L % 4 == 0) && (grid_batch_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (vec) {
        const int64_t total = (int64_t)B * Dq * Hq * Wq * k * k * (L / 4);
        mic::launch((block_permute_kernel<4>), dim3(grid_for(total, 256)), dim3(256), 0, st, src, dst, Dq, Hq, Wq, k, C, grid_batch_stride, to_rows, total);
    } else {
        const int64_t total = (int64_t)B * Dq * Hq * Wq * k * k * L;
        mic::launch((block_permute_kernel<1>), dim3(grid_for(total, 256)), dim3(256), 0, st, src, dst, Dq, Hq, Wq, k, C, grid_batch_stride, to_rows, total);
    }
    return check_launch("block_permute_kernel");
}

static int dice_partial_launch(const float* logits, const void* target, int target_u8, double* sums, int B, int C, int64_t S,
                               void* stream) {
    MIC_REQUIRE(logits && target && sums && B > 0 && C > 0 && S > 0, "dice_bce_partial: bad arguments");
    int chunks = (int)ceil_div64(S, 256 * 16);
    const int cap = ceil_div(num_sms() * 8, B * C);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    if (target_u8)
        mic::launch(dice_partial_kernel<uint8_t>, dim3(chunks, B * C), dim3(256), 0, (cudaStream_t)stream, logits,
                    (const uint8_t*)target, sums, C, S);
    else
        mic::launch(dice_partial_kernel<float>, dim3(chunks, B * C), dim3(256), 0, (cudaStream_t)stream, logits,
                    (const float*)target, sums, C, S);
    return check_launch("dice_partial_kernel");
}

extern "C" int mic_dice_bce_partial(const float* logits, const float* target, double* sums, int B, int C, int64_t S,
                                    void* stream) {
    return dice_partial_launch(logits, target, 0, sums, B, C, S, stream);
}
extern "C" int mic_dice_bce_partial_u8(const float* logits, const uint8_t* target, double* sums, int B, int C, int64_t S,
                                       void* stream) {
    return dice_partial_launch(logits, target, 1, sums, B, C, S, stream);
}

extern "C" int mic_dice_bce_finalize_weighted(const double* sums, float* loss, float* coef, int C, double n_per_channel,
                                              double w_dice, double w_bce, void* stream) {
    MIC_REQUIRE(sums && loss && coef && C > 0 && n_per_channel > 0, "dice_bce_finalize: bad arguments");
    mic::launch(dice_finalize_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, sums, loss, coef, C, n_per_channel, w_dice,
                w_bce);
    return check_launch("dice_finalize_kernel");
}
extern "C" int mic_dice_bce_finalize(const double* sums, float* loss, float* coef, int C, double n_per_channel,
                                     void* stream) {
    return mic_dice_bce_finalize_weighted(sums, loss, coef, C, n_per_channel, 0.7, 0.3, stream);
}

static int dice_bwd_launch(const float* logits, const void* target, int target_u8, const float* coef, const float* dloss,
                           float* dlogits, int B, int C, int64_t S, void* stream) {
    MIC_REQUIRE(logits && target && coef && dlogits && B > 0 && C > 0 && S > 0, "dice_bce_bwd: bad arguments");
    int chunks = (int)ceil_div64(S, 256 * 8);
    const int cap = ceil_div(num_sms() * 16, B * C);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    if (target_u8)
        mic::launch(dice_bwd_kernel<uint8_t>, dim3(chunks, B * C), dim3(256), 0, (cudaStream_t)stream, logits,
                    (const uint8_t*)target, coef, dloss, dlogits, C, S);
    else
        mic::launch(dice_bwd_kernel<float>, dim3(chunks, B * C), dim3(256), 0, (cudaStream_t)stream, logits,
                    (const float*)target, coef, dloss, dlogits, C, S);
    return check_launch("dice_bwd_kernel");
}
extern "C" int mic_dice_bce_bwd(const float* logits, const float* target, const float* coef, const float* dloss,
                                float* dlogits, int B, int C, int64_t S, double n_per_channel, void* stream) {
    (void)n_per_channel;
    return dice_bwd_launch(logits, target, 0, coef, dloss, dlogits, B, C, S, stream);
}
extern "C" int mic_dice_bce_bwd_u8(const float* logits, const uint8_t* target, const float* coef, const float* dloss,
                                   float* dlogits, int B, int C, int64_t S, void* stream) {
    return dice_bwd_launch(logits, target, 1, coef, dloss, dlogits, B, C, S, stream);
}

extern "C" int mic_crop_residual(const float* res, const float* branch, const float* rowscale, float* y, int B, int D,
                                 int H, int W, int Dp, int Hp, int Wp, int C, void* stream) {
    MIC_REQUIRE(res && branch && y && (C & 3) == 0, "crop_residual: bad arguments");
    const int64_t total = (int64_t)B * D * H * W * (C / 4);
    mic::launch(crop_residual_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, res, branch, rowscale, y, D, H, W, Dp, Hp, Wp,
                                                                                C / 4, total);
    return check_launch("crop_residual_kernel");
}

extern "C" int mic_crop_residual_bwd(const float* dy, const float* rowscale, float* dbranch, int B, int D, int H, int W,
                                     int Dp, int Hp, int Wp, int C, void* stream) {
    MIC_REQUIRE(dy && dbranch && (C & 3) == 0, "crop_residual_bwd: bad arguments");
    const int64_t total = (int64_t)B * Dp * Hp * Wp * (C / 4);
    mic::launch(crop_residual_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dy, rowscale, dbranch, D, H, W, Dp, Hp, Wp,
                                                                                    C / 4, total);
    return check_launch("crop_residual_bwd_kernel");
}
