// Fused MLP half-block for the DEEP stages (few tokens, wide channels: C = 96 / 192 / 384 in the train config):
//     y = x + rowscale * fc2( GELU( fc1( LayerNorm(x) ) ) )                 reference M:28-34, 403-404, 419-424
// The deep stages have too few 128-token tiles to fill the GPU (8 at stage 2, 1 at stage 3) and weights too large for shared
// memory, so the work is split over the HIDDEN axis as well: CTA (tile i, chunk j) owns 64 hidden units,
//     hpre_j = LN(x_i) W1_j^T  ->  h_j = GELU(hpre_j + b1_j)  ->  y_i += rowscale * h_j W2_j^T   (atomic; chunk 0 adds x + rowscale b2)
// with every operand STREAMED through small shared-memory rings in 64-wide K panels: the LayerNorm output panels are
// produced on the fly by the row threads (split bf16), the weight-image panels arrive by bulk copies.  y must be zero on
// entry.  One launch replaces LayerNorm + fc1 + fc2 (3 launches, the 4C-wide hidden tensor written and read twice).
//
// Roles (320 threads): warp 0 = weight loader, warp 1 = MMA issuer, warps 2-9 = row threads (lane quarter = warp & 3, the
// two warps of a quarter split the columns).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"
#include "block_mlp_split.cuh"

namespace mic {
using namespace t5;

__global__ void __launch_bounds__(MS_THREADS, 1) mlp_split_fwd_kernel(const MlpSplitFwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sXN = smem + MsF::XN;          // [2 slots][hi 16 KB | lo 16 KB]
    uint8_t* sW1 = smem + MsF::W1;          // [2 slots][hi 8 KB | lo 8 KB]      64 hidden rows x one K panel
    uint8_t* sH = smem + MsF::H;            // hi 16 KB | lo 16 KB
    uint8_t* sW2 = smem + MsF::W2;          // [2 slots][hi 16 KB | lo 16 KB]    up to 128 output rows x 64 hidden
    float* spart = reinterpret_cast<float*>(smem + MsF::PART);      // [128 rows][2 halves]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MsF::BAR);
    uint64_t* xn_full = bars + 0;   // [2]
    uint64_t* xn_empty = bars + 2;  // [2]
    uint64_t* w1_full = bars + 4;   // [2]
    uint64_t* w1_empty = bars + 6;  // [2]
    uint64_t* w2_full = bars + 8;   // [2]
    uint64_t* w2_empty = bars + 10; // [2]
    uint64_t* hp_full = bars + 12;
    uint64_t* h_full = bars + 13;
    uint64_t* y_full = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, j = blockIdx.y;
    const int C = a.C, CP = a.CP;
    const int KP = (C + 63) >> 6;                 // K panels of fc1
    const int NCK = (CP + 127) >> 7;              // 128-row output chunks of fc2
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            bar_init(&xn_full[s], 8); bar_init(&xn_empty[s], 1); bar_init(&w1_full[s], 1); bar_init(&w1_empty[s], 1);
            bar_init(&w2_full[s], 1); bar_init(&w2_empty[s], 1);
        }
        bar_init(hp_full, 1); bar_init(h_full, 8); bar_init(y_full, 1);
        bar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_hp = tmem, t_y = tmem + 64;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            for (int p = 0; p < KP; ++p) {
                const int s = p & 1;
                bar_wait(&w1_empty[s], ((p >> 1) & 1) ^ 1);
                bar_expect_tx(&w1_full[s], 2 * 8192);
                const size_t off = (size_t)p * a.n_pad1 * 128 + (size_t)j * 8192;
                bulk_g2s(sW1 + s * 16384, a.w1_hi + off, 8192, &w1_full[s]);
                bulk_g2s(sW1 + s * 16384 + 8192, a.w1_lo + off, 8192, &w1_full[s]);
            }
            for (int n = 0; n < NCK; ++n) {
                const int s = n & 1;
                const int rows = min(128, CP - 128 * n);
                bar_wait(&w2_empty[s], ((n >> 1) & 1) ^ 1);
                bar_expect_tx(&w2_full[s], 2 * rows * 128);
                const size_t off = (size_t)j * CP * 128 + (size_t)n * 128 * 128;
                bulk_g2s(sW2 + s * 32768, a.w2_hi + off, rows * 128, &w2_full[s]);
                bulk_g2s(sW2 + s * 32768 + 16384, a.w2_lo + off, rows * 128, &w2_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id1 = idesc_bf16(128, 64, false, false);
            for (int p = 0; p < KP; ++p) {
                const int s = p & 1;
                bar_wait(&xn_full[s], (p >> 1) & 1);
                bar_wait(&w1_full[s], (p >> 1) & 1);
                fence_after();
                const uint32_t ah = s32(sXN + s * 32768), al = ah + 16384, bh = s32(sW1 + s * 16384), bl = bh + 8192;
                const int ksteps = (min(64, CP - 64 * p) + 15) >> 4;
                for (int ks = 0; ks < ksteps; ++ks)
                    mma3(t_hp, desc_k(ah + ks * 32), desc_k(al + ks * 32), desc_k(bh + ks * 32), desc_k(bl + ks * 32), id1,
                         (p | ks) ? 1u : 0u);
                commit(&xn_empty[s]);
                commit(&w1_empty[s]);
            }
            commit(hp_full);
            bar_wait(h_full, 0);
            fence_after();
            const uint32_t hh = s32(sH), hl = hh + 16384;
            for (int n = 0; n < NCK; ++n) {
                const int s = n & 1;
                const int rows = min(128, CP - 128 * n);
                bar_wait(&w2_full[s], (n >> 1) & 1);
                fence_after();
                const uint32_t id2 = idesc_bf16(128, rows, false, false);
                const uint32_t bh = s32(sW2 + s * 32768), bl = bh + 16384;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma3(t_y + 128 * n, desc_k(hh + ks * 32), desc_k(hl + ks * 32), desc_k(bh + ks * 32), desc_k(bl + ks * 32), id2,
                         ks ? 1u : 0u);
                commit(&w2_empty[s]);
            }
            commit(y_full);
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int64_t grow = (int64_t)tile * 128 + row;
        const bool ok = grow < a.T;
        const float* xr = a.x + grow * C;
        // ---- LayerNorm statistics: two passes over the row, each thread owns half of the columns
        const int h0 = half * (C >> 1), h1 = h0 + (C >> 1);
        float s = 0.f;
        if (ok)
            for (int c = h0; c < h1; c += 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c)); s += (v.x + v.y) + (v.z + v.w); }
        spart[row * 2 + half] = s;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mean = (spart[row * 2] + spart[row * 2 + 1]) / (float)C;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float ss = 0.f;
        if (ok)
            for (int c = h0; c < h1; c += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
                const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
                ss += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
        spart[row * 2 + half] = ss;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float rstd = rsqrtf((spart[row * 2] + spart[row * 2 + 1]) / (float)C + a.eps);
        // ---- LayerNorm output, one 64-column panel at a time into the 2-slot ring
        for (int p = 0; p < KP; ++p) {
            const int sl = p & 1;
            bar_wait(&xn_empty[sl], ((p >> 1) & 1) ^ 1);
            uint8_t* th = sXN + sl * 32768;
            uint8_t* tl = th + 16384;
            const int c0 = 64 * p + 32 * half;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float v8[8];
                const int c = c0 + 8 * cc;
                if (ok && c < C) {
                    const float4 xa = __ldg(reinterpret_cast<const float4*>(xr + c)), xb = __ldg(reinterpret_cast<const float4*>(xr + c + 4));
                    const float4 ga = __ldg(reinterpret_cast<const float4*>(a.gamma + c)), gb = __ldg(reinterpret_cast<const float4*>(a.gamma + c + 4));
                    const float4 ba = __ldg(reinterpret_cast<const float4*>(a.beta + c)), bb = __ldg(reinterpret_cast<const float4*>(a.beta + c + 4));
                    v8[0] = (xa.x - mean) * rstd * ga.x + ba.x; v8[1] = (xa.y - mean) * rstd * ga.y + ba.y;
                    v8[2] = (xa.z - mean) * rstd * ga.z + ba.z; v8[3] = (xa.w - mean) * rstd * ga.w + ba.w;
                    v8[4] = (xb.x - mean) * rstd * gb.x + bb.x; v8[5] = (xb.y - mean) * rstd * gb.y + bb.y;
                    v8[6] = (xb.z - mean) * rstd * gb.z + bb.z; v8[7] = (xb.w - mean) * rstd * gb.w + bb.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v8[e] = 0.f;
                }
                store_chunk(th, tl, row, half * 4 + cc, v8);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(&xn_full[sl]);
        }
        // ---- h_j = GELU(hpre_j + b1_j) -> A tile of fc2
        bar_wait(hp_full, 0);
        fence_after();
        {
            float v[32];
            ld32(t_hp + lane_base + half * 32, v);
            ld_wait();
            const int hid0 = j * 64 + half * 32;
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = gelu_fast(v[e] + (hid0 + e < a.HID ? __ldg(a.b1 + hid0 + e) : 0.f));
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) store_chunk(sH, sH + 16384, row, half * 4 + cc, v + 8 * cc);
        }
        fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bar_arrive(h_full);
        // ---- y += rowscale * partial (chunk 0 also adds the residual and rowscale * b2)
        bar_wait(y_full, 0);
        fence_after();
        const float rs = (ok && a.rowscale) ? a.rowscale[grow / a.rps] : 1.f;
        float* yr = a.y + grow * C;
        for (int g = half; g * 32 < C; g += 2) {
            float v[32];
            ld32(t_y + lane_base + g * 32, v);
            ld_wait();
            if (ok) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const int c = g * 32 + e;
                    if (c < C) {
                        float4 o = make_float4(rs * v[e], rs * v[e + 1], rs * v[e + 2], rs * v[e + 3]);
                        if (j == 0) {
                            const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + c));
                            const float4 b = __ldg(reinterpret_cast<const float4*>(a.b2 + c));
                            o.x += xv.x + rs * b.x; o.y += xv.y + rs * b.y; o.z += xv.z + rs * b.z; o.w += xv.w + rs * b.w;
                        }
                        atomicAdd(reinterpret_cast<float4*>(yr + c), o);
                    }
                }
            }
        }
        fence_before();
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace mic

using namespace mic;

extern "C" int mic_mlp_split_fwd(const float* x, float* y_zeroed, const float* gamma, const float* beta, const float* b1,
                                 const float* b2, const void* w1_hi, const void* w1_lo, const void* w2_hi, const void* w2_lo,
                                 const float* rowscale, int rows_per_sample, int T, int C, int HID, float eps, void* stream) {
    MIC_REQUIRE(x && y_zeroed && gamma && beta && b1 && b2 && w1_hi && w1_lo && w2_hi && w2_lo && T > 0, "mlp_split_fwd: bad arguments");
    if (C % 8 || C < 64 || C > 384 || HID % 64 || HID != 4 * C)
        return fail(MIC_ERR_UNSUPPORTED, "mlp_split_fwd: C=%d HID=%d not taken (C %% 8 == 0, 64 <= C <= 384, HID = 4C, HID %% 64 == 0)", C, HID);
    MlpSplitFwdArgs a;
    a.x = x; a.y = y_zeroed; a.gamma = gamma; a.beta = beta; a.b1 = b1; a.b2 = b2;
    a.w1_hi = (const uint8_t*)w1_hi; a.w1_lo = (const uint8_t*)w1_lo; a.w2_hi = (const uint8_t*)w2_hi; a.w2_lo = (const uint8_t*)w2_lo;
    a.rowscale = rowscale; a.rps = rows_per_sample > 0 ? rows_per_sample : 1;
    a.T = T; a.C = C; a.HID = HID; a.CP = (C + 15) / 16 * 16; a.n_pad1 = (HID + 63) / 64 * 64; a.eps = eps;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(mlp_split_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MsF::SMEM) != cudaSuccess) {
            cudaGetLastError();
            return MIC_ERR_UNSUPPORTED;
        }
        attr = true;
    }
    dim3 grid((T + 127) / 128, HID / 64);
    mic::launch(mlp_split_fwd_kernel, grid, dim3(MS_THREADS), (size_t)MsF::SMEM, (cudaStream_t)stream, a);
    return check_launch("mlp_split_fwd_kernel");
}
