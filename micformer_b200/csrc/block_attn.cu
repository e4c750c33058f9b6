// Fused attention half of a transformer block for 2x2x2 windows (the train config; reference M:473-499 self,
// M:339-401 + M:179-203 cross):
//     x1 = x + rowscale * proj( window_attention( q = Wq LN(x) + bq,  [k|v] = Wkv src + bkv ) )
// with src = LN(x) (TransformerBlock3D) or the deformably resampled other modality (CrossTransformerBlock3D; only the
// query stream is normalised).  One persistent tcgen05 kernel: LayerNorm, the q / kv projections, the 8-token
// softmax attention and the output projection + residual never leave the SM.  Nothing is saved for the backward
// (block_attn_bwd.cu recomputes from x).
//
// A tile is 16 windows = 128 rows; row r = 8 * window + token, token = (dz, dy, dx) of the 2x2x2 window, so the 8 tokens
// of a window are 8 consecutive lanes of one warp and keys / values are exchanged with warp shuffles (no shared memory).
// Roles: warp 0 = weight loader (pre-swizzled bf16 images, bulk copies, once per CTA), warp 1 = MMA issuer, then 4*HEADS
// row warps: TMEM lane quarter = warp & 3, one head per warp.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"
#include "block_attn.cuh"

namespace mic {
using namespace t5;

template <int C, int HD>
__global__ void __launch_bounds__(AttnCfg<C, HD>::THREADS, 1) attn_block_fwd_kernel(const AttnFwdArgs a) {
    using K = AttnCfg<C, HD>;
    constexpr int CP = K::CP, HEADS = K::HEADS, TILE = K::TILE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sXNh = smem + K::F_XN;  uint8_t* sXNl = sXNh + TILE;      // LN(x); reused for the attention output o
    uint8_t* sSPh = smem + K::F_SP;  uint8_t* sSPl = sSPh + TILE;      // k/v source of a cross block
    uint8_t* sWq = smem + K::F_WQ;   uint8_t* sWkv = smem + K::F_WKV;  uint8_t* sWp = smem + K::F_WP;   // hi then lo each
    float* spar = reinterpret_cast<float*>(smem + K::F_PAR);            // gamma[C] beta[C] bq[C] bkv[2C] bp[C]
    float* sg = spar; float* sbt = sg + C; float* sbq = sbt + C; float* sbkv = sbq + C; float* sbp = sbkv + 2 * C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::F_BAR);
    uint64_t* w_full = bars + 0; uint64_t* a_full = bars + 1; uint64_t* qkv_full = bars + 2; uint64_t* o_full = bars + 3;
    uint64_t* y_full = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool cross = a.kvsrc != nullptr;
    if (threadIdx.x == 0) {
        bar_init(w_full, 1); bar_init(a_full, K::ROW_WARPS); bar_init(qkv_full, 1); bar_init(o_full, K::ROW_WARPS); bar_init(y_full, 1);
        bar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, K::F_TCOLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            bar_expect_tx(w_full, 4 * K::WC_BYTES + 2 * K::WKV_BYTES);
            bulk_g2s(sWq, a.wq_hi, K::WC_BYTES, w_full);   bulk_g2s(sWq + K::WC_BYTES, a.wq_lo, K::WC_BYTES, w_full);
            bulk_g2s(sWkv, a.wkv_hi, K::WKV_BYTES, w_full); bulk_g2s(sWkv + K::WKV_BYTES, a.wkv_lo, K::WKV_BYTES, w_full);
            bulk_g2s(sWp, a.wp_hi, K::WC_BYTES, w_full);   bulk_g2s(sWp + K::WC_BYTES, a.wp_lo, K::WC_BYTES, w_full);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t id_c = idesc_bf16(128, CP, false, false);
            constexpr uint32_t id_kv = idesc_bf16(128, 2 * C, false, false);
            bar_wait(w_full, 0);
            uint32_t n = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                bar_wait(a_full, n & 1);
                fence_after();
                const uint32_t kvh = s32(cross ? sSPh : sXNh), kvl = s32(cross ? sSPl : sXNl);
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks) {
                    mma3(tmem + K::T_Q, desc_k(s32(sXNh) + ks * 32), desc_k(s32(sXNl) + ks * 32), desc_k(s32(sWq) + ks * 32),
                         desc_k(s32(sWq) + K::WC_BYTES + ks * 32), id_c, ks ? 1u : 0u);
                    mma3(tmem + K::T_K, desc_k(kvh + ks * 32), desc_k(kvl + ks * 32), desc_k(s32(sWkv) + ks * 32),
                         desc_k(s32(sWkv) + K::WKV_BYTES + ks * 32), id_kv, ks ? 1u : 0u);
                }
                commit(qkv_full);
                bar_wait(o_full, n & 1);
                fence_after();
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks)
                    mma3(tmem + K::T_Q, desc_k(s32(sXNh) + ks * 32), desc_k(s32(sXNl) + ks * 32), desc_k(s32(sWp) + ks * 32),
                         desc_k(s32(sWp) + K::WC_BYTES + ks * 32), id_c, ks ? 1u : 0u);
                commit(y_full);
            }
        }
    } else {
        const int q = warp & 3, hh = (warp - 2) >> 2;        // TMEM lane quarter, head of this warp
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int wbase = lane & ~7;                         // first lane of this thread's window
        for (int i = threadIdx.x - 64; i < 6 * C; i += K::THREADS - 64) {
            float v;
            if (i < C) v = a.gamma[i];
            else if (i < 2 * C) v = a.beta[i - C];
            else if (i < 3 * C) v = a.bq[i - 2 * C];
            else if (i < 5 * C) v = a.bkv[i - 3 * C];
            else v = a.bp[i - 5 * C];
            spar[i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(K::ROW_WARPS * 32) : "memory");
        WinGeom wg(a.D, a.H, a.W);
        uint32_t n = 0;
        for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
            const int64_t grow = wg.row_of((int64_t)t * 16 + (row >> 3), row & 7, a.nwin_total);
            const bool ok = grow >= 0;
            if (hh == 0 || (hh == 1 && cross)) {          // the next tile's rows start their way to L2 now
                const int64_t gnext = t + (int)gridDim.x < a.ntiles ? wg.row_of((int64_t)(t + gridDim.x) * 16 + (row >> 3), row & 7, a.nwin_total) : -1;
                if (gnext >= 0) prefetch_l2((hh == 0 ? a.x : a.kvsrc) + gnext * C, C * 4);
            }
            // ---- operand tiles: head-0 warps LayerNorm x, head-1 warps stage the k/v source of a cross block
            if (hh == 0 || (hh == 1 && cross)) {
                float r[C];
                const float* src = hh == 0 ? a.x : a.kvsrc;
                load_row<C>(src, grow, ok, r);
                if (hh == 0) {
                    float mean, rstd;
                    ln_stats<C>(r, a.eps, mean, rstd);
#pragma unroll
                    for (int i = 0; i < C; ++i) r[i] = ok ? (r[i] - mean) * rstd * sg[i] + sbt[i] : 0.f;
                }
                store_row_tile<C>(hh == 0 ? sXNh : sSPh, hh == 0 ? sXNl : sSPl, row, r);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(a_full);
            // ---- attention of (row, head): q, k, v of this head from TMEM, keys / values of the window via shuffles
            bar_wait(qkv_full, n & 1);
            fence_after();
            float qv[HD], kv_[HD], vv[HD];
            ld_cols<HD>(tmem + lane_base + K::T_Q + hh * HD, qv);
            ld_cols<HD>(tmem + lane_base + K::T_K + hh * HD, kv_);
            ld_cols<HD>(tmem + lane_base + K::T_V + hh * HD, vv);
            ld_wait();
            fence_before();
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                qv[d] = (qv[d] + sbq[hh * HD + d]) * a.scale;
                kv_[d] += sbkv[hh * HD + d];
                vv[d] += sbkv[C + hh * HD + d];
            }
            float s[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < HD; ++d) acc = fmaf(qv[d], __shfl_sync(0xffffffffu, kv_[d], wbase + j), acc);
                s[j] = acc;
            }
            float mx = s[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) mx = fmaxf(mx, s[j]);
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] = __expf(s[j] - mx); den += s[j]; }
            const float inv = 1.f / den;
            float o[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = s[j] * inv;
#pragma unroll
                for (int d = 0; d < HD; ++d) o[d] = fmaf(p, __shfl_sync(0xffffffffu, vv[d], wbase + j), o[d]);
            }
            // o -> A tile of the projection (the LN tile is free: the q/kv MMAs completed before qkv_full)
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) store_chunk(sXNh, sXNl, row, (hh * HD) / 8 + c, o + 8 * c);
            if (hh == 0 && CP > C) {                  // zero the padding columns of the row once per tile
                float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = C / 8; c < CP / 8; ++c) store_chunk(sXNh, sXNl, row, c, z);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(o_full);
            // ---- x1 = x + rowscale * (proj + bp): head-0 warps drain the row
            bar_wait(y_full, n & 1);
            fence_after();
            if (hh == 0) {
                float pr[CP];
#pragma unroll
                for (int c0 = 0; c0 < CP; c0 += 16) ld16(tmem + lane_base + K::T_Q + c0, pr + c0);
                ld_wait();
                if (ok) {
                    const float rs = a.rowscale ? a.rowscale[grow / a.rps] : 1.f;
                    const float4* xp = reinterpret_cast<const float4*>(a.x + grow * C);
                    float4* yp = reinterpret_cast<float4*>(a.y + grow * C);
#pragma unroll
                    for (int i = 0; i < C / 4; ++i) {
                        const float4 xv = __ldg(xp + i);
                        float4 r;
                        r.x = xv.x + rs * (pr[4 * i] + sbp[4 * i]);
                        r.y = xv.y + rs * (pr[4 * i + 1] + sbp[4 * i + 1]);
                        r.z = xv.z + rs * (pr[4 * i + 2] + sbp[4 * i + 2]);
                        r.w = xv.w + rs * (pr[4 * i + 3] + sbp[4 * i + 3]);
                        yp[i] = r;
                    }
                }
            }
            fence_before();
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        tmem_dealloc(tmem, K::F_TCOLS);
    }
}

template <int C, int HD>
static int launch_attn_fwd(const AttnFwdArgs& a, cudaStream_t st) {
    using K = AttnCfg<C, HD>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_block_fwd_kernel<C, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::F_SMEM) != cudaSuccess) {
            cudaGetLastError();
            return MIC_ERR_UNSUPPORTED;
        }
        attr = true;
    }
    int grid = num_sms();
    if (grid > a.ntiles) grid = a.ntiles;
    mic::launch(attn_block_fwd_kernel<C, HD>, dim3(grid), dim3(K::THREADS), (size_t)K::F_SMEM, st, a);
    return check_launch("attn_block_fwd_kernel");
}

}  // namespace mic

using namespace mic;

extern "C" int mic_attn_block_fwd(const float* x, const float* kvsrc, float* y, const float* gamma, const float* beta,
                                  const float* bq, const float* bkv, const float* bp, const void* wq_hi, const void* wq_lo,
                                  const void* wkv_hi, const void* wkv_lo, const void* wp_hi, const void* wp_lo,
                                  const float* rowscale, int B, int D, int H, int W, int C, int heads, float scale, float eps,
                                  void* stream) {
    MIC_REQUIRE(x && y && gamma && beta && bq && bkv && bp && wq_hi && wq_lo && wkv_hi && wkv_lo && wp_hi && wp_lo,
                "attn_block_fwd: null pointer");
    MIC_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && heads > 0 && C % heads == 0, "attn_block_fwd: bad geometry");
    if ((D | H | W) & 1) return fail(MIC_ERR_UNSUPPORTED, "attn_block_fwd: the fused kernel takes even grids (2x2x2 windows, no pad)");
    AttnFwdArgs a;
    a.x = x; a.kvsrc = kvsrc; a.y = y; a.gamma = gamma; a.beta = beta; a.bq = bq; a.bkv = bkv; a.bp = bp;
    a.wq_hi = (const uint8_t*)wq_hi; a.wq_lo = (const uint8_t*)wq_lo; a.wkv_hi = (const uint8_t*)wkv_hi;
    a.wkv_lo = (const uint8_t*)wkv_lo; a.wp_hi = (const uint8_t*)wp_hi; a.wp_lo = (const uint8_t*)wp_lo;
    a.rowscale = rowscale; a.rps = D * H * W;
    a.D = D; a.H = H; a.W = W;
    a.nwin_total = (int64_t)B * (D / 2) * (H / 2) * (W / 2);
    a.ntiles = (int)((a.nwin_total + 15) / 16);
    a.scale = scale; a.eps = eps;
    const int hd = C / heads;
    if (C == 48 && hd == 16) return launch_attn_fwd<48, 16>(a, (cudaStream_t)stream);
    if (C == 48 && hd == 24) return launch_attn_fwd<48, 24>(a, (cudaStream_t)stream);
    return fail(MIC_ERR_UNSUPPORTED, "attn_block_fwd: C=%d head_dim=%d is not built ((48,16), (48,24))", C, hd);
}
