// Fused backward of the attention half-block (block_attn.cu) from dy (gradient w.r.t. x1), x and -- for a cross block --
// the k/v source alone: LayerNorm and the q / kv projections are recomputed on chip, the 8-token softmax attention is
// differentiated in registers (keys / values / queries of a window travel by warp shuffles), and ONE persistent tcgen05
// kernel produces
//     dx = dy + LN'(dq Wq [+ dkv Wkv]),   dsrc = dkv Wkv (cross),
//     dWp += (rs dy)^T o,  dWq += dq^T xn,  dWkv += dkv^T src,  dbp, dbq, dbkv, dgamma, dbeta.
// Weight-gradient accumulators stay in TMEM across all tiles of the CTA and are flushed once.
//
// Roles: warp 0 streams weight images (forward images for the recompute, then -- into the same shared memory -- the
// transposed images of the data-gradient products), warps 1-2 issue MMAs, 4*HEADS row warps (lane quarter = warp & 3, one
// head per warp).  Shared-memory regions are reused inside a tile:  dy tile -> dq tile,  o tile -> dkv tile (hi).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"
#include "block_attn.cuh"

namespace mic {
using namespace t5;

template <int C, int HD>
struct AttnBwdCfg {
    using F = AttnCfg<C, HD>;
    static constexpr int CP = F::CP, HEADS = F::HEADS, ROW_WARPS = F::ROW_WARPS, TILE = F::TILE, WC = F::WC_BYTES, WKV = F::WKV_BYTES;
    static constexpr int THREADS = 96 + 32 * ROW_WARPS;
    static constexpr int R_XN = 0, R_SP = 2 * TILE, R_DX = 4 * TILE, R_O = 6 * TILE, R_E = 8 * TILE;
    static constexpr int W_A = 10 * TILE;
    static constexpr int WA_BYTES = (2 * WC + 2 * WKV) > 6 * WC ? (2 * WC + 2 * WKV) : 6 * WC;
    static constexpr int W_P = W_A + WA_BYTES;
    static constexpr int PAR = W_P + 2 * WC;                 // gamma[C] beta[C] bq[C] bkv[2C]
    static constexpr int BAR = PAR + 4 * 5 * C;
    static constexpr int SMEM = BAR + 256 + 1024;
    static constexpr int NB = CP + 16;                       // weight-gradient tiles carry one extra "ones" column: the bias gradient
    static constexpr int T_Q = 0, T_K = CP, T_V = CP + C, T_DO = CP + 2 * C, T_DWP = 192, T_DWQ = 192 + NB, T_DWKV = 192 + 2 * NB;
    static_assert(T_DO + CP <= 192 && T_DWKV + NB <= 512 && NB <= 64, "TMEM budget / ones column inside the 64-feature panel");
    static_assert(SMEM <= 232448, "shared memory budget");
    static_assert(C % 16 == 0 && C <= 64, "the chunked LayerNorm backward walks 16-column chunks (4 accumulators)");
    static_assert(W_A % 1024 == 0 && W_P % 1024 == 0 && WC % 1024 == 0, "swizzle alignment");
};

// 8x8 transpose inside each group of 8 lanes: in: v[c] = M[l][c] at lane l -> out: v[c] = M[c][l]
__device__ __forceinline__ void transpose8(float (&v)[8], int l) {
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
        const bool up = (l & s) != 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if ((r & s) == 0) {
                const float snd = up ? v[r] : v[r | s];
                const float rcv = __shfl_xor_sync(0xffffffffu, snd, s);
                if (up) v[r] = rcv; else v[r | s] = rcv;
            }
        }
    }
}

template <int C, int HD>
__global__ void __launch_bounds__(AttnBwdCfg<C, HD>::THREADS, 1) attn_block_bwd_kernel(const AttnBwdArgs a) {
    using K = AttnBwdCfg<C, HD>;
    constexpr int CP = K::CP, HEADS = K::HEADS, TILE = K::TILE, WC = K::WC, WKV = K::WKV, NB = K::NB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sXN = smem + K::R_XN;    // hi, lo (+TILE)
    uint8_t* sSP = smem + K::R_SP;
    uint8_t* sDX = smem + K::R_DX;    // rs*dy tile, later the dq tile
    uint8_t* sO = smem + K::R_O;      // o tile (hi, lo), later dkv hi (two panels)
    uint8_t* sE = smem + K::R_E;      // dkv lo (two panels)
    uint8_t* sWA = smem + K::W_A;
    uint8_t* sWP = smem + K::W_P;
    float* sg = reinterpret_cast<float*>(smem + K::PAR);
    float* sbt = sg + C; float* sbq = sbt + C; float* sbkv = sbq + C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::BAR);
    uint64_t* wnk_full = bars + 0; uint64_t* wkn_full = bars + 1; uint64_t* wp_full = bars + 2; uint64_t* a_full = bars + 3;
    uint64_t* g1 = bars + 4; uint64_t* o_full = bars + 5; uint64_t* dwp_done = bars + 6; uint64_t* dqkv_full = bars + 7;
    uint64_t* g2 = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool cross = a.kvsrc != nullptr;
    if (threadIdx.x == 0) {
        bar_init(wnk_full, 1); bar_init(wkn_full, 1); bar_init(wp_full, 1); bar_init(a_full, K::ROW_WARPS); bar_init(g1, 2);
        bar_init(o_full, K::ROW_WARPS); bar_init(dwp_done, 1); bar_init(dqkv_full, K::ROW_WARPS); bar_init(g2, 2);
        bar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_sync();

    constexpr uint32_t id_c = idesc_bf16(128, CP, false, false);
    constexpr uint32_t id_kv = idesc_bf16(128, 2 * C, false, false);
    constexpr uint32_t id_dw = idesc_bf16(128, NB, true, true);

    if (warp == 0) {
        if (lane == 0) {
            bar_expect_tx(wp_full, 2 * WC);
            bulk_g2s(sWP, a.wpT_hi, WC, wp_full); bulk_g2s(sWP + WC, a.wpT_lo, WC, wp_full);
            uint32_t n = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                if (n > 0) bar_wait(g2, (n - 1) & 1);                 // the transposed images of the previous tile are consumed
                bar_expect_tx(wnk_full, 2 * WC + 2 * WKV);
                bulk_g2s(sWA, a.wq_hi, WC, wnk_full);               bulk_g2s(sWA + WC, a.wq_lo, WC, wnk_full);
                bulk_g2s(sWA + 2 * WC, a.wkv_hi, WKV, wnk_full);    bulk_g2s(sWA + 2 * WC + WKV, a.wkv_lo, WKV, wnk_full);
                bar_wait(g1, n & 1);                                  // recompute products done: the forward images are dead
                bar_expect_tx(wkn_full, 6 * WC);
                bulk_g2s(sWA, a.wqT_hi, WC, wkn_full);              bulk_g2s(sWA + WC, a.wqT_lo, WC, wkn_full);
                bulk_g2s(sWA + 2 * WC, a.wkvT_hi, 2 * WC, wkn_full); bulk_g2s(sWA + 4 * WC, a.wkvT_lo, 2 * WC, wkn_full);
            }
        }
    } else if (warp == 1) {
        // ---------------- issuer A: q / kv recompute, then dxn (and dsrc), dWq
        if (lane == 0) {
            uint32_t n = 0;
            const uint32_t xh = s32(sXN), xl = xh + TILE;
            const uint32_t kvh = cross ? s32(sSP) : xh, kvl = kvh + TILE;
            const uint32_t wa = s32(sWA);
            const uint32_t qh = s32(sDX), ql = qh + TILE;            // dq tile
            const uint32_t dh = s32(sO), dl = s32(sE);               // dkv tile (two panels each)
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                bar_wait(a_full, n & 1);
                bar_wait(wnk_full, n & 1);
                fence_after();
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks) {
                    mma3(tmem + K::T_Q, desc_k(xh + ks * 32), desc_k(xl + ks * 32), desc_k(wa + ks * 32), desc_k(wa + WC + ks * 32),
                         id_c, ks ? 1u : 0u);
                    mma3(tmem + K::T_K, desc_k(kvh + ks * 32), desc_k(kvl + ks * 32), desc_k(wa + 2 * WC + ks * 32),
                         desc_k(wa + 2 * WC + WKV + ks * 32), id_kv, ks ? 1u : 0u);
                }
                commit(g1);
                bar_wait(dqkv_full, n & 1);
                bar_wait(wkn_full, n & 1);
                fence_after();
                // dxn = dq WqT (+ dkv WkvT for a self block);  cross: dsrc = dkv WkvT into the (dead) k columns
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks)
                    mma3(tmem + K::T_Q, desc_k(qh + ks * 32), desc_k(ql + ks * 32), desc_k(wa + ks * 32), desc_k(wa + WC + ks * 32),
                         id_c, ks ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < (2 * C) / 16; ++ks) {
                    const uint32_t ao = (ks >> 2) * TILE + (ks & 3) * 32, bo = (ks >> 2) * WC + (ks & 3) * 32;
                    mma3(tmem + (cross ? K::T_K : K::T_Q), desc_k(dh + ao), desc_k(dl + ao), desc_k(wa + 2 * WC + bo),
                         desc_k(wa + 4 * WC + bo), id_c, (cross && ks == 0) ? 0u : 1u);
                }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma3(tmem + K::T_DWQ, desc_mn(qh + ks * 2048, TILE), desc_mn(ql + ks * 2048, TILE), desc_mn(xh + ks * 2048, TILE),
                         desc_mn(xl + ks * 2048, TILE), id_dw, (n | ks) ? 1u : 0u);
                commit(g2);
            }
        }
    } else if (warp == 2) {
        // ---------------- issuer B: do = (rs dy) WpT, dWp, dWkv
        if (lane == 0) {
            uint32_t n = 0;
            const uint32_t xh = s32(sXN);
            const uint32_t sph = cross ? s32(sSP) : xh, spl = sph + TILE;
            const uint32_t yh = s32(sDX), yl = yh + TILE;
            const uint32_t oh = s32(sO), ol = oh + TILE;
            const uint32_t dh = s32(sO), dl = s32(sE);
            const uint32_t wp = s32(sWP);
            bar_wait(wp_full, 0);
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                bar_wait(a_full, n & 1);
                fence_after();
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks)
                    mma3(tmem + K::T_DO, desc_k(yh + ks * 32), desc_k(yl + ks * 32), desc_k(wp + ks * 32), desc_k(wp + WC + ks * 32),
                         id_c, ks ? 1u : 0u);
                commit(g1);
                bar_wait(o_full, n & 1);
                fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma3(tmem + K::T_DWP, desc_mn(yh + ks * 2048, TILE), desc_mn(yl + ks * 2048, TILE), desc_mn(oh + ks * 2048, TILE),
                         desc_mn(ol + ks * 2048, TILE), id_dw, (n | ks) ? 1u : 0u);
                commit(dwp_done);
                bar_wait(dqkv_full, n & 1);
                fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma3(tmem + K::T_DWKV, desc_mn(dh + ks * 2048, TILE), desc_mn(dl + ks * 2048, TILE), desc_mn(sph + ks * 2048, TILE),
                         desc_mn(spl + ks * 2048, TILE), id_dw, (n | ks) ? 1u : 0u);
                commit(g2);
            }
        }
    } else {
        // ---------------- row threads: (row, head)
        const int q = warp & 3, hh = (warp - 3) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int wbase = lane & ~7, wl = lane & 7;
        constexpr int STG_X = 0, STG_DY = 1, STG_SP = HEADS - 1 == 1 ? 0 : HEADS - 1;       // which head's warps stage which tile
        for (int i = threadIdx.x - 96; i < 5 * C; i += K::THREADS - 96) {
            float v;
            if (i < C) v = a.gamma[i];
            else if (i < 2 * C) v = a.beta[i - C];
            else if (i < 3 * C) v = a.bq[i - 2 * C];
            else v = a.bkv[i - 3 * C];
            sg[i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(K::ROW_WARPS * 32) : "memory");
        WinGeom wg(a.D, a.H, a.W);
        float cacc0[2] = {0.f, 0.f}, cacc1[2] = {0.f, 0.f};      // STG_X warps: dgamma, dbeta partial sums
        uint32_t n = 0;
        for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
            const int64_t grow = wg.row_of((int64_t)t * 16 + (row >> 3), row & 7, a.nwin_total);
            const bool ok = grow >= 0;
            float mean = 0.f, rstd = 0.f;
            const bool tr = threadIdx.x == 96 + 32;       // (warp 4: quarter 0, head 0, lane 0 stamps the phase trace)
            if (tr) t5_trace(n, 0);
            {   // the next tile's rows start their way to L2 now
                const int64_t gnext = t + (int)gridDim.x < a.ntiles ? wg.row_of((int64_t)(t + gridDim.x) * 16 + (row >> 3), row & 7, a.nwin_total) : -1;
                if (gnext >= 0) {
                    if (hh == STG_X) prefetch_l2(a.x + gnext * C, C * 4);
                    if (hh == STG_DY) prefetch_l2(a.dy + gnext * C, C * 4);
                    if (hh == STG_SP && cross) prefetch_l2(a.kvsrc + gnext * C, C * 4);
                }
            }
            // ---- stage the operand tiles (xn and the k/v source carry the ones column of the bias gradients)
            if (hh == STG_X) {
                float r[C];
                load_row<C>(a.x, grow, ok, r);
                ln_stats<C>(r, a.eps, mean, rstd);
#pragma unroll
                for (int i = 0; i < C; ++i) r[i] = ok ? (r[i] - mean) * rstd * sg[i] + sbt[i] : 0.f;
                store_row_tile<C>(sXN, sXN + TILE, row, r);
                store_ones_chunk(sXN, sXN + TILE, row, CP / 8);
            }
            if (hh == STG_DY) {
                float r[C];
                load_row<C>(a.dy, grow, ok, r);
                const float rs = (ok && a.rowscale) ? a.rowscale[grow / a.rps] : 1.f;
#pragma unroll
                for (int i = 0; i < C; ++i) r[i] *= rs;
                store_row_tile<C>(sDX, sDX + TILE, row, r);
            }
            if (hh == STG_SP && cross) {
                float r[C];
                load_row<C>(a.kvsrc, grow, ok, r);
                store_row_tile<C>(sSP, sSP + TILE, row, r);
                store_ones_chunk(sSP, sSP + TILE, row, CP / 8);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(a_full);
            if (tr) t5_trace(n, 1);
            // ---- q, k, v, do of (row, head)
            bar_wait(g1, n & 1);
            fence_after();
            if (tr) t5_trace(n, 2);
            float qv[HD], kv_[HD], vv[HD], dov[HD];
            ld_cols<HD>(tmem + lane_base + K::T_Q + hh * HD, qv);
            ld_cols<HD>(tmem + lane_base + K::T_K + hh * HD, kv_);
            ld_cols<HD>(tmem + lane_base + K::T_V + hh * HD, vv);
            ld_cols<HD>(tmem + lane_base + K::T_DO + hh * HD, dov);
            ld_wait();
            fence_before();
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                qv[d] = (qv[d] + sbq[hh * HD + d]) * a.scale;
                kv_[d] += sbkv[hh * HD + d];
                vv[d] += sbkv[C + hh * HD + d];
            }
            // ---- as query i: scores, probabilities, dS; o and dq
            float p[8], ds[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s = 0.f, dp = 0.f;
#pragma unroll
                for (int d = 0; d < HD; ++d) {
                    s = fmaf(qv[d], __shfl_sync(0xffffffffu, kv_[d], wbase + j), s);
                    dp = fmaf(dov[d], __shfl_sync(0xffffffffu, vv[d], wbase + j), dp);
                }
                p[j] = s; ds[j] = dp;
            }
            {
                float mx = p[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) mx = fmaxf(mx, p[j]);
                float den = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { p[j] = __expf(p[j] - mx); den += p[j]; }
                const float inv = 1.f / den;
                float delta = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { p[j] *= inv; delta = fmaf(p[j], ds[j], delta); }
#pragma unroll
                for (int j = 0; j < 8; ++j) ds[j] = p[j] * (ds[j] - delta);
            }
            float dq[HD];
            {
                float o[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) { o[d] = 0.f; dq[d] = 0.f; }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int d = 0; d < HD; ++d) {
                        o[d] = fmaf(p[j], __shfl_sync(0xffffffffu, vv[d], wbase + j), o[d]);
                        dq[d] = fmaf(ds[j], __shfl_sync(0xffffffffu, kv_[d], wbase + j), dq[d]);
                    }
                }
#pragma unroll
                for (int c = 0; c < HD / 8; ++c) store_chunk(sO, sO + TILE, row, (hh * HD) / 8 + c, o + 8 * c);
                if (hh == 0) {
                    float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = C / 8; c < CP / 8; ++c) store_chunk(sO, sO + TILE, row, c, z);
                    store_ones_chunk(sO, sO + TILE, row, CP / 8);          // dbp = column sums of rs*dy
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(o_full);
            if (tr) t5_trace(n, 3);
#pragma unroll
            for (int d = 0; d < HD; ++d) dq[d] *= a.scale;           // q = scale * (Wq xn + bq)
            // ---- as key j: dk_j = sum_i dS_ij q_i,  dv_j = sum_i P_ij do_i  (P, dS transposed inside the window)
            transpose8(p, wl);
            transpose8(ds, wl);
            float dk[HD], dv[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int d = 0; d < HD; ++d) {
                    dk[d] = fmaf(ds[i], __shfl_sync(0xffffffffu, qv[d], wbase + i), dk[d]);
                    dv[d] = fmaf(p[i], __shfl_sync(0xffffffffu, dov[d], wbase + i), dv[d]);
                }
            }
            // ---- dq / dk / dv -> operand tiles (they reuse the dy and o tiles: wait until dWp has read those)
            if (tr) t5_trace(n, 4);
            bar_wait(dwp_done, n & 1);
            if (tr) t5_trace(n, 5);
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                const int cq = (hh * HD) / 8 + c;                    // chunk of dq / dk inside [0, C); dv sits at C + ...
                store_chunk(sDX, sDX + TILE, row, cq, dq + 8 * c);
                const int ck = cq, cv = C / 8 + cq;
                store_chunk(sO + (ck >> 3) * TILE, sE + (ck >> 3) * TILE, row, ck & 7, dk + 8 * c);
                store_chunk(sO + (cv >> 3) * TILE, sE + (cv >> 3) * TILE, row, cv & 7, dv + 8 * c);
            }
            if (hh == 0) {
                float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = C / 8; c < CP / 8; ++c) store_chunk(sDX, sDX + TILE, row, c, z);
#pragma unroll
                for (int c = (2 * C) / 8; c < 16; ++c) store_chunk(sO + (c >> 3) * TILE, sE + (c >> 3) * TILE, row, c & 7, z);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(dqkv_full);
            if (tr) t5_trace(n, 6);
            // ---- tile end: dxn (and dsrc) complete
            bar_wait(g2, n & 1);
            fence_after();
            if (tr) t5_trace(n, 7);
            if (hh == STG_X) {
                // LayerNorm backward in 16-column chunks (two passes: the row means first), so that only a few dozen registers are
                // live: the one-shot version spilled under the 128-register cap and took a third of the tile time
                const float* xp = a.x + grow * C;
                const float* dp_ = a.dy + grow * C;
                float m1 = 0.f, m2 = 0.f;
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 16) {
                    float gx[16];
                    ld16(tmem + lane_base + K::T_Q + c0, gx);
                    ld_wait();
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + c0 + i));
                            const float d0 = gx[i] * sg[c0 + i], d1 = gx[i + 1] * sg[c0 + i + 1], d2 = gx[i + 2] * sg[c0 + i + 2],
                                        d3 = gx[i + 3] * sg[c0 + i + 3];
                            m1 += (d0 + d1) + (d2 + d3);
                            m2 = fmaf(d0, (xv.x - mean) * rstd, m2); m2 = fmaf(d1, (xv.y - mean) * rstd, m2);
                            m2 = fmaf(d2, (xv.z - mean) * rstd, m2); m2 = fmaf(d3, (xv.w - mean) * rstd, m2);
                        }
                    }
                }
                m1 *= (1.f / C); m2 *= (1.f / C);
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 16) {
                    float gx[16], v[32];
                    ld16(tmem + lane_base + K::T_Q + c0, gx);
                    ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv4 = xv;
                        if (ok) { xv = __ldg(reinterpret_cast<const float4*>(xp + c0 + i)); dv4 = __ldg(reinterpret_cast<const float4*>(dp_ + c0 + i)); }
                        const float xh0 = (xv.x - mean) * rstd, xh1 = (xv.y - mean) * rstd, xh2 = (xv.z - mean) * rstd, xh3 = (xv.w - mean) * rstd;
                        const float g0 = ok ? gx[i] : 0.f, g1_ = ok ? gx[i + 1] : 0.f, g2_ = ok ? gx[i + 2] : 0.f, g3 = ok ? gx[i + 3] : 0.f;
                        if (ok) {
                            float4 o;
                            o.x = dv4.x + rstd * (g0 * sg[c0 + i] - m1 - xh0 * m2);
                            o.y = dv4.y + rstd * (g1_ * sg[c0 + i + 1] - m1 - xh1 * m2);
                            o.z = dv4.z + rstd * (g2_ * sg[c0 + i + 2] - m1 - xh2 * m2);
                            o.w = dv4.w + rstd * (g3 * sg[c0 + i + 3] - m1 - xh3 * m2);
                            *reinterpret_cast<float4*>(a.dx + grow * C + c0 + i) = o;
                        }
                        v[i] = g0 * xh0; v[i + 1] = g1_ * xh1; v[i + 2] = g2_ * xh2; v[i + 3] = g3 * xh3;      // dgamma terms
                        v[16 + i] = g0; v[16 + i + 1] = g1_; v[16 + i + 2] = g2_; v[16 + i + 3] = g3;             // dbeta terms
                    }
                    // lanes 0..15 end up with the dgamma column sums of this chunk, lanes 16..31 with the dbeta ones
                    const float cs = warp_colsum32(v, lane);
                    if (c0 == 0) cacc0[0] += cs; else if (c0 == 16) cacc0[1] += cs; else if (c0 == 32) cacc1[0] += cs; else cacc1[1] += cs;
                }
            } else if (hh == STG_DY && cross) {
                float gs[CP];
#pragma unroll
                for (int c0 = 0; c0 < CP; c0 += 16) ld16(tmem + lane_base + K::T_K + c0, gs + c0);
                ld_wait();
                if (ok) {
                    float4* po = reinterpret_cast<float4*>(a.dkvsrc + grow * C);
#pragma unroll
                    for (int i = 0; i < C / 4; ++i) po[i] = make_float4(gs[4 * i], gs[4 * i + 1], gs[4 * i + 2], gs[4 * i + 3]);
                }
            }
            fence_before();
            if (tr) t5_trace(n, 8);
        }
        if (threadIdx.x == 96 + 32) t5_trace(7, 0);
        // ---------------- flush
        if (hh == STG_X) {                   // chunk k of 16 columns: lanes 0..15 hold dgamma, lanes 16..31 dbeta
            const float accs[4] = {cacc0[0], cacc0[1], cacc1[0], cacc1[1]};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = 16 * k + (lane & 15);
                if (16 * k < C && col < C) atomicAdd((lane < 16 ? a.dgamma : a.dbeta) + col, accs[k]);
            }
        }
        {
            // weight-gradient accumulators: rows = output feature (TMEM lane), columns 0..C-1 = input feature, column CP = the
            // product with the ones column = the bias gradient
            float v[NB];
            auto flush = [&](uint32_t tcol, float* d, float* db, int rows) {
#pragma unroll
                for (int c0 = 0; c0 < NB; c0 += 16) ld16(tmem + lane_base + tcol + c0, v + c0);
                ld_wait();
                if (row < rows) {
#pragma unroll
                    for (int i = 0; i < C; ++i) atomicAdd(d + (int64_t)row * C + i, v[i]);
                    atomicAdd(db + row, v[CP]);
                }
            };
            if (HEADS >= 3) {
                if (hh == 0) flush(K::T_DWP, a.dWp, a.dbp, C);
                else if (hh == 1) flush(K::T_DWQ, a.dWq, a.dbq, C);
                else flush(K::T_DWKV, a.dWkv, a.dbkv, 2 * C);
            } else {                                   // two heads: head-0 warps flush dWp and dWkv
                if (hh == 0) { flush(K::T_DWP, a.dWp, a.dbp, C); flush(K::T_DWKV, a.dWkv, a.dbkv, 2 * C); }
                else flush(K::T_DWQ, a.dWq, a.dbq, C);
            }
        }
    }
    if (threadIdx.x == 96 + 32) t5_trace(7, 1);
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int C, int HD>
static int launch_attn_bwd(const AttnBwdArgs& a, cudaStream_t st) {
    using K = AttnBwdCfg<C, HD>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_block_bwd_kernel<C, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM) != cudaSuccess) {
            cudaGetLastError();
            return MIC_ERR_UNSUPPORTED;
        }
        attr = true;
    }
    int grid = num_sms();
    if (grid > a.ntiles) grid = a.ntiles;
    mic::launch(attn_block_bwd_kernel<C, HD>, dim3(grid), dim3(K::THREADS), (size_t)K::SMEM, st, a);
    return check_launch("attn_block_bwd_kernel");
}

}  // namespace mic

using namespace mic;

extern "C" int mic_debug_t5_trace(void* buf) {       // phase trace of the attention backward kernel (this translation unit)
    unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(mic::t5::g_t5_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}

extern "C" int mic_attn_block_bwd(const float* x, const float* kvsrc, const float* dy, float* dx, float* dkvsrc,
                                  const float* gamma, const float* beta, const float* bq, const float* bkv, const void* const* imgs,
                                  const float* rowscale, float* dgamma, float* dbeta, float* dWq, float* dbq, float* dWkv,
                                  float* dbkv, float* dWp, float* dbp, int B, int D, int H, int W, int C, int heads, float scale,
                                  float eps, void* stream) {
    MIC_REQUIRE(x && dy && dx && gamma && beta && bq && bkv && imgs && dgamma && dbeta && dWq && dbq && dWkv && dbkv && dWp && dbp,
                "attn_block_bwd: null pointer");
    MIC_REQUIRE((kvsrc == nullptr) == (dkvsrc == nullptr), "attn_block_bwd: kvsrc and dkvsrc go together");
    MIC_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && heads > 0 && C % heads == 0, "attn_block_bwd: bad geometry");
    if ((D | H | W) & 1) return fail(MIC_ERR_UNSUPPORTED, "attn_block_bwd: the fused kernel takes even grids (2x2x2 windows, no pad)");
    for (int i = 0; i < 10; ++i) MIC_REQUIRE(imgs[i] && (reinterpret_cast<uintptr_t>(imgs[i]) & 15) == 0, "attn_block_bwd: image %d", i);
    AttnBwdArgs a;
    a.x = x; a.kvsrc = kvsrc; a.dy = dy; a.dx = dx; a.dkvsrc = dkvsrc; a.gamma = gamma; a.beta = beta; a.bq = bq; a.bkv = bkv;
    const uint8_t* const* im = reinterpret_cast<const uint8_t* const*>(imgs);
    a.wq_hi = im[0]; a.wq_lo = im[1]; a.wkv_hi = im[2]; a.wkv_lo = im[3]; a.wpT_hi = im[4]; a.wpT_lo = im[5];
    a.wqT_hi = im[6]; a.wqT_lo = im[7]; a.wkvT_hi = im[8]; a.wkvT_lo = im[9];
    a.rowscale = rowscale; a.rps = D * H * W;
    a.dgamma = dgamma; a.dbeta = dbeta; a.dWq = dWq; a.dbq = dbq; a.dWkv = dWkv; a.dbkv = dbkv; a.dWp = dWp; a.dbp = dbp;
    a.D = D; a.H = H; a.W = W;
    a.nwin_total = (int64_t)B * (D / 2) * (H / 2) * (W / 2);
    a.ntiles = (int)((a.nwin_total + 15) / 16);
    a.scale = scale; a.eps = eps;
    const int hd = C / heads;
    if (C == 48 && hd == 16) return launch_attn_bwd<48, 16>(a, (cudaStream_t)stream);
    if (C == 48 && hd == 24) return launch_attn_bwd<48, 24>(a, (cudaStream_t)stream);
    return fail(MIC_ERR_UNSUPPORTED, "attn_block_bwd: C=%d head_dim=%d is not built ((48,16), (48,24))", C, hd);
}
