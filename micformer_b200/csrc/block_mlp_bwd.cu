// Fused backward of the MLP half of a transformer block (see block_mlp.cu; reference M:28-34, 403-404, 419-424 under
// autograd): from dy and the block input x ONLY -- LayerNorm, fc1 and GELU are recomputed on chip -- one persistent
// tcgen05 kernel produces
//     dx = dy + LN'(dxn),  dW1 += dh^T xn,  db1 += colsum dh,  dW2 += (rs dy)^T h,  db2 += colsum rs dy,  dgamma, dbeta
// with   xn = LN(x),  hpre = xn W1^T + b1,  h = GELU(hpre),  dh = ((rs dy) W2) * GELU'(hpre),  dxn = dh W1.
// The 4C-wide tensors hpre / h / dh exist only per 64-column chunk in TMEM / shared memory; the weight-gradient
// accumulators stay in TMEM across all tiles of the CTA and are flushed once (atomics) at the end.
//
// Roles (384 threads): warp 0 streams the weight-image chunks (bulk copies, 2-slot ring); warps 1-3 are three MMA issuers
// (a tcgen05.mma costs its issuing thread >= ~100 clocks whatever its size, profiles/r02_ubench_b200.txt, so the independent
// accumulation chains are spread over three threads); warps 4-11 are row threads (TMEM lane quarter = warp & 3, two warps
// per quarter split the columns).  Per tile and chunk c:
//     issuer A: hpre_c = xn w1nk_c^T              then  dxn  += dh_c w1kn_c^T
//     issuer B: dhacc_c = (rs dy) w2kn_c^T        then  dW1_c += dh_c^T xn        (token-reduction: MN-major views of the tiles)
//     issuer C:                                         dW2_c^T += h_c^T (rs dy)
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"

namespace mic {
using namespace t5;

struct MlpBwdArgs {
    const float* dy; const float* x; float* dx;
    const float* gamma; const float* beta; const float* b1;
    const uint8_t* w1nk_hi; const uint8_t* w1nk_lo;      // [hidden rows (padded to 64s)][C]      one 64-k panel
    const uint8_t* w2kn_hi; const uint8_t* w2kn_lo;      // [hidden rows (padded to 64s)][C]      W2 transposed
    const uint8_t* w1kn_hi; const uint8_t* w1kn_lo;      // per hidden chunk: [CP rows][64 hidden] W1 transposed
    const float* rowscale; int rps;
    float* dW1; float* db1; float* dW2; float* db2; float* dgamma; float* dbeta;
    int T, ntiles;
    float eps;
};

constexpr int MB_THREADS = 384;

template <int C>
struct MlpBwdCfg {
    static constexpr int HID = 4 * C;
    static constexpr int CP = (C + 15) / 16 * 16;
    static constexpr int NCH = (HID + 63) / 64;              // hidden chunks of 64
    static constexpr int TILE = 128 * 128;                   // one 64-feature panel of 128 rows (16 KB)
    static constexpr int OFF_XN = 0, OFF_DY = 2 * TILE, OFF_H = 4 * TILE, OFF_DH = 6 * TILE;      // hi then lo each
    static constexpr int W1KN_BYTES = CP * 128;
    static constexpr int SLOT = 4 * 8192 + 2 * W1KN_BYTES;   // w1nk hi/lo, w2kn hi/lo (64 rows x 128 B), w1kn hi/lo
    static constexpr int OFF_RING = 8 * TILE;
    static constexpr int OFF_PAR = OFF_RING + 2 * SLOT;      // b1[NCH*64] gamma[C] beta[C]
    static constexpr int OFF_BAR = OFF_PAR + 4 * (NCH * 64 + 2 * C);
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int NB = CP + 16;                       // dW1 tiles carry the "ones" column of xn: column CP = db1
    static constexpr int T_HP = 0, T_DH = 64, T_DXN = 128, T_DW1 = 128 + CP, T_DW2 = T_DW1 + NCH * NB;
    static constexpr int TCOLS = 512;
    static_assert(T_DW2 + NCH * CP <= 512 && NB <= 64, "TMEM budget / ones column inside the 64-feature panel");
    static_assert(C % 8 == 0 && C <= 64, "fused MLP backward: C multiple of 8, at most 64");
    static_assert(SMEM <= 232448, "shared memory budget");
};

template <int C>
__global__ void __launch_bounds__(MB_THREADS, 1) mlp_block_bwd_kernel(const MlpBwdArgs a) {
    using K = MlpBwdCfg<C>;
    constexpr int HID = K::HID, CP = K::CP, NCH = K::NCH, TILE = K::TILE, NB = K::NB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sXNh = smem + K::OFF_XN;  uint8_t* sXNl = sXNh + TILE;
    uint8_t* sDYh = smem + K::OFF_DY;  uint8_t* sDYl = sDYh + TILE;
    uint8_t* sHh = smem + K::OFF_H;    uint8_t* sHl = sHh + TILE;
    uint8_t* sDHh = smem + K::OFF_DH;  uint8_t* sDHl = sDHh + TILE;
    uint8_t* ring = smem + K::OFF_RING;
    float* sb1 = reinterpret_cast<float*>(smem + K::OFF_PAR);
    float* sg = sb1 + NCH * 64;
    float* sbt = sg + C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
    uint64_t* w_full = bars + 0;       // [2] ring slot landed (tx bytes)
    uint64_t* w_empty = bars + 2;      // [2] ring slot consumed (issuers A and B commit)
    uint64_t* a_full = bars + 4;       // xn / dy tiles written (8 warps), once per tile
    uint64_t* g1_done = bars + 5;      // hpre_c and dhacc_c complete (2 commits)
    uint64_t* g1_free = bars + 6;      // their TMEM columns have been read (8 warps)
    uint64_t* hd_full = bars + 7;      // h_c / dh_c tiles written (8 warps)
    uint64_t* g2_done = bars + 8;      // the three accumulation chains of chunk c complete (3 commits)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        bar_init(&w_full[0], 1); bar_init(&w_full[1], 1); bar_init(&w_empty[0], 2); bar_init(&w_empty[1], 2);
        bar_init(a_full, 8); bar_init(g1_done, 2); bar_init(g1_free, 8); bar_init(hd_full, 8); bar_init(g2_done, 3);
        bar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, K::TCOLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_sync();

    constexpr uint32_t id_g1 = idesc_bf16(128, 64, false, false);
    constexpr uint32_t id_dx = idesc_bf16(128, CP, false, false);
    constexpr uint32_t id_dw = idesc_bf16(128, CP, true, true);
    constexpr uint32_t id_dw1 = idesc_bf16(128, NB, true, true);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
                for (int c = 0; c < NCH; ++c, ++g) {
                    const uint32_t s = g & 1;
                    bar_wait(&w_empty[s], ((g >> 1) & 1) ^ 1);
                    bar_expect_tx(&w_full[s], K::SLOT);
                    uint8_t* d = ring + s * K::SLOT;
                    bulk_g2s(d, a.w1nk_hi + c * 8192, 8192, &w_full[s]);
                    bulk_g2s(d + 8192, a.w1nk_lo + c * 8192, 8192, &w_full[s]);
                    bulk_g2s(d + 16384, a.w2kn_hi + c * 8192, 8192, &w_full[s]);
                    bulk_g2s(d + 24576, a.w2kn_lo + c * 8192, 8192, &w_full[s]);
                    bulk_g2s(d + 32768, a.w1kn_hi + c * K::W1KN_BYTES, K::W1KN_BYTES, &w_full[s]);
                    bulk_g2s(d + 32768 + K::W1KN_BYTES, a.w1kn_lo + c * K::W1KN_BYTES, K::W1KN_BYTES, &w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- issuer A: hpre_c, then dxn += dh_c w1kn_c^T
        if (lane == 0) {
            uint32_t g = 0, n = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                bar_wait(a_full, n & 1);
                for (int c = 0; c < NCH; ++c, ++g) {
                    const uint32_t s = g & 1;
                    const uint32_t wb = s32(ring + s * K::SLOT);
                    bar_wait(&w_full[s], (g >> 1) & 1);
                    bar_wait(g1_free, (g & 1) ^ 1);
                    fence_after();
#pragma unroll
                    for (int ks = 0; ks < CP / 16; ++ks)
                        mma3(tmem + K::T_HP, desc_k(s32(sXNh) + ks * 32), desc_k(s32(sXNl) + ks * 32), desc_k(wb + ks * 32),
                             desc_k(wb + 8192 + ks * 32), id_g1, ks ? 1u : 0u);
                    commit(g1_done);
                    bar_wait(hd_full, g & 1);
                    fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma3(tmem + K::T_DXN, desc_k(s32(sDHh) + ks * 32), desc_k(s32(sDHl) + ks * 32),
                             desc_k(wb + 32768 + ks * 32), desc_k(wb + 32768 + K::W1KN_BYTES + ks * 32), id_dx,
                             (c | ks) ? 1u : 0u);
                    commit(&w_empty[s]);
                    commit(g2_done);
                }
            }
        }
    } else if (warp == 2) {
        // ---------------- issuer B: dhacc_c, then dW1_c += dh_c^T xn
        if (lane == 0) {
            uint32_t g = 0, n = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                bar_wait(a_full, n & 1);
                for (int c = 0; c < NCH; ++c, ++g) {
                    const uint32_t s = g & 1;
                    const uint32_t wb = s32(ring + s * K::SLOT);
                    bar_wait(&w_full[s], (g >> 1) & 1);
                    bar_wait(g1_free, (g & 1) ^ 1);
                    fence_after();
#pragma unroll
                    for (int ks = 0; ks < CP / 16; ++ks)
                        mma3(tmem + K::T_DH, desc_k(s32(sDYh) + ks * 32), desc_k(s32(sDYl) + ks * 32),
                             desc_k(wb + 16384 + ks * 32), desc_k(wb + 24576 + ks * 32), id_g1, ks ? 1u : 0u);
                    commit(g1_done);
                    commit(&w_empty[s]);
                    bar_wait(hd_full, g & 1);
                    fence_after();
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        mma3(tmem + K::T_DW1 + c * NB, desc_mn(s32(sDHh) + ks * 2048, TILE), desc_mn(s32(sDHl) + ks * 2048, TILE),
                             desc_mn(s32(sXNh) + ks * 2048, TILE), desc_mn(s32(sXNl) + ks * 2048, TILE), id_dw1, (n | ks) ? 1u : 0u);
                    commit(g2_done);
                }
            }
        }
    } else if (warp == 3) {
        // ---------------- issuer C: dW2_c^T += h_c^T (rs dy)
        if (lane == 0) {
            uint32_t g = 0, n = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                bar_wait(a_full, n & 1);
                for (int c = 0; c < NCH; ++c, ++g) {
                    bar_wait(hd_full, g & 1);
                    fence_after();
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        mma3(tmem + K::T_DW2 + c * CP, desc_mn(s32(sHh) + ks * 2048, TILE), desc_mn(s32(sHl) + ks * 2048, TILE),
                             desc_mn(s32(sDYh) + ks * 2048, TILE), desc_mn(s32(sDYl) + ks * 2048, TILE), id_dw, (n | ks) ? 1u : 0u);
                    commit(g2_done);
                }
            }
        }
    } else {
        // ---------------- row threads
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        for (int i = threadIdx.x - 128; i < NCH * 64 + 2 * C; i += MB_THREADS - 128) {
            float v;
            if (i < NCH * 64) v = i < HID ? a.b1[i] : 0.f;
            else if (i < NCH * 64 + C) v = a.gamma[i - NCH * 64];
            else v = a.beta[i - NCH * 64 - C];
            sb1[i] = v;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float cacc0[2] = {0.f, 0.f};                                // half 1: db2 partial sums
        float lnacc[4] = {0.f, 0.f, 0.f, 0.f};                      // half 0: per 16-column chunk, lanes 0..15 dgamma / 16..31 dbeta
        constexpr int NCHK = CP / 8;
        uint32_t g = 0, n = 0;
        for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
            const int64_t grow = (int64_t)t * 128 + row;
            const bool ok = grow < a.T;
            float mean = 0.f, rstd = 0.f;
            {
                float r[C];
                const float* src = half == 0 ? a.x : a.dy;
                if (grow + (int64_t)gridDim.x * 128 < a.T) prefetch_l2(src + (grow + (int64_t)gridDim.x * 128) * C, C * 4);
                if (ok) {
                    const float4* p = reinterpret_cast<const float4*>(src + grow * C);
#pragma unroll
                    for (int i = 0; i < C / 4; ++i) {
                        const float4 v = __ldg(p + i);
                        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < C; ++i) r[i] = 0.f;
                }
                if (half == 0) {
                    // LayerNorm(x) -> xn tile (all chunks of the row)
#pragma unroll
                    for (int i = 0; i < C; ++i) mean += r[i];
                    mean *= (1.f / C);
                    float var = 0.f;
#pragma unroll
                    for (int i = 0; i < C; ++i) { const float d = r[i] - mean; var = fmaf(d, d, var); }
                    rstd = rsqrtf(var * (1.f / C) + a.eps);
#pragma unroll
                    for (int i = 0; i < C; ++i) r[i] = ok ? (r[i] - mean) * rstd * sg[i] + sbt[i] : 0.f;
                } else {
                    // rs * dy -> dy tile; db2 += column sums
                    const float rs = (ok && a.rowscale) ? a.rowscale[grow / a.rps] : 1.f;
#pragma unroll
                    for (int i = 0; i < C; ++i) r[i] *= rs;
                }
                uint8_t* th = half == 0 ? sXNh : sDYh;
                uint8_t* tl = half == 0 ? sXNl : sDYl;
#pragma unroll
                for (int c = 0; c < NCHK; ++c) {
                    float v8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v8[e] = (c * 8 + e) < C ? r[(c * 8 + e) < C ? c * 8 + e : 0] : 0.f;
                    store_chunk(th, tl, row, c, v8);
                }
                if (half == 0) store_ones_chunk(sXNh, sXNl, row, NCHK);       // db1 = dh^T 1 comes out of the dW1 product
                if (half == 1) {
#pragma unroll
                    for (int gq = 0; gq < (C + 31) / 32; ++gq) {
                        float v[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = (gq * 32 + e) < C ? r[(gq * 32 + e) < C ? gq * 32 + e : 0] : 0.f;
                        cacc0[gq] += warp_colsum32(v, lane);
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(a_full);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++g) {
                bar_wait(g1_done, g & 1);
                fence_after();
                float hp[32], dv[32];
                ld32(tmem + lane_base + K::T_HP + half * 32, hp);
                ld32(tmem + lane_base + K::T_DH + half * 32, dv);
                ld_wait();
                fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(g1_free);
                const int col0 = c * 64 + half * 32;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    float gl, dg;
                    gelu_both(hp[e] + sb1[col0 + e], gl, dg);
                    hp[e] = gl;
                    dv[e] *= dg;
                }
                if (g > 0) bar_wait(g2_done, (g - 1) & 1);          // the chains of the previous chunk have read h / dh
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    store_chunk(sHh, sHl, row, half * 4 + j, hp + 8 * j);
                    store_chunk(sDHh, sDHl, row, half * 4 + j, dv + 8 * j);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) bar_arrive(hd_full);
            }
            // ---- tile end: dxn complete -> LayerNorm backward, dx = dy + LN'(dxn); dgamma / dbeta partial sums
            bar_wait(g2_done, (g - 1) & 1);
            fence_after();
            if (half == 0) {
                // LayerNorm backward in 16-column chunks, two passes (row means first): few live registers, no spills
                const float* xp = a.x + grow * C;
                const float* dp_ = a.dy + grow * C;
                float m1 = 0.f, m2 = 0.f;
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 16) {
                    float gx[16];
                    ld16(tmem + lane_base + K::T_DXN + c0, gx);
                    ld_wait();
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            if (c0 + i < C) {
                                const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + c0 + i));
                                const float d0 = gx[i] * sg[c0 + i], d1 = gx[i + 1] * sg[c0 + i + 1], d2 = gx[i + 2] * sg[c0 + i + 2],
                                            d3 = gx[i + 3] * sg[c0 + i + 3];
                                m1 += (d0 + d1) + (d2 + d3);
                                m2 = fmaf(d0, (xv.x - mean) * rstd, m2); m2 = fmaf(d1, (xv.y - mean) * rstd, m2);
                                m2 = fmaf(d2, (xv.z - mean) * rstd, m2); m2 = fmaf(d3, (xv.w - mean) * rstd, m2);
                            }
                        }
                    }
                }
                m1 *= (1.f / C); m2 *= (1.f / C);
#pragma unroll
                for (int c0 = 0; c0 < C; c0 += 16) {
                    float gx[16], v[32];
                    ld16(tmem + lane_base + K::T_DXN + c0, gx);
                    ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const bool in = ok && (c0 + i < C);
                        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv4 = xv;
                        if (in) { xv = __ldg(reinterpret_cast<const float4*>(xp + c0 + i)); dv4 = __ldg(reinterpret_cast<const float4*>(dp_ + c0 + i)); }
                        const float xh0 = (xv.x - mean) * rstd, xh1 = (xv.y - mean) * rstd, xh2 = (xv.z - mean) * rstd, xh3 = (xv.w - mean) * rstd;
                        const float g0 = in ? gx[i] : 0.f, g1_ = in ? gx[i + 1] : 0.f, g2_ = in ? gx[i + 2] : 0.f, g3 = in ? gx[i + 3] : 0.f;
                        if (in) {
                            const int ci = (c0 + i) < C ? c0 + i : 0;
                            float4 o;
                            o.x = dv4.x + rstd * (g0 * sg[ci] - m1 - xh0 * m2);
                            o.y = dv4.y + rstd * (g1_ * sg[ci + 1] - m1 - xh1 * m2);
                            o.z = dv4.z + rstd * (g2_ * sg[ci + 2] - m1 - xh2 * m2);
                            o.w = dv4.w + rstd * (g3 * sg[ci + 3] - m1 - xh3 * m2);
                            *reinterpret_cast<float4*>(a.dx + grow * C + c0 + i) = o;
                        }
                        v[i] = g0 * xh0; v[i + 1] = g1_ * xh1; v[i + 2] = g2_ * xh2; v[i + 3] = g3 * xh3;      // dgamma terms
                        v[16 + i] = g0; v[16 + i + 1] = g1_; v[16 + i + 2] = g2_; v[16 + i + 3] = g3;             // dbeta terms
                    }
                    const float cs = warp_colsum32(v, lane);      // lanes 0..15: dgamma of this chunk, lanes 16..31: dbeta
                    if (c0 == 0) lnacc[0] += cs; else if (c0 == 16) lnacc[1] += cs; else if (c0 == 32) lnacc[2] += cs; else lnacc[3] += cs;
                }
            }
            fence_before();
        }
        // ---------------- flush: bias / LayerNorm gradients from registers, weight gradients from TMEM
#pragma unroll
        for (int gq = 0; gq < (C + 31) / 32; ++gq) {
            const int col = gq * 32 + lane;
            if (col < C && half == 1) atomicAdd(a.db2 + col, cacc0[gq]);
        }
        if (half == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = 16 * k + (lane & 15);
                if (16 * k < C && col < C) atomicAdd((lane < 16 ? a.dgamma : a.dbeta) + col, lnacc[k]);
            }
        }
        if (q < 2) {                      // accumulator rows 0..63 = hidden index inside the chunk (rows 64..127 are not used)
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                float v[NB];
                const uint32_t col = half == 0 ? K::T_DW1 + c * NB : K::T_DW2 + c * CP;
#pragma unroll
                for (int c0 = 0; c0 < (half == 0 ? NB : CP); c0 += 16) ld16(tmem + lane_base + col + c0, v + c0);
                ld_wait();
                const int hid = c * 64 + row;
                if (hid < HID) {
                    if (half == 0) {
#pragma unroll
                        for (int i = 0; i < C; ++i) atomicAdd(a.dW1 + (int64_t)hid * C + i, v[i]);
                        atomicAdd(a.db1 + hid, v[CP]);
                    } else {
#pragma unroll
                        for (int i = 0; i < C; ++i) atomicAdd(a.dW2 + (int64_t)i * HID + hid, v[i]);
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        tmem_dealloc(tmem, K::TCOLS);
    }
}

template <int C>
static int launch_mlp_bwd(const MlpBwdArgs& a, cudaStream_t st) {
    using K = MlpBwdCfg<C>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(mlp_block_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM) != cudaSuccess) {
            cudaGetLastError();
            return MIC_ERR_UNSUPPORTED;
        }
        attr = true;
    }
    int grid = num_sms();
    if (grid > a.ntiles) grid = a.ntiles;
    mic::launch(mlp_block_bwd_kernel<C>, dim3(grid), dim3(MB_THREADS), (size_t)K::SMEM, st, a);
    return check_launch("mlp_block_bwd_kernel");
}

}  // namespace mic

using namespace mic;

extern "C" int mic_mlp_block_bwd(const float* dy, const float* x, float* dx, const float* gamma, const float* beta,
                                 const float* b1, const void* w1nk_hi, const void* w1nk_lo, const void* w2kn_hi,
                                 const void* w2kn_lo, const void* w1kn_hi, const void* w1kn_lo, const float* rowscale,
                                 int rows_per_sample, float* dW1, float* db1, float* dW2, float* db2, float* dgamma,
                                 float* dbeta, int T, int C, float eps, void* stream) {
    MIC_REQUIRE(dy && x && dx && gamma && beta && b1 && w1nk_hi && w1nk_lo && w2kn_hi && w2kn_lo && w1kn_hi && w1kn_lo && dW1 &&
                    db1 && dW2 && db2 && dgamma && dbeta && T > 0, "mlp_block_bwd: bad arguments");
    MIC_REQUIRE(!rowscale || rows_per_sample > 0, "mlp_block_bwd: rows_per_sample");
    if (((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
          reinterpret_cast<uintptr_t>(w1nk_hi) | reinterpret_cast<uintptr_t>(w1nk_lo) | reinterpret_cast<uintptr_t>(w2kn_hi) |
          reinterpret_cast<uintptr_t>(w2kn_lo) | reinterpret_cast<uintptr_t>(w1kn_hi) | reinterpret_cast<uintptr_t>(w1kn_lo)) & 15) != 0)
        return fail(MIC_ERR_UNSUPPORTED, "mlp_block_bwd: pointers must be 16-byte aligned");
    MlpBwdArgs a;
    a.dy = dy; a.x = x; a.dx = dx; a.gamma = gamma; a.beta = beta; a.b1 = b1;
    a.w1nk_hi = (const uint8_t*)w1nk_hi; a.w1nk_lo = (const uint8_t*)w1nk_lo;
    a.w2kn_hi = (const uint8_t*)w2kn_hi; a.w2kn_lo = (const uint8_t*)w2kn_lo;
    a.w1kn_hi = (const uint8_t*)w1kn_hi; a.w1kn_lo = (const uint8_t*)w1kn_lo;
    a.rowscale = rowscale; a.rps = rows_per_sample > 0 ? rows_per_sample : 1;
    a.dW1 = dW1; a.db1 = db1; a.dW2 = dW2; a.db2 = db2; a.dgamma = dgamma; a.dbeta = dbeta;
    a.T = T; a.ntiles = (T + 127) / 128; a.eps = eps;
    switch (C) {
        case 24: return launch_mlp_bwd<24>(a, (cudaStream_t)stream);
        case 48: return launch_mlp_bwd<48>(a, (cudaStream_t)stream);
        default: return fail(MIC_ERR_UNSUPPORTED, "mlp_block_bwd: C=%d is not built (24, 48)", C);
    }
}
