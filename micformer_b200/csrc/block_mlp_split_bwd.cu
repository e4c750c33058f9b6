// Backward of the hidden-split fused MLP (block_mlp_split.cu) for the deep stages.  CTA (tile i, chunk j) recomputes its
// 64 hidden units from x and produces, all through streamed 64-wide panels and ONE TMEM output region that is reused:
//     L1  hpre_j  = LN(x) W1_j^T                      (A: xn panels, made on the fly;      B: fc1 image rows of chunk j)
//     L2  dhacc_j = (rs dy) W2[:, j]                  (A: dy panels;                        B: fc2-transposed image rows of chunk j)
//     E1  h_j = GELU(hpre_j + b1_j), dh_j = dhacc_j * GELU'(.)   -> two operand tiles;  db1_j += column sums of dh_j
//     L3  dxn += dh_j W1_j        (atomic into the zeroed dxn)    (A: dh tile; B: fc1-transposed image panel j, 128-row chunks)
//     L4  dW1_j += dh_j^T xn      (token reduction: MN-major views; xn panels are produced a second time)
//     L5  dW2[:, j] += (rs dy)^T h_j                              (dy panels a second time)
// The LayerNorm backward (dx = dy + LN'(dxn), dgamma, dbeta) is the existing mic_layernorm_bwd on dxn with the statistics
// this kernel writes; db2 (column sums of rs dy) is accumulated by the chunk-0 CTAs.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"
#include "block_mlp_split.cuh"

namespace mic {
using namespace t5;

struct MlpSplitBwdArgs {
    const float* dy; const float* x;
    float* dxn;                                       // (T, C) zero on entry: gradient w.r.t. LN(x), accumulated
    float* mean; float* rstd;                         // (T) written by the chunk-0 CTAs for the LayerNorm backward
    const float* gamma; const float* beta; const float* b1;
    const uint8_t* w1nk_hi; const uint8_t* w1nk_lo;   // fc1:            N = HID (n_pad1) x K = C
    const uint8_t* w2kn_hi; const uint8_t* w2kn_lo;   // fc2 transposed: N = HID (n_pad1) x K = C
    const uint8_t* w1kn_hi; const uint8_t* w1kn_lo;   // fc1 transposed: N = C (CP) x K = HID: panel j = hidden chunk j
    const float* rowscale; int rps;
    float* dW1; float* db1; float* dW2; float* db2;
    int T, C, HID, CP, n_pad1;
    float eps;
};

struct MsB {                                          // backward shared-memory map (bytes)
    // (the h / dh tiles come first: as MN-major A operands of an M = 128 MMA their second 64-row block is addressed 16 KB
    //  further on -- rows that are never read back, but the address must stay inside the allocation)
    static constexpr int H = 0;                       // h_j tile  (hi, lo)
    static constexpr int DH = 32768;                  // dh_j tile (hi, lo)
    static constexpr int AR = 65536;                  // A ring: 2 slots x (hi 16 KB + lo 16 KB)
    static constexpr int WR = 131072;                 // W ring: 2 slots x (hi 16 KB + lo 16 KB)
    static constexpr int PART = 196608;
    static constexpr int BAR = PART + 1024;
    static constexpr int SMEM = BAR + 256 + 1024;
};

__global__ void __launch_bounds__(MS_THREADS, 1) mlp_split_bwd_kernel(const MlpSplitBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sAR = smem + MsB::AR;
    uint8_t* sWR = smem + MsB::WR;
    uint8_t* sH = smem + MsB::H;
    uint8_t* sDH = smem + MsB::DH;
    float* spart = reinterpret_cast<float*>(smem + MsB::PART);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MsB::BAR);
    uint64_t* a_full = bars + 0;    // [2] A-ring slot written (8 warps)
    uint64_t* a_empty = bars + 2;   // [2] A-ring slot consumed (MMA commit)
    uint64_t* w_full = bars + 4;    // [2] W-ring slot landed (tx)
    uint64_t* w_empty = bars + 6;   // [2]
    uint64_t* g1_done = bars + 8;   // hpre_j and dhacc_j complete
    uint64_t* hd_full = bars + 9;   // h_j / dh_j tiles written (8 warps)
    uint64_t* out_full = bars + 10; // the TMEM output region holds dxn (phase 0), dW1_j (1), dW2_j (0)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, j = blockIdx.y;
    const int C = a.C, CP = a.CP;
    const int KP = (C + 63) >> 6;
    const int NCK = (CP + 127) >> 7;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { bar_init(&a_full[s], 8); bar_init(&a_empty[s], 1); bar_init(&w_full[s], 1); bar_init(&w_empty[s], 1); }
        bar_init(g1_done, 1); bar_init(hd_full, 8); bar_init(out_full, 1);
        bar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_hp = tmem, t_dh = tmem + 64, t_out = tmem + 128;
    pdl_sync();

    if (warp == 0) {
        // ---------------- weight loader: L1 and L2 take 64-row panels of chunk j, L3 takes 128-row chunks of panel j
        if (lane == 0) {
            uint32_t w = 0;
            for (int pass = 0; pass < 2; ++pass) {
                const uint8_t* hi = pass == 0 ? a.w1nk_hi : a.w2kn_hi;
                const uint8_t* lo = pass == 0 ? a.w1nk_lo : a.w2kn_lo;
                for (int p = 0; p < KP; ++p, ++w) {
                    const int s = w & 1;
                    bar_wait(&w_empty[s], ((w >> 1) & 1) ^ 1);
                    bar_expect_tx(&w_full[s], 2 * 8192);
                    const size_t off = (size_t)p * a.n_pad1 * 128 + (size_t)j * 8192;
                    bulk_g2s(sWR + s * 32768, hi + off, 8192, &w_full[s]);
                    bulk_g2s(sWR + s * 32768 + 16384, lo + off, 8192, &w_full[s]);
                }
            }
            for (int n = 0; n < NCK; ++n, ++w) {
                const int s = w & 1;
                const int rows = min(128, CP - 128 * n);
                bar_wait(&w_empty[s], ((w >> 1) & 1) ^ 1);
                bar_expect_tx(&w_full[s], 2 * rows * 128);
                const size_t off = (size_t)j * CP * 128 + (size_t)n * 128 * 128;
                bulk_g2s(sWR + s * 32768, a.w1kn_hi + off, rows * 128, &w_full[s]);
                bulk_g2s(sWR + s * 32768 + 16384, a.w1kn_lo + off, rows * 128, &w_full[s]);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            const uint32_t id_k64 = idesc_bf16(128, 64, false, false);
            const uint32_t id_mn64 = idesc_bf16(128, 64, true, true);
            uint32_t u = 0, w = 0;
            for (int pass = 0; pass < 2; ++pass) {                       // L1 (hpre), L2 (dhacc)
                for (int p = 0; p < KP; ++p, ++u, ++w) {
                    const int sa = u & 1, sw = w & 1;
                    bar_wait(&a_full[sa], (u >> 1) & 1);
                    bar_wait(&w_full[sw], (w >> 1) & 1);
                    fence_after();
                    const uint32_t ah = s32(sAR + sa * 32768), al = ah + 16384, bh = s32(sWR + sw * 32768), bl = bh + 16384;
                    const int ksteps = (min(64, CP - 64 * p) + 15) >> 4;
                    for (int ks = 0; ks < ksteps; ++ks)
                        mma3(pass == 0 ? t_hp : t_dh, desc_k(ah + ks * 32), desc_k(al + ks * 32), desc_k(bh + ks * 32),
                             desc_k(bl + ks * 32), id_k64, (p | ks) ? 1u : 0u);
                    commit(&a_empty[sa]);
                    commit(&w_empty[sw]);
                }
            }
            commit(g1_done);
            bar_wait(hd_full, 0);
            fence_after();
            const uint32_t hh = s32(sH), hl = hh + 16384, dh = s32(sDH), dl = dh + 16384;
            for (int n = 0; n < NCK; ++n, ++w) {                          // L3: dxn columns [128 n, ...)
                const int sw = w & 1;
                const int rows = min(128, CP - 128 * n);
                bar_wait(&w_full[sw], (w >> 1) & 1);
                fence_after();
                const uint32_t idn = idesc_bf16(128, rows, false, false);
                const uint32_t bh = s32(sWR + sw * 32768), bl = bh + 16384;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma3(t_out + 128 * n, desc_k(dh + ks * 32), desc_k(dl + ks * 32), desc_k(bh + ks * 32), desc_k(bl + ks * 32), idn,
                         ks ? 1u : 0u);
                commit(&w_empty[sw]);
            }
            commit(out_full);
            for (int pass = 0; pass < 2; ++pass) {                       // L4 (dW1_j = dh^T xn), L5 (dW2_j^T = h^T dy)
                const uint32_t mh = pass == 0 ? dh : hh, ml = pass == 0 ? dl : hl;
                for (int p = 0; p < KP; ++p, ++u) {
                    const int sa = u & 1;
                    bar_wait(&a_full[sa], (u >> 1) & 1);
                    fence_after();
                    const uint32_t bh = s32(sAR + sa * 32768), bl = bh + 16384;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        mma3(t_out + 64 * p, desc_mn(mh + ks * 2048, 16384), desc_mn(ml + ks * 2048, 16384),
                             desc_mn(bh + ks * 2048, 16384), desc_mn(bl + ks * 2048, 16384), id_mn64, ks ? 1u : 0u);
                    commit(&a_empty[sa]);
                }
                commit(out_full);
            }
        }
    } else {
        // ---------------- row threads
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int64_t grow = (int64_t)tile * 128 + row;
        const bool ok = grow < a.T;
        const float* xr = a.x + grow * C;
        const float* dr = a.dy + grow * C;
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
        if (j == 0 && half == 0 && ok) { a.mean[grow] = mean; a.rstd[grow] = rstd; }
        const float rs = (ok && a.rowscale) ? a.rowscale[grow / a.rps] : 1.f;

        uint32_t u = 0;
        // one panel of LN(x) (kind 0) or rs*dy (kind 1) into the A ring; chunk-0 CTAs add the column sums of rs*dy to db2
        auto produce = [&](int kind, int p, bool want_db2) {
            const int sl = u & 1;
            bar_wait(&a_empty[sl], ((u >> 1) & 1) ^ 1);
            ++u;
            uint8_t* th = sAR + sl * 32768;
            uint8_t* tl = th + 16384;
            const int c0 = 64 * p + 32 * half;
            float v[32];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const int c = c0 + 4 * cc;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok && c < C) {
                    if (kind == 0) {
                        const float4 xa = __ldg(reinterpret_cast<const float4*>(xr + c));
                        const float4 ga = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
                        const float4 ba = __ldg(reinterpret_cast<const float4*>(a.beta + c));
                        o.x = (xa.x - mean) * rstd * ga.x + ba.x; o.y = (xa.y - mean) * rstd * ga.y + ba.y;
                        o.z = (xa.z - mean) * rstd * ga.z + ba.z; o.w = (xa.w - mean) * rstd * ga.w + ba.w;
                    } else {
                        const float4 d = __ldg(reinterpret_cast<const float4*>(dr + c));
                        o.x = rs * d.x; o.y = rs * d.y; o.z = rs * d.z; o.w = rs * d.w;
                    }
                }
                v[4 * cc] = o.x; v[4 * cc + 1] = o.y; v[4 * cc + 2] = o.z; v[4 * cc + 3] = o.w;
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) store_chunk(th, tl, row, half * 4 + cc, v + 8 * cc);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(&a_full[sl]);
            if (want_db2) {
                const float cs = warp_colsum32(v, lane);
                if (c0 + lane < C) atomicAdd(a.db2 + c0 + lane, cs);
            }
        };
        for (int p = 0; p < KP; ++p) produce(0, p, false);                  // L1
        for (int p = 0; p < KP; ++p) produce(1, p, j == 0);                 // L2
        // ---- E1: h_j, dh_j
        bar_wait(g1_done, 0);
        fence_after();
        {
            float hp[32], dv[32];
            ld32(t_hp + lane_base + half * 32, hp);
            ld32(t_dh + lane_base + half * 32, dv);
            ld_wait();
            const int hid0 = j * 64 + half * 32;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                float gl, dg;
                gelu_both(hp[e] + (hid0 + e < a.HID ? __ldg(a.b1 + hid0 + e) : 0.f), gl, dg);
                hp[e] = gl;
                dv[e] *= dg;
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                store_chunk(sH, sH + 16384, row, half * 4 + cc, hp + 8 * cc);
                store_chunk(sDH, sDH + 16384, row, half * 4 + cc, dv + 8 * cc);
            }
            fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(hd_full);
            const float cs = warp_colsum32(dv, lane);
            if (hid0 + lane < a.HID) atomicAdd(a.db1 + hid0 + lane, cs);
        }
        // ---- E3: dxn partial -> global (atomic)
        bar_wait(out_full, 0);
        fence_after();
        for (int g = half; g * 32 < C; g += 2) {
            float v[32];
            ld32(t_out + lane_base + g * 32, v);
            ld_wait();
            if (ok) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const int c = g * 32 + e;
                    if (c < C) atomicAdd(reinterpret_cast<float4*>(a.dxn + grow * C + c), make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
                }
            }
        }
        fence_before();
        // ---- L4 panels (xn again), E4: dW1_j rows = hidden unit (TMEM lanes 0..63), columns = input channel
        for (int p = 0; p < KP; ++p) produce(0, p, false);
        bar_wait(out_full, 1);
        fence_after();
        const int hid = j * 64 + row;
        if (q < 2 && hid < a.HID) {
            for (int g = half; g * 32 < C; g += 2) {
                float v[32];
                ld32(t_out + lane_base + g * 32, v);
                ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const int c = g * 32 + e;
                    if (c < C) atomicAdd(reinterpret_cast<float4*>(a.dW1 + (int64_t)hid * C + c), make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
                }
            }
        }
        fence_before();
        // ---- L5 panels (dy again), E5: dW2[c][hid] += out[hid][c]
        for (int p = 0; p < KP; ++p) produce(1, p, false);
        bar_wait(out_full, 0);
        fence_after();
        if (q < 2 && hid < a.HID) {
            for (int g = half; g * 32 < C; g += 2) {
                float v[32];
                ld32(t_out + lane_base + g * 32, v);
                ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int c = g * 32 + e;
                    if (c < C) atomicAdd(a.dW2 + (int64_t)c * a.HID + hid, v[e]);
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

extern "C" int mic_mlp_split_bwd(const float* dy, const float* x, float* dxn_zeroed, float* mean, float* rstd, const float* gamma,
                                 const float* beta, const float* b1, const void* w1nk_hi, const void* w1nk_lo,
                                 const void* w2kn_hi, const void* w2kn_lo, const void* w1kn_hi, const void* w1kn_lo,
                                 const float* rowscale, int rows_per_sample, float* dW1, float* db1, float* dW2, float* db2, int T,
                                 int C, int HID, float eps, void* stream) {
    MIC_REQUIRE(dy && x && dxn_zeroed && mean && rstd && gamma && beta && b1 && w1nk_hi && w1nk_lo && w2kn_hi && w2kn_lo && w1kn_hi &&
                    w1kn_lo && dW1 && db1 && dW2 && db2 && T > 0, "mlp_split_bwd: bad arguments");
    if (C % 8 || C < 64 || C > 384 || HID % 64 || HID != 4 * C)
        return fail(MIC_ERR_UNSUPPORTED, "mlp_split_bwd: C=%d HID=%d not taken", C, HID);
    MlpSplitBwdArgs a;
    a.dy = dy; a.x = x; a.dxn = dxn_zeroed; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.b1 = b1;
    a.w1nk_hi = (const uint8_t*)w1nk_hi; a.w1nk_lo = (const uint8_t*)w1nk_lo; a.w2kn_hi = (const uint8_t*)w2kn_hi;
    a.w2kn_lo = (const uint8_t*)w2kn_lo; a.w1kn_hi = (const uint8_t*)w1kn_hi; a.w1kn_lo = (const uint8_t*)w1kn_lo;
    a.rowscale = rowscale; a.rps = rows_per_sample > 0 ? rows_per_sample : 1;
    a.dW1 = dW1; a.db1 = db1; a.dW2 = dW2; a.db2 = db2;
    a.T = T; a.C = C; a.HID = HID; a.CP = (C + 15) / 16 * 16; a.n_pad1 = (HID + 63) / 64 * 64; a.eps = eps;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(mlp_split_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MsB::SMEM) != cudaSuccess) {
            cudaGetLastError();
            return MIC_ERR_UNSUPPORTED;
        }
        attr = true;
    }
    dim3 grid((T + 127) / 128, HID / 64);
    mic::launch(mlp_split_bwd_kernel, grid, dim3(MS_THREADS), (size_t)MsB::SMEM, (cudaStream_t)stream, a);
    return check_launch("mlp_split_bwd_kernel");
}
