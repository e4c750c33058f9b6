// tcgen05 / TMA windowed multi-head attention BACKWARD for large windows (128 <= tokens <= 352, head_dim 32): the
// autograd of `softmax(q k^T * scale) v` inside WindowAttention3D / CrossWindowAttention3D
// (reference models/MICFormer_self.py:179-203 and :237-261) for the 343-token windows of BASELINE.json's config 4 and
// the window-7 model configs.  One CTA per (window, head); dQ, dK, dV from Q, K, V, O, dO and the forward's
// log-sum-exp -- the score matrices are recomputed on the tensor cores and never leave TMEM.
//
// A tcgen05.mma takes its A operand from TMEM with one accumulator lane per A row, so a product whose *rows* are keys
// (dV = P^T dO, dK = dS^T Q) needs the scores transposed.  Instead of transposing through shared memory the kernel
// computes the cheap K=32 score products in both orientations:
//
//   phase 0 (dV), per 128-key tile:   S^T = K Q^T                   P^T  = exp(S^T * scale - lse[col])
//                                                                   dV  += P^T  . dO        (A = P^T  in TMEM)
//   phase 1 (dQ), per 128-query tile: S = Q K^T,   dP = dO V^T      dS   = P * (dP - delta[row])
//                                                                   dQ  += dS   . K         (A = dS   in TMEM)
//   phase 2 (dK), per 128-key tile:   S^T = K Q^T, dP^T = V dO^T    dS^T = P^T * (dP^T - delta[col])
//                                                                   dK  += dS^T . Q         (A = dS^T in TMEM)
//
// with delta[i] = sum_c dO[i,c] O[i,c].  The score columns are processed in chunks of 96; two TMEM buffers
// (score | second score, 192 columns each) and two accumulators alternate, so that while one group of four warps turns
// a chunk of scores into P / dS the tensor core already runs the next chunk's score products and the previous chunk's
// accumulating product.
//
//   warp 0      lane 0: TMA producer (5-D boxes gather the window's head slice out of the token grid, as the forward)
//               and MMA issuer
//   warps 1-4   element-wise group 0 (even chunks): thread = TMEM lane; tcgen05.ld scores, exp2 / fma in registers,
//               tcgen05.st the TF32-rounded result in place; epilogue of a pass (accumulator -> global, window_reverse
//               is the store address)
//   warps 5-8   element-wise group 1 (odd chunks)
//
// Shared memory: Q, K, V, dO as K-major 128B-swizzled tiles (operands of the score products) and one slot that holds
// the MN-major (32B-atom swizzle) copy the current phase needs as B operand of the accumulating product: dO, K, Q in
// turn.  5 x 44 KB.  fp32 in / fp32 out, operands read as TF32 (Q, K, V, dO tiles rounded to nearest in shared memory,
// as the forward does), fp32 accumulation.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace mic {
namespace {

constexpr int AB_THREADS = 320;
constexpr int AB_SLOT_BYTES = 352 * 128;        // up to 352 token rows x 32 fp32
constexpr int AB_SLOTS = 5;                     // 0: Q, 1: K, 2: V, 3: dO (K-major)   4: MN-major operand of the phase
constexpr int AB_CW = 96;                       // score columns per chunk
constexpr int AB_BUF_COLS = 2 * AB_CW;          // first | second score matrix of a chunk
constexpr int AB_ACC_COL = 2 * AB_BUF_COLS;     // four 32-column accumulators (issuer x pass parity) behind the two buffers
constexpr int AB_STAT = 384;                    // per-token statistics kept in shared memory

struct AttnBwdArgs {
    const float* out; const float* dout; int ldo;
    const float* lse;
    float* dq; int lddq;
    float* dk; float* dv; int lddkv;
    int Dp, Hp, Wp, heads, C, wd, wh, ww, nwd, nwh, nww;
    int N;            // tokens per window
    int Nk;           // N rounded up to 32
    float scale_log2; // head_dim^-0.5 * log2(e)
    float scale;
};

// phase trace (scripts/trace_attn_bwd.py): SM clock stamps of CTA 0, slot layout documented in the script
__device__ long long* g_ab_trace = nullptr;
__device__ __forceinline__ void btrace_(long long* t, int slot) {
    if (t && slot < 1024) t[slot] = clock64();
}
#define btrace(slot) btrace_(trp, slot)

__device__ __forceinline__ uint32_t bsmem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bbar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bsmem(b)), "r"(c));
}
__device__ __forceinline__ void bbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bsmem(b)) : "memory");
}
__device__ __forceinline__ void bbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bsmem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BW_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BW_DONE;\n"
        "bra BW_LOOP;\n"
        "BW_DONE:\n"
        "}\n" ::"r"(bsmem(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bcommit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bsmem(b)) : "memory");
}
__device__ __forceinline__ void btma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                             int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(bsmem(dst)), "l"(map), "r"(bsmem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ uint64_t bdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void bmma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void bmma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void btld32(uint32_t taddr, uint32_t* r) {      // no wait
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void btst32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ float bex2(float x) {
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(x));
    return p;
}
__device__ __forceinline__ uint32_t rn_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ void btld16(uint32_t taddr, uint32_t* r) {      // no wait
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void btst16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__global__ void __launch_bounds__(AB_THREADS, 1)
window_attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap mapQK,    // qkv rows, 128B swizzle      (K-major tiles)
                          const __grid_constant__ CUtensorMap mapQM,    // qkv rows, 32B-atom swizzle  (MN-major tiles)
                          const __grid_constant__ CUtensorMap mapGK,    // dO rows, 128B swizzle
                          const __grid_constant__ CUtensorMap mapGM,    // dO rows, 32B-atom swizzle
                          const __grid_constant__ CUtensorMap mapOK,    // O rows, 128B swizzle
                          AttnBwdArgs a) {
    extern __shared__ uint8_t ab_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = sm;
    uint8_t* tail = sm + AB_SLOTS * AB_SLOT_BYTES;                 // nothing is read beyond slot 4 (it is never an A tile)
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);            // [5] TMA landed
    uint64_t* rdy = full + AB_SLOTS;                               // [2] (Q,K) / (V,dO) rounded to TF32: 256 arrivals
    uint64_t* s_full = rdy + 2;                                    // [2] score products of a chunk complete (per buffer)
    uint64_t* p_ready = s_full + 2;                                // [2] P / dS of a chunk written to TMEM: 256 arrivals
    uint64_t* acc_full = p_ready + 2;                              // [2 issuers][2] accumulating products of a pass complete
    uint64_t* acc_free = acc_full + 4;                             // [2] accumulators read by the epilogue: 128 arrivals
    uint64_t* o_done = acc_free + 2;                               // delta computed, slot 2 (O) may be overwritten: 256 arrivals
    uint32_t* tslot = reinterpret_cast<uint32_t*>(o_done + 1);
    float* lse2 = reinterpret_cast<float*>(tail + 256);            // [AB_STAT] lse * log2(e)   (16-byte aligned)
    float* del = lse2 + AB_STAT;                                   // [AB_STAT] delta

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nmt = (a.N + 127) / 128;
    const int nch = (a.Nk + AB_CW - 1) / AB_CW;                    // >= 2 (N >= 128)
    const int G = 3 * nmt * nch;                                   // chunks of this item, in issue order
    const uint32_t box_bytes = (uint32_t)a.N * 128u;

    // item -> (batch, window, head)
    const uint32_t it = blockIdx.x;
    const uint32_t head = it % (uint32_t)a.heads;
    uint32_t w = it / (uint32_t)a.heads;
    const uint32_t wx = w % (uint32_t)a.nww; w /= (uint32_t)a.nww;
    const uint32_t wy = w % (uint32_t)a.nwh; w /= (uint32_t)a.nwh;
    const uint32_t wz = w % (uint32_t)a.nwd; w /= (uint32_t)a.nwd;     // w = batch index
    const int64_t row0 = (((int64_t)w * a.Dp + wz * a.wd) * a.Hp + wy * a.wh) * a.Wp + wx * a.ww;

    // rows [N, 352) of every slot are never written by TMA: zero them once so they stay finite (0 * garbage = NaN)
    {
        const int per = (352 - a.N) * 8;                           // float4 per slot
        for (int i = threadIdx.x; i < AB_SLOTS * per; i += AB_THREADS) {
            const int s = i / per, o = i - s * per;
            reinterpret_cast<float4*>(ring + (size_t)s * AB_SLOT_BYTES + (size_t)a.N * 128)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < AB_SLOTS; ++s) bbar_init(&full[s], 1);
        for (int x = 0; x < 2; ++x) {
            bbar_init(&rdy[x], 256); bbar_init(&s_full[x], 1); bbar_init(&p_ready[x], 256);
            bbar_init(&acc_full[2 * x], 1); bbar_init(&acc_full[2 * x + 1], 1); bbar_init(&acc_free[x], 128);
        }
        bbar_init(o_done, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bsmem(tslot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    pdl_sync();
    long long* const trp = blockIdx.x == gridDim.x / 2 ? g_ab_trace : nullptr;      // read once: not on the issue path

    if (warp < 2) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer x (chunks g = x, x+2, ..: TMEM buffer x,
            // accumulators x); issuer 0 is also the TMA producer
            const int x = warp;
            const int cx = (int)(wx * a.ww), cy = (int)(wy * a.wh), cz = (int)(wz * a.wd), cb = (int)w;
            const int hc = (int)head * 32;
            auto load = [&](int slot, const CUtensorMap* map, int c0) {
                bbar_expect(&full[slot], box_bytes);
                btma_load_5d(ring + (size_t)slot * AB_SLOT_BYTES, map, &full[slot], c0, cx, cy, cz, cb);
            };
            if (x == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQK) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQM) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&mapGK) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&mapGM) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&mapOK) : "memory");
                load(0, &mapQK, hc);                     // Q
                load(1, &mapQK, a.C + hc);               // K
                load(4, &mapGM, hc);                     // dO, MN-major (phase 0)
                load(3, &mapGK, hc);                     // dO
                load(2, &mapOK, hc);                     // O: visits slot 2 until delta is computed, then V takes the slot
            }
            const uint32_t ring_a = bsmem(ring);
            const uint32_t sQ = ring_a, sK = ring_a + AB_SLOT_BYTES, sV = ring_a + 2 * AB_SLOT_BYTES,
                           sG = ring_a + 3 * AB_SLOT_BYTES, sM = ring_a + 4 * AB_SLOT_BYTES;
            const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_o = idesc_base | (1u << 16) | ((uint32_t)(32 >> 3) << 17);
            const uint32_t buf = tmem + (uint32_t)(x * AB_BUF_COLS);
            // score products of a chunk: M = 128 lanes (tile t of the A slot), N = chunk columns (rows of the B slot)
            auto issue_s = [&](int ph, int t, int c) {
                const int col0 = c * AB_CW;
                const int wdt = min(AB_CW, a.Nk - col0);
                const uint32_t idesc = idesc_base | ((uint32_t)(wdt >> 3) << 17);
                const uint32_t a1 = (ph == 1 ? sQ : sK) + (uint32_t)(t * 128 * 128);
                const uint32_t b1 = (ph == 1 ? sK : sQ) + (uint32_t)(col0 * 128);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    bmma_ss(buf, bdesc(a1 + (uint32_t)(ks * 32), 16, 1024, 2), bdesc(b1 + (uint32_t)(ks * 32), 16, 1024, 2), idesc,
                            ks ? 1u : 0u);
                if (ph > 0) {
                    const uint32_t a2 = (ph == 1 ? sG : sV) + (uint32_t)(t * 128 * 128);
                    const uint32_t b2 = (ph == 1 ? sV : sG) + (uint32_t)(col0 * 128);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        bmma_ss(buf + AB_CW, bdesc(a2 + (uint32_t)(ks * 32), 16, 1024, 2),
                                bdesc(b2 + (uint32_t)(ks * 32), 16, 1024, 2), idesc, ks ? 1u : 0u);
                }
                bcommit(&s_full[x]);
            };
            // (pass, c) of a chunk index, advanced incrementally (no divisions in the loop)
            int c = x, pass = 0, t = 0, ph = 0;           // current chunk g
            auto norm = [&](int& c_, int& pass_, int& t_, int& ph_) {
                while (c_ >= nch) { c_ -= nch; ++pass_; if (++t_ == nmt) { t_ = 0; ++ph_; } }
            };
            norm(c, pass, t, ph);
            if (x == 0) btrace(1000);
            bbar_wait(&rdy[0], 0u);                       // Q and K rounded
            if (x == 0) btrace(1001);
            bool vg_ready = false;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_s(ph, t, c);
            if (x == 0) {
                bbar_wait(o_done, 0u);
                load(2, &mapQK, 2 * a.C + hc);           // V
            }
            uint32_t k = 0;                               // own chunks done
            bool first_in_pass = true;                    // the next OUT is this issuer's first of its pass
            for (int g = x; g < G; g += 2, ++k) {
                bbar_wait(&p_ready[x], k & 1u);
                btrace(g * 8 + 2);
                if (first_in_pass) {
                    if (t == 0 && c < 2) {                // this issuer's first chunk of a phase
                        if (ph > 0 && x == 0) {
                            // the previous phase's accumulating products (both issuers) have read slot 4: bring in this
                            // phase's MN-major operand
                            const uint32_t pp = (uint32_t)(pass - 1);
                            bbar_wait(&acc_full[(pp & 1)], (pp >> 1) & 1u);
                            bbar_wait(&acc_full[2 + (pp & 1)], (pp >> 1) & 1u);
                            if (ph == 1) load(4, &mapQM, a.C + hc);      // K, MN-major
                            else load(4, &mapQM, hc);                     // Q, MN-major
                        }
                        bbar_wait(&full[4], (uint32_t)(ph & 1));
                    }
                    if (pass >= 2) bbar_wait(&acc_free[pass & 1], (uint32_t)(((pass >> 1) - 1) & 1));
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                {
                    const int col0 = c * AB_CW;
                    const int ksteps = min(AB_CW, a.Nk - col0) >> 3;
                    const uint32_t acc = tmem + (uint32_t)(AB_ACC_COL + 32 * (2 * x + (pass & 1)));
                    uint64_t bd = bdesc(sM + (uint32_t)((col0 >> 3) * 1024), 4096, 512, 1);
                    for (int kk = 0; kk < ksteps; ++kk, bd += 1024 >> 4)
                        bmma_ts(acc, buf + (uint32_t)(kk * 8), bd, idesc_o, (!first_in_pass || kk) ? 1u : 0u);
                }
                btrace(g * 8 + 3);
                // next own chunk
                int c2 = c + 2, pass2 = pass, t2 = t, ph2 = ph;
                norm(c2, pass2, t2, ph2);
                const bool more = g + 2 < G;
                if (pass2 != pass || !more) bcommit(&acc_full[2 * x + (pass & 1)]);     // that was my last chunk of the pass
                first_in_pass = pass2 != pass;
                if (more) {
                    if (!vg_ready && ph2 > 0) { bbar_wait(&rdy[1], 0u); vg_ready = true; }   // V and dO rounded
                    btrace(g * 8);
                    issue_s(ph2, t2, c2);                  // same buffer: ordered behind OUT(g) on this thread
                    btrace(g * 8 + 1);
                }
                c = c2; pass = pass2; t = t2; ph = ph2;
            }
        }
    } else {
        // ---------------------------------------------------------------------- element-wise warps: both groups work on every
        // chunk, group x on the x-th half of its columns
        const int x = warp >= 6 ? 1 : 0;
        const int q = warp & 3;                       // TMEM lane quarter of this warp
        const int r = q * 32 + lane;                  // lane inside the 128-row tile
        const int tid = threadIdx.x - 64;             // 0..255
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const bool tr = lane == 0 && (warp == 2 || warp == 6);
        if (tr && warp == 2) btrace(1002);
        const int wplane = a.ww * a.wh;
        auto tok_row = [&](int i) -> int64_t {        // token i of the window -> row of the (P, .) buffers
            const int iz = i / wplane, rem = i - iz * wplane;
            const int iy = rem / a.ww, ix = rem - iy * a.ww;
            return row0 + ((int64_t)iz * a.Hp + iy) * a.Wp + ix;
        };
        // ---- item start.  lse first (its global latency hides behind the tile loads) ...
        float lreg[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = tid + 256 * u;
            lreg[u] = i < a.N ? a.lse[tok_row(i) * a.heads + head] * 1.4426950408889634f : 0.f;
        }
        // operand conditioning: the tensor core truncates fp32 to TF32; round the K-major tiles to nearest instead
        const int nvec = a.N * 8;
        auto round_slot = [&](int s) {
            uint4* p4 = reinterpret_cast<uint4*>(ring + (size_t)s * AB_SLOT_BYTES);
#pragma unroll 4
            for (int i = tid; i < nvec; i += 256) {
                uint4 v = p4[i];
                v.x = (v.x + 0x1000u) & 0xFFFFE000u; v.y = (v.y + 0x1000u) & 0xFFFFE000u;
                v.z = (v.z + 0x1000u) & 0xFFFFE000u; v.w = (v.w + 0x1000u) & 0xFFFFE000u;
                p4[i] = v;
            }
        };
        bbar_wait(&full[0], 0u);
        round_slot(0);
        bbar_wait(&full[1], 0u);
        round_slot(1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bbar_arrive(&rdy[0]);                                     // the issuers start the first score products
        if (tr && warp == 2) btrace(1003);
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (tid + 256 * u < AB_STAT) lse2[tid + 256 * u] = lreg[u];
        // ... delta = <dO, O> from the dO tile (slot 3) and the O tile that visits slot 2: 8 threads per token; both tiles
        // carry the same swizzle, so chunk `sub` of a row holds the same 4 channels in both
        bbar_wait(&full[2], 0u);
        bbar_wait(&full[3], 0u);
        {
            const int sub = tid & 7;
#pragma unroll 4
            for (int i0 = 0; i0 < AB_STAT; i0 += 32) {
                const int i = i0 + (tid >> 3);
                float v = 0.f;
                if (i < a.N) {
                    const float4 gv = *reinterpret_cast<const float4*>(ring + 3 * (size_t)AB_SLOT_BYTES + (size_t)i * 128 + sub * 16);
                    const float4 ov = *reinterpret_cast<const float4*>(ring + 2 * (size_t)AB_SLOT_BYTES + (size_t)i * 128 + sub * 16);
                    v = (gv.x * ov.x + gv.y * ov.y) + (gv.z * ov.z + gv.w * ov.w);
                }
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (sub == 0) del[i] = v;
            }
        }
        bbar_arrive(o_done);                                      // slot 2 may take V now
        round_slot(3);
        asm volatile("bar.sync 1, 256;" ::: "memory");           // statistics visible to all element-wise warps
        if (tr && warp == 2) btrace(1004);

        // epilogue of a pass: sum of the two issuers' accumulators -> global (dV unscaled; dQ, dK x scale), `ncol` columns
        // from column `cb` of this thread's lane
        auto epilogue = [&](int pass, int cb, int nhalf, bool release) {
            const int ph = pass / nmt, t = pass - ph * nmt;
            const uint32_t pp = (uint32_t)pass & 1u, par = ((uint32_t)pass >> 1) & 1u;
            bbar_wait(&acc_full[pp], par);
            bbar_wait(&acc_full[2 + pp], par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int li = t * 128 + r;
            float* dst = nullptr;
            if (li < a.N) {
                const int64_t row = tok_row(li);
                dst = (ph == 0 ? a.dv + row * a.lddkv : (ph == 1 ? a.dq + row * a.lddq : a.dk + row * a.lddkv)) + head * 32;
            }
            const float f = ph == 0 ? 1.f : a.scale;
            for (int hh = 0; hh < nhalf; ++hh) {
                uint32_t e0[16], e1[16];
                const int cc = cb + 16 * hh;
                btld16(lane_addr + (uint32_t)(AB_ACC_COL + 32 * pp + cc), e0);
                btld16(lane_addr + (uint32_t)(AB_ACC_COL + 32 * (2 + pp) + cc), e1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (dst) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        *reinterpret_cast<float4*>(dst + cc + 4 * e) =
                            make_float4((__uint_as_float(e0[4 * e]) + __uint_as_float(e1[4 * e])) * f,
                                        (__uint_as_float(e0[4 * e + 1]) + __uint_as_float(e1[4 * e + 1])) * f,
                                        (__uint_as_float(e0[4 * e + 2]) + __uint_as_float(e1[4 * e + 2])) * f,
                                        (__uint_as_float(e0[4 * e + 3]) + __uint_as_float(e1[4 * e + 3])) * f);
                }
            }
            if (release) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                bbar_arrive(&acc_free[pp]);
            }
        };

        uint32_t va[48], vb[48];
        int c = 0, pass = 0, t = 0, ph = 0;
        for (int g = 0; g < G; ++g) {
            const int b = g & 1;
            const int col0 = c * AB_CW;
            const int wdt = min(AB_CW, a.Nk - col0);
            const int half = wdt >> 1;                           // multiple of 16
            const int nun = half >> 4;                           // 16-column units per half
            const int li = t * 128 + r;                          // token of this lane (query in phase 1, key otherwise)
            const float my_l2 = lse2[li], my_del = del[li];
            // Normally group x turns the x-th half of the chunk's columns into P / dS.  On the second chunk of a pass one
            // group stores the previous pass's result (its accumulators completed a chunk ago) while the other group takes
            // both halves: the epilogue's scattered 128-byte row stores stay off the chunk pipeline.
            const bool epi = c == 1 && pass > 0;
            const int egrp = (pass - 1) & 1;
            if (tr) btrace(g * 8 + 4);
            if (epi && x == egrp) {
                // nothing of this chunk is mine -- but its s_full phase is still observed: a waiter that skips a phase
                // would take the NEXT use of this parity for complete the moment it looks
                bbar_wait(&s_full[b], (uint32_t)((g >> 1) & 1));
                bbar_arrive(&p_ready[b]);
                epilogue(pass - 1, 0, 2, true);
                if (tr) btrace(g * 8 + 7);
            } else {
                bbar_wait(&s_full[b], (uint32_t)((g >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tr) btrace(g * 8 + 5);
                const int h0 = epi ? 0 : x, h1 = epi ? 2 : x + 1;
                for (int hh = h0; hh < h1; ++hh) {
                    const int cbeg = hh * half;                  // first column (inside the chunk) of this half
                    const uint32_t sbase = lane_addr + (uint32_t)(b * AB_BUF_COLS + cbeg);
                    // all TMEM loads of the half are issued before the first use: one load latency
#pragma unroll
                    for (int u = 0; u < 3; ++u)
                        if (u < nun) {
                            btld16(sbase + (uint32_t)(u * 16), va + 16 * u);
                            if (ph > 0) btld16(sbase + (uint32_t)(AB_CW + u * 16), vb + 16 * u);
                        }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        if (u < nun) {
                            uint32_t* sa = va + 16 * u;
                            uint32_t* sb = vb + 16 * u;
                            const int j0 = col0 + cbeg + u * 16;         // token of the first column of this unit
                            if (ph == 1) {
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    const float p = bex2(fmaf(__uint_as_float(sa[e]), a.scale_log2, -my_l2));
                                    sa[e] = rn_tf32(p * (__uint_as_float(sb[e]) - my_del));
                                }
                            } else if (ph == 0) {
#pragma unroll
                                for (int e4 = 0; e4 < 4; ++e4) {
                                    const float4 l4 = *reinterpret_cast<const float4*>(lse2 + j0 + 4 * e4);
                                    sa[4 * e4] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4]), a.scale_log2, -l4.x)));
                                    sa[4 * e4 + 1] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4 + 1]), a.scale_log2, -l4.y)));
                                    sa[4 * e4 + 2] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4 + 2]), a.scale_log2, -l4.z)));
                                    sa[4 * e4 + 3] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4 + 3]), a.scale_log2, -l4.w)));
                                }
                            } else {
#pragma unroll
                                for (int e4 = 0; e4 < 4; ++e4) {
                                    const float4 l4 = *reinterpret_cast<const float4*>(lse2 + j0 + 4 * e4);
                                    const float4 d4 = *reinterpret_cast<const float4*>(del + j0 + 4 * e4);
                                    sa[4 * e4] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4]), a.scale_log2, -l4.x)) *
                                                         (__uint_as_float(sb[4 * e4]) - d4.x));
                                    sa[4 * e4 + 1] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4 + 1]), a.scale_log2, -l4.y)) *
                                                             (__uint_as_float(sb[4 * e4 + 1]) - d4.y));
                                    sa[4 * e4 + 2] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4 + 2]), a.scale_log2, -l4.z)) *
                                                             (__uint_as_float(sb[4 * e4 + 2]) - d4.z));
                                    sa[4 * e4 + 3] = rn_tf32(bex2(fmaf(__uint_as_float(sa[4 * e4 + 3]), a.scale_log2, -l4.w)) *
                                                             (__uint_as_float(sb[4 * e4 + 3]) - d4.w));
                                }
                            }
                            if (j0 + 16 > a.N) {                         // columns of padding tokens contribute nothing
#pragma unroll
                                for (int e = 0; e < 16; ++e) if (j0 + e >= a.N) sa[e] = 0u;
                            }
                            btst16(sbase + (uint32_t)(u * 16), sa);
                        }
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");    // also: the registers are reloaded next
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                bbar_arrive(&p_ready[b]);
                if (tr) btrace(g * 8 + 6);
            }
            // V replaces O in slot 2 during phase 0: round it once it has landed (phase 1 starts at chunk nmt * nch >= 4)
            if (g == 1) {
                bbar_wait(&full[2], 1u);
                round_slot(2);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bbar_arrive(&rdy[1]);
            }
            if (++c == nch) { c = 0; ++pass; if (++t == nmt) { t = 0; ++ph; } }
        }
        epilogue(3 * nmt - 1, 16 * x, 1, false);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}


typedef CUresult (*EncodeTiledFn5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int tc_window_attn_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* out,
                       const float* dout, int ldo, const float* lse, float* dq, int lddq, float* dk, float* dv, int lddkv,
                       int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh, int ww, float scale,
                       cudaStream_t st) {
    const int N = wd * wh * ww;
    const int C = heads * hd;
    // taken only for the fused (P, 3C) layout: q | k | v column blocks of one buffer (what the forward takes)
    if (hd != 32 || N < 128 || N > 352 || wd > 256 || wh > 256 || ww > 256) return MIC_ERR_UNSUPPORTED;
    if (ldq != ldkv || k != q + C || v != q + 2 * C || (ldq & 3) || (reinterpret_cast<uintptr_t>(q) & 15) || (ldo & 3) ||
        (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(dout) & 15) || (lddq & 3) || (lddkv & 3) ||
        (reinterpret_cast<uintptr_t>(dq) & 15) || (reinterpret_cast<uintptr_t>(dk) & 15) || (reinterpret_cast<uintptr_t>(dv) & 15))
        return MIC_ERR_UNSUPPORTED;
    if (Dp % wd || Hp % wh || Wp % ww) return MIC_ERR_UNSUPPORTED;
    const int64_t items = (int64_t)B * (Dp / wd) * (Hp / wh) * (Wp / ww) * heads;
    if (items <= 0 || items >= ((int64_t)1 << 31)) return MIC_ERR_UNSUPPORTED;
    static EncodeTiledFn5 enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return MIC_ERR_UNSUPPORTED;
        enc = reinterpret_cast<EncodeTiledFn5>(p);
    }
    CUtensorMap mQK, mQM, mGK, mGM, mOK;
    cuuint32_t box[5] = {32, (cuuint32_t)ww, (cuuint32_t)wh, (cuuint32_t)wd, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    auto encode = [&](CUtensorMap* m, const float* base, int cols, int ld, CUtensorMapSwizzle sw) {
        cuuint64_t dims[5] = {(cuuint64_t)cols, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)Dp, (cuuint64_t)B};
        cuuint64_t strides[4] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * Wp, (cuuint64_t)ld * 4 * Wp * Hp,
                                 (cuuint64_t)ld * 4 * Wp * Hp * Dp};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!encode(&mQK, q, 3 * C, ldq, CU_TENSOR_MAP_SWIZZLE_128B) || !encode(&mQM, q, 3 * C, ldq, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !encode(&mGK, dout, C, ldo, CU_TENSOR_MAP_SWIZZLE_128B) || !encode(&mGM, dout, C, ldo, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) || !encode(&mOK, out, C, ldo, CU_TENSOR_MAP_SWIZZLE_128B))
        return MIC_ERR_UNSUPPORTED;
    AttnBwdArgs a{};
    a.out = out; a.dout = dout; a.ldo = ldo; a.lse = lse; a.dq = dq; a.lddq = lddq; a.dk = dk; a.dv = dv; a.lddkv = lddkv;
    a.Dp = Dp; a.Hp = Hp; a.Wp = Wp; a.heads = heads; a.C = C; a.wd = wd; a.wh = wh; a.ww = ww;
    a.nwd = Dp / wd; a.nwh = Hp / wh; a.nww = Wp / ww; a.N = N;
    a.Nk = ((N + 31) / 32) * 32;
    a.scale = scale;
    a.scale_log2 = scale * 1.4426950408889634f;
    const size_t smem = 1024 + (size_t)AB_SLOTS * AB_SLOT_BYTES + 256 + 2 * AB_STAT * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(window_attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    mic::launch(window_attn_tc_bwd_kernel, dim3((unsigned)items), dim3(AB_THREADS), smem, st, mQK, mQM, mGK, mGM, mOK, a);
    return check_launch("window_attn_tc_bwd_kernel");
}

}  // namespace mic
extern "C" int mic_debug_attn_bwd_trace(void* buf) {
    long long* p = reinterpret_cast<long long*>(buf);
    return cudaMemcpyToSymbol(mic::g_ab_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
