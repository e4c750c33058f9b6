// tcgen05 / TMA windowed multi-head attention forward for large windows (128 <= tokens <= 352, head_dim 32):
// the Cross-Modal Q(CT) K(MR)^T V(MR) kernel of BASELINE.json's isolation config (343-token 7x7x7 windows, 96 ch,
// 3 heads) and the window-7 model configs.  FlashAttention-style, one persistent CTA per SM:
//
//   warp 0    TMA producer: ONE 5-D tensor-map box {32 ch, ww, wh, wd, 1} gathers the Q / K / V head slice of a whole
//             window straight out of the token grid (window_partition is the box shape) into a 5-slot smem ring.
//   warp 1    MMA issuer:  S = Q K^T   tcgen05.mma.kind::tf32, A,B from smem (K-major, 128B swizzle), S in TMEM
//                          O = P V     A = P read back from TMEM, B = V from smem (MN-major, 128B/32B-atom swizzle)
//   warps 2-5 softmax + epilogue: thread = query row; tcgen05.ld S, max / exp2 / sum in registers, tcgen05.st P in
//             place, then O / l and the log-sum-exp to global (window_reverse is the store address).
//
// TMEM: S/P 352 columns + O 32 columns.  Scores never touch shared or global memory.  fp32 in / fp32 out, operands
// read as TF32, fp32 accumulation (the HBM floor with fp32 I/O is above the TF32 tensor time).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace mic {

constexpr int AT_THREADS = 192;
constexpr int AT_SLOTS = 5;
constexpr int AT_SLOT_BYTES = 352 * 128;        // 45056: up to 352 token rows x 32 fp32
constexpr int AT_OCOL = 352;                    // TMEM column of the O accumulator

__device__ unsigned long long* g_at_trace = nullptr;
__device__ __forceinline__ void atrace(int slot) {
    unsigned long long* t = g_at_trace;
    if (t && blockIdx.x == 0 && slot < 64) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        t[slot] = now;
    }
}

struct AttnTcArgs {
    float* out; int ldo;
    float* lse;
    int B, Dp, Hp, Wp, heads, C, wd, wh, ww, nwd, nwh, nww;
    int N;            // tokens per window
    int Nk;           // key columns (N rounded up to 16 / 32)
    int64_t items;    // windows * heads
    float scale_log2; // head_dim^-0.5 * log2(e)
    float scale;
};

__device__ __forceinline__ uint32_t asmem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void abar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(asmem(b)), "r"(c));
}
__device__ __forceinline__ void abar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(asmem(b)) : "memory");
}
__device__ __forceinline__ void abar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(asmem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void abar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AW_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra AW_DONE;\n"
        "bra AW_LOOP;\n"
        "AW_DONE:\n"
        "}\n" ::"r"(asmem(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool abar_test(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(asmem(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void acommit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(asmem(b)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(asmem(dst)), "l"(map), "r"(asmem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ uint64_t adesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void amma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void amma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tld32_nowait(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tst32(uint32_t taddr, const uint32_t* r) {
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

__global__ void __launch_bounds__(AT_THREADS, 1)
window_attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapV, AttnTcArgs a) {
    extern __shared__ uint8_t at_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = sm;
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + AT_SLOTS * AT_SLOT_BYTES + 4096);   // +4 KB: Q tile over-read pad
    uint64_t* empty = full + AT_SLOTS;
    uint64_t* s_full = empty + AT_SLOTS;
    uint64_t* p_full = s_full + 1;
    uint64_t* o_full = p_full + 1;
    uint64_t* o_empty = o_full + 1;
    uint64_t* rdy_qk = o_empty + 1;          // Q and K tiles rounded to nearest TF32 (128 arrivals per item)
    uint64_t* rdy_v = rdy_qk + 1;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(rdy_v + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nmt = (a.N + 127) / 128;
    const uint32_t box_bytes = (uint32_t)a.N * 128u;

    // rows beyond the window box are never written by TMA: zero the ring once so they stay finite (0 * garbage = NaN)
    for (int i = threadIdx.x; i < (AT_SLOTS * AT_SLOT_BYTES + 4096) / 16; i += AT_THREADS)
        reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < AT_SLOTS; ++s) { abar_init(&full[s], 1); abar_init(&empty[s], 1); }
        abar_init(s_full, 1); abar_init(p_full, 128); abar_init(o_full, 1); abar_init(o_empty, 128);
        abar_init(rdy_qk, 128); abar_init(rdy_v, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(asmem(tslot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    pdl_sync();            // prologue above (ring clear, barriers, TMEM) overlaps the previous kernel

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQK) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
            int64_t u = 0;                                          // slot use counter
            for (int64_t it = blockIdx.x; it < a.items; it += gridDim.x) {
                const int head = (int)(it % a.heads);
                int64_t w = it / a.heads;
                const int wx = (int)(w % a.nww); w /= a.nww;
                const int wy = (int)(w % a.nwh); w /= a.nwh;
                const int wz = (int)(w % a.nwd); w /= a.nwd;
                const int b = (int)w;
                const int cx = wx * a.ww, cy = wy * a.wh, cz = wz * a.wd;
#pragma unroll
                for (int part = 0; part < 3; ++part, ++u) {        // 0: Q, 1: K, 2: V
                    const int s = (int)(u % AT_SLOTS);
                    abar_wait(&empty[s], (uint32_t)((u / AT_SLOTS) & 1) ^ 1u);
                    abar_expect(&full[s], box_bytes);
                    tma_load_5d(ring + s * AT_SLOT_BYTES, part == 2 ? &mapV : &mapQK, &full[s], part * a.C + head * 32, cx, cy,
                                cz, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // S = Q K^T : M=128, N = Nk or Nk/2, both K-major;  O = P V : M=128, N=32, A from TMEM, B MN-major
            const int nsplit = a.Nk > 256 ? 2 : 1;
            const int nhalf = a.Nk / nsplit;
            const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nhalf >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                                     ((uint32_t)(128 >> 4) << 24);
            const int ksteps_pv = (a.N + 7) / 8;
            int64_t u = 0, tile = 0, item = 0;
            for (int64_t it = blockIdx.x; it < a.items; it += gridDim.x, u += 3, ++item) {
                const int sq = (int)(u % AT_SLOTS), sk = (int)((u + 1) % AT_SLOTS), sv = (int)((u + 2) % AT_SLOTS);
                if (item == 2) atrace(0);
                abar_wait(rdy_qk, (uint32_t)(item & 1));
                if (item == 2) atrace(1);
                const uint32_t qa = asmem(ring + sq * AT_SLOT_BYTES), ka = asmem(ring + sk * AT_SLOT_BYTES),
                               va = asmem(ring + sv * AT_SLOT_BYTES);
                for (int mt = 0; mt < nmt; ++mt, ++tile) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int h = 0; h < nsplit; ++h)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t ad = adesc(qa + (uint32_t)(mt * 128 * 128 + ks * 32), 16, 1024, 2);
                            const uint64_t bd = adesc(ka + (uint32_t)(h * nhalf * 128 + ks * 32), 16, 1024, 2);
                            amma_ss(tmem + (uint32_t)(h * nhalf), ad, bd, idesc_s, ks ? 1u : 0u);
                        }
                    acommit(s_full);
                    if (item == 2) atrace(2 + mt * 4);
                    if (mt == nmt - 1) { acommit(&empty[sq]); acommit(&empty[sk]); }
                    if (mt == 0) abar_wait(rdy_v, (uint32_t)(item & 1));
                    abar_wait(p_full, (uint32_t)(tile & 1));
                    if (item == 2) atrace(3 + mt * 4);
                    if (tile > 0) abar_wait(o_empty, (uint32_t)((tile - 1) & 1));
                    if (item == 2) atrace(4 + mt * 4);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kk = 0; kk < ksteps_pv; ++kk) {
                        const uint64_t bd = adesc(va + (uint32_t)(kk * 1024), 4096, 512, 1);
                        amma_ts(tmem + AT_OCOL, tmem + (uint32_t)(kk * 8), bd, idesc_o, kk ? 1u : 0u);
                    }
                    acommit(o_full);
                    if (item == 2) atrace(5 + mt * 4);
                    if (mt == nmt - 1) acommit(&empty[sv]);
                }
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter of this warp
        const int r = q * 32 + lane;                  // query row inside the 128-row tile
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        int64_t tile = 0, u = 0;
        const int et = (warp - 2) * 32 + lane;       // 0..127
        const int nvec = a.N * 8;                     // float4 per operand tile
        auto round_tile = [&](uint8_t* base) {
            float4* p4 = reinterpret_cast<float4*>(base);
#pragma unroll 4
            for (int i = et; i < nvec; i += 128) {
                // round-to-nearest TF32 with integer ALU ops (add half an ulp of the 10-bit mantissa, clear 13 bits);
                // cvt.rna.tf32 goes through the quarter-rate conversion pipe
                uint4 t = reinterpret_cast<uint4*>(p4)[i];
                t.x = (t.x + 0x1000u) & 0xFFFFE000u; t.y = (t.y + 0x1000u) & 0xFFFFE000u;
                t.z = (t.z + 0x1000u) & 0xFFFFE000u; t.w = (t.w + 0x1000u) & 0xFFFFE000u;
                reinterpret_cast<uint4*>(p4)[i] = t;
            }
        };
        const int nch = a.Nk / 32;                    // 32-column chunks of S (Nk is a multiple of 32 when > 256)
        int64_t sitem = 0;
        for (int64_t it = blockIdx.x; it < a.items; it += gridDim.x, u += 3, ++sitem) {
            const bool tr = sitem == 2 && warp == 2 && lane == 0;
            const int head = (int)(it % a.heads);
            int64_t w = it / a.heads;
            const int wx = (int)(w % a.nww); w /= a.nww;
            const int wy = (int)(w % a.nwh); w /= a.nwh;
            const int wz = (int)(w % a.nwd); w /= a.nwd;
            const int b = (int)w;
            if (tr) atrace(20);
            // operand conditioning: the tensor core truncates fp32 to TF32; round to nearest instead (unbiased)
            {
                const int sq = (int)(u % AT_SLOTS), sk = (int)((u + 1) % AT_SLOTS), sv = (int)((u + 2) % AT_SLOTS);
                abar_wait(&full[sq], (uint32_t)((u / AT_SLOTS) & 1));
                round_tile(ring + sq * AT_SLOT_BYTES);
                abar_wait(&full[sk], (uint32_t)(((u + 1) / AT_SLOTS) & 1));
                round_tile(ring + sk * AT_SLOT_BYTES);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                abar_arrive(rdy_qk);
                abar_wait(&full[sv], (uint32_t)(((u + 2) / AT_SLOTS) & 1));
                round_tile(ring + sv * AT_SLOT_BYTES);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                abar_arrive(rdy_v);
                if (tr) atrace(21);
            }
            for (int mt = 0; mt < nmt; ++mt, ++tile) {
                abar_wait(s_full, (uint32_t)(tile & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tr) atrace(22 + mt * 5);
                // TMEM reads are software-pipelined: chunk c+1 is in flight while chunk c is reduced
                uint32_t va[32], vb[32];
                // pass 1: row maximum over the valid key columns
                float m = -INFINITY;
                tld32_nowait(lane_addr, va);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int c = 0; c < nch; c += 2) {
                    if (c + 1 < nch) tld32_nowait(lane_addr + (uint32_t)((c + 1) * 32), vb);
                    if ((c + 1) * 32 <= a.N) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) m = fmaxf(m, __uint_as_float(va[t]));
                    } else {
#pragma unroll
                        for (int t = 0; t < 32; ++t) if (c * 32 + t < a.N) m = fmaxf(m, __uint_as_float(va[t]));
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c + 1 < nch) {
                        if (c + 2 < nch) tld32_nowait(lane_addr + (uint32_t)((c + 2) * 32), va);
                        if ((c + 2) * 32 <= a.N) {
#pragma unroll
                            for (int t = 0; t < 32; ++t) m = fmaxf(m, __uint_as_float(vb[t]));
                        } else {
#pragma unroll
                            for (int t = 0; t < 32; ++t) if ((c + 1) * 32 + t < a.N) m = fmaxf(m, __uint_as_float(vb[t]));
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                }
                if (tr) atrace(23 + mt * 5);
                // pass 2: p = exp2((s - m) * scale * log2 e) rounded to TF32, row sum, P written in place
                const float mb = m * a.scale_log2;
                float l = 0.f;
                auto expchunk = [&](uint32_t (&v)[32], int c) {
                    const bool full_chunk = (c + 1) * 32 <= a.N;
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        float p;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(__uint_as_float(v[t]), a.scale_log2, -mb)));
                        if (!full_chunk && c * 32 + t >= a.N) p = 0.f;
                        const uint32_t pr = (__float_as_uint(p) + 0x1000u) & 0xFFFFE000u;     // RN to TF32
                        l += __uint_as_float(pr);
                        v[t] = pr;
                    }
                    tst32(lane_addr + (uint32_t)(c * 32), v);
                };
                tld32_nowait(lane_addr, va);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int c = 0; c < nch; c += 2) {
                    if (c + 1 < nch) tld32_nowait(lane_addr + (uint32_t)((c + 1) * 32), vb);
                    expchunk(va, c);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c + 1 < nch) {
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");   // va is reloaded next
                        if (c + 2 < nch) tld32_nowait(lane_addr + (uint32_t)((c + 2) * 32), va);
                        expchunk(vb, c + 1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                abar_arrive(p_full);
                if (tr) atrace(24 + mt * 5);
                // epilogue of this tile
                abar_wait(o_full, (uint32_t)(tile & 1));
                if (tr) atrace(25 + mt * 5);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t o[32];
                tld32(lane_addr + AT_OCOL, o);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                abar_arrive(o_empty);
                const int i = mt * 128 + r;           // token index inside the window
                if (i < a.N) {
                    const int ix = i % a.ww, iy = (i / a.ww) % a.wh, iz = i / (a.ww * a.wh);
                    const int64_t row = (((int64_t)b * a.Dp + wz * a.wd + iz) * a.Hp + wy * a.wh + iy) * a.Wp + wx * a.ww + ix;
                    const float inv = 1.f / l;
                    float* dst = a.out + row * a.ldo + head * 32;
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        *reinterpret_cast<float4*>(dst + 4 * t) =
                            make_float4(__uint_as_float(o[4 * t]) * inv, __uint_as_float(o[4 * t + 1]) * inv,
                                        __uint_as_float(o[4 * t + 2]) * inv, __uint_as_float(o[4 * t + 3]) * inv);
                    a.lse[row * a.heads + head] = m * a.scale + logf(l);
                }
                if (tr) atrace(26 + mt * 5);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// v2: two softmax pipelines per CTA.  v1 serialises  S-MMA -> softmax -> PV-MMA  per 128-row tile because one fp32 S
// tile (352 columns) nearly fills TMEM; the SFU (ex2) then idles during the MMAs and the tensor core during the
// softmax.  Here the keys are split into two blocks (192 | N-192) handled with a two-block online softmax whose
// partial outputs O0 / O1 are kept in separate accumulators and merged in the epilogue (out = (O0*alpha + O1) / l),
// so one tile needs only 192 + 2*32 = 256 TMEM columns and TWO tiles are in flight: while pipeline A's four warps
// exponentiate a block, the tensor core runs pipeline B's QK^T / PV, and vice versa.
//   warp 0       TMA producer (as v1)
//   warp 1       MMA issuer for both pipelines (fixed interleave A/B)
//   warps 2-5    softmax + epilogue, pipeline A (even tiles)
//   warps 6-9    softmax + epilogue, pipeline B (odd tiles)
//   warps 10-11  operand conditioning: round Q/K/V to nearest TF32 in shared memory (off the softmax warps' path)
constexpr int A2_THREADS = 384;            // 12 warps -> 168 registers per thread
constexpr int A2_B0 = 192;                      // keys in block 0 (6 x 32-column chunks)
constexpr int A2_PIPE_COLS = 256;               // S/P 192 + O0 32 + O1 32

__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], float m, int valid) {
    float m0 = m, m1 = m, m2 = m, m3 = m;                  // four chains instead of one 32-deep dependency
    if (valid >= 32) {
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[t])); m1 = fmaxf(m1, __uint_as_float(v[t + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[t + 2])); m3 = fmaxf(m3, __uint_as_float(v[t + 3]));
        }
    } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) if (t < valid) m0 = fmaxf(m0, __uint_as_float(v[t]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// v <- RN_tf32(exp2(v * sc - mb)) (0 beyond `valid`); returns the sum of the rounded values
__device__ __forceinline__ float chunk_exp(uint32_t (&v)[32], float sc, float mb, int valid) {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        float p;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(__uint_as_float(v[t]), sc, -mb)));
        if (valid < 32 && t >= valid) p = 0.f;
        const uint32_t pr = (__float_as_uint(p) + 0x1000u) & 0xFFFFE000u;
        if ((t & 3) == 0) l0 += __uint_as_float(pr);
        else if ((t & 3) == 1) l1 += __uint_as_float(pr);
        else if ((t & 3) == 2) l2 += __uint_as_float(pr);
        else l3 += __uint_as_float(pr);
        v[t] = pr;
    }
    return (l0 + l1) + (l2 + l3);
}

// One key block of the two-block softmax for this thread's row: S (nch 32-column chunks at taddr, `valid` real columns)
// -> m = max(m_in, row max), P = RN_tf32(exp2((S - m) * sc)) written in place, l = sum(P).  The max pass streams the
// chunks through va (even) / vb (odd); the last two chunks are still in registers when the exp pass starts, so only
// nch - 2 chunks are read from TMEM twice (TMEM reads are the scarcest port of this kernel).
__device__ __forceinline__ void softmax_block(uint32_t taddr, int nch, int valid, float m_in, float sc, float& m_out,
                                              float& l_out, uint32_t (&va)[32], uint32_t (&vb)[32]) {
    float m = m_in;
    tld32_nowait(taddr, va);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < nch; c += 2) {
        if (c + 1 < nch) tld32_nowait(taddr + (uint32_t)((c + 1) * 32), vb);
        m = chunk_max(va, m, valid - c * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c + 1 < nch) {
            if (c + 2 < nch) tld32_nowait(taddr + (uint32_t)((c + 2) * 32), va);
            m = chunk_max(vb, m, valid - (c + 1) * 32);
            if (c + 2 < nch) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
    }
    const float mb = m * sc;
    float l = 0.f;
    const int L = nch - 1;                            // va holds the last even chunk, vb the last odd one
    if (L & 1) {
        l += chunk_exp(vb, sc, mb, valid - L * 32);
        tst32(taddr + (uint32_t)(L * 32), vb);
        l += chunk_exp(va, sc, mb, valid - (L - 1) * 32);
        tst32(taddr + (uint32_t)((L - 1) * 32), va);
    } else {
        l += chunk_exp(va, sc, mb, valid - L * 32);
        tst32(taddr + (uint32_t)(L * 32), va);
        if (L >= 1) {
            l += chunk_exp(vb, sc, mb, valid - (L - 1) * 32);
            tst32(taddr + (uint32_t)((L - 1) * 32), vb);
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const int nrem = L - 1;                           // chunks [0, nrem) are read again
    if (nrem > 0) {
        tld32_nowait(taddr, va);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < nrem; c += 2) {
            if (c + 1 < nrem) tld32_nowait(taddr + (uint32_t)((c + 1) * 32), vb);
            l += chunk_exp(va, sc, mb, 32);
            tst32(taddr + (uint32_t)(c * 32), va);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (c + 1 < nrem) {
                if (c + 2 < nrem) tld32_nowait(taddr + (uint32_t)((c + 2) * 32), va);
                l += chunk_exp(vb, sc, mb, 32);
                tst32(taddr + (uint32_t)((c + 1) * 32), vb);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
    }
    m_out = m;
    l_out = l;
}

// Single-pass variant for a row whose shift mb is known up front (see the Cauchy-Schwarz bound in the kernel): every chunk
// is read from TMEM exactly once.  Chunk 0 may already sit in va.
__device__ __forceinline__ float exp_block(uint32_t taddr, int nch, int valid, float sc, float mb, uint32_t (&va)[32],
                                           uint32_t (&vb)[32], bool va_loaded, uint64_t* pub = nullptr) {
    // pub (two-issuer kernel): arrive there once the first three chunks (96 columns of P) are in TMEM, so that their
    // P V products are issued while the rest of the block is still being exponentiated
    float l = 0.f;
    if (!va_loaded) {
        tld32_nowait(taddr, va);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    for (int c = 0; c < nch; c += 2) {
        if (c + 1 < nch) tld32_nowait(taddr + (uint32_t)((c + 1) * 32), vb);
        l += chunk_exp(va, sc, mb, valid - c * 32);
        tst32(taddr + (uint32_t)(c * 32), va);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (pub != nullptr && c == 2 && nch > 3) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            abar_arrive(pub);
        }
        if (c + 1 < nch) {
            if (c + 2 < nch) tld32_nowait(taddr + (uint32_t)((c + 2) * 32), va);
            l += chunk_exp(vb, sc, mb, valid - (c + 1) * 32);
            tst32(taddr + (uint32_t)((c + 1) * 32), vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    return l;
}

// ISSUERS = 1 (MICFORMER_ATTN_ISSUERS=1): warp 0 is the TMA producer and warp 1 issues the MMAs of both pipelines.
// ISSUERS = 2 (default since round 2: 1.449 -> 1.155 ms on config 4, gpurun_out r2a; all window-attention parity tests pass
// with it): one issuing thread per pipeline -- a K=8 TF32 tcgen05.mma costs its
// issuing thread ~100 clocks whatever N is (profiles/r01_ubench_b200.txt), and the 51 MMAs per tile of both pipelines on
// one thread are about half of the kernel's time.  Warp 0 lane 0 then drives pipeline 0 AND the TMA loads (both are
// non-blocking polls), warp 1 lane 0 drives pipeline 1, and the operand slots are released by per-tile commits counted by
// the `empty` barriers (nmt arrivals) instead of by counters private to a single issuer.
template <int ISSUERS>
__global__ void __launch_bounds__(A2_THREADS, 1)
window_attn_tc2_fwd_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapV, AttnTcArgs a) {
    extern __shared__ uint8_t at_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = sm;
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + AT_SLOTS * AT_SLOT_BYTES + 4096);   // +4 KB: Q tile over-read pad
    uint64_t* empty = full + AT_SLOTS;
    uint64_t* rdy = empty + AT_SLOTS;          // slot rounded to TF32 (64 arrivals)
    uint64_t* s_full = rdy + AT_SLOTS;         // [2]
    uint64_t* p_full = s_full + 2;             // [2] 128 arrivals
    uint64_t* o_full = p_full + 2;             // [2]
    uint64_t* pc_full = o_full + 2;            // [2][4] two-issuer kernel: P published in 96-column chunks (block 0: 2, block 1: 1-2)
    uint32_t* tslot = reinterpret_cast<uint32_t*>(pc_full + 8);
    float* kmax2 = reinterpret_cast<float*>(tslot + 2);   // [AT_SLOTS][2]: max_j |k_j|^2 of the K tile in a slot (per conditioning warp)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nmt = (a.N + 127) / 128;
    const uint32_t box_bytes = (uint32_t)a.N * 128u;
    const int64_t nit = (a.items - blockIdx.x + gridDim.x - 1) / gridDim.x;     // items of this CTA
    const int64_t G = nit * nmt;                                                 // tiles of this CTA
    const int w1 = a.Nk - A2_B0;               // key columns of block 1 (multiple of 32)
    const int v1 = a.N - A2_B0;                // valid ones

    for (int i = threadIdx.x; i < (AT_SLOTS * AT_SLOT_BYTES + 4096) / 16; i += A2_THREADS)
        reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < AT_SLOTS; ++s) { abar_init(&full[s], 1); abar_init(&empty[s], ISSUERS == 1 ? 1 : nmt); abar_init(&rdy[s], 64); }
        for (int x = 0; x < 2; ++x) { abar_init(&s_full[x], 1); abar_init(&p_full[x], 128); abar_init(&o_full[x], 1); }
        for (int i = 0; i < 8; ++i) abar_init(&pc_full[i], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(asmem(tslot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    pdl_sync();            // prologue above (ring clear, barriers, TMEM) overlaps the previous kernel
    const int nchunk1 = w1 > 96 ? 2 : 1;       // 96-column chunks of key block 1
    // phase trace of CTA 0 (scripts/trace_attn2.py): SM clock stamps, pointer read once (not on the issue path)
    unsigned long long* const trp = blockIdx.x == 0 ? g_at_trace : nullptr;
#define ATR2(slot) do { if (trp != nullptr && (slot) < 512) trp[(slot)] = (unsigned long long)clock64(); } while (0)
#define ATR2I(n, k) do { if ((n) < 16) ATR2((n) * 16 + (k)); } while (0)      /* issuer stamps: first 16 tiles only */

    if (ISSUERS == 2 && warp <= 1) {
        if (lane == 0) {
            // ---- one issuing thread per pipeline; warp 0's thread also feeds the TMA ring
            const int x = warp;
            const uint32_t idesc_s0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(A2_B0 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_s1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(w1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                                     ((uint32_t)(128 >> 4) << 24);
            const int ks_pv1 = (v1 + 7) / 8;
            const uint32_t ring_a = asmem(ring);
            const uint32_t tcol = tmem + (uint32_t)(x * A2_PIPE_COLS);
            auto slot_addr = [&](int u) { return ring_a + (uint32_t)(u % AT_SLOTS) * AT_SLOT_BYTES; };
            auto issue_s = [&](int j, int mt, int h) {
                const uint32_t qa = slot_addr(3 * j), ka = slot_addr(3 * j + 1);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t ad = adesc(qa + (uint32_t)(mt * 128 * 128 + ks * 32), 16, 1024, 2);
                    const uint64_t bd = adesc(ka + (uint32_t)(h * A2_B0 * 128 + ks * 32), 16, 1024, 2);
                    amma_ss(tcol, ad, bd, h ? idesc_s1 : idesc_s0, ks ? 1u : 0u);
                }
                acommit(&s_full[x]);
            };
            // k-steps [k0, k1) of P V for key block h (8 keys per step)
            auto issue_pv = [&](int j, int h, int k0, int k1) {
                const uint32_t ocol = tcol + (uint32_t)(A2_B0 + 32 * h);
                uint64_t bd = adesc(slot_addr(3 * j + 2) + (uint32_t)((h * (A2_B0 / 8) + k0) * 1024), 4096, 512, 1);
                for (int kk = k0; kk < k1; ++kk, bd += 1024 >> 4)
                    amma_ts(ocol, tcol + (uint32_t)(kk * 8), bd, idesc_o, kk ? 1u : 0u);
            };
            uint64_t* pc = pc_full + 4 * x;
            if (x == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQK) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
            }
            const int n_loads = x == 0 ? (int)(3 * nit) : 0;
            int u_load = 0;                               // next slot use to load (item u_load / 3, part u_load % 3)
            int j = 0, mt = x;                            // nmt >= 2: tile g = x is (item 0, tile x)
            int64_t g = x;
            int st = g < G ? 0 : 3;
            uint32_t pw = 0;
            while (st != 3 || u_load < n_loads) {
                if (u_load < n_loads) {
                    const int s = u_load % AT_SLOTS;
                    if (abar_test(&empty[s], (uint32_t)((u_load / AT_SLOTS) & 1) ^ 1u)) {
                        const int part = u_load % 3;
                        const uint32_t it = blockIdx.x + (uint32_t)(u_load / 3) * gridDim.x;
                        const uint32_t head = it % (uint32_t)a.heads;
                        uint32_t w = it / (uint32_t)a.heads;
                        const uint32_t wx = w % (uint32_t)a.nww; w /= (uint32_t)a.nww;
                        const uint32_t wy = w % (uint32_t)a.nwh; w /= (uint32_t)a.nwh;
                        const uint32_t wz = w % (uint32_t)a.nwd; w /= (uint32_t)a.nwd;
                        abar_expect(&full[s], box_bytes);
                        tma_load_5d(ring + s * AT_SLOT_BYTES, part == 2 ? &mapV : &mapQK, &full[s], part * a.C + (int)head * 32,
                                    (int)(wx * a.ww), (int)(wy * a.wh), (int)(wz * a.wd), (int)w);
                        ++u_load;
                    }
                }
                if (st == 3) continue;
                const int u0 = 3 * j;
                if (st == 0) {
                    if (abar_test(&rdy[u0 % AT_SLOTS], (uint32_t)((u0 / AT_SLOTS) & 1)) &&
                        abar_test(&rdy[(u0 + 1) % AT_SLOTS], (uint32_t)(((u0 + 1) / AT_SLOTS) & 1))) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        ATR2I((int)pw, x * 8 + 0);
                        issue_s(j, mt, 0);
                        st = 1;
                    }
                } else if (st == 1) {
                    // (each chunk barrier completes once per tile of this pipeline: parity = tiles done & 1)
                    if (abar_test(&pc[0], pw & 1u) &&
                        abar_test(&rdy[(u0 + 2) % AT_SLOTS], (uint32_t)(((u0 + 2) / AT_SLOTS) & 1))) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        ATR2I((int)pw, x * 8 + 1);
                        issue_pv(j, 0, 0, 12);
                        st = 4;
                    }
                } else if (st == 4) {
                    if (abar_test(&pc[1], pw & 1u)) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        ATR2I((int)pw, x * 8 + 2);
                        issue_pv(j, 0, 12, A2_B0 / 8);
                        issue_s(j, mt, 1);
                        acommit(&empty[u0 % AT_SLOTS]);               // one of the nmt releases of Q and K
                        acommit(&empty[(u0 + 1) % AT_SLOTS]);
                        st = 2;
                    }
                } else if (st == 2) {
                    if (abar_test(&pc[2], pw & 1u)) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        ATR2I((int)pw, x * 8 + 3);
                        issue_pv(j, 1, 0, nchunk1 == 2 ? 12 : ks_pv1);
                        st = nchunk1 == 2 ? 5 : 6;
                    }
                } else if (st == 5) {
                    if (abar_test(&pc[3], pw & 1u)) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        ATR2I((int)pw, x * 8 + 4);
                        issue_pv(j, 1, 12, ks_pv1);
                        st = 6;
                    }
                }
                if (st == 6) {
                    ATR2I((int)pw, x * 8 + 5);
                    ++pw;
                    acommit(&o_full[x]);
                    acommit(&empty[(u0 + 2) % AT_SLOTS]);             // one of the nmt releases of V
                    g += 2;
                    mt += 2;
                    if (mt >= nmt) { mt -= nmt; ++j; }
                    st = g < G ? 0 : 3;
                }
            }
        }
    } else
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQK) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
            int64_t u = 0;
            for (int64_t it = blockIdx.x; it < a.items; it += gridDim.x) {
                const int head = (int)(it % a.heads);
                int64_t w = it / a.heads;
                const int wx = (int)(w % a.nww); w /= a.nww;
                const int wy = (int)(w % a.nwh); w /= a.nwh;
                const int wz = (int)(w % a.nwd); w /= a.nwd;
                const int b = (int)w;
#pragma unroll
                for (int part = 0; part < 3; ++part, ++u) {        // 0: Q, 1: K, 2: V
                    const int s = (int)(u % AT_SLOTS);
                    abar_wait(&empty[s], (uint32_t)((u / AT_SLOTS) & 1) ^ 1u);
                    abar_expect(&full[s], box_bytes);
                    tma_load_5d(ring + s * AT_SLOT_BYTES, part == 2 ? &mapV : &mapQK, &full[s], part * a.C + head * 32,
                                wx * a.ww, wy * a.wh, wz * a.wd, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(A2_B0 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_s1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(w1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                                     ((uint32_t)(128 >> 4) << 24);
            const int ks_pv1 = (v1 + 7) / 8;
            uint32_t pw[2] = {0u, 0u};                  // p_full phases consumed per pipeline
            const uint32_t ring_a = asmem(ring);
            auto slot_addr = [&](int u) { return ring_a + (uint32_t)(u % AT_SLOTS) * AT_SLOT_BYTES; };
            auto issue_s = [&](int x, int j, int mt, int h) {
                const uint32_t qa = slot_addr(3 * j), ka = slot_addr(3 * j + 1);
                const uint32_t d = tmem + (uint32_t)(x * A2_PIPE_COLS);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t ad = adesc(qa + (uint32_t)(mt * 128 * 128 + ks * 32), 16, 1024, 2);
                    const uint64_t bd = adesc(ka + (uint32_t)(h * A2_B0 * 128 + ks * 32), 16, 1024, 2);
                    amma_ss(d, ad, bd, h ? idesc_s1 : idesc_s0, ks ? 1u : 0u);
                }
                acommit(&s_full[x]);
            };
            auto issue_pv = [&](int x, int j, int h) {
                const uint32_t va = slot_addr(3 * j + 2);
                const uint32_t pcol = tmem + (uint32_t)(x * A2_PIPE_COLS);
                const uint32_t ocol = pcol + (uint32_t)(A2_B0 + 32 * h);
                const int n = h ? ks_pv1 : A2_B0 / 8;
                uint64_t bd = adesc(va + (uint32_t)(h * (A2_B0 / 8) * 1024), 4096, 512, 1);
                for (int kk = 0; kk < n; ++kk, bd += 1024 >> 4)      // next 8 key rows: start address + 1 KB
                    amma_ts(ocol, pcol + (uint32_t)(kk * 8), bd, idesc_o, kk ? 1u : 0u);
            };
            // Each pipeline x walks its tiles g = x, x+2, ... through three issue steps; the issuer polls both pipelines
            // and issues whichever step has its inputs ready (a pipeline waiting for the next item's V does not stall
            // the other one):
            //   step 0: S(block 0)                       needs Q, K of the item rounded (first tile of the item)
            //   step 1: PV(block 0) -> O0, S(block 1)    needs P(block 0) published [+ V rounded]
            //   step 2: PV(block 1) -> O1                needs P(block 1) published
            // Q/K (V) slots are released when all nmt tiles of the item have issued step 1 (step 2).
            int jx[2] = {0, 1 / nmt}, mx[2] = {0, 1 % nmt};   // (item, tile) of each pipeline's current tile: tile index g = x, x+2, ..
            int64_t gx[2] = {0, 1};
            int st[2] = {0, 0};
            int done1[2] = {0, 0}, done2[2] = {0, 0};   // per item parity (j & 1): tiles that issued step 1 / step 2
            int live = (G > 0) + (G > 1);
            if (G <= 1) st[1] = 3;
            if (G <= 0) st[0] = 3;
            while (live > 0) {
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    if (st[x] == 3) continue;
                    const int j = jx[x], mt = mx[x];
                    const int u0 = 3 * j;               // slot-use index of this item's Q (K = +1, V = +2)
                    if (st[x] == 0) {
                        // (re-testing a completed phase is cheap; the slot cannot be re-armed while this item is live)
                        if (!abar_test(&rdy[u0 % AT_SLOTS], (uint32_t)((u0 / AT_SLOTS) & 1))) continue;
                        if (!abar_test(&rdy[(u0 + 1) % AT_SLOTS], (uint32_t)(((u0 + 1) / AT_SLOTS) & 1))) continue;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        issue_s(x, j, mt, 0);
                        st[x] = 1;
                    } else if (st[x] == 1) {
                        if (!abar_test(&p_full[x], pw[x] & 1u)) continue;
                        if (!abar_test(&rdy[(u0 + 2) % AT_SLOTS], (uint32_t)(((u0 + 2) / AT_SLOTS) & 1))) continue;
                        ++pw[x];
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        issue_pv(x, j, 0);
                        issue_s(x, j, mt, 1);
                        if (++done1[j & 1] == nmt) {
                            done1[j & 1] = 0;
                            acommit(&empty[u0 % AT_SLOTS]); acommit(&empty[(u0 + 1) % AT_SLOTS]);
                        }
                        st[x] = 2;
                    } else {
                        if (!abar_test(&p_full[x], pw[x] & 1u)) continue;
                        ++pw[x];
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        issue_pv(x, j, 1);
                        acommit(&o_full[x]);
                        if (++done2[j & 1] == nmt) {
                            done2[j & 1] = 0;
                            acommit(&empty[(u0 + 2) % AT_SLOTS]);
                        }
                        gx[x] += 2;
                        mx[x] += 2;
                        if (mx[x] >= nmt) { mx[x] -= nmt; ++jx[x]; }     // nmt >= 2: at most one wrap
                        st[x] = 0;
                        if (gx[x] >= G) { st[x] = 3; --live; }
                    }
                }
            }
        }
    } else if (warp < 10) {
        // ------------------------------------------------------------------ softmax + epilogue, pipeline x
        const int x = warp >= 6 ? 1 : 0;
        const int q = warp & 3;                       // TMEM lane quarter of this warp
        const int r = q * 32 + lane;                  // query row inside the 128-row tile
        const uint32_t sbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(x * A2_PIPE_COLS);
        const int nch1 = w1 / 32;
        uint32_t va[32], vb[32];
        // token offset inside the window for this thread's row of each M-tile (nmt <= 3), computed once
        int64_t tokoff[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int i = t * 128 + r;
            const int ix = i % a.ww, iy = (i / a.ww) % a.wh, iz = i / (a.ww * a.wh);
            tokoff[t] = i < a.N ? ((int64_t)iz * a.Hp + iy) * a.Wp + ix : -1;
        }
        uint32_t un = 0;                              // tiles done by this pipeline
        int j = 0, mt = x;                            // nmt >= 2: tile g = x is (item 0, tile x)
        for (int64_t g = x; g < G; g += 2, ++un) {
            const uint32_t it = blockIdx.x + (uint32_t)j * gridDim.x;
            // ---- block 0: keys [0, 192), all valid
            const bool trs = lane == 0 && (warp == 2 || warp == 6);
            if (trs) ATR2(256 + (int)un * 16 + x * 8 + 0);
            abar_wait(&s_full[x], 0u);                // two s_full phases per tile: parities 0, 1
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trs) ATR2(256 + (int)un * 16 + x * 8 + 1);
            // Softmax is invariant to the shift m; it only has to keep exp2 in range.  s_ij <= |q_i| max_j |k_j|
            // (Cauchy-Schwarz) is known without reading S, so when that bound is within 2^64 of an actual score of the
            // row (taken from the first chunk), the max pass is skipped and S is read from TMEM ONCE.  Otherwise (norms
            // far above the scores) the warp takes the exact two-pass path for this tile.
            float bound;
            {
                const uint8_t* qrow = ring + (size_t)((3 * j) % AT_SLOTS) * AT_SLOT_BYTES + (size_t)(mt * 128 + r) * 128;
                float q2 = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    // rows are 128 B apart: start each lane at a different 16-byte column (sum order is irrelevant)
                    const float4 t = *reinterpret_cast<const float4*>(qrow + (((c + lane) & 7) << 4));
                    q2 += (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
                }
                const int ks = (3 * j + 1) % AT_SLOTS;
                bound = sqrtf(q2 * fmaxf(kmax2[2 * ks], kmax2[2 * ks + 1])) * 1.0001f;
            }
            const float mbB = bound * a.scale_log2;
            tld32_nowait(sbase, va);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float mest = chunk_max(va, -INFINITY, 32);
            const bool row_live = mt * 128 + r < a.N;
            const bool fast = __all_sync(0xffffffffu, !row_live || (mbB - mest * a.scale_log2 <= 64.f));
            float m0, l0, m1, l1;
            uint64_t* pc = pc_full + 4 * x;
            if (fast) {
                l0 = exp_block(sbase, A2_B0 / 32, A2_B0, a.scale_log2, mbB, va, vb, true, ISSUERS == 2 ? &pc[0] : nullptr);
                m0 = bound;
            } else {
                softmax_block(sbase, A2_B0 / 32, A2_B0, -INFINITY, a.scale_log2, m0, l0, va, vb);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (ISSUERS == 2) {
                if (!fast) abar_arrive(&pc[0]);       // two-pass rows publish the whole block at once
                abar_arrive(&pc[1]);
            } else {
                abar_arrive(&p_full[x]);
            }
            // ---- block 1: keys [192, N)
            if (trs) ATR2(256 + (int)un * 16 + x * 8 + 2);
            abar_wait(&s_full[x], 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trs) ATR2(256 + (int)un * 16 + x * 8 + 3);
            if (fast) {
                l1 = exp_block(sbase, nch1, v1, a.scale_log2, mbB, va, vb, false, ISSUERS == 2 ? &pc[2] : nullptr);
                m1 = bound;
            } else {
                softmax_block(sbase, nch1, v1, m0, a.scale_log2, m1, l1, va, vb);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (ISSUERS == 2) {
                if (nchunk1 == 2) {
                    if (!fast) abar_arrive(&pc[2]);
                    abar_arrive(&pc[3]);
                } else {
                    abar_arrive(&pc[2]);
                }
            } else {
                abar_arrive(&p_full[x]);
            }
            const float mb0 = m0 * a.scale_log2, mb1 = m1 * a.scale_log2;
            // ---- epilogue: out = (O0 * alpha + O1) / (l0 * alpha + l1), alpha = 2^((m0 - m1) * scale * log2 e)
            float alpha;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"(mb0 - mb1));
            const float l = fmaf(l0, alpha, l1);
            if (trs) ATR2(256 + (int)un * 16 + x * 8 + 4);
            abar_wait(&o_full[x], (uint32_t)(un & 1));
            if (trs) ATR2(256 + (int)un * 16 + x * 8 + 5);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tld32_nowait(sbase + A2_B0, va);
            tld32_nowait(sbase + A2_B0 + 32, vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int64_t toff = mt == 0 ? tokoff[0] : (mt == 1 ? tokoff[1] : tokoff[2]);
            if (toff >= 0) {
                const uint32_t head = it % (uint32_t)a.heads;
                uint32_t w = it / (uint32_t)a.heads;
                const uint32_t wx = w % (uint32_t)a.nww; w /= (uint32_t)a.nww;
                const uint32_t wy = w % (uint32_t)a.nwh; w /= (uint32_t)a.nwh;
                const uint32_t wz = w % (uint32_t)a.nwd; w /= (uint32_t)a.nwd;   // w = batch index
                const int64_t row = (((int64_t)w * a.Dp + wz * a.wd) * a.Hp + wy * a.wh) * a.Wp + wx * a.ww + toff;
                const float inv = 1.f / l;
                const float ai = alpha * inv;
                float* dst = a.out + row * a.ldo + head * 32;
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    *reinterpret_cast<float4*>(dst + 4 * t) =
                        make_float4(fmaf(__uint_as_float(va[4 * t]), ai, __uint_as_float(vb[4 * t]) * inv),
                                    fmaf(__uint_as_float(va[4 * t + 1]), ai, __uint_as_float(vb[4 * t + 1]) * inv),
                                    fmaf(__uint_as_float(va[4 * t + 2]), ai, __uint_as_float(vb[4 * t + 2]) * inv),
                                    fmaf(__uint_as_float(va[4 * t + 3]), ai, __uint_as_float(vb[4 * t + 3]) * inv));
                a.lse[row * a.heads + head] = fmaf(__log2f(l), 0.6931471805599453f, m1 * a.scale);
            }
            if (trs) ATR2(256 + (int)un * 16 + x * 8 + 6);
            mt += 2;
            if (mt >= nmt) { mt -= nmt; ++j; }
        }
    } else {
        // ------------------------------------------------------------------ operand conditioning (warps 10-11)
        const int et = (warp - 10) * 32 + lane;      // 0..63
        const int nvec = a.N * 8;                     // float4 per operand tile
        for (int64_t u = 0; u < 3 * nit; ++u) {
            const int s = (int)(u % AT_SLOTS);
            abar_wait(&full[s], (uint32_t)((u / AT_SLOTS) & 1));
            uint4* p4 = reinterpret_cast<uint4*>(ring + s * AT_SLOT_BYTES);
            const bool is_k = u % 3 == 1;
            float nmax = 0.f;                        // largest squared row norm seen by this thread's 8-lane groups
#pragma unroll 4
            for (int i0 = 0; i0 < nvec; i0 += 64) {
                const int i = i0 + et;               // float4 index: row = i >> 3 (8 consecutive lanes share a row)
                float sq = 0.f;
                if (i < nvec) {
                    uint4 t = p4[i];
                    t.x = (t.x + 0x1000u) & 0xFFFFE000u; t.y = (t.y + 0x1000u) & 0xFFFFE000u;
                    t.z = (t.z + 0x1000u) & 0xFFFFE000u; t.w = (t.w + 0x1000u) & 0xFFFFE000u;
                    p4[i] = t;
                    const float a0 = __uint_as_float(t.x), a1 = __uint_as_float(t.y), a2 = __uint_as_float(t.z), a3 = __uint_as_float(t.w);
                    sq = (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
                }
                if (is_k) {
                    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
                    nmax = fmaxf(nmax, sq);
                }
            }
            if (is_k) {
                nmax = warp_max(nmax);
                if (lane == 0) kmax2[2 * s + (warp - 10)] = nmax;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            abar_arrive(&rdy[s]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

typedef CUresult (*EncodeTiledFn5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tc_window_attn_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo, float* lse,
                       int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh, int ww, float scale,
                       cudaStream_t st) {
    const int N = wd * wh * ww;
    const int C = heads * hd;
    // taken only for the fused (P, 3C) layout: q | k | v column blocks of one buffer
    if (hd != 32 || N < 128 || N > 352 || wd > 256 || wh > 256 || ww > 256) return MIC_ERR_UNSUPPORTED;
    if (ldq != ldkv || k != q + C || v != q + 2 * C || (ldq & 3) || (reinterpret_cast<uintptr_t>(q) & 15) || (ldo & 3) ||
        (reinterpret_cast<uintptr_t>(out) & 15))
        return MIC_ERR_UNSUPPORTED;
    if (Dp % wd || Hp % wh || Wp % ww) return MIC_ERR_UNSUPPORTED;
    static EncodeTiledFn5 enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return MIC_ERR_UNSUPPORTED;
        enc = reinterpret_cast<EncodeTiledFn5>(p);
    }
    CUtensorMap mQK, mV;
    cuuint64_t dims[5] = {(cuuint64_t)(3 * C), (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)Dp, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)ldq * 4, (cuuint64_t)ldq * 4 * Wp, (cuuint64_t)ldq * 4 * Wp * Hp,
                             (cuuint64_t)ldq * 4 * Wp * Hp * Dp};
    cuuint32_t box[5] = {32, (cuuint32_t)ww, (cuuint32_t)wh, (cuuint32_t)wd, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(&mQK, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(q), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return MIC_ERR_UNSUPPORTED;
    if (enc(&mV, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(q), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return MIC_ERR_UNSUPPORTED;
    AttnTcArgs a{};
    a.out = out; a.ldo = ldo; a.lse = lse; a.B = B; a.Dp = Dp; a.Hp = Hp; a.Wp = Wp; a.heads = heads; a.C = C;
    a.wd = wd; a.wh = wh; a.ww = ww; a.nwd = Dp / wd; a.nwh = Hp / wh; a.nww = Wp / ww; a.N = N;
    a.Nk = ((N + 31) / 32) * 32;
    a.items = (int64_t)B * a.nwd * a.nwh * a.nww * heads;
    a.scale = scale;
    a.scale_log2 = scale * 1.4426950408889634f;
    const size_t smem = 1024 + (size_t)AT_SLOTS * AT_SLOT_BYTES + 4096 + 512;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(window_attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    int64_t grid = num_sms();
    if (grid > a.items) grid = a.items;
    static const bool force_v1 = getenv("MICFORMER_ATTN_V1") != nullptr;
    if (N > A2_B0 && !force_v1) {
        static const bool two_issuers = []() { const char* v = getenv("MICFORMER_ATTN_ISSUERS"); return !(v && v[0] == '1'); }();
        static bool attr2 = false;
        if (!attr2) {
            cudaFuncSetAttribute(window_attn_tc2_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(window_attn_tc2_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr2 = true;
        }
        if (a.items >= ((int64_t)1 << 30)) return MIC_ERR_UNSUPPORTED;
        if (two_issuers) mic::launch(window_attn_tc2_fwd_kernel<2>, dim3((unsigned)grid), dim3(A2_THREADS), smem, st, mQK, mV, a);
        else mic::launch(window_attn_tc2_fwd_kernel<1>, dim3((unsigned)grid), dim3(A2_THREADS), smem, st, mQK, mV, a);
        return check_launch("window_attn_tc2_fwd_kernel");
    }
    mic::launch(window_attn_tc_fwd_kernel, dim3((unsigned)grid), dim3(AT_THREADS), smem, st, mQK, mV, a);
    return check_launch("window_attn_tc_fwd_kernel");
}

}  // namespace mic
extern "C" int mic_debug_attn_trace(void* buf) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(mic::g_at_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
