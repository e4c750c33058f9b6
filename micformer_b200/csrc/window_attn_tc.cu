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
    const size_t smem = 1024 + (size_t)AT_SLOTS * AT_SLOT_BYTES + 4096 + 256;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(window_attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    int64_t grid = num_sms();
    if (grid > a.items) grid = a.items;
    window_attn_tc_fwd_kernel<<<(unsigned)grid, AT_THREADS, smem, st>>>(mQK, mV, a);
    return check_launch("window_attn_tc_fwd_kernel");
}

}  // namespace mic
extern "C" int mic_debug_attn_trace(void* buf) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(mic::g_at_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
