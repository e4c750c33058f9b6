// tcgen05 (5th-gen tensor core) TF32 GEMM for sm_100a: C[i,j] = sum_r A(i,r) * B(r,j), fp32 in / fp32 out,
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) into a 4-stage shared-memory ring, accumulated by
// tcgen05.mma.kind::tf32 into TMEM, drained by 4 epilogue warps with tcgen05.ld and the same fused epilogues as
// the CUDA-core path (bias, erf-GELU + saved pre-activation, GELU', per-sample DropPath scale, residual,
// accumulate, split-R atomic accumulate).
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one
// elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
//
// Operand layouts.  Each operand is either reduction-contiguous ("K-major": activations X[m,k], weights W[n,k])
// or output-index-contiguous ("MN-major": needed by backward-data for W and by backward-weight for both
// operands).  Both are fed straight from the row-major tensors -- no transposed copies:
//   K-major : TMA box {32 r, ROWS i}  -> smem [ROWS][128 B], canonical SW128 K-major, SBO = 1024 B
//   MN-major: TMA box {32 i, 32 r} per 32-wide index group (swizzle 128B with 32B atoms) -> smem
//             [group][32 r][128 B], canonical SW128_BASE32B MN-major (the only MN-major layout tcgen05 takes for
//             32-bit operands), LBO = 4096 B (group stride), SBO = 512 B (4 r-rows)
// One MMA consumes 8 reduction elements (32 B of fp32 read as tf32): K-major advances the descriptor start
// address by 32 B inside the swizzle atom, MN-major by 1024 B.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"

namespace mic {

constexpr int TM = 128;          // output rows per CTA (UMMA M)
constexpr int TKB = 32;          // reduction elements per stage (128 B of fp32)
constexpr int STAGES = 3;           // ring depth with the 3xTF32 lo tiles (12 slots of 16 KB)
constexpr int MAXST = 6;            // single-pass kernels on a grid of at most one CTA per SM: the same 12 slots = 6 stages
constexpr int TC_THREADS = 320;        // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (+ converters), warps 6-9 converters
constexpr int NCONV = 256;             // converter threads (warps 2..9)

struct TcEpi {
    float* C; int64_t ldc;
    int I, J;
    const float* bias;
    int act; float* pre; int64_t ldpre;
    const float* mulgrad; int64_t ldmg;
    const float* res; int64_t ldres;
    const float* rowscale_i; int rps_i;
    const float* rowscale_r; int rps_r;     // uniform per split chunk (checked on the host)
    int accumulate;                         // 0 store, 1 +=, 2 atomicAdd
    int kb_total, kb_per_split;
    int round_rn;                           // 1: operands rounded to nearest TF32 in shared memory before the MMA
                                            // (the tensor core itself truncates; truncation is biased toward zero)
    int split3;                             // 1: 3xTF32 -- x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi);
                                            // D = Ahi*Bhi + Alo*Bhi + Ahi*Blo (fp32-faithful, ~2^-21).  Implies round_rn
                                            // 2: split-bf16 (K-major operands, one-shot kernel): the landed fp32 tile
                                            // [rows][32 k] is rewritten IN PLACE as [rows][32 k hi | 32 k lo] bf16 and the
                                            // same three products run as kind::f16 MMAs (K = 16): 6 instead of 12 MMAs per
                                            // 32-wide k-block (these kernels are MMA-issue bound), ~2^-17, no lo slots
    int nst;                                // pipeline stages in use (3, or 2 when split3 doubles the tiles)
    float* db;                              // weight-gradient GEMMs (A = dY^T, MN-major): db[i] += sum_r A(i, r), folded into the
                                            // operand pass of the CTAs with blockIdx.y == 0 (one-shot kernel only)
    // "unpatch view" (decoder tail): one operand is the (T cells) x (64 * ch) matrix of ConvTranspose3d(k4, s4) rows, but it
    // LIVES as the (B, 4 dc, 4 hc, 4 wc, ch) channels-last grid it is the 4^3-block permutation of.  Its tensor map is 5-D
    // {cc = 4 ch, wc, 4 (y sub-position), hc, B * dc * 4} and (column, row) of the logical matrix become the coordinates
    // below -- the block_permute pass (0.22 ms each way at 128^3) disappears.  vw_op: 0 off, 1 C (store), 2 A K-major
    // (rows = cells), 3 B MN-major (32 columns x 32 cells per box)
    int vw_op, vw_cc, vw_wc, vw_hc;
};

// (column, row) of the logical rows matrix -> 5-D coordinates of the fine grid (see TcEpi::vw_op)
__device__ __forceinline__ void view_coords(const TcEpi& e, int col, int row, int (&c)[5]) {
    const int g = col / e.vw_cc;
    c[0] = col - g * e.vw_cc;               // (x sub-position, channel)
    c[2] = g & 3;                           // y sub-position
    const int r1 = row / e.vw_wc;
    c[1] = row - r1 * e.vw_wc;              // cell x
    const int bd = r1 / e.vw_hc;
    c[3] = r1 - bd * e.vw_hc;               // cell y
    c[4] = bd * 4 + (g >> 2);               // (batch, cell z, z sub-position)
}

// optional per-CTA phase trace (debug): 16 x u64 globaltimer stamps per CTA when a buffer is registered
__device__ unsigned long long* g_tc_trace = nullptr;
__device__ __forceinline__ void trace(int slot) {
    unsigned long long* t = g_tc_trace;
    if (t) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        if (cta < 4096) t[cta * 16 + slot] = now;
    }
}

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, const int (&c)[5]) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, const int (&c)[5]) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
// operand / output tile moves that honour the unpatch view
__device__ __forceinline__ void load_a_k(void* dst, const CUtensorMap* map, uint64_t* bar, const TcEpi& e, int r0, int i0) {
    if (e.vw_op == 2) { int c[5]; view_coords(e, r0, i0, c); tma_load_5d(dst, map, bar, c); }
    else tma_load_2d(dst, map, bar, r0, i0);
}
__device__ __forceinline__ void load_b_mn(void* dst, const CUtensorMap* map, uint64_t* bar, const TcEpi& e, int j, int r0) {
    if (e.vw_op == 3) { int c[5]; view_coords(e, j, r0, c); tma_load_5d(dst, map, bar, c); }
    else tma_load_2d(dst, map, bar, j, r0);
}
__device__ __forceinline__ void store_c(const CUtensorMap* map, const void* src, const TcEpi& e, int gj0, int i0) {
    if (e.vw_op == 1) { int c[5]; view_coords(e, gj0, i0, c); tma_store_5d(map, src, c); }
    else tma_store_2d(map, src, gj0, i0);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
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

// shared-memory matrix descriptor (sm_100 format: version 1 at bit 46, layout type at bits 61..63)
// layout: 2 = SWIZZLE_128B (16-byte atoms; K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte atoms: the only
// layout tcgen05 accepts for MN-major 32-bit (tf32) operands; TMA side = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;
    return d;
}

// Bias gradient folded into a weight-gradient GEMM (the 256 converter threads, A = dY^T MN-major): thread et holds, per group
// g of 32 output features, the running sum of float4 (row r = et / 8, chunk c = et % 8) of every A tile of this CTA.  The
// tile is 128B-swizzled with 32-byte atoms (Swizzle<2,5,2>: address bits 5..6 ^= bits 7..8), so chunk c of row r holds
// features 8 * ((c >> 1) ^ (r & 3)) + 4 * (c & 1) ... + 3 of the group: the four lanes {r & 3 = 0..3} that hold the same
// features are lane ^ 10 and lane ^ 20 apart (butterfly), the eight warps meet in shared memory (no shared-memory float
// atomics: those are CAS loops), then one global atomicAdd per feature, scaled by the split's DropPath factor like dW.
__device__ __forceinline__ void colsum_flush(float4 (&acc)[4], float* sred, int et, int i0, const TcEpi& e, int kb0) {
    const int lane = et & 31, w = et >> 5;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int m = 10; m <= 20; m += 10) {
            acc[g].x += __shfl_xor_sync(0xffffffffu, acc[g].x, m);
            acc[g].y += __shfl_xor_sync(0xffffffffu, acc[g].y, m);
            acc[g].z += __shfl_xor_sync(0xffffffffu, acc[g].z, m);
            acc[g].w += __shfl_xor_sync(0xffffffffu, acc[g].w, m);
        }
        if (lane < 8) *reinterpret_cast<float4*>(sred + w * TM + g * 32 + 4 * lane) = acc[g];     // r & 3 == 0: features 4c..4c+3
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (et < TM && i0 + et < e.I) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < NCONV / 32; ++k) v += sred[k * TM + et];
        const float rs = e.rowscale_r ? e.rowscale_r[(int64_t)kb0 * TKB / e.rps_r] : 1.f;
        v *= rs;
        if (v != 0.f) atomicAdd(e.db + i0 + et, v);
    }
}

// BNT: B-tile columns (UMMA N, multiple of 16, <= 128); TCOLS: TMEM columns (power of two >= BNT)
// EPI (compile-time epilogue kind, keeps the drain loop small and branch-free):
//   0 plain store (+bias, *rowscale)   1 bias + save pre-activation + erf-GELU   2 residual: res + rowscale*(acc+bias)
//   3 multiply by GELU'(aux)           4 accumulate (C += acc)                   5 atomic accumulate (split-R)
template <bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapP, TcEpi e, int BNT,
               int TCOLS) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = TM * TKB * 4;                     // 16 KB
    const int b_groups = (BNT + 31) / 32;
    const int B_BYTES = (B_MN ? b_groups * 32 : BNT) * TKB * 4;
    const int B_STRIDE = 128 * TKB * 4;                       // fixed slot size (16 KB) keeps every slot 1024-aligned
    const int NST = e.nst;
    uint8_t* sA = smem;
    const int NSLOT = e.split3 == 1 ? 4 * STAGES : 2 * NST;   // 16 KB slots in the ring
    uint8_t* sB = smem + (e.split3 == 1 ? STAGES : NST) * A_BYTES;
    // split3 keeps the "lo" tiles in a second set of slots:  [A0 A1 A2 | B0 B1 B2 | Alo0..2 | Blo0..2]  (16 KB each)
    uint8_t* sAlo = smem + 2 * STAGES * A_BYTES;
    uint8_t* sBlo = sAlo + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSLOT * A_BYTES);
    uint64_t* empty = full + MAXST;
    uint64_t* tmem_full = empty + MAXST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    uint64_t* conv = tmem_full + 2;          // [MAXST] operands rounded (arrived by the 8 converter warps)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * TM, j0 = blockIdx.y * BNT;
    const int kb0 = blockIdx.z * e.kb_per_split;
    const int kb1 = min(e.kb_total, kb0 + e.kb_per_split);
    const int nkb = kb1 - kb0;
    // bias gradient of a weight-gradient GEMM: the converter warps add up the rows of the dY^T tiles as they pass
    const bool do_sum = A_MN && B_MN && EPI == 5 && e.db != nullptr && blockIdx.y == 0;
    float* sdb = reinterpret_cast<float*>(smem + NSLOT * A_BYTES + 256 + 512);     // [8 warps][128] partial sums (4 KB)

    if (threadIdx.x == 0) trace(0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], NCONV / 32); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();            // everything above touched only shared memory / TMEM: it overlaps the previous kernel
    if (threadIdx.x == 0) trace(1);

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            const uint32_t tx = A_BYTES + B_BYTES;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NST;
                const uint32_t ph = (kb / NST) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], tx);
                const int r0 = (kb0 + kb) * TKB;
                uint8_t* a = sA + s * A_BYTES;
                uint8_t* b = sB + s * B_STRIDE;
                if (A_MN) {
#pragma unroll
                    for (int g = 0; g < TM / 32; ++g) tma_load_2d(a + g * 4096, &mapA, &full[s], i0 + g * 32, r0);
                } else {
                    load_a_k(a, &mapA, &full[s], e, r0, i0);
                }
                if (B_MN) {
                    for (int g = 0; g < b_groups; ++g) load_b_mn(b + g * 4096, &mapB, &full[s], e, j0 + g * 32, r0);
                } else {
                    tma_load_2d(b, &mapB, &full[s], r0, j0);
                }
                if (kb == 0) trace(2);
            }
            trace(3);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=tf32, majors, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BNT >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NST;
                const uint32_t ph = (kb / NST) & 1;
                mbar_wait((e.round_rn || do_sum) ? &conv[s] : &full[s], ph);
                tc_fence_after();
                if (kb == 0) trace(4);
                const uint32_t a_addr = smem_u32(sA + s * A_BYTES);
                const uint32_t b_addr = smem_u32(sB + s * B_STRIDE);
                if (!A_MN && !B_MN && e.split3 == 2) {
                    // rows are [32 k hi | 32 k lo] bf16: hi k-step ks at +32 ks bytes, lo at +64 + 32 ks
                    const uint32_t id16 = t5::idesc_bf16(TM, BNT, false, false);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        t5::mma3(tmem_base, make_desc(a_addr + ks * 32, 16, 1024, 2), make_desc(a_addr + 64 + ks * 32, 16, 1024, 2),
                                 make_desc(b_addr + ks * 32, 16, 1024, 2), make_desc(b_addr + 64 + ks * 32, 16, 1024, 2), id16,
                                 (kb | ks) ? 1u : 0u);
                    umma_commit(&empty[s]);
                    continue;
                }
#pragma unroll
                for (int k = 0; k < TKB / 8; ++k) {
                    // MN-major: [group][32 r][128 B]; atom = 4 r-rows (512 B): LBO = group stride, SBO = 512 B, 8 r per MMA
                    const uint64_t ad = A_MN ? make_desc(a_addr + k * 1024, 4096, 512, 1) : make_desc(a_addr + k * 32, 16, 1024, 2);
                    const uint64_t bd = B_MN ? make_desc(b_addr + k * 1024, 4096, 512, 1) : make_desc(b_addr + k * 32, 16, 1024, 2);
                    umma_tf32(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
                    if (e.split3 == 1) {
                        const uint32_t al = smem_u32(sAlo + s * A_BYTES), bl = smem_u32(sBlo + s * B_STRIDE);
                        const uint64_t adl = A_MN ? make_desc(al + k * 1024, 4096, 512, 1) : make_desc(al + k * 32, 16, 1024, 2);
                        const uint64_t bdl = B_MN ? make_desc(bl + k * 1024, 4096, 512, 1) : make_desc(bl + k * 32, 16, 1024, 2);
                        umma_tf32(tmem_base, adl, bd, idesc, 1u);
                        umma_tf32(tmem_base, ad, bdl, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);          // frees the smem slot once these MMAs retire
            }
            umma_commit(tmem_full);              // accumulator complete
            trace(5);
        }
    } else {
        // ---------------- epilogue: warp q = warp % 4 owns TMEM lanes [32q, 32q+32) = output rows i0 + 32q + lane
        const int q = warp & 3;
        const int gi = i0 + q * 32 + lane;
        if (!e.round_rn && do_sum) {
            // single-pass TF32 (operands used as they landed): the converter warps only add up dY^T.  A tile = 4 groups of
            // [32 r][32 i] floats; float4 (et + 256 g) is row r = et / 8, 16-byte chunk c = et % 8 of group g
            const int et = (warp - 2) * 32 + lane;
            float4 acc[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NST;
                const uint32_t ph = (kb / NST) & 1;
                mbar_wait(&full[s], ph);
                const float4* a4 = reinterpret_cast<const float4*>(sA + s * A_BYTES);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 t = a4[et + g * NCONV];
                    acc[g].x += t.x; acc[g].y += t.y; acc[g].z += t.z; acc[g].w += t.w;
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv[s])) : "memory");
            }
            colsum_flush(acc, sdb, et, i0, e, kb0);
        }
        if (e.round_rn) {
            // converter role during the main loop: round the freshly landed A/B tiles to nearest-even TF32 in
            // place (element-wise, so the swizzle is irrelevant), then hand the stage to the MMA warp
            const int et = (warp - 2) * 32 + lane;                       // 0..NCONV-1
            const int b_vec = B_BYTES / 16;
            float4 acc[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NST;
                const uint32_t ph = (kb / NST) & 1;
                mbar_wait(&full[s], ph);
                float4* a4 = reinterpret_cast<float4*>(sA + s * A_BYTES);
                float4* b4 = reinterpret_cast<float4*>(sB + s * B_STRIDE);
                if (!A_MN && !B_MN && e.split3 == 2) {
                    // in place, row by row: the 4 threads of a row (adjacent lanes) each read two fp32 chunks (8 k), all
                    // reads of the row complete (__syncwarp), then each writes one hi chunk and one lo chunk
                    auto rewrite = [&](uint8_t* tile, int item) {
                        const int r = item >> 2, j = item & 3, x = r & 7;
                        uint8_t* row = tile + r * 128;
                        const float4 v0 = *reinterpret_cast<const float4*>(row + (((2 * j) ^ x) << 4));
                        const float4 v1 = *reinterpret_cast<const float4*>(row + (((2 * j + 1) ^ x) << 4));
                        uint4 h, l;
                        t5::split2(v0.x, v0.y, h.x, l.x); t5::split2(v0.z, v0.w, h.y, l.y);
                        t5::split2(v1.x, v1.y, h.z, l.z); t5::split2(v1.z, v1.w, h.w, l.w);
                        __syncwarp();
                        *reinterpret_cast<uint4*>(row + ((j ^ x) << 4)) = h;
                        *reinterpret_cast<uint4*>(row + (((4 + j) ^ x) << 4)) = l;
                    };
                    rewrite(sA + s * A_BYTES, et);
                    rewrite(sA + s * A_BYTES, et + NCONV);
                    for (int item = et; item < BNT * 4; item += NCONV) rewrite(sB + s * B_STRIDE, item);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv[s])) : "memory");
                    continue;
                }
                if (e.split3) {
                    float4* al4 = reinterpret_cast<float4*>(sAlo + s * A_BYTES);
                    float4* bl4 = reinterpret_cast<float4*>(sBlo + s * B_STRIDE);
                    auto split = [](float x, float& hi, float& lo) {
                        hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
                        lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
                    };
#pragma unroll 4
                    for (int i = et; i < A_BYTES / 16; i += NCONV) {
                        const float4 t = a4[i];
                        float4 h, l;
                        split(t.x, h.x, l.x); split(t.y, h.y, l.y); split(t.z, h.z, l.z); split(t.w, h.w, l.w);
                        a4[i] = h; al4[i] = l;
                    }
                    for (int i = et; i < b_vec; i += NCONV) {
                        const float4 t = b4[i];
                        float4 h, l;
                        split(t.x, h.x, l.x); split(t.y, h.y, l.y); split(t.z, h.z, l.z); split(t.w, h.w, l.w);
                        b4[i] = h; bl4[i] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv[s])) : "memory");
                    continue;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int i = et + g * NCONV;
                    float4 t = a4[i];
                    if (do_sum) { acc[g].x += t.x; acc[g].y += t.y; acc[g].z += t.z; acc[g].w += t.w; }
                    uint32_t r0, r1, r2, r3;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r0) : "f"(t.x));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r1) : "f"(t.y));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r2) : "f"(t.z));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r3) : "f"(t.w));
                    a4[i] = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
                }
                for (int i = et; i < b_vec; i += NCONV) {
                    float4 t = b4[i];
                    uint32_t r0, r1, r2, r3;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r0) : "f"(t.x));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r1) : "f"(t.y));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r2) : "f"(t.z));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r3) : "f"(t.w));
                    b4[i] = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv[s])) : "memory");
            }
            if (do_sum) colsum_flush(acc, sdb, et, i0, e, kb0);
        }
        if (warp >= 6) goto teardown;           // converter-only warps
        // Drain: TMEM -> registers -> fused epilogue math -> 128B-swizzled staging tile in shared memory (the operand
        // ring is free once tmem_full fired) -> one TMA store (or TMA reduce-add for accumulate / split-R) per
        // 32-column chunk.  TMA clips rows >= I and columns >= J, so ragged edges need no store guards.
        // Global operands of the epilogue are fetched BEFORE waiting for the accumulator: the bias tile goes to shared
        // memory once, the residual / pre-activation row is software-pipelined one chunk ahead.
        const bool row_ok = gi < e.I;
        float rs = e.rowscale_r ? e.rowscale_r[(int64_t)kb0 * TKB / e.rps_r] : 1.f;
        if (row_ok && e.rowscale_i) rs *= e.rowscale_i[gi / e.rps_i];
        const bool use_bias = e.bias != nullptr && blockIdx.z == 0;
        float* sbias = reinterpret_cast<float*>(smem + NSLOT * A_BYTES + 256);
        if (use_bias) {
            const int tt = (warp - 2) * 32 + lane;               // 0..127 >= BNT columns of this tile
            if (tt < BNT) sbias[tt] = j0 + tt < e.J ? e.bias[j0 + tt] : 0.f;
        }
        const int64_t gi64 = gi;
        const float* aux = nullptr;          // EPI 2: residual row, EPI 3: pre-activation row
        if (EPI == 2) aux = e.res + gi64 * e.ldres;
        if (EPI == 3) aux = e.mulgrad + gi64 * e.ldmg;
        const int64_t ldaux = EPI == 2 ? e.ldres : e.ldmg;
        const bool aux_al = aux && ((reinterpret_cast<uintptr_t>(aux) | (uintptr_t)(ldaux * 4)) & 15) == 0 && (j0 & 3) == 0;
        auto load_aux = [&](int c0, float4 (&av)[8]) {
            const int gj0 = j0 + c0;
            const int ncol = min(32, min(BNT - c0, e.J - gj0));
            if ((EPI == 2 || EPI == 3) && row_ok) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (aux_al && 4 * t + 3 < ncol) av[t] = __ldg(reinterpret_cast<const float4*>(aux + gj0 + 4 * t));
                    else {
                        av[t].x = 4 * t + 0 < ncol ? aux[gj0 + 4 * t + 0] : 0.f;
                        av[t].y = 4 * t + 1 < ncol ? aux[gj0 + 4 * t + 1] : 0.f;
                        av[t].z = 4 * t + 2 < ncol ? aux[gj0 + 4 * t + 2] : 0.f;
                        av[t].w = 4 * t + 3 < ncol ? aux[gj0 + 4 * t + 3] : 0.f;
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) av[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 av[8];
        load_aux(0, av);
        if ((EPI == 2 || EPI == 3) && row_ok) {
            // the rest of this thread's row segment: pull the lines towards L1 while the accumulator finishes
            for (int c = 32; c < BNT && j0 + c < e.J; c += 32)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(aux + j0 + c));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");           // bias tile visible to the four epilogue warps
        if (nkb > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        if (warp == 2 && lane == 0) trace(6);
        const int row = q * 32 + lane;                    // row inside the 128-row tile
        const bool want_pre = EPI == 1 && e.pre != nullptr;
        int nchunk = 0;
        for (int c0 = 0; c0 < BNT; c0 += 32, ++nchunk) {
            float v[32];
            if (nkb > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] = 0.f;
            }
            if (c0 == 0 && warp == 2 && lane == 0) trace(10);
            const int gj0 = j0 + c0;
            const int ncol = min(32, min(BNT - c0, e.J - gj0));      // valid columns in this chunk (may be <= 0)
            // staging buffers: chunk c -> 16 KB slot c of the operand ring (6 slots: sA[0..2], sB[0..2]);
            // EPI 1 additionally stages the pre-activation in slot 3 + c (BNT <= 96 there)
            uint8_t* cbuf = smem + nchunk * 16384;
            uint8_t* pbuf = smem + (3 + nchunk) * 16384;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float x[4] = {v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]};
                if (use_bias) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + 4 * t);
                    x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w;
                }
                const uint32_t soff = (uint32_t)row * 128u + ((uint32_t)(t ^ (row & 7)) << 4);
                if (EPI == 1) {
                    if (want_pre) *reinterpret_cast<float4*>(pbuf + soff) = make_float4(x[0], x[1], x[2], x[3]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[u] = gelu_erf(x[u]);
                }
                if (EPI == 3) {
                    x[0] *= gelu_erf_grad(av[t].x); x[1] *= gelu_erf_grad(av[t].y);
                    x[2] *= gelu_erf_grad(av[t].z); x[3] *= gelu_erf_grad(av[t].w);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] *= rs;
                if (EPI == 2) { x[0] += av[t].x; x[1] += av[t].y; x[2] += av[t].z; x[3] += av[t].w; }
                *reinterpret_cast<float4*>(cbuf + soff) = make_float4(x[0], x[1], x[2], x[3]);
            }
            if (c0 + 32 < BNT) load_aux(c0 + 32, av);               // next chunk's row segment: in flight across the store
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && lane == 0 && ncol > 0) {
                if (EPI == 4 || EPI == 5) tma_reduce_add_2d(&mapC, cbuf, gj0, i0);
                else store_c(&mapC, cbuf, e, gj0, i0);
                if (want_pre) tma_store_2d(&mapP, pbuf, gj0, i0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
teardown:
    if (warp == 2 && lane == 0) trace(7);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 32) trace(8);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
    }
    if (threadIdx.x == 32) trace(9);
}

// ------------------------------------------------------------------------------------------------ persistent variant
// Large problems (several tiles per SM): one CTA per SM walks its tiles; the drain of tile n (TMEM -> epilogue ->
// staging -> TMA store) overlaps the loads / operand conditioning / MMAs of tile n+1 through two TMEM accumulators and
// a dedicated double-buffered staging area, so the per-tile latency chain of the one-shot kernel (prologue, first TMA
// round trip, epilogue) is paid once per CTA instead of once per tile.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-5: epilogue   warps 6-13: operand converters
constexpr int TCP_THREADS = 448;
constexpr int TCP_CONV = 256;            // converter threads (warps 6..13)

struct TcpSched {
    int tiles_i, tiles_j, splits, ntiles;
    int nst;                             // ring stages
    int stage_slots;                     // 16 KB slots per stage (2, or 4 with the 3xTF32 lo tiles)
};

template <bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tcp_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapP, TcEpi e, TcpSched sc,
                int BNT, int TCOLS) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int SLOT = TM * TKB * 4;                        // 16 KB
    const int b_groups = (BNT + 31) / 32;
    const int B_BYTES = (B_MN ? b_groups * 32 : BNT) * TKB * 4;
    const int NST = sc.nst;
    const int STG = sc.stage_slots * SLOT;                    // bytes per stage: [A | B | (Alo | Blo)]
    uint8_t* ring = smem;
    uint8_t* stage_c = smem + NST * STG;                      // 2 output staging slots
    uint8_t* stage_p = stage_c + 2 * SLOT;                    // 2 pre-activation staging slots (EPI 1)
    uint8_t* tail = stage_p + ((EPI == 1 && e.pre) ? 2 * SLOT : 0);
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty = full + 8;
    uint64_t* conv = empty + 8;
    uint64_t* tmem_full = conv + 8;           // [2]
    uint64_t* tmem_empty = tmem_full + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* sbias = reinterpret_cast<float*>(tail + 256);      // [2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_ij = sc.tiles_i * sc.tiles_j;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], TCP_CONV / 32); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();            // everything above touched only shared memory / TMEM: it overlaps the previous kernel

    // tile t -> (j fastest: CTAs that run concurrently share the A tile in L2)
    auto decode = [&](int t, int& i0, int& j0, int& kb0, int& nkb) {
        const int z = t / tiles_ij;
        const int r = t - z * tiles_ij;
        const int ti = r / sc.tiles_j, tj = r - ti * sc.tiles_j;
        i0 = ti * TM; j0 = tj * BNT;
        kb0 = z * e.kb_per_split;
        nkb = min(e.kb_total, kb0 + e.kb_per_split) - kb0;
    };

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            const uint32_t tx = SLOT + B_BYTES;
            uint32_t cnt = 0;
            for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x) {
                int i0, j0, kb0, nkb;
                decode(t, i0, j0, kb0, nkb);
                for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                    const int s = cnt % NST;
                    mbar_wait(&empty[s], ((cnt / NST) & 1) ^ 1);
                    mbar_expect_tx(&full[s], tx);
                    const int r0 = (kb0 + kb) * TKB;
                    uint8_t* a = ring + s * STG;
                    uint8_t* b = a + SLOT;
                    if (A_MN) {
#pragma unroll
                        for (int g = 0; g < TM / 32; ++g) tma_load_2d(a + g * 4096, &mapA, &full[s], i0 + g * 32, r0);
                    } else {
                        load_a_k(a, &mapA, &full[s], e, r0, i0);
                    }
                    if (B_MN) {
                        for (int g = 0; g < b_groups; ++g) load_b_mn(b + g * 4096, &mapB, &full[s], e, j0 + g * 32, r0);
                    } else {
                        tma_load_2d(b, &mapB, &full[s], r0, j0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BNT >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            uint32_t cnt = 0, n = 0;
            for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, ++n) {
                int i0, j0, kb0, nkb;
                decode(t, i0, j0, kb0, nkb);
                const uint32_t buf = n & 1;
                mbar_wait(&tmem_empty[buf], ((n >> 1) & 1) ^ 1);        // epilogue drained the previous use
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * (uint32_t)TCOLS;
                for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                    const int s = cnt % NST;
                    mbar_wait(e.round_rn ? &conv[s] : &full[s], (cnt / NST) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + s * STG);
                    const uint32_t b_addr = a_addr + SLOT;
#pragma unroll
                    for (int k = 0; k < TKB / 8; ++k) {
                        const uint64_t ad = A_MN ? make_desc(a_addr + k * 1024, 4096, 512, 1) : make_desc(a_addr + k * 32, 16, 1024, 2);
                        const uint64_t bd = B_MN ? make_desc(b_addr + k * 1024, 4096, 512, 1) : make_desc(b_addr + k * 32, 16, 1024, 2);
                        umma_tf32(dcol, ad, bd, idesc, (kb | k) ? 1u : 0u);
                        if (e.split3) {
                            const uint32_t al = a_addr + 2 * SLOT, bl = a_addr + 3 * SLOT;
                            const uint64_t adl = A_MN ? make_desc(al + k * 1024, 4096, 512, 1) : make_desc(al + k * 32, 16, 1024, 2);
                            const uint64_t bdl = B_MN ? make_desc(bl + k * 1024, 4096, 512, 1) : make_desc(bl + k * 32, 16, 1024, 2);
                            umma_tf32(dcol, adl, bd, idesc, 1u);
                            umma_tf32(dcol, ad, bdl, idesc, 1u);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else if (warp >= 6) {
        // ---------------- operand converters (only when round_rn): nearest TF32 / hi-lo split of the landed tiles
        if (e.round_rn) {
            const int et = (warp - 6) * 32 + lane;
            const int b_vec = B_BYTES / 16;
            uint32_t cnt = 0;
            for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x) {
                int i0, j0, kb0, nkb;
                decode(t, i0, j0, kb0, nkb);
                for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                    const int s = cnt % NST;
                    mbar_wait(&full[s], (cnt / NST) & 1);
                    float4* a4 = reinterpret_cast<float4*>(ring + s * STG);
                    float4* b4 = a4 + SLOT / 16;
                    auto rn = [](float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); };
                    if (e.split3) {
                        float4* al4 = a4 + 2 * (SLOT / 16);
                        float4* bl4 = a4 + 3 * (SLOT / 16);
#pragma unroll 4
                        for (int i = et; i < SLOT / 16; i += TCP_CONV) {
                            const float4 v = a4[i];
                            const float4 h = make_float4(rn(v.x), rn(v.y), rn(v.z), rn(v.w));
                            a4[i] = h;
                            al4[i] = make_float4(rn(v.x - h.x), rn(v.y - h.y), rn(v.z - h.z), rn(v.w - h.w));
                        }
                        for (int i = et; i < b_vec; i += TCP_CONV) {
                            const float4 v = b4[i];
                            const float4 h = make_float4(rn(v.x), rn(v.y), rn(v.z), rn(v.w));
                            b4[i] = h;
                            bl4[i] = make_float4(rn(v.x - h.x), rn(v.y - h.y), rn(v.z - h.z), rn(v.w - h.w));
                        }
                    } else {
#pragma unroll 4
                        for (int i = et; i < SLOT / 16; i += TCP_CONV) {
                            const float4 v = a4[i];
                            a4[i] = make_float4(rn(v.x), rn(v.y), rn(v.z), rn(v.w));
                        }
                        for (int i = et; i < b_vec; i += TCP_CONV) {
                            const float4 v = b4[i];
                            b4[i] = make_float4(rn(v.x), rn(v.y), rn(v.z), rn(v.w));
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv[s])) : "memory");
                }
            }
        }
    } else {
        // ---------------- epilogue warps 2..5: warp q = warp % 4 owns TMEM lanes [32q, 32q+32)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool elected = warp == 2 && lane == 0;
        const bool want_pre = EPI == 1 && e.pre != nullptr;
        uint32_t n = 0, chunk = 0;
        for (int t = blockIdx.x; t < sc.ntiles; t += gridDim.x, ++n) {
            int i0, j0, kb0, nkb;
            decode(t, i0, j0, kb0, nkb);
            const uint32_t buf = n & 1;
            const int gi = i0 + row;
            const bool row_ok = gi < e.I;
            float rs = e.rowscale_r ? e.rowscale_r[(int64_t)kb0 * TKB / e.rps_r] : 1.f;
            if (row_ok && e.rowscale_i) rs *= e.rowscale_i[gi / e.rps_i];
            const bool use_bias = e.bias != nullptr && kb0 == 0;
            float* sb = sbias + buf * 128;
            if (use_bias) {
                const int tt = (warp - 2) * 32 + lane;
                if (tt < BNT) sb[tt] = j0 + tt < e.J ? e.bias[j0 + tt] : 0.f;
            }
            const float* aux = nullptr;
            if (EPI == 2) aux = e.res + (int64_t)gi * e.ldres;
            if (EPI == 3) aux = e.mulgrad + (int64_t)gi * e.ldmg;
            const int64_t ldaux = EPI == 2 ? e.ldres : e.ldmg;
            const bool aux_al = aux && ((reinterpret_cast<uintptr_t>(aux) | (uintptr_t)(ldaux * 4)) & 15) == 0 && (j0 & 3) == 0;
            auto load_aux = [&](int c0, float4 (&av)[8]) {
                const int gj0 = j0 + c0;
                const int ncol = min(32, min(BNT - c0, e.J - gj0));
                if ((EPI == 2 || EPI == 3) && row_ok) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (aux_al && 4 * u + 3 < ncol) av[u] = __ldg(reinterpret_cast<const float4*>(aux + gj0 + 4 * u));
                        else {
                            av[u].x = 4 * u + 0 < ncol ? aux[gj0 + 4 * u + 0] : 0.f;
                            av[u].y = 4 * u + 1 < ncol ? aux[gj0 + 4 * u + 1] : 0.f;
                            av[u].z = 4 * u + 2 < ncol ? aux[gj0 + 4 * u + 2] : 0.f;
                            av[u].w = 4 * u + 3 < ncol ? aux[gj0 + 4 * u + 3] : 0.f;
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            float4 av[8];
            load_aux(0, av);                                   // in flight while the accumulator finishes
            mbar_wait(&tmem_full[buf], (n >> 1) & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < BNT; c0 += 32, ++chunk) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)TCOLS + (uint32_t)c0, v);
                if (c0 + 32 >= BNT) {
                    // accumulator fully read: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                }
                const int gj0 = j0 + c0;
                const int ncol = min(32, min(BNT - c0, e.J - gj0));
                uint8_t* cbuf = stage_c + (chunk & 1) * SLOT;
                uint8_t* pbuf = stage_p + (chunk & 1) * SLOT;
                // the staging slot was last used two chunks ago: its TMA store must have finished reading it
                if (elected) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");     // (also publishes the bias tile on the first chunk)
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float x[4] = {v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]};
                    if (use_bias) {
                        const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + 4 * u);
                        x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w;
                    }
                    const uint32_t soff = (uint32_t)row * 128u + ((uint32_t)(u ^ (row & 7)) << 4);
                    if (EPI == 1) {
                        if (want_pre) *reinterpret_cast<float4*>(pbuf + soff) = make_float4(x[0], x[1], x[2], x[3]);
#pragma unroll
                        for (int w = 0; w < 4; ++w) x[w] = gelu_erf(x[w]);
                    }
                    if (EPI == 3) {
                        x[0] *= gelu_erf_grad(av[u].x); x[1] *= gelu_erf_grad(av[u].y);
                        x[2] *= gelu_erf_grad(av[u].z); x[3] *= gelu_erf_grad(av[u].w);
                    }
#pragma unroll
                    for (int w = 0; w < 4; ++w) x[w] *= rs;
                    if (EPI == 2) { x[0] += av[u].x; x[1] += av[u].y; x[2] += av[u].z; x[3] += av[u].w; }
                    *reinterpret_cast<float4*>(cbuf + soff) = make_float4(x[0], x[1], x[2], x[3]);
                }
                if (c0 + 32 < BNT) load_aux(c0 + 32, av);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (elected) {
                    if (ncol > 0) {
                        if (EPI == 4 || EPI == 5) tma_reduce_add_2d(&mapC, cbuf, gj0, i0);
                        else store_c(&mapC, cbuf, e, gj0, i0);
                        if (want_pre) tma_store_2d(&mapP, pbuf, gj0, i0);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TCOLS));
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor map: inner (contiguous) extent d0, outer extent d1 with row stride ld (floats)
static bool make_map(CUtensorMap* m, const float* base, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t box0, uint32_t box1,
                     bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 5-D fp32 tensor map of the unpatch view (TcEpi::vw_op): dims {cc, wc, 4, hc, Z}; box = 32 columns x `rows_w` cells x
// `rows_h` cell rows
static bool make_view_map(CUtensorMap* m, const float* base, int cc, int wc, int hc, int64_t Z, uint32_t rows_w, uint32_t rows_h,
                          bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[5] = {(cuuint64_t)cc, (cuuint64_t)wc, 4, (cuuint64_t)hc, (cuuint64_t)Z};
    cuuint64_t strides[4] = {(cuuint64_t)cc * 4, (cuuint64_t)cc * 4 * wc, (cuuint64_t)cc * 4 * wc * 4, (cuuint64_t)cc * 4 * wc * 4 * hc};
    cuuint32_t box[5] = {32, rows_w, 1, rows_h, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// the view requested for the NEXT mic_linear_* call of this host thread (mic_linear_unpatch_view); consumed by that call
struct UnpatchView { int ch, dc, hc, wc; };
static thread_local UnpatchView g_view = {0, 0, 0, 0};
bool tc_view_pending() { return g_view.ch != 0; }
void tc_view_set(int ch, int dc, int hc, int wc) { g_view = {ch, dc, hc, wc}; }
static UnpatchView take_view() { UnpatchView v = g_view; g_view = {0, 0, 0, 0}; return v; }
// rows = cells, cols = 64 * ch of the logical matrix; the kernels move 128-row / 32-row boxes, so a box must be 4 / 1 whole
// cell rows
static bool view_ok(const UnpatchView& v, int rows, int cols) {
    return v.ch > 0 && v.ch % 8 == 0 && v.wc == 32 && v.hc % 4 == 0 && v.dc > 0 && cols == 64 * v.ch &&
           rows % (v.dc * v.hc * v.wc) == 0 && rows % 128 == 0;
}

struct TcOperand {
    const float* p;
    bool mn_major;     // true: output index contiguous, rows = reduction index
    int64_t ld;        // row stride in floats
};

static bool operand_ok(const TcOperand& o) { return (reinterpret_cast<uintptr_t>(o.p) & 15) == 0 && o.ld % 4 == 0 && o.ld > 0; }

// generic launcher: C[I,J] = sum_R A(i,r) B(r,j)
static int tc_gemm(const TcOperand& A, const TcOperand& B, TcEpi e, int R, int kb_per_split, cudaStream_t st,
                   bool* db_done = nullptr) {
    if (!operand_ok(A) || !operand_ok(B)) return MIC_ERR_UNSUPPORTED;
    const int I = e.I, J = e.J;
    int BNT = J <= 128 ? ((J + 15) / 16) * 16 : 128;
    if (J > 128) {   // balance column tiles (e.g. 192 -> 2 x 96); multiples of 32 so that no 32-column store box
                     // straddles two tiles
        const int nt = (J + 127) / 128;
        BNT = (((J + nt - 1) / nt) + 31) / 32 * 32;
    }
    int epi = 0;
    if (e.accumulate == 2) epi = 5;
    else if (e.accumulate == 1) epi = 4;
    else if (e.act == 1) epi = 1;
    else if (e.res) epi = 2;
    else if (e.mulgrad) epi = 3;
    if (epi == 1 && BNT > 96) {       // the pre-activation needs its own staging slots: at most 3 chunks per tile
        const int nt = (J + 95) / 96;
        BNT = nt > 1 ? (((J + nt - 1) / nt) + 31) / 32 * 32 : BNT;
    }
    int TCOLS = 32;
    while (TCOLS < BNT) TCOLS <<= 1;
    if ((reinterpret_cast<uintptr_t>(e.C) & 15) || e.ldc % 4) return MIC_ERR_UNSUPPORTED;
    if (epi == 1 && e.pre && ((reinterpret_cast<uintptr_t>(e.pre) & 15) || e.ldpre % 4)) return MIC_ERR_UNSUPPORTED;
    CUtensorMap mA, mB, mC, mP;
    bool ok = A.mn_major ? make_map(&mA, A.p, (uint64_t)I, (uint64_t)R, (uint64_t)A.ld, 32, TKB, true)
                         : make_map(&mA, A.p, (uint64_t)R, (uint64_t)I, (uint64_t)A.ld, TKB, TM, false);
    ok = ok && (B.mn_major ? make_map(&mB, B.p, (uint64_t)J, (uint64_t)R, (uint64_t)B.ld, 32, TKB, true)
                           : make_map(&mB, B.p, (uint64_t)R, (uint64_t)J, (uint64_t)B.ld, TKB, (uint32_t)BNT, false));
    ok = ok && make_map(&mC, e.C, (uint64_t)J, (uint64_t)I, (uint64_t)e.ldc, 32, TM, false);
    if (ok && epi == 1 && e.pre) ok = make_map(&mP, e.pre, (uint64_t)J, (uint64_t)I, (uint64_t)e.ldpre, 32, TM, false);
    else mP = mC;
    if (ok && e.vw_op) {
        // the viewed operand: (cells) x (64 ch) logical matrix stored as the fine channels-last grid
        const int64_t cells = e.vw_op == 1 ? I : (e.vw_op == 2 ? I : R);
        const int64_t Z = cells / ((int64_t)e.vw_hc * e.vw_wc) * 4;
        if (e.vw_op == 1) ok = epi == 0 && make_view_map(&mC, e.C, e.vw_cc, e.vw_wc, e.vw_hc, Z, 32, 4, false);
        else if (e.vw_op == 2) ok = !A.mn_major && make_view_map(&mA, A.p, e.vw_cc, e.vw_wc, e.vw_hc, Z, 32, 4, false);
        else ok = B.mn_major && make_view_map(&mB, B.p, e.vw_cc, e.vw_wc, e.vw_hc, Z, 32, 1, true);
    }
    if (!ok) return MIC_ERR_UNSUPPORTED;
    e.nst = STAGES;
    if (e.split3) e.round_rn = 1;
    e.kb_total = (R + TKB - 1) / TKB;
    e.kb_per_split = (kb_per_split > 0 && kb_per_split < e.kb_total) ? kb_per_split : e.kb_total;
    const int splits = (e.kb_total + e.kb_per_split - 1) / e.kb_per_split;
    dim3 grid((I + TM - 1) / TM, (J + BNT - 1) / BNT, splits);
    // small grids (the deep, latency-bound stages) never share an SM: single-pass kernels then run a 6-stage ring in the shared
    // memory the 3xTF32 variant needs anyway, which halves the exposed TMA round trips of a long reduction
    static const bool deep_ring = []() { const char* v = getenv("MICFORMER_GEMM_DEEP_RING"); return !(v && v[0] == '0'); }();
    if (e.split3 == 2 && (A.mn_major || B.mn_major)) e.split3 = 1;
    if (deep_ring && e.split3 != 1 && e.kb_per_split > STAGES && (int64_t)grid.x * grid.y * grid.z <= (int64_t)num_sms()) e.nst = MAXST;
    const size_t smem = 1024 + (size_t)(e.split3 == 1 ? 12 : 2 * e.nst) * (TM * TKB * 4) + 256 + 512 +
                        (e.db ? 4096 : 0);                                     // ring + barriers + bias tile (+ db partial sums)
    // several tiles per SM: persistent kernel (epilogue of tile n overlaps the main loop of tile n+1)
    static const bool no_persist = getenv("MICFORMER_GEMM_ONESHOT") != nullptr;
    const int64_t ntiles = (int64_t)grid.x * grid.y * grid.z;
    if (db_done) *db_done = false;
    if (!no_persist && ntiles >= 2 * (int64_t)num_sms() && ntiles < (1 << 30)) {
        TcpSched sc{};
        e.db = nullptr;
        if (e.split3 == 2) e.split3 = 1;
        sc.tiles_i = (int)grid.x; sc.tiles_j = (int)grid.y; sc.splits = splits; sc.ntiles = (int)ntiles;
        sc.stage_slots = e.split3 ? 4 : 2;
        const int stage_bytes = sc.stage_slots * TM * TKB * 4;
        const int staging = ((epi == 1 && e.pre) ? 4 : 2) * TM * TKB * 4;
        int nst = (int)((227 * 1024 - 1024 - 2048 - staging) / stage_bytes);
        if (nst > 6) nst = 6;
        if (nst > e.kb_per_split * 2) nst = e.kb_per_split * 2 > 2 ? e.kb_per_split * 2 : 2;
        sc.nst = nst;
        const size_t psmem = 1024 + (size_t)nst * stage_bytes + staging + 256 + 1024 + 256;
        const int pgrid = num_sms();
#define PLAUNCH(AM, BM_, EP)                                                                                         \
    do {                                                                                                             \
        static size_t attr_sz = 0;                                                                                   \
        if (attr_sz < psmem) {                                                                                       \
            cudaFuncSetAttribute(gemm_tcp_kernel<AM, BM_, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
            attr_sz = 227 * 1024;                                                                                    \
        }                                                                                                            \
        mic::launch((gemm_tcp_kernel<AM, BM_, EP>), dim3(pgrid), dim3(TCP_THREADS), psmem, st, mA, mB, mC, mP, e, sc, BNT, TCOLS);            \
    } while (0)
#define PLAUNCH_EPI(AM, BM_)                                                           \
    switch (epi) {                                                                     \
        case 0: PLAUNCH(AM, BM_, 0); break;                                            \
        case 1: PLAUNCH(AM, BM_, 1); break;                                            \
        case 2: PLAUNCH(AM, BM_, 2); break;                                            \
        case 3: PLAUNCH(AM, BM_, 3); break;                                            \
        case 4: PLAUNCH(AM, BM_, 4); break;                                            \
        default: PLAUNCH(AM, BM_, 5); break;                                           \
    }
        if (!A.mn_major && !B.mn_major) { PLAUNCH_EPI(false, false) }
        else if (!A.mn_major && B.mn_major) { PLAUNCH_EPI(false, true) }
        else if (A.mn_major && B.mn_major) { PLAUNCH(true, true, 5); }
        else return MIC_ERR_UNSUPPORTED;
#undef PLAUNCH_EPI
#undef PLAUNCH
        return check_launch("gemm_tcp_kernel");
    }
#define LAUNCH(AM, BM_, EP)                                                                                      \
    do {                                                                                                         \
        static bool attr_done = false;                                                                           \
        if (!attr_done) {                                                                                        \
            cudaFuncSetAttribute(gemm_tc_kernel<AM, BM_, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                 (int)(1024 + 12 * TM * TKB * 4 + 256 + 512 + 4096));                                  \
            attr_done = true;                                                                                    \
        }                                                                                                        \
        mic::launch((gemm_tc_kernel<AM, BM_, EP>), grid, dim3(TC_THREADS), smem, st, mA, mB, mC, mP, e, BNT, TCOLS);                       \
    } while (0)
#define LAUNCH_EPI(AM, BM_)                                                            \
    switch (epi) {                                                                     \
        case 0: LAUNCH(AM, BM_, 0); break;                                             \
        case 1: LAUNCH(AM, BM_, 1); break;                                             \
        case 2: LAUNCH(AM, BM_, 2); break;                                             \
        case 3: LAUNCH(AM, BM_, 3); break;                                             \
        case 4: LAUNCH(AM, BM_, 4); break;                                             \
        default: LAUNCH(AM, BM_, 5); break;                                            \
    }
    if (!(A.mn_major && B.mn_major && epi == 5)) e.db = nullptr;
    if (db_done) *db_done = e.db != nullptr;
    if (!A.mn_major && !B.mn_major) { LAUNCH_EPI(false, false) }
    else if (!A.mn_major && B.mn_major) { LAUNCH_EPI(false, true) }
    else if (A.mn_major && B.mn_major) { LAUNCH(true, true, 5); }
    else return MIC_ERR_UNSUPPORTED;
#undef LAUNCH_EPI
#undef LAUNCH
    return check_launch("gemm_tc_kernel");
}

int colsum(const float* X, int64_t ldx, int M, int N, const float* rowscale, int rps, float* out, cudaStream_t st);

}  // namespace mic
extern "C" int mic_debug_tc_trace(void* buf) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(mic::g_tc_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
namespace mic {

int tc_linear_fwd(const float* X, int ldx, const float* W, int ldw, int w_is_kn, const float* bias, float* Y, int ldy,
                  int M, int N, int K, int act, float* pre, int ldpre, const float* res, int ldres, const float* rowscale,
                  int rps, int accumulate, int mode, cudaStream_t st) {
    (void)mode;
    const UnpatchView vw = take_view();
    if (N < 16 || K < 8) return MIC_ERR_UNSUPPORTED;
    if (vw.ch && (!view_ok(vw, M, N) || act || pre || res || accumulate || rowscale)) return MIC_ERR_UNSUPPORTED;
    TcEpi e{};
    if (vw.ch) { e.vw_op = 1; e.vw_cc = 4 * vw.ch; e.vw_wc = vw.wc; e.vw_hc = vw.hc; }
    e.C = Y; e.ldc = ldy; e.I = M; e.J = N; e.bias = bias; e.act = act; e.pre = pre; e.ldpre = ldpre; e.res = res;
    e.ldres = ldres; e.rowscale_i = rowscale; e.rps_i = rps > 0 ? rps : 1; e.accumulate = accumulate ? 1 : 0;
    // forward GEMMs carry the logits parity bar (1e-3 vs the fp32 CPU path): 3xTF32 split keeps them fp32-faithful
    // (the MMAs are not the bottleneck of these HBM/latency-bound shapes); MICFORMER_TF32_FWD=1 selects single-pass
    // nearest-rounded TF32 (measured 8e-4..1e-3 logits error at 64^3/128^3)
    static const int fwd_single = []() { const char* v = getenv("MICFORMER_TF32_FWD"); return v && v[0] == '1'; }();
    // MICFORMER_TF32_FWD_SMALL_M=<rows>: GEMMs with at most that many rows (the deep, latency-bound stages) run
    // single-pass nearest-rounded TF32; the large-M stages and the decoder tail keep the 3xTF32 split
    static const int small_m = []() { const char* v = getenv("MICFORMER_TF32_FWD_SMALL_M"); return v ? atoi(v) : 0; }();
    // MICFORMER_FWD_BF16X3=0: 3xTF32 for every forward GEMM (the round-1 path); default: split-bf16 where both operands are K-major
    static const int bf3 = []() { const char* v = getenv("MICFORMER_FWD_BF16X3"); return !(v && v[0] == '0'); }();
    e.round_rn = 1;
    e.split3 = (fwd_single || M <= small_m) ? 0 : (bf3 && !w_is_kn ? 2 : 1);
    TcOperand A{X, false, ldx};
    TcOperand B{W, w_is_kn != 0, ldw};      // W[n,k]: K-major; W[k,n]: MN-major
    // Long reductions on a grid that leaves most SMs idle (fc2 of the deep stages: 16 CTAs x 24..48 k-blocks, each k-block paced
    // by the CTA's shared-memory bandwidth): split the reduction over blockIdx.z.  The output is initialised with the residual
    // (or zero) by a copy node and every split adds rowscale * (partial [+ bias in split 0]) with TMA reduce-add.
    static const bool splitk = []() { const char* v = getenv("MICFORMER_FWD_SPLITK"); return !(v && v[0] == '0'); }();
    if (splitk && !act && !accumulate && !pre && M > 0 && !vw.ch) {
        const int kb_total = (K + TKB - 1) / TKB;
        int bnt = N <= 128 ? ((N + 15) / 16) * 16 : 128;
        if (N > 128) { const int nt = (N + 127) / 128; bnt = (((N + nt - 1) / nt) + 31) / 32 * 32; }
        const int tiles = ((M + TM - 1) / TM) * ((N + bnt - 1) / bnt);
        int smax = kb_total / 4;
        if (smax > num_sms() / tiles) smax = num_sms() / tiles;
        if (smax > 8) smax = 8;
        if (kb_total >= 12 && smax >= 2) {
            cudaError_t ce = cudaSuccess;
            if (res && res != Y) ce = cudaMemcpy2DAsync(Y, (size_t)ldy * 4, res, (size_t)ldres * 4, (size_t)N * 4, (size_t)M, cudaMemcpyDeviceToDevice, st);
            else if (!res) ce = cudaMemset2DAsync(Y, (size_t)ldy * 4, 0, (size_t)N * 4, (size_t)M, st);
            if (ce != cudaSuccess) return MIC_ERR_CUDA;
            e.res = nullptr; e.accumulate = 2;
            return tc_gemm(A, B, e, K, (kb_total + smax - 1) / smax, st);
        }
    }
    return tc_gemm(A, B, e, K, 0, st);
}

int tc_linear_bwd_data(const float* dY, int lddy, const float* W, int ldw, int w_is_kn, float* dX, int lddx, int M, int N,
                       int K, const float* gelu_pre, int ldpre, const float* rowscale, int rps, int accumulate, int mode,
                       cudaStream_t st) {
    (void)mode;
    const UnpatchView vw = take_view();
    if (K < 16 || N < 8) return MIC_ERR_UNSUPPORTED;
    if (vw.ch && !view_ok(vw, M, N)) return MIC_ERR_UNSUPPORTED;
    TcEpi e{};
    if (vw.ch) { e.vw_op = 2; e.vw_cc = 4 * vw.ch; e.vw_wc = vw.wc; e.vw_hc = vw.hc; }     // dY is the viewed operand
    e.C = dX; e.ldc = lddx; e.I = M; e.J = K; e.mulgrad = gelu_pre; e.ldmg = ldpre; e.rowscale_i = rowscale;
    e.rps_i = rps > 0 ? rps : 1; e.accumulate = accumulate ? 1 : 0;
    TcOperand A{dY, false, lddy};
    // B(r = n, j = k): W[n,k] has j contiguous -> MN-major; W[k,n] has r contiguous -> K-major
    TcOperand B{W, w_is_kn == 0, ldw};
    return tc_gemm(A, B, e, N, 0, st);
}

int tc_linear_bwd_weight(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, int w_is_kn, float* db,
                         int M, int N, int K, const float* rowscale, int rps, int mode, cudaStream_t st) {
    (void)mode;
    const UnpatchView vw = take_view();
    if (N < 16 || K < 16) return MIC_ERR_UNSUPPORTED;
    // dY viewed: only as the MN-major B operand (w_is_kn).  db then receives the column sums of the buffer AS LAID OUT IN
    // MEMORY (rows of 64 ch floats): the sums over the 64 block positions of each channel are exact, which is all a bias
    // tiled over the block positions (br.repeat(64)) needs
    if (vw.ch && (!view_ok(vw, M, N) || !w_is_kn || rowscale)) return MIC_ERR_UNSUPPORTED;
    TcEpi e{};
    if (vw.ch) { e.vw_op = 3; e.vw_cc = 4 * vw.ch; e.vw_wc = vw.wc; e.vw_hc = vw.hc; }
    e.C = dW; e.ldc = lddw; e.accumulate = 2;
    TcOperand A{}, B{};
    if (!w_is_kn) { A = {dY, true, lddy}; B = {X, true, ldx}; e.I = N; e.J = K; }
    else { A = {X, true, ldx}; B = {dY, true, lddy}; e.I = K; e.J = N; }
    const int kb_total = (M + TKB - 1) / TKB;
    const int BNT = e.J <= 128 ? e.J : 128;
    const int tiles = ((e.I + TM - 1) / TM) * ((e.J + BNT - 1) / BNT);
    int splits = (num_sms() * 2 + tiles - 1) / tiles;
    if (splits > kb_total) splits = kb_total;
    if (splits < 1) splits = 1;
    int kb_per = (kb_total + splits - 1) / splits;
    if (rowscale) {
        // the DropPath scale must be uniform inside a split chunk: chunk rows must divide rows-per-sample
        if (rps <= 0 || rps % TKB) return MIC_ERR_UNSUPPORTED;
        const int kb_rps = rps / TKB;
        if (kb_per > kb_rps) kb_per = kb_rps;
        while (kb_rps % kb_per) --kb_per;
        e.rowscale_r = rowscale; e.rps_r = rps;
    }
    // the bias gradient rides along when dY is the A operand (its tiles pass the converter warps of the blockIdx.y == 0 CTAs)
    static const bool fold = []() { const char* v = getenv("MICFORMER_FOLD_COLSUM"); return !(v && v[0] == '0'); }();
    if (db && !w_is_kn && fold) e.db = db;
    bool db_done = false;
    int rc = tc_gemm(A, B, e, M, kb_per, st, &db_done);
    if (rc) return rc;
    if (db && !db_done) return colsum(dY, lddy, M, N, rowscale, rps > 0 ? rps : 1, db, st);
    return MIC_OK;
}

}  // namespace mic
