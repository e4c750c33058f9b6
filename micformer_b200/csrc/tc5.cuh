// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the fused block kernels (sm_100a).
//
// Operand convention of the fused kernels ("split bf16"): every fp32 GEMM operand x is carried as two bf16 tiles
// hi = bf16_rn(x), lo = bf16_rn(x - hi) and every product is issued as three kind::f16 MMAs hi*hi + lo*hi + hi*lo with fp32
// accumulation in TMEM: relative error ~2^-17 per product (16 mantissa bits kept), i.e. ~30x tighter than one TF32 pass,
// at 3 MMAs per 16 reduction elements (3xTF32 needs 6) and the same shared-memory footprint as one fp32 tile.
//
// Tile layout: activation tiles are [row = token][feature] with the feature axis split into 64-element panels; a panel is
// 128 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7) -- the canonical SWIZZLE_128B layout.  The same bytes are
//   * a K-major operand   (MN = token,   K = feature): SBO = 1024 B, one MMA (K = 16) advances the start address by 32 B;
//   * an MN-major operand (MN = feature, K = token)  : LBO = panel stride, SBO = 1024 B, one MMA (16 tokens) advances 2048 B,
// so a tile written once serves X*W^T-type products and the token-reduction products of the weight gradients.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace mic {
namespace t5 {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// optional phase trace (debug / profiling scripts): when a buffer is registered (mic_debug_t5_trace), the calling thread
// stamps %globaltimer into slot `slot` of record `rec` (32 u64 slots per record) of CTA blockIdx.x (first 4 CTAs only)
static __device__ unsigned long long* g_t5_trace = nullptr;      // (one copy per translation unit; no -rdc)
__device__ __forceinline__ void t5_trace(int rec, int slot) {
    unsigned long long* t = g_t5_trace;
    if (t && blockIdx.x < 4 && blockIdx.y == 0 && rec < 8) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        t[(blockIdx.x * 8 + rec) * 32 + slot] = now;
    }
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "T5_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra T5_DONE;\n"
        "bra T5_WAIT;\n"
        "T5_DONE:\n"
        "}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------- bulk copy (1-D TMA)
// global -> shared, completion counted in bytes on an mbarrier; size and both addresses multiples of 16 B
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive columns (thread = lane = tile row); no wait inside: issue several, then ld_wait()
__device__ __forceinline__ void ld32(uint32_t taddr, float* v) {
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
}
__device__ __forceinline__ void ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

// ---------------------------------------------------------------------------------------------- descriptors
// shared-memory matrix descriptor (sm_100: version 1 at bit 46, swizzle mode at bits 61..63; 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major view of a panel: rows = MN, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) { return desc_sw128(saddr, 16, 1024); }
// MN-major view: K = tile row (token), MN = feature; 64-feature panels `panel_stride` bytes apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t panel_stride) { return desc_sw128(saddr, panel_stride, 1024); }

// instruction descriptor, kind::f16: D = f32 (1 << 4), A = B = bf16 (1 << 7, 1 << 10), majors at bits 15 / 16
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// the split product: D (+)= Ahi*Bhi + Alo*Bhi + Ahi*Blo
__device__ __forceinline__ void mma3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc) {
    mma_bf16(d, ah, bh, idesc, acc);
    mma_bf16(d, al, bh, idesc, 1u);
    mma_bf16(d, ah, bl, idesc, 1u);
}

// ---------------------------------------------------------------------------------------------- operand tiles
// byte offset of the 16-byte chunk holding features [8c, 8c+8) of row r inside a 64-feature panel
__device__ __forceinline__ uint32_t chunk_off(int r, int c) { return (uint32_t)r * 128u + ((uint32_t)(c ^ (r & 7)) << 4); }

// x -> (hi, lo) bf16 pair, round-to-nearest-even both
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// two floats -> packed (hi, hi) and (lo, lo) bf16x2 words: cvt.rn.bf16x2.f32 twice, the hi values are recovered as floats
// by a mask / shift of the packed word (6 instructions per pair)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));          // upper half <- b, lower half <- a
    const float ra = a - __uint_as_float(hi << 16);
    const float rb = b - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
// 8 consecutive features -> one 16-byte chunk each of the hi and the lo tile
__device__ __forceinline__ void store_chunk(uint8_t* hi_panel, uint8_t* lo_panel, int r, int c, const float* x8) {
    uint4 h, l;
    split2(x8[0], x8[1], h.x, l.x);
    split2(x8[2], x8[3], h.y, l.y);
    split2(x8[4], x8[5], h.z, l.z);
    split2(x8[6], x8[7], h.w, l.w);
    const uint32_t off = chunk_off(r, c);
    *reinterpret_cast<uint4*>(hi_panel + off) = h;
    *reinterpret_cast<uint4*>(lo_panel + off) = l;
}

// pull a row of `bytes` bytes towards L2 (the next tile's rows, one tile ahead of their use)
__device__ __forceinline__ void prefetch_l2(const void* p, int bytes) {
    const char* c = reinterpret_cast<const char*>(p);
    for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + bytes - 4));
}
// a "ones" feature column: chunk c of row r holds {1, 0, 0, 0, 0, 0, 0, 0} (hi) / zeros (lo).  With the column inside the
// N range of a token-reduction product D = A^T B (B = this tile), column 8c of D is the column sum of A: the bias gradient
// of a linear layer comes out of the weight-gradient MMA for free.
__device__ __forceinline__ void store_ones_chunk(uint8_t* hi_panel, uint8_t* lo_panel, int r, int c) {
    const uint32_t off = chunk_off(r, c);
    *reinterpret_cast<uint4*>(hi_panel + off) = make_uint4(0x00003F80u, 0u, 0u, 0u);      // bf16(1.0) = 0x3F80 in the low half
    *reinterpret_cast<uint4*>(lo_panel + off) = make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------------------------------------- GELU (erf form)
// nn.GELU(approximate='none') and its derivative from ONE exponential: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7,
// below the fp32 accumulation noise of the surrounding GEMMs), evaluated without cancellation on the negative side:
//   cdf(x) = 0.5 * erfc(-x / sqrt 2);  for z = |x| / sqrt 2:  0.5 * erfc(z) = 0.5 * poly(t) * exp(-z^2),  t = 1 / (1 + p z)
//   gelu(x) = x * cdf(x),   gelu'(x) = cdf(x) + x * exp(-x^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;                                                          // MUFU.RCP (the IEEE __frcp_rn is an 11-instruction sequence)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    float e;                                                          // exp(-x^2 / 2) = 2^(-x^2 * log2(e) / 2): one FMUL + MUFU.EX2
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));
    const float tail = 0.5f * (p * t) * e;          // 0.5 * erfc(z)
    const float cdf = x >= 0.f ? 1.0f - tail : tail;
    g = x * cdf;
    dg = fmaf(x * e, 0.39894228040143267794f, cdf);
}
__device__ __forceinline__ float gelu_fast(float x) {
    float g, dg;
    gelu_both(x, g, dg);
    return g;
}

// column sums over the 32 lanes of a warp: lane l ends up with sum over lanes of v[l] (31 shuffles for 32 columns)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

}  // namespace t5
}  // namespace mic
