// tcgen05 implicit-GEMM 3x3x3 convolution (stride 1, zero pad 1) on channels-last token grids -- forward and
// backward-data (the same kernel: backward-data is the forward conv of dy with the taps mirrored and the roles of
// the two channel axes swapped, so Wt[tap][ci][co] is read as the [tap][n][k] operand with tap -> 26 - tap).
//
//   y[p, n] = bias[n] + sum_{tap} sum_{c} x[p + tap, c] * Wsrc[tap'][n][c]     (M = positions, N tile = 16 | 32, K = 27 * Cin)
//
// No im2col and no per-tap re-fetch: a CTA owns an (8 x 16) x/y footprint and marches over a z segment.  Each
// input z-plane of the footprint (+1 halo: 10 x 18 positions) is staged ONCE per 8-channel chunk into an 8-slot
// shared-memory ring in the tcgen05 "no swizzle" K-major core-matrix layout
//     plane[kj][pos][4 floats]      (kj = 4-channel group, pos = y'*10 + x' in the haloed plane)
// so that all 27 taps are just different START ADDRESSES of the same staged data: tap (tz,ty,tx) reads ring slot
// z+tz at byte offset (ty*10 + tx)*16 with SBO = 160 B (next y row = next 8-row core-matrix group) and LBO = 2880 B
// (next 4-channel group).  One tcgen05.mma (M=128, N=NT, K=8 tf32) per (tap, chunk); the Dz output planes of
// the segment accumulate in Dz*NT TMEM columns across all channel chunks.
//
// Producers (4 warps): channels-last inputs arrive by TMA -- one 5-D box {4 channels, 10 x, 18 y, 1 z, 1 b} per 4-channel
// group lands the haloed plane in exactly the layout above, out-of-volume positions zero-filled by the TMA unit (round 1
// moved every 16-byte element with cp.async: 360 requests per plane-chunk, L1 request-rate bound at 5-9 us per plane;
// that path remains for NCDHW inputs, whose channels are not contiguous).  The producer threads then round their own
// elements to nearest TF32 in shared memory (the tensor core truncates) and publish the plane through an mbarrier; the
// MMA issuer is one elected lane of warp 4; the producer warps drain TMEM at the end.  Small grids are split over channel
// chunks (grid.z) with an atomic epilogue.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace mic {

constexpr int TX = 8, TY = 16;                 // footprint (x, y) -> 128 output positions per plane (UMMA M)
constexpr int HXS = TX + 2, HYS = TY + 2;      // haloed plane
constexpr int PPOS = HXS * HYS;                // 180 positions per staged plane
constexpr int KJ = 2;                          // 4-channel groups per K chunk (8 input channels = one MMA K step)
constexpr int CCH = 4 * KJ;
constexpr int SLAB = (PPOS * 16 + 127) / 128 * 128;   // one 4-channel group of a plane: 2880 B padded to 2944 (TMA wants 128-byte aligned destinations)
constexpr int PLANE_BYTES = KJ * SLAB;         // 5888
constexpr int RING = 8;
constexpr int AHEAD = RING - 3;                // plane-chunks in flight ahead of the one being published
constexpr int PL = (PPOS * KJ + 127) / 128;    // 16-byte elements per producer thread per plane-chunk (3)
constexpr int CT_THREADS = 160;

// optional phase trace of CTA 0 (debug): u64 globaltimer stamps, [event][index]
__device__ unsigned long long* g_conv_trace = nullptr;
__device__ __forceinline__ void ctrace(int ev, int idx) {
    unsigned long long* t = g_conv_trace;
    if (t && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && idx < 256) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        t[ev * 256 + idx] = now;
    }
}

struct ConvTcGeom {
    int B, D, H, W;
    int C0, C1;            // input channels: x0 | x1 (channels-last), or C0 planes of an NCDHW tensor (in_ncdhw)
    int N0, N1;            // output channels: y0 | y1 (channels-last), or N0 planes of an NCDHW tensor (out_ncdhw)
    int acc0, acc1;        // accumulate into y0 / y1 instead of storing
    int in_ncdhw, out_ncdhw, flip;
    int Dz, nseg, nfy, nfx;
    int ksplit, cps;       // channel-chunk splits (grid.z) and chunks per split; ksplit > 1 -> atomic epilogue
    int use_tma;           // channels-last inputs with 8-channel-aligned tensors: planes are loaded by TMA boxes
};

__device__ __forceinline__ void ctma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void cbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint32_t csmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(csmem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void cbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(csmem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CW_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CW_DONE;\n"
        "bra CW_LOOP;\n"
        "CW_DONE:\n"
        "}\n" ::"r"(csmem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ccommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(csmem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t cdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // sm_100 descriptor version; layout type 0 = no swizzle
    return d;
}
__device__ __forceinline__ void cmma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 rna4(float4 v) { return make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w)); }


__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void round_smem16(uint8_t* p) {
    float4* q = reinterpret_cast<float4*>(p);
    *q = rna4(*q);
}

template <int NT>
__global__ void __launch_bounds__(CT_THREADS, NT == 16 ? 3 : 2)
conv3_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ Wsrc,
                const float* __restrict__ bias, float* __restrict__ y0, float* __restrict__ y1, ConvTcGeom g, int tcols) {
    constexpr int W_BYTES = 27 * KJ * NT * 16;
    constexpr int WL = (27 * NT * KJ + 127) / 128;         // 16-byte weight elements per producer thread per chunk
    extern __shared__ __align__(128) uint8_t csm[];
    uint8_t* ring = csm;                                   // RING plane-chunks
    uint8_t* wbuf = csm + RING * PLANE_BYTES;              // 2 weight buffers
    uint64_t* pfull = reinterpret_cast<uint64_t*>(wbuf + 2 * W_BYTES);
    uint64_t* pempty = pfull + RING;
    uint64_t* wfull = pempty + RING;
    uint64_t* wempty = wfull + 2;
    uint64_t* accdone = wempty + 2;
    uint64_t* pland = accdone + 1;                         // [RING] TMA plane landed (tx bytes)
    uint32_t* tslot = reinterpret_cast<uint32_t*>(pland + RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cin = g.C0 + g.C1;
    const int Ntot = g.N0 + g.N1;
    const int n0 = blockIdx.y * NT;                        // first output channel of this CTA
    const int nchunks_all = (Cin + CCH - 1) / CCH;
    const int ch_lo = blockIdx.z * g.cps;
    const int nchunks = min(g.cps, nchunks_all - ch_lo);   // channel chunks of this CTA (>= 1 by construction)
    // tile decode: blockIdx.x = ((b*nseg + seg)*nfy + fy)*nfx + fx
    int t = blockIdx.x;
    const int fx = t % g.nfx; t /= g.nfx;
    const int fy = t % g.nfy; t /= g.nfy;
    const int seg = t % g.nseg; t /= g.nseg;
    const int b = t;
    const int xb = fx * TX, yb = fy * TY, zs = seg * g.Dz;
    const int nz = min(g.Dz, g.D - zs);                    // output planes of this segment
    const int ppc = nz + 2;                                // staged planes per chunk
    const int total = nchunks * ppc;                       // plane-chunks of this CTA
    const int64_t S = (int64_t)g.D * g.H * g.W;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RING; ++s) { cbar_init(&pfull[s], 128); cbar_init(&pempty[s], 1); cbar_init(&pland[s], 1); }
        for (int s = 0; s < 2; ++s) { cbar_init(&wfull[s], 128); cbar_init(&wempty[s], 1); }
        cbar_init(accdone, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(csmem_u32(tslot)), "r"(tcols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    pdl_sync();            // prologue above (barriers, TMEM) overlaps the previous kernel

    if (warp < 4) {
        // ------------------------------------------------------------------ producers
        const int tid = threadIdx.x;     // 0..127
        const uint32_t ring_u32 = csmem_u32(ring), wbuf_u32 = csmem_u32(wbuf);
        // this thread's elements of a plane-chunk: element e = tid + i*128 -> (kj = e % KJ, pos = e / KJ)
        int e_off[PL], e_dy[PL], e_dx[PL];
#pragma unroll
        for (int i = 0; i < PL; ++i) {
            const int e = tid + i * 128;
            const int kj = e % KJ, pos = e / KJ;
            e_off[i] = e < PPOS * KJ ? kj * SLAB + pos * 16 : -1;
            e_dy[i] = yb + pos / HXS - 1;
            e_dx[i] = xb + pos % HXS - 1;
        }
        // issue the copies of plane-chunk n (and, on the first plane of a chunk, that chunk's weights)
        auto issue = [&](int n) {
            const int ch = n / ppc, pz = n - ch * ppc;
            const int c0 = (ch_lo + ch) * CCH;
            const int z = zs - 1 + pz;
            const int slot = n % RING;
            cbar_wait(&pempty[slot], ((n / RING) & 1) ^ 1);
            if (tid == 0) ctrace(0, n);                          // slot free, copies of plane-chunk n issued
            const bool zok = z >= 0 && z < g.D;
            const uint32_t sbase = ring_u32 + slot * PLANE_BYTES;
            if (g.use_tma) {
                // one box per 4-channel group: {4 c, 10 x, 18 y, 1 z, 1 b} at (c, xb-1, yb-1, z, b); the TMA unit zero-fills
                // everything outside the tensor (the conv padding, and the whole plane for z = -1 / D)
                if (tid == 0) {
                    cbar_expect_tx(&pland[slot], KJ * PPOS * 16);
#pragma unroll
                    for (int kj = 0; kj < KJ; ++kj) {
                        const int c = c0 + kj * 4;
                        if (c < g.C0) ctma_load_5d(sbase + kj * SLAB, &map0, &pland[slot], c, xb - 1, yb - 1, z, b);
                        else ctma_load_5d(sbase + kj * SLAB, &map1, &pland[slot], c - g.C0, xb - 1, yb - 1, z, b);
                    }
                }
            } else
#pragma unroll
            for (int i = 0; i < PL; ++i) {
                if (e_off[i] < 0) continue;
                const int c = c0 + ((tid + i * 128) % KJ) * 4;
                const int yy = e_dy[i], xx = e_dx[i];
                const bool ok = zok && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W && c < Cin;
                if (g.in_ncdhw) {
                    const float* src = ok ? x0 + ((int64_t)b * Cin + c) * S + ((int64_t)z * g.H + yy) * g.W + xx : x0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) cp_async4(sbase + e_off[i] + 4 * j, ok && c + j < Cin ? src + j * S : x0, ok && c + j < Cin);
                } else {
                    const float* src = x0;
                    if (ok) {
                        const int64_t row = (((int64_t)b * g.D + z) * g.H + yy) * g.W + xx;
                        src = c < g.C0 ? x0 + row * g.C0 + c : x1 + row * g.C1 + (c - g.C0);
                    }
                    cp_async16(sbase + e_off[i], src, ok);
                }
            }
            if (pz == 0) {
                const int wb = ch & 1;
                cbar_wait(&wempty[wb], ((ch >> 1) & 1) ^ 1);
#pragma unroll
                for (int i = 0; i < WL; ++i) {
                    const int idx = tid + i * 128;
                    if (idx >= 27 * NT * KJ) continue;
                    const int kj = idx % KJ, o = (idx / KJ) % NT, tap = idx / (KJ * NT);
                    const int c = c0 + kj * 4;
                    const bool ok = n0 + o < Ntot && c < Cin;
                    const int ts = g.flip ? 26 - tap : tap;
                    cp_async16(wbuf_u32 + wb * W_BYTES + ((tap * KJ + kj) * NT + o) * 16,
                               ok ? Wsrc + ((int64_t)ts * Ntot + n0 + o) * Cin + c : Wsrc, ok);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // look-ahead: a chunk's weights are copied with its first plane and wait for the chunk two before to retire, so
        // the producer may not run further ahead than one chunk plus one plane (else it would wait on its own output)
        const int ahead = min(AHEAD, ppc + 1);                 // 4 or 5
        for (int n = 0; n < ahead; ++n) {
            if (n < total) issue(n);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int n = 0; n < total; ++n) {
            // groups committed so far: ahead + n; plane-chunk n is group n (with TMA planes the groups carry only the weights)
            if (ahead == AHEAD) asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD - 1) : "memory");
            else asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD - 2) : "memory");
            if (g.use_tma) cbar_wait(&pland[n % RING], (n / RING) & 1);
            if (tid == 0) ctrace(1, n);                          // plane-chunk n landed (this thread's part)
            const int ch = n / ppc, pz = n - ch * ppc;
            const int slot = n % RING;
            uint8_t* sbase = ring + slot * PLANE_BYTES;
#pragma unroll
            for (int i = 0; i < PL; ++i)
                if (e_off[i] >= 0) round_smem16(sbase + e_off[i]);
            if (pz == 0) {
                const int wb = ch & 1;
#pragma unroll
                for (int i = 0; i < WL; ++i) {
                    const int idx = tid + i * 128;
                    if (idx >= 27 * NT * KJ) continue;
                    const int kj = idx % KJ, o = (idx / KJ) % NT, tap = idx / (KJ * NT);
                    round_smem16(wbuf + wb * W_BYTES + ((tap * KJ + kj) * NT + o) * 16);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                cbar_arrive(&wfull[wb]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            cbar_arrive(&pfull[slot]);
            if (tid == 0) ctrace(2, n);                          // published
            if (n + ahead < total) issue(n + ahead);
            else asm volatile("cp.async.commit_group;" ::: "memory");          // keep the group count uniform
        }
        // ------------------------------------------------------------------ epilogue
        if (tid == 0) ctrace(5, 0);
        cbar_wait(accdone, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) ctrace(5, 1);
        const int q = warp;                                   // TMEM lane quarter
        const int r = q * 32 + lane;                          // row in the 128-position plane tile
        const int yy = yb + r / TX, xx = xb + r % TX;
        const bool ok = yy < g.H && xx < g.W;
        const bool atomic = g.ksplit > 1;
        const bool add_bias = bias != nullptr && blockIdx.z == 0;
        for (int zi = 0; zi < nz; ++zi) {
            const int z = zs + zi;
            const int64_t sp = ((int64_t)z * g.H + yy) * g.W + xx;
#pragma unroll
            for (int gi = 0; gi < NT / 16; ++gi) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(zi * NT + gi * 16)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (!ok) continue;
                const int nb = n0 + gi * 16;
                if (g.out_ncdhw) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) {
                        if (nb + o >= Ntot) continue;
                        float* dst = y0 + ((int64_t)b * Ntot + nb + o) * S + sp;
                        const float val = __uint_as_float(v[o]) + (add_bias ? bias[nb + o] : 0.f);
                        if (atomic) atomicAdd(dst, val);
                        else if (g.acc0) *dst += val;
                        else *dst = val;
                    }
                } else {
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        const int c = nb + o4 * 4;
                        if (c >= Ntot) continue;
                        float4 val = make_float4(__uint_as_float(v[o4 * 4]), __uint_as_float(v[o4 * 4 + 1]),
                                                 __uint_as_float(v[o4 * 4 + 2]), __uint_as_float(v[o4 * 4 + 3]));
                        if (add_bias) {
                            const float4 bb = *reinterpret_cast<const float4*>(bias + c);
                            val.x += bb.x; val.y += bb.y; val.z += bb.z; val.w += bb.w;
                        }
                        float4* dst;
                        int accf;
                        if (c < g.N0) { dst = reinterpret_cast<float4*>(y0 + ((int64_t)b * S + sp) * g.N0 + c); accf = g.acc0; }
                        else { dst = reinterpret_cast<float4*>(y1 + ((int64_t)b * S + sp) * g.N1 + (c - g.N0)); accf = g.acc1; }
                        if (atomic) atomicAdd(dst, val);
                        else {
                            if (accf) {
                                const float4 old = *dst;
                                val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                            }
                            *dst = val;
                        }
                    }
                }
            }
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t ring_addr = csmem_u32(ring), w_addr = csmem_u32(wbuf);
        int n0p = 0;
        for (int ch = 0; ch < nchunks; ++ch, n0p += ppc) {
            const int wb = ch & 1;
            cbar_wait(&wfull[wb], (ch >> 1) & 1);
            int waited = n0p - 1;                                 // highest plane-chunk index already waited for
            for (int zi = 0; zi < nz; ++zi) {
                const int need = n0p + zi + 2;                    // plane-chunks n0p+zi .. n0p+zi+2
                while (waited < need) {
                    ++waited;
                    cbar_wait(&pfull[waited % RING], (waited / RING) & 1);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                ctrace(3, ch * nz + zi);                          // MMA: planes of output plane (ch, zi) ready
                const uint32_t dcol = tmem + (uint32_t)(zi * NT);
                for (int tz = 0; tz < 3; ++tz) {
                    const uint32_t pbase = ring_addr + (uint32_t)(((n0p + zi + tz) % RING) * PLANE_BYTES);
#pragma unroll
                    for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                        for (int tx = 0; tx < 3; ++tx) {
                            const int tap = (tz * 3 + ty) * 3 + tx;
                            const uint64_t ad = cdesc(pbase + (uint32_t)((ty * HXS + tx) * 16), SLAB, HXS * 16);
                            const uint64_t bd = cdesc(w_addr + (uint32_t)(wb * W_BYTES + tap * KJ * NT * 16), NT * 16, 128);
                            cmma(dcol, ad, bd, idesc, (ch | tap) ? 1u : 0u);
                        }
                }
                ccommit(&pempty[(n0p + zi) % RING]);              // plane z-1 is no longer needed
                ctrace(4, ch * nz + zi);                          // MMAs issued
            }
            ccommit(&pempty[(n0p + nz) % RING]);                  // the last two planes of this chunk
            ccommit(&pempty[(n0p + nz + 1) % RING]);
            ccommit(&wempty[wb]);
        }
        ccommit(accdone);
    }
    if (threadIdx.x == 0) ctrace(5, 2);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tcols));
    }
}

typedef CUresult (*ConvEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ConvEncodeFn conv_get_encode() {
    static ConvEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<ConvEncodeFn>(p);
    }
    return fn;
}
// channels-last (B, D, H, W, C) fp32 tensor as a 5-D map, box {4 c, 10 x, 18 y, 1, 1}, zero fill outside
static bool conv_make_map(CUtensorMap* m, const float* base, int B, int D, int H, int W, int C) {
    ConvEncodeFn enc = conv_get_encode();
    if (!enc || !base) return false;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)D * H * W * C * 4};
    cuuint32_t box[5] = {4, (cuuint32_t)HXS, (cuuint32_t)HYS, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// zero_out: y must be all zeros on entry when the launch splits the channel chunks (atomic epilogue); returned through
// *needs_zero so that callers that own a fresh buffer can clear it themselves
template <int NT>
static int launch_conv_tc(const float* x0, const float* x1, const float* Wsrc, const float* bias, float* y0, float* y1,
                          ConvTcGeom g, bool allow_split, cudaStream_t st, const char* who) {
    const int Cin = g.C0 + g.C1, Ntot = g.N0 + g.N1;
    const int nchunks = (Cin + CCH - 1) / CCH;
    const int nych = (Ntot + NT - 1) / NT;
    g.nfy = (g.H + TY - 1) / TY; g.nfx = (g.W + TX - 1) / TX;
    const int foot = g.B * g.nfy * g.nfx;
    // CTAs wanted before the z segments stop shrinking / the channel chunks are split (MICFORMER_CONV_TARGET_PCT: % of the SM count)
    static const int target_pct = []() { const char* v = getenv("MICFORMER_CONV_TARGET_PCT"); return v ? atoi(v) : 200; }();
    const int target = num_sms() * target_pct / 100;
    // z segment length: enough CTAs to fill the GPU; Dz * NT accumulator columns, 3 (NT=16) or 2 (NT=32) CTAs per SM
    int Dz = 8;
    while (Dz > 2 && (int64_t)foot * ((g.D + Dz - 1) / Dz) * nych < target) Dz >>= 1;
    if (Dz > g.D) Dz = g.D;
    g.Dz = Dz; g.nseg = (g.D + Dz - 1) / Dz;
    const int64_t base = (int64_t)foot * g.nseg * nych;
    int ksplit = 1;
    if (allow_split && base < target) {
        ksplit = (int)((target + base - 1) / base);
        if (ksplit > nchunks) ksplit = nchunks;
    }
    g.cps = (nchunks + ksplit - 1) / ksplit;
    g.ksplit = (nchunks + g.cps - 1) / g.cps;
    int tcols = 32;
    while (tcols < Dz * NT) tcols <<= 1;
    const size_t smem = RING * PLANE_BYTES + 2 * (27 * KJ * NT * 16) + 256 + RING * 8;
    // TMA planes: channels-last inputs whose channel counts are multiples of 8 (a 4-channel group never straddles x0 | x1
    // and no K padding is needed); MICFORMER_CONV_TMA=0 keeps the cp.async producers
    static const bool tma_on = []() { const char* v = getenv("MICFORMER_CONV_TMA"); return !(v && v[0] == '0'); }();
    CUtensorMap m0, m1;
    memset(&m0, 0, sizeof(m0)); memset(&m1, 0, sizeof(m1));
    g.use_tma = 0;
    if (tma_on && !g.in_ncdhw && (g.C0 % 8) == 0 && (g.C1 % 8) == 0) {
        bool ok = conv_make_map(&m0, x0, g.B, g.D, g.H, g.W, g.C0);
        if (ok && g.C1 > 0) ok = conv_make_map(&m1, x1, g.B, g.D, g.H, g.W, g.C1);
        else if (ok) m1 = m0;
        g.use_tma = ok ? 1 : 0;
    }
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv3_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    if (g.ksplit > 1) {
        // atomic epilogue: the destination starts from zero (bias is added by split 0)
        const size_t bytes = (size_t)g.B * g.D * g.H * g.W * sizeof(float);
        cudaError_t e = cudaMemsetAsync(y0, 0, bytes * g.N0, st);
        if (e == cudaSuccess && y1) e = cudaMemsetAsync(y1, 0, bytes * g.N1, st);
        if (e != cudaSuccess) return fail(MIC_ERR_CUDA, "%s memset: %s", who, cudaGetErrorString(e));
    }
    dim3 grid((unsigned)((int64_t)foot * g.nseg), (unsigned)nych, (unsigned)g.ksplit);
    mic::launch((conv3_tc_kernel<NT>), grid, dim3(CT_THREADS), smem, st, m0, m1, x0, x1, Wsrc, bias, y0, y1, g, tcols);
    return check_launch(who);
}

static bool misaligned16(const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15); }

// returns MIC_ERR_UNSUPPORTED when the arguments are not taken (caller falls back to the CUDA-core kernel)
int tc_conv3_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wk, const float* bias, float* y, int B,
                 int D, int H, int W, int Co, int out_ncdhw, cudaStream_t st) {
    if (Co > 16 || Co < 1 || (C0 & 3) || (C1 & 3) || (!out_ncdhw && (Co & 3))) return MIC_ERR_UNSUPPORTED;
    if (misaligned16(x0) || misaligned16(x1) || misaligned16(Wk) || misaligned16(y) || misaligned16(bias)) return MIC_ERR_UNSUPPORTED;
    ConvTcGeom g{};
    g.B = B; g.D = D; g.H = H; g.W = W; g.C0 = C0; g.C1 = C1; g.N0 = Co; g.N1 = 0;
    g.out_ncdhw = out_ncdhw;
    return launch_conv_tc<16>(x0, x1, Wk, bias, y, nullptr, g, true, st, "conv3_tc_kernel<fwd>");
}

// dx0 | dx1 (channels-last, C0 | C1 channels) (+)= conv3^T(dy; Wt) with Wt = [27][C0 + C1][Co]; dy has Co channels,
// channels-last or NCDHW
int tc_conv3_bwd_data(const float* dy, const float* Wt, float* dx0, int C0, int acc0, float* dx1, int C1, int acc1, int B,
                      int D, int H, int W, int Co, int dy_ncdhw, cudaStream_t st) {
    if ((Co != 8 && Co != 16) || (C0 & 3) || (C1 & 3) || C0 + C1 < 16) return MIC_ERR_UNSUPPORTED;
    if (misaligned16(dy) || misaligned16(Wt) || misaligned16(dx0) || misaligned16(dx1)) return MIC_ERR_UNSUPPORTED;
    ConvTcGeom g{};
    g.B = B; g.D = D; g.H = H; g.W = W; g.C0 = Co; g.C1 = 0; g.N0 = C0; g.N1 = C1; g.acc0 = acc0; g.acc1 = acc1;
    g.in_ncdhw = dy_ncdhw; g.flip = 1;
    // the accumulate flags rule out a zero-initialised atomic epilogue: no channel-chunk split (K = Co is 1-2 chunks anyway)
    return launch_conv_tc<32>(dy, nullptr, Wt, nullptr, dx0, dx1, g, false, st, "conv3_tc_kernel<bwd_data>");
}

}  // namespace mic
extern "C" int mic_debug_conv_trace(void* buf) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
    return cudaMemcpyToSymbol(mic::g_conv_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}

extern "C" int mic_conv3_tc_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wk, const float* bias,
                                float* y, int B, int D, int H, int W, int Co, int out_ncdhw, void* stream) {
    MIC_REQUIRE(x0 && Wk && y && (C1 == 0 || x1), "conv3_tc_fwd: null pointer");
    int rc = mic::tc_conv3_fwd(x0, C0, x1, C1, Wk, bias, y, B, D, H, W, Co, out_ncdhw, (cudaStream_t)stream);
    if (rc == MIC_ERR_UNSUPPORTED) return mic::fail(MIC_ERR_UNSUPPORTED, "conv3_tc_fwd: geometry (%d,%d,%d) Co=%d not taken", D, H, W, Co);
    return rc;
}

extern "C" int mic_conv3_tc_bwd_data(const float* dy, const float* Wt, float* dx0, int C0, int acc0, float* dx1, int C1,
                                     int acc1, int B, int D, int H, int W, int Co, int dy_ncdhw, void* stream) {
    MIC_REQUIRE(dy && Wt && dx0 && (C1 == 0 || dx1), "conv3_tc_bwd_data: null pointer");
    int rc = mic::tc_conv3_bwd_data(dy, Wt, dx0, C0, acc0, dx1, C1, acc1, B, D, H, W, Co, dy_ncdhw, (cudaStream_t)stream);
    if (rc == MIC_ERR_UNSUPPORTED) return mic::fail(MIC_ERR_UNSUPPORTED, "conv3_tc_bwd_data: geometry (%d,%d,%d) Co=%d not taken", D, H, W, Co);
    return rc;
}
