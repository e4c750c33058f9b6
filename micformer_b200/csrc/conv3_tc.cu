// tcgen05 implicit-GEMM 3x3x3 convolution (stride 1, zero pad 1) on channels-last token grids -- forward.
//
//   y[p, o] = bias[o] + sum_{tap} sum_{c} x[p + tap, c] * Wk[tap][o][c]        (M = positions, N = 16, K = 27 * Cin)
//
// No im2col and no per-tap re-fetch: a CTA owns an (8 x 16) x/y footprint and marches over a z segment.  Each
// input z-plane of the footprint (+1 halo: 10 x 18 positions) is staged ONCE per 16-channel chunk into a 4-slot
// shared-memory ring in the tcgen05 "no swizzle" K-major core-matrix layout
//     plane[kj][pos][4 floats]      (kj = 4-channel group, pos = y'*10 + x' in the haloed plane)
// so that all 27 taps are just different START ADDRESSES of the same staged data: tap (tz,ty,tx) reads ring slot
// z+tz at byte offset (ty*10 + tx)*16 with SBO = 160 B (next y row = next 8-row core-matrix group) and LBO = 2880 B
// (next 4-channel group).  One tcgen05.mma (M=128, N=16, K=8 tf32) per (tap, 8 channels); the Dz output planes of
// the segment accumulate in Dz*16 TMEM columns across all channel chunks.  Operands are rounded to nearest TF32
// while being staged.  Producer / epilogue: 4 warps (coalesced float4 gathers, zero fill outside the volume);
// MMA: 1 elected lane of warp 4.
#include "common.cuh"

namespace mic {

constexpr int TX = 8, TY = 16;                 // footprint (x, y) -> 128 output positions per plane (UMMA M)
constexpr int HXS = TX + 2, HYS = TY + 2;      // haloed plane
constexpr int PPOS = HXS * HYS;                // 180 positions per staged plane
constexpr int CCH = 16;                        // channels per chunk
constexpr int KJ = CCH / 4;                    // 4-channel groups per chunk
constexpr int PLANE_BYTES = KJ * PPOS * 16;    // 11520
constexpr int RING = 4;
constexpr int W_BYTES = 27 * KJ * 16 * 16;     // 27648: [tap][kj][o=16][4 floats]
constexpr int CT_THREADS = 160;

struct ConvTcGeom {
    int B, D, H, W, C0, C1, Co, Dz, nseg, nfy, nfx;
};

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

__global__ void __launch_bounds__(CT_THREADS, 2)
conv3_tc_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ Wk,
                    const float* __restrict__ bias, float* __restrict__ y, ConvTcGeom g, int out_ncdhw, int tcols) {
    extern __shared__ __align__(128) uint8_t csm[];
    uint8_t* ring = csm;                                   // RING planes
    uint8_t* wbuf = csm + RING * PLANE_BYTES;              // 2 weight buffers
    uint64_t* pfull = reinterpret_cast<uint64_t*>(wbuf + 2 * W_BYTES);
    uint64_t* pempty = pfull + RING;
    uint64_t* wfull = pempty + RING;
    uint64_t* wempty = wfull + 2;
    uint64_t* accdone = wempty + 2;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(accdone + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cin = g.C0 + g.C1;
    const int nchunks = (Cin + CCH - 1) / CCH;
    // tile decode: blockIdx.x = ((b*nseg + seg)*nfy + fy)*nfx + fx
    int t = blockIdx.x;
    const int fx = t % g.nfx; t /= g.nfx;
    const int fy = t % g.nfy; t /= g.nfy;
    const int seg = t % g.nseg; t /= g.nseg;
    const int b = t;
    const int xb = fx * TX, yb = fy * TY, zs = seg * g.Dz;
    const int nz = min(g.Dz, g.D - zs);                    // output planes of this segment
    const int planes_per_chunk = nz + 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RING; ++s) { cbar_init(&pfull[s], 128); cbar_init(&pempty[s], 1); }
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

    if (warp < 4) {
        // ------------------------------------------------------------------ producers
        const int tid = threadIdx.x;     // 0..127
        int n = 0;                       // running plane counter (ring position)
        constexpr int PL = (PPOS * KJ + 127) / 128;          // float4 gathers per thread per plane (6)
        constexpr int WL = (27 * 16 * KJ + 127) / 128;       // float4 loads per thread per weight chunk (14)
        // gather one haloed plane chunk into registers (all loads in flight before the first use)
        auto gather = [&](int z, int c0, float4 (&r)[PL]) {
            const bool zok = z >= 0 && z < g.D;
#pragma unroll
            for (int i = 0; i < PL; ++i) {
                const int idx = tid + i * 128;
                const int kj = idx % KJ, pos = idx / KJ;
                const int yy = yb + pos / HXS - 1, xx = xb + pos % HXS - 1;
                const int c = c0 + kj * 4;
                r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < PPOS * KJ && zok && yy >= 0 && yy < g.H && xx >= 0 && xx < g.W && c < Cin) {
                    const int64_t row = (((int64_t)b * g.D + z) * g.H + yy) * g.W + xx;
                    r[i] = c < g.C0 ? *reinterpret_cast<const float4*>(x0 + row * g.C0 + c)
                                    : *reinterpret_cast<const float4*>(x1 + row * g.C1 + (c - g.C0));
                }
            }
        };
        float4 cur[PL], nxt[PL];
        gather(zs - 1, 0, cur);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * CCH;
            // weights of this chunk: [tap][kj][o][4]
            {
                const int wb = ch & 1;
                float4 wr[WL];
#pragma unroll
                for (int i = 0; i < WL; ++i) {
                    const int idx = tid + i * 128;
                    const int kj = idx % KJ, o = (idx / KJ) % 16, tap = idx / (KJ * 16);
                    const int c = c0 + kj * 4;
                    wr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < 27 * 16 * KJ && o < g.Co && c < Cin)
                        wr[i] = *reinterpret_cast<const float4*>(Wk + ((int64_t)tap * g.Co + o) * Cin + c);
                }
                cbar_wait(&wempty[wb], ((ch >> 1) & 1) ^ 1);
                float4* wd = reinterpret_cast<float4*>(wbuf + wb * W_BYTES);
#pragma unroll
                for (int i = 0; i < WL; ++i) {
                    const int idx = tid + i * 128;
                    const int kj = idx % KJ, o = (idx / KJ) % 16, tap = idx / (KJ * 16);
                    if (idx < 27 * 16 * KJ) wd[(tap * KJ + kj) * 16 + o] = rna4(wr[i]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                cbar_arrive(&wfull[wb]);
            }
            for (int pz = 0; pz < planes_per_chunk; ++pz, ++n) {
                // prefetch the next plane (possibly the first plane of the next chunk) while this one is stored
                const bool last_plane = pz + 1 == planes_per_chunk;
                if (!last_plane) gather(zs + pz, c0, nxt);
                else if (ch + 1 < nchunks) gather(zs - 1, c0 + CCH, nxt);
                const int slot = n % RING;
                cbar_wait(&pempty[slot], ((n / RING) & 1) ^ 1);
                float4* pd = reinterpret_cast<float4*>(ring + slot * PLANE_BYTES);
#pragma unroll
                for (int i = 0; i < PL; ++i) {
                    const int idx = tid + i * 128;
                    const int kj = idx % KJ, pos = idx / KJ;
                    if (idx < PPOS * KJ) pd[kj * PPOS + pos] = rna4(cur[i]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                cbar_arrive(&pfull[slot]);
#pragma unroll
                for (int i = 0; i < PL; ++i) cur[i] = nxt[i];
            }
        }
        // ------------------------------------------------------------------ epilogue
        cbar_wait(accdone, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp;                                   // TMEM lane quarter
        const int r = q * 32 + lane;                          // row in the 128-position plane tile
        const int yy = yb + r / TX, xx = xb + r % TX;
        const bool ok = yy < g.H && xx < g.W;
        const int64_t S = (int64_t)g.D * g.H * g.W;
        for (int zi = 0; zi < nz; ++zi) {
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(zi * 16)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (!ok) continue;
            const int z = zs + zi;
            const int64_t sp = ((int64_t)z * g.H + yy) * g.W + xx;
            if (out_ncdhw) {
                for (int o = 0; o < g.Co; ++o)
                    y[((int64_t)b * g.Co + o) * S + sp] = __uint_as_float(v[o]) + (bias ? bias[o] : 0.f);
            } else {
                float* dst = y + ((int64_t)b * S + sp) * g.Co;
                if (g.Co == 16) {
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        float4 bb = bias ? *reinterpret_cast<const float4*>(bias + o4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4*>(dst + o4 * 4) =
                            make_float4(__uint_as_float(v[o4 * 4]) + bb.x, __uint_as_float(v[o4 * 4 + 1]) + bb.y,
                                        __uint_as_float(v[o4 * 4 + 2]) + bb.z, __uint_as_float(v[o4 * 4 + 3]) + bb.w);
                    }
                } else {
                    for (int o = 0; o < g.Co; ++o) dst[o] = __uint_as_float(v[o]) + (bias ? bias[o] : 0.f);
                }
            }
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t ring_addr = csmem_u32(ring), w_addr = csmem_u32(wbuf);
        int n0 = 0;
        for (int ch = 0; ch < nchunks; ++ch, n0 += planes_per_chunk) {
            const int wb = ch & 1;
            cbar_wait(&wfull[wb], (ch >> 1) & 1);
            const int cvalid = min(CCH, Cin - ch * CCH);
            const int ksteps = (cvalid + 7) / 8;                 // K = 8 channels per MMA
            int waited = n0 - 1;                                  // highest plane index already waited for
            for (int zi = 0; zi < nz; ++zi) {
                const int need = n0 + zi + 2;                     // planes n0+zi .. n0+zi+2
                while (waited < need) {
                    ++waited;
                    cbar_wait(&pfull[waited % RING], (waited / RING) & 1);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem + (uint32_t)(zi * 16);
                for (int tz = 0; tz < 3; ++tz) {
                    const uint32_t pbase = ring_addr + (uint32_t)(((n0 + zi + tz) % RING) * PLANE_BYTES);
#pragma unroll
                    for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                        for (int tx = 0; tx < 3; ++tx) {
                            const int tap = (tz * 3 + ty) * 3 + tx;
                            for (int ks = 0; ks < ksteps; ++ks) {
                                const uint64_t ad = cdesc(pbase + (uint32_t)((ty * HXS + tx) * 16 + 2 * ks * PPOS * 16), PPOS * 16, HXS * 16);
                                const uint64_t bd = cdesc(w_addr + (uint32_t)(wb * W_BYTES + (tap * KJ + 2 * ks) * 256), 256, 128);
                                cmma(dcol, ad, bd, idesc, (ch | tap | ks) ? 1u : 0u);
                            }
                        }
                }
                ccommit(&pempty[(n0 + zi) % RING]);               // plane z-1 is no longer needed
            }
            ccommit(&pempty[(n0 + nz) % RING]);                   // the last two planes of this chunk
            ccommit(&pempty[(n0 + nz + 1) % RING]);
            ccommit(&wempty[wb]);
        }
        ccommit(accdone);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tcols));
    }
}

// returns MIC_ERR_UNSUPPORTED when the geometry is not taken (caller falls back to the CUDA-core kernel)
int tc_conv3_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wk, const float* bias, float* y, int B,
                 int D, int H, int W, int Co, int out_ncdhw, cudaStream_t st) {
    if (W % TX || H % TY || Co > 16 || Co < 1 || (C0 & 3) || (C1 & 3)) return MIC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x0) & 15) || (x1 && (reinterpret_cast<uintptr_t>(x1) & 15)) ||
        (reinterpret_cast<uintptr_t>(Wk) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
        return MIC_ERR_UNSUPPORTED;
    ConvTcGeom g{};
    g.B = B; g.D = D; g.H = H; g.W = W; g.C0 = C0; g.C1 = C1; g.Co = Co;
    g.nfy = H / TY; g.nfx = W / TX;
    const int foot = B * g.nfy * g.nfx;
    // z segment length: enough CTAs to fill the GPU, at most 16 planes (16 TMEM columns each)
    int Dz = 16;
    while (Dz > 2 && (int64_t)foot * ((D + Dz - 1) / Dz) < 2 * num_sms()) Dz >>= 1;
    if (Dz > D) Dz = D;
    g.Dz = Dz; g.nseg = (D + Dz - 1) / Dz;
    int tcols = 32;
    while (tcols < Dz * 16) tcols <<= 1;
    const size_t smem = RING * PLANE_BYTES + 2 * W_BYTES + 256;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv3_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const int64_t ctas = (int64_t)foot * g.nseg;
    conv3_tc_fwd_kernel<<<(unsigned)ctas, CT_THREADS, smem, st>>>(x0, x1, Wk, bias, y, g, out_ncdhw, tcols);
    return check_launch("conv3_tc_fwd_kernel");
}

}  // namespace mic

extern "C" int mic_conv3_tc_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wk, const float* bias,
                                float* y, int B, int D, int H, int W, int Co, int out_ncdhw, void* stream) {
    MIC_REQUIRE(x0 && Wk && y && (C1 == 0 || x1), "conv3_tc_fwd: null pointer");
    int rc = mic::tc_conv3_fwd(x0, C0, x1, C1, Wk, bias, y, B, D, H, W, Co, out_ncdhw, (cudaStream_t)stream);
    if (rc == MIC_ERR_UNSUPPORTED) return mic::fail(MIC_ERR_UNSUPPORTED, "conv3_tc_fwd: geometry (%d,%d,%d) Co=%d not taken", D, H, W, Co);
    return rc;
}
