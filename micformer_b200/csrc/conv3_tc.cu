// tcgen05 implicit-GEMM 3x3x3 convolution (stride 1, zero pad 1) on channels-last token grids -- forward and
// backward-data (the same kernel: backward-data is the forward conv of dy with the taps mirrored and the roles of
// the two channel axes swapped, so Wt[tap][ci][co] is read as the [tap][n][k] operand with tap -> 26 - tap).
//
//   y[p, n] = bias[n] + sum_{tap} sum_{c} x[p + tap, c] * Wsrc[tap'][n][c]     (M = positions, N tile = 16 | 32, K = 27 * Cin)
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
constexpr int RING = 4;
constexpr int CT_THREADS = 160;
// template parameters of the kernel: KJ = 4-channel groups per K chunk (chunk = 4*KJ input channels: 16, or 8 for the
// 8-channel dy of out_conv's backward), NT = output channels per CTA (UMMA N).
//   plane slot  : KJ * PPOS * 16 bytes            [kj][pos][4 floats]
//   weight chunk: 27 * KJ * NT * 16 bytes         [tap][kj][n][4 floats]

struct ConvTcGeom {
    int B, D, H, W;
    int C0, C1;            // input channels: x0 | x1 (channels-last), or C0 planes of an NCDHW tensor (in_ncdhw)
    int N0, N1;            // output channels: y0 | y1 (channels-last), or N0 planes of an NCDHW tensor (out_ncdhw)
    int acc0, acc1;        // accumulate into y0 / y1 instead of storing
    int in_ncdhw, out_ncdhw, flip;
    int Dz, nseg, nfy, nfx, nwb;
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

template <int KJ, int NT>
__global__ void __launch_bounds__(CT_THREADS, 2)
conv3_tc_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ Wsrc,
                const float* __restrict__ bias, float* __restrict__ y0, float* __restrict__ y1, ConvTcGeom g, int tcols) {
    constexpr int CCH = 4 * KJ;
    constexpr int PLANE_BYTES = KJ * PPOS * 16;
    constexpr int W_BYTES = 27 * KJ * NT * 16;
    extern __shared__ __align__(128) uint8_t csm[];
    uint8_t* ring = csm;                                   // RING planes
    uint8_t* wbuf = csm + RING * PLANE_BYTES;              // nwb weight buffers
    uint64_t* pfull = reinterpret_cast<uint64_t*>(wbuf + g.nwb * W_BYTES);
    uint64_t* pempty = pfull + RING;
    uint64_t* wfull = pempty + RING;
    uint64_t* wempty = wfull + 2;
    uint64_t* accdone = wempty + 2;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(accdone + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cin = g.C0 + g.C1;
    const int Ntot = g.N0 + g.N1;
    const int n0 = blockIdx.y * NT;                        // first output channel of this CTA
    const int nchunks = (Cin + CCH - 1) / CCH;
    const int nwb = g.nwb;
    // tile decode: blockIdx.x = ((b*nseg + seg)*nfy + fy)*nfx + fx
    int t = blockIdx.x;
    const int fx = t % g.nfx; t /= g.nfx;
    const int fy = t % g.nfy; t /= g.nfy;
    const int seg = t % g.nseg; t /= g.nseg;
    const int b = t;
    const int xb = fx * TX, yb = fy * TY, zs = seg * g.Dz;
    const int nz = min(g.Dz, g.D - zs);                    // output planes of this segment
    const int planes_per_chunk = nz + 2;
    const int64_t S = (int64_t)g.D * g.H * g.W;

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
        constexpr int PL = (PPOS * KJ + 127) / 128;          // float4 gathers per thread per plane
        constexpr int WL = (27 * NT * KJ + 127) / 128;       // float4 loads per thread per weight chunk
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
                    if (g.in_ncdhw) {
                        const float* src = x0 + ((int64_t)b * Cin + c) * S + ((int64_t)z * g.H + yy) * g.W + xx;
                        r[i].x = src[0];
                        if (c + 1 < Cin) r[i].y = src[S];
                        if (c + 2 < Cin) r[i].z = src[2 * S];
                        if (c + 3 < Cin) r[i].w = src[3 * S];
                    } else {
                        const int64_t row = (((int64_t)b * g.D + z) * g.H + yy) * g.W + xx;
                        r[i] = c < g.C0 ? *reinterpret_cast<const float4*>(x0 + row * g.C0 + c)
                                        : *reinterpret_cast<const float4*>(x1 + row * g.C1 + (c - g.C0));
                    }
                }
            }
        };
        float4 cur[PL], nxt[PL];
        gather(zs - 1, 0, cur);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * CCH;
            // weights of this chunk: smem [tap][kj][n][4] <- Wsrc[tap'][n0 + n][c0 + 4 kj ..]
            {
                const int wb = ch % nwb;
                float4 wr[WL];
#pragma unroll
                for (int i = 0; i < WL; ++i) {
                    const int idx = tid + i * 128;
                    const int kj = idx % KJ, o = (idx / KJ) % NT, tap = idx / (KJ * NT);
                    const int c = c0 + kj * 4;
                    wr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < 27 * NT * KJ && n0 + o < Ntot && c < Cin) {
                        const int ts = g.flip ? 26 - tap : tap;
                        const float* src = Wsrc + ((int64_t)ts * Ntot + n0 + o) * Cin + c;
                        if ((Cin & 3) == 0) wr[i] = *reinterpret_cast<const float4*>(src);
                        else {
                            wr[i].x = src[0];
                            if (c + 1 < Cin) wr[i].y = src[1];
                            if (c + 2 < Cin) wr[i].z = src[2];
                            if (c + 3 < Cin) wr[i].w = src[3];
                        }
                    }
                }
                cbar_wait(&wempty[wb], ((ch / nwb) & 1) ^ 1);
                float4* wd = reinterpret_cast<float4*>(wbuf + wb * W_BYTES);
#pragma unroll
                for (int i = 0; i < WL; ++i) {
                    const int idx = tid + i * 128;
                    const int kj = idx % KJ, o = (idx / KJ) % NT, tap = idx / (KJ * NT);
                    if (idx < 27 * NT * KJ) wd[(tap * KJ + kj) * NT + o] = rna4(wr[i]);
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
                    for (int o = 0; o < 16 && nb + o < Ntot; ++o) {
                        float* dst = y0 + ((int64_t)b * Ntot + nb + o) * S + sp;
                        float val = __uint_as_float(v[o]) + (bias ? bias[nb + o] : 0.f);
                        if (g.acc0) val += *dst;
                        *dst = val;
                    }
                } else {
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        const int c = nb + o4 * 4;
                        if (c >= Ntot) continue;
                        float4 val = make_float4(__uint_as_float(v[o4 * 4]), __uint_as_float(v[o4 * 4 + 1]),
                                                 __uint_as_float(v[o4 * 4 + 2]), __uint_as_float(v[o4 * 4 + 3]));
                        if (bias) {
                            const float4 bb = *reinterpret_cast<const float4*>(bias + c);
                            val.x += bb.x; val.y += bb.y; val.z += bb.z; val.w += bb.w;
                        }
                        float4* dst;
                        int accf;
                        if (c < g.N0) { dst = reinterpret_cast<float4*>(y0 + ((int64_t)b * S + sp) * g.N0 + c); accf = g.acc0; }
                        else { dst = reinterpret_cast<float4*>(y1 + ((int64_t)b * S + sp) * g.N1 + (c - g.N0)); accf = g.acc1; }
                        if (accf) {
                            const float4 old = *dst;
                            val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                        }
                        *dst = val;
                    }
                }
            }
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t ring_addr = csmem_u32(ring), w_addr = csmem_u32(wbuf);
        int n0p = 0;
        for (int ch = 0; ch < nchunks; ++ch, n0p += planes_per_chunk) {
            const int wb = ch % nwb;
            cbar_wait(&wfull[wb], (ch / nwb) & 1);
            const int cvalid = min(CCH, Cin - ch * CCH);
            const int ksteps = (cvalid + 7) / 8;                 // K = 8 channels per MMA
            int waited = n0p - 1;                                 // highest plane index already waited for
            for (int zi = 0; zi < nz; ++zi) {
                const int need = n0p + zi + 2;                    // planes n0p+zi .. n0p+zi+2
                while (waited < need) {
                    ++waited;
                    cbar_wait(&pfull[waited % RING], (waited / RING) & 1);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t dcol = tmem + (uint32_t)(zi * NT);
                for (int tz = 0; tz < 3; ++tz) {
                    const uint32_t pbase = ring_addr + (uint32_t)(((n0p + zi + tz) % RING) * PLANE_BYTES);
#pragma unroll
                    for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                        for (int tx = 0; tx < 3; ++tx) {
                            const int tap = (tz * 3 + ty) * 3 + tx;
                            for (int ks = 0; ks < ksteps; ++ks) {
                                const uint64_t ad = cdesc(pbase + (uint32_t)((ty * HXS + tx) * 16 + 2 * ks * PPOS * 16), PPOS * 16, HXS * 16);
                                const uint64_t bd = cdesc(w_addr + (uint32_t)(wb * W_BYTES + (tap * KJ + 2 * ks) * NT * 16), NT * 16, 128);
                                cmma(dcol, ad, bd, idesc, (ch | tap | ks) ? 1u : 0u);
                            }
                        }
                }
                ccommit(&pempty[(n0p + zi) % RING]);              // plane z-1 is no longer needed
            }
            ccommit(&pempty[(n0p + nz) % RING]);                  // the last two planes of this chunk
            ccommit(&pempty[(n0p + nz + 1) % RING]);
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

template <int KJ, int NT>
static int launch_conv_tc(const float* x0, const float* x1, const float* Wsrc, const float* bias, float* y0, float* y1,
                          ConvTcGeom g, cudaStream_t st, const char* who) {
    const int Cin = g.C0 + g.C1, Ntot = g.N0 + g.N1;
    const int nchunks = (Cin + 4 * KJ - 1) / (4 * KJ);
    const int nych = (Ntot + NT - 1) / NT;
    g.nfy = g.H / TY; g.nfx = g.W / TX;
    g.nwb = nchunks > 1 ? 2 : 1;
    const int foot = g.B * g.nfy * g.nfx;
    // z segment length: enough CTAs to fill the GPU; Dz * NT accumulator columns, two CTAs per SM share 512
    int Dz = 256 / NT;
    while (Dz > 2 && (int64_t)foot * ((g.D + Dz - 1) / Dz) * nych < 2 * num_sms()) Dz >>= 1;
    if (Dz > g.D) Dz = g.D;
    g.Dz = Dz; g.nseg = (g.D + Dz - 1) / Dz;
    int tcols = 32;
    while (tcols < Dz * NT) tcols <<= 1;
    const size_t smem = RING * (KJ * PPOS * 16) + g.nwb * (27 * KJ * NT * 16) + 256;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv3_tc_kernel<KJ, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             RING * (KJ * PPOS * 16) + 2 * (27 * KJ * NT * 16) + 256);
        attr = true;
    }
    dim3 grid((unsigned)((int64_t)foot * g.nseg), (unsigned)nych);
    conv3_tc_kernel<KJ, NT><<<grid, CT_THREADS, smem, st>>>(x0, x1, Wsrc, bias, y0, y1, g, tcols);
    return check_launch(who);
}

static bool misaligned16(const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15); }

// returns MIC_ERR_UNSUPPORTED when the geometry is not taken (caller falls back to the CUDA-core kernel)
int tc_conv3_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wk, const float* bias, float* y, int B,
                 int D, int H, int W, int Co, int out_ncdhw, cudaStream_t st) {
    if (W % TX || H % TY || Co > 16 || Co < 1 || (C0 & 3) || (C1 & 3) || (!out_ncdhw && (Co & 3))) return MIC_ERR_UNSUPPORTED;
    if (misaligned16(x0) || misaligned16(x1) || misaligned16(Wk) || misaligned16(y) || misaligned16(bias)) return MIC_ERR_UNSUPPORTED;
    ConvTcGeom g{};
    g.B = B; g.D = D; g.H = H; g.W = W; g.C0 = C0; g.C1 = C1; g.N0 = Co; g.N1 = 0;
    g.out_ncdhw = out_ncdhw;
    return launch_conv_tc<4, 16>(x0, x1, Wk, bias, y, nullptr, g, st, "conv3_tc_kernel<fwd>");
}

// dx0 | dx1 (channels-last, C0 | C1 channels) (+)= conv3^T(dy; Wt) with Wt = [27][C0 + C1][Co]; dy has Co channels,
// channels-last or NCDHW
int tc_conv3_bwd_data(const float* dy, const float* Wt, float* dx0, int C0, int acc0, float* dx1, int C1, int acc1, int B,
                      int D, int H, int W, int Co, int dy_ncdhw, cudaStream_t st) {
    if (W % TX || H % TY || (Co != 8 && Co != 16) || (C0 & 3) || (C1 & 3) || C0 + C1 < 16) return MIC_ERR_UNSUPPORTED;
    if (misaligned16(dy) || misaligned16(Wt) || misaligned16(dx0) || misaligned16(dx1)) return MIC_ERR_UNSUPPORTED;
    ConvTcGeom g{};
    g.B = B; g.D = D; g.H = H; g.W = W; g.C0 = Co; g.C1 = 0; g.N0 = C0; g.N1 = C1; g.acc0 = acc0; g.acc1 = acc1;
    g.in_ncdhw = dy_ncdhw; g.flip = 1;
    if (Co == 8) return launch_conv_tc<2, 32>(dy, nullptr, Wt, nullptr, dx0, dx1, g, st, "conv3_tc_kernel<bwd_data,8>");
    return launch_conv_tc<4, 32>(dy, nullptr, Wt, nullptr, dx0, dx1, g, st, "conv3_tc_kernel<bwd_data,16>");
}

}  // namespace mic

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
