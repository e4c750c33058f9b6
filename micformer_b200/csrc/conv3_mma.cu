// 3x3x3 convolution backward-weight on the tensor cores (warp-level mma.sync m16n8k8 TF32, fp32 accumulate):
//
//   dWt[tap][ci][co] += sum_p x[p + tap, ci] * dy[p, co]            (M = ci, N = co, K = positions)
//
// The reduction index is the POSITION, so every tap is the same dy tile against an x tile shifted by the tap
// offset -- sub-16-byte shifts of a K-contiguous operand, which the tcgen05 shared-memory descriptors cannot
// express (start addresses are 16-byte granular) without staging several shifted copies.  Warp-level MMA reads
// its fragments with ordinary shared-memory loads, so the shift is index arithmetic.
//
// CTA = 9 warps, one per (tz, ty) tap row (3 taps each); a CTA owns a 32-input-channel chunk (blockIdx.y) and walks
// a strided set of 4x4x8 position bricks.  Per brick the x halo (6x6x10 positions x 32 channels) and the dy brick
// (128 x Co) are staged once in shared memory, rounded to nearest TF32; each warp then runs 16 k-steps (one 8-wide
// x row of the brick each) of 3 taps x 2 channel tiles x (Co/8) MMAs on register accumulators that live across all
// bricks and are flushed once with atomics.  Row strides (40 / 24 floats) make all fragment loads conflict-free.
//
// Measured bound (round 2, gpurun r2bj): the kernel sits on the legacy mma.sync TF32 issue rate of this chip (~1
// m16n8k8 per 9-11 SM clocks: 864 MMAs per brick in ~9.4k clocks), NOT on its shared-memory fragment loads -- a variant
// that applied every x row to all nine (tz, ty) taps (2 LDS per MMA instead of 4.3, cp.async double-buffered bricks, one
// CTA per SM) ran out_conv's gradient in 1.34 ms against 1.06 ms here and was dropped.
#include "common.cuh"

namespace mic {

namespace {

constexpr int MB_Z = 4, MB_Y = 4, MB_X = 8;
constexpr int MH_Z = MB_Z + 2, MH_Y = MB_Y + 2, MH_X = MB_X + 2;
constexpr int M_NB = MB_Z * MB_Y * MB_X;     // 128 positions per brick
constexpr int M_NH = MH_Z * MH_Y * MH_X;     // 360 halo positions
constexpr int XS = 40;                       // x row stride (floats): bank = 8*pos + ci  -> conflict-free A fragments
constexpr int DS = 24;                       // dy row stride: bank = 24*pos + co -> conflict-free B fragments
constexpr int M_THREADS = 288;

struct MGeom {
    int B, D, H, W, C0, C1, Co;
};

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int COP>   // 8 or 16 output channels (padded)
__global__ void __launch_bounds__(M_THREADS, 2)
conv3_mma_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x0, const float* __restrict__ x1,
                            float* __restrict__ dWt, float* __restrict__ dbias, MGeom g, int dy_ncdhw, int nbz, int nby,
                            int nbx, int native) {
    pdl_sync();
    constexpr int NTN = COP / 8;
    extern __shared__ __align__(16) float msm[];
    float* Xs = msm;                        // [M_NH][XS]
    float* dYs = msm + M_NH * XS;           // [M_NB][DS]
    float* bsum = dYs + M_NB * DS;          // [16]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int c0 = blockIdx.y * 32;
    const int Cin = g.C0 + g.C1;
    const int64_t S = (int64_t)g.D * g.H * g.W;
    const int64_t nbricks = (int64_t)g.B * nbz * nby * nbx;
    const int tz = warp / 3, ty = warp % 3;
    const bool want_bias = dbias != nullptr && blockIdx.y == 0;

    float acc[3][2][NTN][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < NTN; ++c)
#pragma unroll
                for (int d = 0; d < 4; ++d) acc[a][b][c][d] = 0.f;
    if (tid < 16) bsum[tid] = 0.f;
    constexpr int DLB = (M_NB * COP + M_THREADS - 1) / M_THREADS;
    float bacc[DLB];
#pragma unroll
    for (int i = 0; i < DLB; ++i) bacc[i] = 0.f;

    for (int64_t brick = blockIdx.x; brick < nbricks; brick += gridDim.x) {
        int64_t t = brick;
        const int bx = (int)(t % nbx); t /= nbx;
        const int by = (int)(t % nby); t /= nby;
        const int bz = (int)(t % nbz); t /= nbz;
        const int b = (int)t;
        const int z0 = bz * MB_Z, y0 = by * MB_Y, x0c = bx * MB_X;
        __syncthreads();                    // previous brick fully consumed
        // x halo (zero outside the volume), 32 channels, rounded to TF32: all loads of the thread in flight before the
        // first store
        {
            constexpr int XL = (M_NH * 8 + M_THREADS - 1) / M_THREADS;      // 10
            float4 xv[XL];
#pragma unroll
            for (int i = 0; i < XL; ++i) {
                const int idx = tid + i * M_THREADS;
                const int hp = idx >> 3, c4 = (idx & 7) * 4;
                const int hx = hp % MH_X, hy = (hp / MH_X) % MH_Y, hz = hp / (MH_X * MH_Y);
                const int z = z0 + hz - 1, yy = y0 + hy - 1, x = x0c + hx - 1;
                xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int c = c0 + c4;
                if (idx < M_NH * 8 && z >= 0 && z < g.D && yy >= 0 && yy < g.H && x >= 0 && x < g.W && c < Cin) {
                    const int64_t row = (((int64_t)b * g.D + z) * g.H + yy) * g.W + x;
                    xv[i] = c < g.C0 ? __ldg(reinterpret_cast<const float4*>(x0 + row * g.C0 + c))
                                     : __ldg(reinterpret_cast<const float4*>(x1 + row * g.C1 + (c - g.C0)));
                }
            }
#pragma unroll
            for (int i = 0; i < XL; ++i) {
                const int idx = tid + i * M_THREADS;
                if (idx < M_NH * 8)
                    *reinterpret_cast<float4*>(Xs + (idx >> 3) * XS + (idx & 7) * 4) =
                        make_float4(to_tf32(xv[i].x), to_tf32(xv[i].y), to_tf32(xv[i].z), to_tf32(xv[i].w));
            }
        }
        // dy brick (zero outside), Co padded to COP
        {
            constexpr int DL = (M_NB * COP + M_THREADS - 1) / M_THREADS;
            float dv[DL];
#pragma unroll
            for (int i = 0; i < DL; ++i) {
                const int idx = tid + i * M_THREADS;
                int pos, o;
                if (dy_ncdhw) { pos = idx % M_NB; o = idx / M_NB; } else { o = idx % COP; pos = idx / COP; }
                const int lx = pos % MB_X, ly = (pos / MB_X) % MB_Y, lz = pos / (MB_X * MB_Y);
                const int z = z0 + lz, yy = y0 + ly, x = x0c + lx;
                dv[i] = 0.f;
                if (idx < M_NB * COP && o < g.Co && z < g.D && yy < g.H && x < g.W) {
                    const int64_t sp = ((int64_t)z * g.H + yy) * g.W + x;
                    dv[i] = __ldg(dy_ncdhw ? dy + ((int64_t)b * g.Co + o) * S + sp : dy + ((int64_t)b * S + sp) * g.Co + o);
                }
            }
#pragma unroll
            for (int i = 0; i < DL; ++i) {
                const int idx = tid + i * M_THREADS;
                if (idx >= M_NB * COP) continue;          // whole warps drop out together (M_NB * COP % 32 == 0)
                int pos, o;
                if (dy_ncdhw) { pos = idx % M_NB; o = idx / M_NB; } else { o = idx % COP; pos = idx / COP; }
                dYs[pos * DS + o] = to_tf32(dv[i]);
                if (want_bias) bacc[i] += dv[i];          // (pos, o) of slot i is the same for every brick: reduce once at the end
            }
        }
        __syncthreads();
        // 16 k-steps: one x row (8 positions) of the brick each
#pragma unroll 2
        for (int ks = 0; ks < MB_Z * MB_Y; ++ks) {
            const int lz = ks / MB_Y, ly = ks % MB_Y;
            uint32_t bf[NTN][2];
            const float* db = dYs + (ks * 8 + tq) * DS + gq;
#pragma unroll
            for (int nt = 0; nt < NTN; ++nt) {
                bf[nt][0] = __float_as_uint(db[nt * 8]);
                bf[nt][1] = __float_as_uint(db[4 * DS + nt * 8]);
            }
            // x row of this warp's (tz, ty): halo positions hx = 0..9; thread needs hx = tq + {0,1,2,4,5,6}
            const float* xr = Xs + (((lz + tz) * MH_Y + (ly + ty)) * MH_X + tq) * XS + gq;
            uint32_t xa[2][2][6];            // [m-tile][row half g / g+8][position offset]
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const int off = j < 3 ? j : j + 1;      // 0,1,2,4,5,6
                        xa[mt][hh][j] = __float_as_uint(xr[off * XS + mt * 16 + hh * 8]);
                    }
#pragma unroll
            for (int tx = 0; tx < 3; ++tx)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NTN; ++nt)
                        mma_tf32(acc[tx][mt][nt], xa[mt][0][tx], xa[mt][1][tx], xa[mt][0][tx + 3], xa[mt][1][tx + 3], bf[nt][0],
                                 bf[nt][1]);
        }
    }
    // flush: c0:(row gq, col 2 tq) c1:(gq, 2tq+1) c2:(gq+8, 2tq) c3:(gq+8, 2tq+1); row = ci within the m-tile, col = co
    if (native) {
        // the Conv3d parameter's own layout [co][ci][tap]: stage the CTA's [COP][32 ci][27] block in shared memory (the x halo is
        // dead) so that the global atomics run over contiguous (ci, tap) spans instead of 27-float strides
        __syncthreads();
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
            const int tap = (tz * 3 + ty) * 3 + tx;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTN; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int cl = mt * 16 + gq + (r >> 1) * 8;
                        const int co = nt * 8 + 2 * tq + (r & 1);
                        Xs[(co * 32 + cl) * 27 + tap] = acc[tx][mt][nt][r];
                    }
        }
        __syncthreads();
        const int nci = min(32, Cin - c0);
        for (int idx = tid; idx < COP * 32 * 27; idx += M_THREADS) {
            const int co = idx / (32 * 27), rem = idx - co * (32 * 27);
            if (co < g.Co && rem < nci * 27) {
                const float v = Xs[idx];
                if (v != 0.f) atomicAdd(&dWt[((int64_t)co * Cin + c0) * 27 + rem], v);
            }
        }
    } else {
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
            const int tap = (tz * 3 + ty) * 3 + tx;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTN; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int ci = c0 + mt * 16 + gq + (r >> 1) * 8;
                        const int co = nt * 8 + 2 * tq + (r & 1);
                        if (ci < Cin && co < g.Co) atomicAdd(&dWt[((int64_t)tap * Cin + ci) * g.Co + co], acc[tx][mt][nt][r]);
                    }
        }
    }
    if (want_bias) {
        // lanes of a warp share o (NCDHW: 128 % 32 == 0) or hold o = lane % COP (channels-last)
#pragma unroll
        for (int i = 0; i < DLB; ++i) {
            const int idx = tid + i * M_THREADS;
            if (idx >= M_NB * COP) continue;
            const int o = dy_ncdhw ? idx / M_NB : idx % COP;
            float s = bacc[i];
            if (dy_ncdhw) {
                s = warp_sum(s);
                if (lane == 0 && o < g.Co) atomicAdd(&bsum[o], s);
            } else {
#pragma unroll
                for (int off = 16; off >= COP; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane < COP && o < g.Co) atomicAdd(&bsum[o], s);
            }
        }
        __syncthreads();
        if (tid < g.Co) atomicAdd(&dbias[tid], bsum[tid]);
    }
}

}  // namespace

int mma_conv3_bwd_weight(const float* dy, const float* x0, int C0, const float* x1, int C1, float* dWt, float* dbias,
                         int B, int D, int H, int W, int Co, int dy_ncdhw, int native, cudaStream_t st) {
    if ((Co != 8 && Co != 16) || (C0 & 3) || (C1 & 3)) return MIC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x0) & 15) || (x1 && (reinterpret_cast<uintptr_t>(x1) & 15))) return MIC_ERR_UNSUPPORTED;
    MGeom g{B, D, H, W, C0, C1, Co};
    const int nbz = ceil_div(D, MB_Z), nby = ceil_div(H, MB_Y), nbx = ceil_div(W, MB_X);
    const int64_t nbricks = (int64_t)B * nbz * nby * nbx;
    const int chunks = ceil_div(C0 + C1, 32);
    int64_t gx = ceil_div64((int64_t)num_sms() * 2, chunks);
    if (gx > nbricks) gx = nbricks;
    if (gx < 1) gx = 1;
    // even out the bricks per CTA
    const int64_t per = ceil_div64(nbricks, gx);
    gx = ceil_div64(nbricks, per);
    dim3 grid((unsigned)gx, chunks);
    const size_t smem = (M_NH * XS + M_NB * DS + 16) * sizeof(float);
    if (Co == 8) {
        static bool once = false;
        if (!once) { cudaFuncSetAttribute(conv3_mma_bwd_weight_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); once = true; }
        mic::launch((conv3_mma_bwd_weight_kernel<8>), grid, dim3(M_THREADS), smem, st, dy, x0, x1, dWt, dbias, g, dy_ncdhw, nbz, nby, nbx, native);
    } else {
        static bool once = false;
        if (!once) { cudaFuncSetAttribute(conv3_mma_bwd_weight_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); once = true; }
        mic::launch((conv3_mma_bwd_weight_kernel<16>), grid, dim3(M_THREADS), smem, st, dy, x0, x1, dWt, dbias, g, dy_ncdhw, nbz, nby, nbx, native);
    }
    return check_launch("conv3_mma_bwd_weight_kernel");
}

}  // namespace mic

extern "C" int mic_conv3_mma_bwd_weight(const float* dy, const float* x0, int C0, const float* x1, int C1, float* dWt,
                                        float* dbias, int B, int D, int H, int W, int Co, int dy_ncdhw, int native_layout,
                                        void* stream) {
    MIC_REQUIRE(dy && x0 && dWt && (C1 == 0 || x1), "conv3_mma_bwd_weight: null pointer");
    int rc = mic::mma_conv3_bwd_weight(dy, x0, C0, x1, C1, dWt, dbias, B, D, H, W, Co, dy_ncdhw, native_layout != 0,
                                       (cudaStream_t)stream);
    if (rc == MIC_ERR_UNSUPPORTED) return mic::fail(MIC_ERR_UNSUPPORTED, "conv3_mma_bwd_weight: Co=%d / channel split not taken", Co);
    return rc;
}
