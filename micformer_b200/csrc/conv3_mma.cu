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
#include <stdlib.h>
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


// ---------------------------------------------------------------------------------------------------------------
// Co = 8 variant (out_conv: 24 -> 8 classes on the 128^3 grid, the single largest weight-gradient launch of the step).
// Legacy mma.sync issues one m16n8k8 TF32 per 2.2 clocks per SM here (scripts/ubench/mma_sync_rate.cu: 272 TFLOP/s), so
// the 864 MMAs of a brick are ~1.9k clocks -- the kernel above spends ~9.4k per brick on dependent shared-memory loads
// (one warp per (tz, ty) re-reads every x row for each of its k-steps, 4.3 LDS per MMA, nothing in flight across the two
// barriers).  Here a warp owns ONE tx shift and a quarter of the 36 halo x-rows and applies each row's A fragments (8 LDS,
// rounded once) to ALL (tz, ty) taps whose output row (hz - tz, hy - ty) lies inside the brick; all fragment loads of a
// row are issued before its first MMA.  Bricks are double-buffered in shared memory with cp.async (zero-fill outside the
// volume), so staging overlaps the MMAs of the previous brick in a single CTA per SM.
constexpr int M8_THREADS = 384;                  // 12 warps: tx = warp % 3, row group = warp / 3 (rows rg, rg + 4, ..)
constexpr int PS8 = 132;                         // NCDHW dy tile: [o][pos] row stride; bank = 4 gq + tq -> conflict-free
constexpr int M8_DY = 128 * DS;                  // floats reserved for the dy tile (>= 8 * PS8)
constexpr int M8_STAGE = M_NH * XS + M8_DY;      // floats per stage

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;               // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ uint32_t rn_tf32_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// The nine halo rows RG, RG + 4, .. of one warp, fully unrolled: (hz, hy), the valid (tz, ty) taps of each row and every
// shared-memory offset are compile-time constants (the first version of this kernel computed them at run time and was
// instruction-issue bound: ~350 instructions per row for 18 MMAs).  xt / dt already carry the thread's fragment offsets.
// MT 16-channel tiles x NT 8-output tiles per tap, MT * NT = 2: (2, 1) for Co <= 8, (1, 2) for Co <= 16.
template <int RG, bool NCDHW, int MT, int NT>
__device__ __forceinline__ void bw8_rows(const float* __restrict__ xt, const float* __restrict__ dt, float (&acc)[3][3][MT][NT][4]) {
    constexpr int XSx = MT == 2 ? XS : 24;       // x row stride: 8 tq + gq / 24 tq + gq are both conflict-free
    constexpr int sp = NCDHW ? 1 : DS, so = NCDHW ? PS8 : 1;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int ir = RG + 4 * i;
        const int hz = ir / MH_Y, hy = ir % MH_Y;
        const float* xr = xt + (hz * MH_Y + hy) * MH_X * XSx;
        uint32_t xa[MT][2][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                xa[mt][hh][0] = rn_tf32_bits(xr[mt * 16 + hh * 8]);
                xa[mt][hh][1] = rn_tf32_bits(xr[4 * XSx + mt * 16 + hh * 8]);
            }
#pragma unroll
        for (int tz = 0; tz < 3; ++tz)
#pragma unroll
            for (int ty = 0; ty < 3; ++ty) {
                const int lz = hz - tz, ly = hy - ty;
                if (lz >= 0 && lz < MB_Z && ly >= 0 && ly < MB_Y) {
                    const float* db = dt + (lz * MB_Y + ly) * 8 * sp;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const uint32_t b0 = rn_tf32_bits(db[nt * 8 * so]), b1 = rn_tf32_bits(db[4 * sp + nt * 8 * so]);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
                            mma_tf32(acc[tz][ty][mt][nt], xa[mt][0][0], xa[mt][1][0], xa[mt][0][1], xa[mt][1][1], b0, b1);
                    }
                }
            }
    }
}

template <bool NCDHW, int MT, int NT>
__global__ void __launch_bounds__(M8_THREADS, 1)
conv3_mma_bwd_weight8_kernel(const float* __restrict__ dy, const float* __restrict__ x0, const float* __restrict__ x1,
                             float* __restrict__ dWt, float* __restrict__ dbias, MGeom g, int nbz, int nby, int nbx,
                             int native) {
    pdl_sync();
    constexpr int CH = 16 * MT, COP = 8 * NT;    // input channels per CTA chunk, padded output channels
    constexpr int XSx = MT == 2 ? XS : 24;
    constexpr int NB = (128 * COP + M8_THREADS - 1) / M8_THREADS;
    extern __shared__ __align__(16) float msm[];
    float* bsum = msm + 2 * M8_STAGE;            // [16]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int c0 = blockIdx.y * CH;
    const int Cin = g.C0 + g.C1;
    const uint32_t S = (uint32_t)g.D * g.H * g.W;            // host checks B * S * max(C) < 2^31
    const uint32_t nbricks = (uint32_t)g.B * nbz * nby * nbx;
    const int tx = warp % 3, rg = warp / 3;
    const bool want_bias = dbias != nullptr && blockIdx.y == 0;
    // thread offsets of the fragment elements: A (row = ci, col = position), B (k = position, n = co)
    const int xoff = (tq + tx) * XSx + gq;
    const int doff = NCDHW ? tq + gq * PS8 : tq * DS + gq;

    float acc[3][3][MT][NT][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < MT * NT; ++c)
#pragma unroll
                for (int d = 0; d < 4; ++d) (&acc[a][b][0][0][0])[c * 4 + d] = 0.f;
    if (tid < 16) bsum[tid] = 0.f;
    float bacc[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) bacc[i] = 0.f;

    auto stage = [&](uint32_t brick, int buf) {
        uint32_t t = brick;
        const int bx = (int)(t % (uint32_t)nbx); t /= (uint32_t)nbx;
        const int by = (int)(t % (uint32_t)nby); t /= (uint32_t)nby;
        const int bz = (int)(t % (uint32_t)nbz);
        const uint32_t b = t / (uint32_t)nbz;
        const int z0 = bz * MB_Z, y0 = by * MB_Y, x0c = bx * MB_X;
        const uint32_t pb = b * S;                            // first position of this batch element
        float* Xs = msm + buf * M8_STAGE;
        float* dYs = Xs + M_NH * XS;
        constexpr int PCS = CH / 4;                           // 16-byte pieces per position
#pragma unroll
        for (int i = 0; i < (M_NH * PCS + M8_THREADS - 1) / M8_THREADS; ++i) {
            const int idx = tid + i * M8_THREADS;
            if (idx < M_NH * PCS) {
                const int hp = idx / PCS, c4 = (idx % PCS) * 4;
                const int hx = hp % MH_X, hyz = hp / MH_X, hy = hyz % MH_Y, hz = hyz / MH_Y;
                const int z = z0 + hz - 1, yy = y0 + hy - 1, x = x0c + hx - 1;
                const int c = c0 + c4;
                const bool ok = (unsigned)z < (unsigned)g.D && (unsigned)yy < (unsigned)g.H && (unsigned)x < (unsigned)g.W && c < Cin;
                const uint32_t row = pb + ((uint32_t)z * g.H + yy) * g.W + x;
                const float* src = !ok ? x0 : (c < g.C0 ? x0 + (size_t)(row * (uint32_t)g.C0 + c) : x1 + (size_t)(row * (uint32_t)g.C1 + (c - g.C0)));
                cp_async16(Xs + hp * XSx + c4, src, ok);
            }
        }
        if (NCDHW) {
            // [o][lz][ly][8 x] <- dy[b][o][z][y][x0c .. x0c+7]: two 16-byte pieces per (o, row)
            for (int idx = tid; idx < COP * 32; idx += M8_THREADS) {
                const int pc = idx & 1, rr = (idx >> 1) & 15, o = idx >> 5;
                const int lz = rr >> 2, ly = rr & 3;
                const int z = z0 + lz, yy = y0 + ly, x = x0c + pc * 4;
                const bool ok = o < g.Co && z < g.D && yy < g.H && x < g.W;       // W % 4 == 0: a piece is all in or all out
                const float* src = ok ? dy + (size_t)((b * (uint32_t)g.Co + o) * S + ((uint32_t)z * g.H + yy) * g.W + x) : dy;
                cp_async16(dYs + o * PS8 + rr * 8 + pc * 4, src, ok);
            }
        } else {
            // [pos][COP co] <- dy[b][z][y][x][0 .. COP-1]   (Co == COP)
            constexpr int PD = COP / 4;
            for (int idx = tid; idx < 128 * PD; idx += M8_THREADS) {
                const int pc = idx % PD, pos = idx / PD;
                const int lx = pos % MB_X, ly = (pos / MB_X) % MB_Y, lz = pos / (MB_X * MB_Y);
                const int z = z0 + lz, yy = y0 + ly, x = x0c + lx;
                const bool ok = z < g.D && yy < g.H && x < g.W;
                const float* src = ok ? dy + (size_t)((pb + ((uint32_t)z * g.H + yy) * g.W + x) * (uint32_t)COP + pc * 4) : dy;
                cp_async16(dYs + pos * DS + pc * 4, src, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int it = 0;
    if (blockIdx.x < nbricks) stage(blockIdx.x, 0);
    for (uint32_t brick = blockIdx.x; brick < nbricks; brick += gridDim.x, ++it) {
        const int buf = it & 1;
        const bool more = brick + gridDim.x < nbricks;
        if (more) stage(brick + gridDim.x, buf ^ 1);       // the other buffer was released by the sync that ended brick it-1
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float* Xs = msm + buf * M8_STAGE;
        const float* dYs = Xs + M_NH * XS;
        if (want_bias) {
            // element idx of the COP x 128 dy tile: co = idx / 128 (NCDHW tile) or idx % COP (channels-last tile)
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int idx = tid + i * M8_THREADS;
                if (idx < 128 * COP) bacc[i] += NCDHW ? dYs[(idx >> 7) * PS8 + (idx & 127)] : dYs[(idx / COP) * DS + (idx % COP)];
            }
        }
        const float* xt = Xs + xoff;
        const float* dt = dYs + doff;
        if (rg == 0) bw8_rows<0, NCDHW, MT, NT>(xt, dt, acc);       // warp-uniform
        else if (rg == 1) bw8_rows<1, NCDHW, MT, NT>(xt, dt, acc);
        else if (rg == 2) bw8_rows<2, NCDHW, MT, NT>(xt, dt, acc);
        else bw8_rows<3, NCDHW, MT, NT>(xt, dt, acc);
        __syncthreads();                                      // buffer `buf` may be restaged by the next iteration's prefetch
    }
    // ---- reduce the four row groups in shared memory ([27][CH ci][COP co], stage 0 is dead), then one atomic pass
    float* Wsm = msm;
    for (int r = 0; r < 4; ++r) {
        if (rg == r) {
#pragma unroll
            for (int tz = 0; tz < 3; ++tz)
#pragma unroll
                for (int ty = 0; ty < 3; ++ty) {
                    const int tap = (tz * 3 + ty) * 3 + tx;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int cl = mt * 16 + gq + (q >> 1) * 8;
                                const int co = nt * 8 + 2 * tq + (q & 1);
                                float* w = Wsm + (tap * CH + cl) * COP + co;
                                *w = r == 0 ? acc[tz][ty][mt][nt][q] : *w + acc[tz][ty][mt][nt][q];
                            }
                }
        }
        __syncthreads();
    }
    const int nci = min(CH, Cin - c0);
    for (int idx = tid; idx < 27 * CH * COP; idx += M8_THREADS) {
        const int co = idx % COP, cl = (idx / COP) % CH, tap = idx / (COP * CH);
        const float v = Wsm[idx];
        if (cl < nci && co < g.Co && v != 0.f) {
            if (native) atomicAdd(&dWt[((int64_t)co * Cin + c0 + cl) * 27 + tap], v);
            else atomicAdd(&dWt[((int64_t)tap * Cin + c0 + cl) * g.Co + co], v);
        }
    }
    if (want_bias) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int idx = tid + i * M8_THREADS;
            if (idx < 128 * COP) {
                const int o = NCDHW ? idx >> 7 : idx % COP;
                if (o < g.Co) atomicAdd(&bsum[o], bacc[i]);
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
    // Co = 8 with 16-byte-addressable dy rows: the row-reuse kernel (one persistent CTA per SM and channel chunk)
    static const bool old8 = []() { const char* e = getenv("MICFORMER_CONV_BW8_OLD"); return e && e[0] == '1'; }();
    // the row-reuse kernel: dy rows must be 16-byte addressable (channels-last: Co == 8 or 16 exactly)
    const int cmax = C0 > C1 ? C0 : C1;
    // Co = 16 (conv_offset): measured no faster than the kernel above (stage 0: 102 vs 82 us, step time equal, gpurun r2bk) --
    // six 16-channel chunk CTAs re-stage the dy tile -- so it is opt-in there (MICFORMER_CONV_BW16_NEW=1)
    static const bool old16 = []() { const char* e = getenv("MICFORMER_CONV_BW16_NEW"); return !(e && e[0] == '1'); }();
    const bool dy_ok = (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (dy_ncdhw ? (W & 3) == 0 : true);
    if (dy_ok && !(Co == 8 ? old8 : old16) && (int64_t)B * D * H * W * (cmax > 16 ? cmax : 16) < ((int64_t)1 << 31) &&
        nbricks < ((int64_t)1 << 31)) {
        const int ch = Co == 8 ? 32 : 16;
        const int chunks8 = ceil_div(C0 + C1, ch);
        int64_t g8 = ceil_div64((int64_t)num_sms(), chunks8);
        if (g8 > nbricks) g8 = nbricks;
        if (g8 < 1) g8 = 1;
        const int64_t per8 = ceil_div64(nbricks, g8);
        g8 = ceil_div64(nbricks, per8);
        const size_t smem8 = (2 * M8_STAGE + 16) * sizeof(float);
        static bool once8 = false;
        if (!once8) {
            cudaFuncSetAttribute(conv3_mma_bwd_weight8_kernel<true, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8);
            cudaFuncSetAttribute(conv3_mma_bwd_weight8_kernel<false, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8);
            cudaFuncSetAttribute(conv3_mma_bwd_weight8_kernel<true, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8);
            cudaFuncSetAttribute(conv3_mma_bwd_weight8_kernel<false, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8);
            once8 = true;
        }
        const dim3 grid8((unsigned)g8, chunks8);
#define LAUNCH8(NC, MT_, NT_) mic::launch(conv3_mma_bwd_weight8_kernel<NC, MT_, NT_>, grid8, dim3(M8_THREADS), smem8, st, dy, x0, x1, dWt, \
                                         dbias, g, nbz, nby, nbx, native)
        if (Co == 8) { if (dy_ncdhw) LAUNCH8(true, 2, 1); else LAUNCH8(false, 2, 1); }
        else { if (dy_ncdhw) LAUNCH8(true, 1, 2); else LAUNCH8(false, 1, 2); }
#undef LAUNCH8
        return check_launch("conv3_mma_bwd_weight8_kernel");
    }
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
