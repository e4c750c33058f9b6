// 3x3x3 stride-1 zero-padded convolution on channels-last token grids, CUDA-core fp32 implicit GEMM
// (no im2col buffer): conv_offset[0] (2C -> 16 on cat[LN(x), xa], read from the two tensors in place) and
// out_conv (E/2 -> num_classes, NCDHW logits).  Inputs live on (D,H,W) and read as zero outside it; outputs are
// produced on the window-padded grid (Dp,Hp,Wp).
//   fwd      : M = output positions (tile 128/256), N = Co (<=16), K = 27 * Cin in chunks of 32 channels
//   bwd_data : M = input positions (tile 128), N = 32 input channels, K = 27 * Co
//   bwd_wgt  : per CTA a 4x4x8 brick of positions with its halo in shared memory, 27x4 register accumulators
//              per thread over a strided set of bricks, one atomic flush at the end.
#include "common.cuh"

namespace mic {

struct ConvGeom {
    int B, D, H, W, Dp, Hp, Wp, C0, C1, Co;
};

__device__ __forceinline__ float4 ld_cat4(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                                          int64_t row, int c) {
    // c multiple of 4; C0 multiple of 4
    if (c < C0) return *reinterpret_cast<const float4*>(x0 + row * C0 + c);
    if (c - C0 < C1) return *reinterpret_cast<const float4*>(x1 + row * C1 + (c - C0));
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------ forward
template <int COP>
__global__ void __launch_bounds__(128) conv3_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                        const float* __restrict__ Wt, const float* __restrict__ bias,
                                                        float* __restrict__ y, ConvGeom g, int out_ncdhw,
                                                        int units_per_split) {
    pdl_sync();
    constexpr int OG = COP / 4;          // output groups of 4
    constexpr int PG = 128 / OG;         // position groups
    constexpr int TP = PG * 4;           // positions per CTA
    constexpr int KC = 32;
    __shared__ __align__(16) float As[KC][TP + 4];
    __shared__ __align__(16) float Ws[KC][COP];
    __shared__ int pcoord[TP];           // packed (b,z,y,x) or -1

    const int tid = threadIdx.x;
    const int64_t P = (int64_t)g.B * g.Dp * g.Hp * g.Wp;
    const int64_t p0 = (int64_t)blockIdx.x * TP;
    const int Cin = g.C0 + g.C1;
    for (int i = tid; i < TP; i += 128) {
        int64_t p = p0 + i;
        int code = -1;
        if (p < P) {
            const int x = (int)(p % g.Wp); p /= g.Wp;
            const int yy = (int)(p % g.Hp); p /= g.Hp;
            const int z = (int)(p % g.Dp); p /= g.Dp;
            code = ((int)p << 24) | (z << 16) | (yy << 8) | x;    // B<128, dims<256 (checked on host)
        }
        pcoord[i] = code;
    }
    __syncthreads();
    const int pg = tid / OG, og = tid % OG;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    // K is split over blockIdx.y in units of (tap, 32-channel chunk); partial sums are atomically accumulated
    const int nchunks = (Cin + KC - 1) / KC;
    const int unit_begin = blockIdx.y * units_per_split;
    const int unit_end = min(27 * nchunks, unit_begin + units_per_split);
    const bool split = gridDim.y > 1;
    for (int unit = unit_begin; unit < unit_end; ++unit) {
        const int tap = unit / nchunks;
        const int tz = tap / 9 - 1, ty = (tap / 3) % 3 - 1, tx = tap % 3 - 1;
        {
            const int c0 = (unit % nchunks) * KC;
            // gather A: TP positions x 8 float4
            for (int idx = tid; idx < TP * 8; idx += 128) {
                const int pos = idx >> 3, c4 = (idx & 7) * 4;
                const int code = pcoord[pos];
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (code >= 0) {
                    const int b = code >> 24, z = ((code >> 16) & 255) + tz, yy = ((code >> 8) & 255) + ty,
                              x = (code & 255) + tx;
                    if (z >= 0 && z < g.D && yy >= 0 && yy < g.H && x >= 0 && x < g.W && c0 + c4 < Cin) {
                        const int64_t row = (((int64_t)b * g.D + z) * g.H + yy) * g.W + x;
                        v = ld_cat4(x0, g.C0, x1, g.C1, row, c0 + c4);
                    }
                }
                As[c4 + 0][pos] = v.x; As[c4 + 1][pos] = v.y; As[c4 + 2][pos] = v.z; As[c4 + 3][pos] = v.w;
            }
            for (int idx = tid; idx < KC * COP; idx += 128) {
                const int c = idx / COP, o = idx % COP;
                Ws[c][o] = (c0 + c < Cin && o < g.Co) ? Wt[((int64_t)tap * Cin + c0 + c) * g.Co + o] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&As[kk][pg * 4]);
                const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][og * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    const int64_t S = (int64_t)g.Dp * g.Hp * g.Wp;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t p = p0 + pg * 4 + i;
        if (p >= P) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = og * 4 + j;
            if (o >= g.Co) continue;
            const float v = acc[i][j] + ((bias && blockIdx.y == 0) ? bias[o] : 0.f);
            float* dst;
            if (out_ncdhw) {
                const int64_t b = p / S;
                dst = y + (b * g.Co + o) * S + (p - b * S);
            } else {
                dst = y + p * g.Co + o;
            }
            if (split) atomicAdd(dst, v); else *dst = v;
        }
    }
}

// ------------------------------------------------------------------------------------------- backward data
template <int COP>
__global__ void __launch_bounds__(256) conv3_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ Wt,
                                                             float* __restrict__ dx0, int acc0,
                                                             float* __restrict__ dx1, int acc1, ConvGeom g,
                                                             int dy_ncdhw) {
    pdl_sync();
    constexpr int TP = 128, TC = 32;
    __shared__ __align__(16) float As[COP][TP + 4];   // dy^T : [o][pos]
    __shared__ __align__(16) float Ws[COP][TC + 4];   // [o][c]
    __shared__ int pcoord[TP];
    const int tid = threadIdx.x;
    const int64_t Q = (int64_t)g.B * g.D * g.H * g.W;
    const int64_t q0 = (int64_t)blockIdx.x * TP;
    const int c0 = blockIdx.y * TC;
    const int Cin = g.C0 + g.C1;
    const int64_t S = (int64_t)g.Dp * g.Hp * g.Wp;
    for (int i = tid; i < TP; i += 256) {
        int64_t q = q0 + i;
        int code = -1;
        if (q < Q) {
            const int x = (int)(q % g.W); q /= g.W;
            const int yy = (int)(q % g.H); q /= g.H;
            const int z = (int)(q % g.D); q /= g.D;
            code = ((int)q << 24) | (z << 16) | (yy << 8) | x;
        }
        pcoord[i] = code;
    }
    __syncthreads();
    const int pg = tid >> 3, cg = tid & 7;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int tap = 0; tap < 27; ++tap) {
        // dx[q] += dy[q - t + 1] * W[t]  ->  source position q + (1 - t)
        const int tz = 1 - tap / 9, ty = 1 - (tap / 3) % 3, tx = 1 - tap % 3;
        for (int idx = tid; idx < TP * COP; idx += 256) {
            int pos, o;
            if (dy_ncdhw) { pos = idx % TP; o = idx / TP; } else { o = idx % COP; pos = idx / COP; }
            const int code = pcoord[pos];
            float v = 0.f;
            if (code >= 0 && o < g.Co) {
                const int b = code >> 24, z = ((code >> 16) & 255) + tz, yy = ((code >> 8) & 255) + ty,
                          x = (code & 255) + tx;
                if (z >= 0 && z < g.Dp && yy >= 0 && yy < g.Hp && x >= 0 && x < g.Wp) {
                    const int64_t sp = ((int64_t)z * g.Hp + yy) * g.Wp + x;
                    v = dy_ncdhw ? dy[((int64_t)b * g.Co + o) * S + sp] : dy[((int64_t)b * S + sp) * g.Co + o];
                }
            }
            As[o][pos] = v;
        }
        for (int idx = tid; idx < TC * COP; idx += 256) {
            const int c = idx / COP, o = idx % COP;
            Ws[o][c] = (c0 + c < Cin && o < g.Co) ? Wt[((int64_t)tap * Cin + c0 + c) * g.Co + o] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < COP; ++o) {
            const float4 a = *reinterpret_cast<const float4*>(&As[o][pg * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[o][cg * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t q = q0 + pg * 4 + i;
        if (q >= Q) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + cg * 4 + j;
            if (c >= Cin) continue;
            float* dst;
            int accf;
            if (c < g.C0) { dst = dx0 + q * g.C0 + c; accf = acc0; }
            else { dst = dx1 + q * g.C1 + (c - g.C0); accf = acc1; }
            if (accf) *dst += acc[i][j]; else *dst = acc[i][j];
        }
    }
}

// ----------------------------------------------------------------------------------------- backward weight
constexpr int BZ = 4, BY = 4, BX = 8;
constexpr int HZ = BZ + 2, HY = BY + 2, HX = BX + 2;
constexpr int NB = BZ * BY * BX;      // 128 positions per brick
constexpr int NH = HZ * HY * HX;      // 360 halo positions

template <int COP>
__global__ void __launch_bounds__(32 * (COP / 4)) conv3_bwd_weight_kernel(const float* __restrict__ dy,
                                                                         const float* __restrict__ x0,
                                                                         const float* __restrict__ x1,
                                                                         float* __restrict__ dWt,
                                                                         float* __restrict__ dbias, ConvGeom g,
                                                                         int dy_ncdhw, int nbz, int nby, int nbx) {
    pdl_sync();
    constexpr int NT = 32 * (COP / 4);
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;                   // [NH][33]
    float* dYs = smem + NH * 33;        // [NB][COP]
    const int tid = threadIdx.x;
    const int c = tid & 31, og = tid >> 5;
    const int c0 = blockIdx.y * 32;
    const int Cin = g.C0 + g.C1;
    const int64_t S = (int64_t)g.Dp * g.Hp * g.Wp;
    const int64_t nbricks = (int64_t)g.B * nbz * nby * nbx;
    float acc[27][4];
#pragma unroll
    for (int t = 0; t < 27; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};

    for (int64_t brick = blockIdx.x; brick < nbricks; brick += gridDim.x) {
        int64_t t = brick;
        const int bx = (int)(t % nbx); t /= nbx;
        const int by = (int)(t % nby); t /= nby;
        const int bz = (int)(t % nbz); t /= nbz;
        const int b = (int)t;
        const int z0 = bz * BZ, y0 = by * BY, x0c = bx * BX;
        // halo of the input (zero outside (D,H,W)), 32 channels
        for (int idx = tid; idx < NH * 8; idx += NT) {
            const int hp = idx >> 3, c4 = (idx & 7) * 4;
            const int hx = hp % HX, hy = (hp / HX) % HY, hz = hp / (HX * HY);
            const int z = z0 + hz - 1, yy = y0 + hy - 1, x = x0c + hx - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (z >= 0 && z < g.D && yy >= 0 && yy < g.H && x >= 0 && x < g.W && c0 + c4 < Cin) {
                const int64_t row = (((int64_t)b * g.D + z) * g.H + yy) * g.W + x;
                v = ld_cat4(x0, g.C0, x1, g.C1, row, c0 + c4);
            }
            float* d = Xs + hp * 33 + c4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        for (int idx = tid; idx < NB * COP; idx += NT) {
            int pos, o;
            if (dy_ncdhw) { pos = idx % NB; o = idx / NB; } else { o = idx % COP; pos = idx / COP; }
            const int lx = pos % BX, ly = (pos / BX) % BY, lz = pos / (BX * BY);
            const int z = z0 + lz, yy = y0 + ly, x = x0c + lx;
            float v = 0.f;
            if (o < g.Co && z < g.Dp && yy < g.Hp && x < g.Wp) {
                const int64_t sp = ((int64_t)z * g.Hp + yy) * g.Wp + x;
                v = dy_ncdhw ? dy[((int64_t)b * g.Co + o) * S + sp] : dy[((int64_t)b * S + sp) * g.Co + o];
            }
            dYs[pos * COP + o] = v;
        }
        __syncthreads();
        for (int pos = 0; pos < NB; ++pos) {
            const float4 d4 = *reinterpret_cast<const float4*>(dYs + pos * COP + og * 4);
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
            const int lx = pos % BX, ly = (pos / BX) % BY, lz = pos / (BX * BY);
            const float* xb = Xs + ((lz * HY + ly) * HX + lx) * 33 + c;
#pragma unroll
            for (int tz = 0; tz < 3; ++tz)
#pragma unroll
                for (int ty = 0; ty < 3; ++ty)
#pragma unroll
                    for (int tx = 0; tx < 3; ++tx) {
                        const float xv = xb[((tz * HY + ty) * HX + tx) * 33];
                        const int tt = (tz * 3 + ty) * 3 + tx;
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[tt][j] = fmaf(xv, dv[j], acc[tt][j]);
                    }
            if (c == 0 && blockIdx.y == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) bsum[j] += dv[j];
            }
        }
        __syncthreads();
    }
    if (c0 + c < Cin) {
#pragma unroll
        for (int tt = 0; tt < 27; ++tt)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = og * 4 + j;
                if (o < g.Co) atomicAdd(&dWt[((int64_t)tt * Cin + c0 + c) * g.Co + o], acc[tt][j]);
            }
    }
    if (dbias && c == 0 && blockIdx.y == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = og * 4 + j;
            if (o < g.Co) atomicAdd(&dbias[o], bsum[j]);
        }
    }
}

static int check_geom(const char* who, int C0, int C1, int B, int D, int H, int W, int Dp, int Hp, int Wp, int Co) {
    if (C0 <= 0 || C1 < 0 || (C0 & 3) || (C1 & 3)) return fail(MIC_ERR_INVALID, "%s: channel counts must be multiples of 4 (%d|%d)", who, C0, C1);
    if (Co <= 0 || Co > 16) return fail(MIC_ERR_UNSUPPORTED, "%s: Co=%d not in 1..16", who, Co);
    if (B <= 0 || B >= 128 || Dp > 255 || Hp > 255 || Wp > 255 || Dp < D || Hp < H || Wp < W || D <= 0 || H <= 0 || W <= 0)
        return fail(MIC_ERR_UNSUPPORTED, "%s: geometry B=%d (%d,%d,%d)->(%d,%d,%d) outside supported range", who, B, D, H, W, Dp, Hp, Wp);
    return MIC_OK;
}

}  // namespace mic

using namespace mic;

extern "C" int mic_conv3_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wt, const float* bias,
                             float* y, int B, int D, int H, int W, int Dp, int Hp, int Wp, int Co, int out_ncdhw,
                             void* stream) {
    MIC_REQUIRE(x0 && Wt && y && (C1 == 0 || x1), "conv3_fwd: null pointer");
    int rc = check_geom("conv3_fwd", C0, C1, B, D, H, W, Dp, Hp, Wp, Co);
    if (rc) return rc;
    ConvGeom g{B, D, H, W, Dp, Hp, Wp, C0, C1, Co};
    const int64_t P = (int64_t)B * Dp * Hp * Wp;
    cudaStream_t st = (cudaStream_t)stream;
    const int TP = Co <= 8 ? 256 : 128;
    const int64_t tiles = ceil_div64(P, TP);
    const int units = 27 * ceil_div(C0 + C1, 32);
    int splits = (int)ceil_div64((int64_t)num_sms() * 3, tiles);
    if (splits > units / 2) splits = units / 2;
    if (splits < 1) splits = 1;
    const int ups = ceil_div(units, splits);
    splits = ceil_div(units, ups);
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(y, 0, (size_t)P * Co * sizeof(float), st);
        if (e != cudaSuccess) return fail(MIC_ERR_CUDA, "conv3_fwd memset: %s", cudaGetErrorString(e));
    }
    dim3 grid((unsigned)tiles, splits);
    if (Co <= 8) mic::launch((conv3_fwd_kernel<8>), grid, dim3(128), 0, st, x0, x1, Wt, bias, y, g, out_ncdhw, ups);
    else mic::launch((conv3_fwd_kernel<16>), grid, dim3(128), 0, st, x0, x1, Wt, bias, y, g, out_ncdhw, ups);
    return check_launch("conv3_fwd_kernel");
}

extern "C" int mic_conv3_bwd_data(const float* dy, const float* Wt, float* dx0, int C0, int acc0, float* dx1, int C1,
                                  int acc1, int B, int D, int H, int W, int Dp, int Hp, int Wp, int Co, int dy_ncdhw,
                                  void* stream) {
    MIC_REQUIRE(dy && Wt && dx0 && (C1 == 0 || dx1), "conv3_bwd_data: null pointer");
    int rc = check_geom("conv3_bwd_data", C0, C1, B, D, H, W, Dp, Hp, Wp, Co);
    if (rc) return rc;
    ConvGeom g{B, D, H, W, Dp, Hp, Wp, C0, C1, Co};
    const int64_t Q = (int64_t)B * D * H * W;
    dim3 grid((unsigned)ceil_div64(Q, 128), ceil_div(C0 + C1, 32));
    cudaStream_t st = (cudaStream_t)stream;
    if (Co <= 8) mic::launch((conv3_bwd_data_kernel<8>), grid, dim3(256), 0, st, dy, Wt, dx0, acc0, dx1, acc1, g, dy_ncdhw);
    else mic::launch((conv3_bwd_data_kernel<16>), grid, dim3(256), 0, st, dy, Wt, dx0, acc0, dx1, acc1, g, dy_ncdhw);
    return check_launch("conv3_bwd_data_kernel");
}

extern "C" int mic_conv3_bwd_weight(const float* dy, const float* x0, int C0, const float* x1, int C1, float* dWt,
                                    float* dbias, int B, int D, int H, int W, int Dp, int Hp, int Wp, int Co,
                                    int dy_ncdhw, void* stream) {
    MIC_REQUIRE(dy && x0 && dWt && (C1 == 0 || x1), "conv3_bwd_weight: null pointer");
    int rc = check_geom("conv3_bwd_weight", C0, C1, B, D, H, W, Dp, Hp, Wp, Co);
    if (rc) return rc;
    ConvGeom g{B, D, H, W, Dp, Hp, Wp, C0, C1, Co};
    const int nbz = ceil_div(Dp, BZ), nby = ceil_div(Hp, BY), nbx = ceil_div(Wp, BX);
    const int64_t nbricks = (int64_t)B * nbz * nby * nbx;
    const int chunks = ceil_div(C0 + C1, 32);
    int64_t gx = ceil_div64((int64_t)num_sms() * 4, chunks);
    if (gx > nbricks) gx = nbricks;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, chunks);
    cudaStream_t st = (cudaStream_t)stream;
    if (Co <= 8) {
        const size_t smem = (NH * 33 + NB * 8) * sizeof(float);
        static bool once8 = false;
        if (!once8) { cudaFuncSetAttribute(conv3_bwd_weight_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); once8 = true; }
        mic::launch((conv3_bwd_weight_kernel<8>), grid, dim3(64), smem, st, dy, x0, x1, dWt, dbias, g, dy_ncdhw, nbz, nby, nbx);
    } else {
        const size_t smem = (NH * 33 + NB * 16) * sizeof(float);
        static bool once16 = false;
        if (!once16) { cudaFuncSetAttribute(conv3_bwd_weight_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); once16 = true; }
        mic::launch((conv3_bwd_weight_kernel<16>), grid, dim3(128), smem, st, dy, x0, x1, dWt, dbias, g, dy_ncdhw, nbz, nby, nbx);
    }
    return check_launch("conv3_bwd_weight_kernel");
}
