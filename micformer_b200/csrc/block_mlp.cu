// Fused MLP half of a transformer block for small channel counts (the train config's stage 0, C = 48):
//     y = x + rowscale * ( fc2( GELU( fc1( LayerNorm(x) ) ) ) )            reference M:28-34, 403-404, 419-424
// as ONE persistent tcgen05 kernel per direction.  Forward reads x and writes y; the 4C-wide hidden activation lives only
// in TMEM / shared memory.  Backward (block_mlp_bwd.cu) recomputes it from x.
//
// Roles (320 threads): warp 0 = weight loader (bulk copies of pre-swizzled bf16 weight images, once per CTA), warp 1 =
// MMA issuer (one elected lane), warps 2..9 = row threads: TMEM lane quarter = warp & 3, thread = tile row, the two
// warps of a quarter split the columns.  GEMM operands use the split-bf16 scheme of tc5.cuh.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"

namespace mic {
using namespace t5;

// ------------------------------------------------------------------------------------------------ weight images
// A weight image is the exact shared-memory picture of a B operand (N rows x K reduction elements, K-major, bf16,
// SWIZZLE_128B): panels of 64 k, each n_pad rows x 128 B, 16-byte chunks XOR-swizzled by (n & 7); rows >= N and k >= K are
// zero.  hi and lo images of the split-bf16 scheme are produced together.  One launch converts all the weights of a model.
struct ImgJob {
    const float* src;            // row-major fp32 weight, leading dimension ld
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    int64_t ld, N, K, transpose; // B[n][k] = transpose ? src[k * ld + n] : src[n * ld + k]
    int64_t n_pad, k_panels;
};
static_assert(sizeof(ImgJob) == 9 * 8, "ImgJob is passed as 9 x int64 from the host side");

__global__ void __launch_bounds__(256) weight_image_kernel(const ImgJob* __restrict__ jobs) {
    pdl_sync();
    const ImgJob j = jobs[blockIdx.y];
    const int64_t total = j.k_panels * j.n_pad * 8;           // 16-byte chunks
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 7);
        const int64_t n = (i >> 3) % j.n_pad;
        const int64_t p = (i >> 3) / j.n_pad;
        __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int64_t k = p * 64 + c * 8 + e;
            float v = 0.f;
            if (n < j.N && k < j.K) v = j.transpose ? j.src[k * j.ld + n] : j.src[n * j.ld + k];
            split_bf16(v, h[e], l[e]);
        }
        const int64_t off = p * j.n_pad * 128 + n * 128 + ((c ^ (int)(n & 7)) << 4);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(j.hi) + off) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(j.lo) + off) = *reinterpret_cast<const uint4*>(l);
    }
}

// The conv kernels' operand layouts of many Conv3d(k=3) weights in one launch: src (Co, Cin, 27) -> [27][Cin][Co], [27][Co][Cin]
struct ConvWJob {
    const float* src; float* tcio; float* toci; int64_t Cin, Co;
};
static_assert(sizeof(ConvWJob) == 5 * 8, "ConvWJob is passed as 5 x int64 from the host side");

// CTA = (job, 32-input-channel chunk): the chunk's [Co][32][27] block goes through shared memory so that the reads and both
// writes are contiguous runs (27-float rows in, 32 x Co / 32-float runs out)
__global__ void __launch_bounds__(256) conv_weight_layout_kernel(const ConvWJob* __restrict__ jobs) {
    pdl_sync();
    const ConvWJob j = jobs[blockIdx.y];
    const int Cin = (int)j.Cin, Co = (int)j.Co;
    const int c0 = blockIdx.x * 32;
    if (c0 >= Cin) return;
    const int nc = min(32, Cin - c0);
    extern __shared__ float cwl[];                 // [Co][nc * 27 (+1 pad)]
    const int row = nc * 27, rowp = row + 1;
    for (int i = threadIdx.x; i < Co * row; i += blockDim.x) {
        const int co = i / row, r = i - co * row;
        cwl[co * rowp + r] = j.src[((int64_t)co * Cin + c0) * 27 + r];
    }
    __syncthreads();
    // tcio[tap][ci][co]: per tap a run of nc * Co floats
    for (int i = threadIdx.x; i < 27 * nc * Co; i += blockDim.x) {
        const int co = i % Co, ci = (i / Co) % nc, tap = i / (Co * nc);
        j.tcio[((int64_t)tap * Cin + c0 + ci) * Co + co] = cwl[co * rowp + ci * 27 + tap];
    }
    // toci[tap][co][ci]: per (tap, co) a run of nc floats
    for (int i = threadIdx.x; i < 27 * Co * nc; i += blockDim.x) {
        const int ci = i % nc, co = (i / nc) % Co, tap = i / (nc * Co);
        j.toci[((int64_t)tap * Co + co) * Cin + c0 + ci] = cwl[co * rowp + ci * 27 + tap];
    }
}

// ------------------------------------------------------------------------------------------------ forward
struct MlpFwdArgs {
    const float* x; float* y;
    const float* gamma; const float* beta; const float* b1; const float* b2;
    const uint8_t* w1_hi; const uint8_t* w1_lo;      // fc1 image: N = HID rows, K = C (one panel)
    const uint8_t* w2_hi; const uint8_t* w2_lo;      // fc2 image: N = CP rows,  K = HID (HID/64 panels)
    const float* rowscale; int rps;
    int T, ntiles;
    float eps;
};

constexpr int MLP_THREADS = 320;

template <int C>
struct MlpCfg {
    static constexpr int HID = 4 * C;
    static constexpr int CP = (C + 15) / 16 * 16;            // channel count rounded to the MMA K / N granularity
    static constexpr int HP = (HID + 63) / 64;               // 64-wide panels of the hidden tile
    static constexpr int W1_BYTES = HID * 128;               // per hi / lo image
    static constexpr int W2_BYTES = HP * CP * 128;
    static constexpr int A1_BYTES = 128 * 128;               // LayerNorm output tile (one panel), per hi / lo
    static constexpr int A2_BYTES = HP * 128 * 128;          // hidden tile, per hi / lo
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = OFF_W1 + 2 * W1_BYTES;
    static constexpr int OFF_A1 = OFF_W2 + 2 * W2_BYTES;
    static constexpr int OFF_A2 = OFF_A1 + 2 * A1_BYTES;
    static constexpr int OFF_PAR = OFF_A2 + 2 * A2_BYTES;    // b1[HID] b2[C] gamma[C] beta[C]
    static constexpr int OFF_BAR = OFF_PAR + 4 * (HID + 3 * C + 4);
    static constexpr int SMEM = OFF_BAR + 128 + 1024;        // + alignment slack
    static constexpr int TCOLS = (HID + CP) <= 128 ? 128 : ((HID + CP) <= 256 ? 256 : 512);
    static_assert(C % 8 == 0 && C <= 64, "fused MLP: C must be a multiple of 8, at most 64");
    static_assert(HID <= 256, "fc1 is one MMA wide");
};

template <int C>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_block_fwd_kernel(const MlpFwdArgs a) {
    using K = MlpCfg<C>;
    constexpr int HID = K::HID, CP = K::CP;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW1h = smem + K::OFF_W1;  uint8_t* sW1l = sW1h + K::W1_BYTES;
    uint8_t* sW2h = smem + K::OFF_W2;  uint8_t* sW2l = sW2h + K::W2_BYTES;
    uint8_t* sA1h = smem + K::OFF_A1;  uint8_t* sA1l = sA1h + K::A1_BYTES;
    uint8_t* sA2h = smem + K::OFF_A2;  uint8_t* sA2l = sA2h + K::A2_BYTES;
    float* sb1 = reinterpret_cast<float*>(smem + K::OFF_PAR);
    float* sb2 = sb1 + HID;
    float* sg = sb2 + C;
    float* sbt = sg + C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
    uint64_t* w_full = bars + 0;       // weight images landed (once)
    uint64_t* a1_full = bars + 1;      // LayerNorm tile written (8 warp arrivals)
    uint64_t* hp_full = bars + 2;      // fc1 accumulator complete (MMA commit)
    uint64_t* h_full = bars + 3;       // hidden tile written (8 warp arrivals)
    uint64_t* y_full = bars + 4;       // fc2 accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        bar_init(w_full, 1); bar_init(a1_full, 8); bar_init(hp_full, 1); bar_init(h_full, 8); bar_init(y_full, 1);
        bar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, K::TCOLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_hp = tmem, t_y = tmem + HID;
    pdl_sync();

    if (warp == 0) {
        if (lane == 0) {
            bar_expect_tx(w_full, 2 * K::W1_BYTES + 2 * K::W2_BYTES);
            bulk_g2s(sW1h, a.w1_hi, K::W1_BYTES, w_full);
            bulk_g2s(sW1l, a.w1_lo, K::W1_BYTES, w_full);
            bulk_g2s(sW2h, a.w2_hi, K::W2_BYTES, w_full);
            bulk_g2s(sW2l, a.w2_lo, K::W2_BYTES, w_full);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t id1 = idesc_bf16(128, HID, false, false);
            constexpr uint32_t id2 = idesc_bf16(128, CP, false, false);
            bar_wait(w_full, 0);
            uint32_t n = 0;
            for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
                const uint32_t ph = n & 1;
                bar_wait(a1_full, ph);
                fence_after();
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks)
                    mma3(t_hp, desc_k(s32(sA1h) + ks * 32), desc_k(s32(sA1l) + ks * 32), desc_k(s32(sW1h) + ks * 32),
                         desc_k(s32(sW1l) + ks * 32), id1, ks ? 1u : 0u);
                commit(hp_full);
                bar_wait(h_full, ph);
                fence_after();
#pragma unroll 4
                for (int ks = 0; ks < HID / 16; ++ks) {
                    const uint32_t ao = (ks >> 2) * 16384 + (ks & 3) * 32, bo = (ks >> 2) * (CP * 128) + (ks & 3) * 32;
                    mma3(t_y, desc_k(s32(sA2h) + ao), desc_k(s32(sA2l) + ao), desc_k(s32(sW2h) + bo), desc_k(s32(sW2l) + bo),
                         id2, ks ? 1u : 0u);
                }
                commit(y_full);
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        for (int i = threadIdx.x - 64; i < HID + 3 * C; i += MLP_THREADS - 64) {
            float v;
            if (i < HID) v = a.b1[i];
            else if (i < HID + C) v = a.b2[i - HID];
            else if (i < HID + 2 * C) v = a.gamma[i - HID - C];
            else v = a.beta[i - HID - 2 * C];
            sb1[i] = v;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t n = 0;
        for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++n) {
            const uint32_t ph = n & 1;
            const int64_t grow = (int64_t)t * 128 + row;
            const bool ok = grow < a.T;
            if (half == 0 && grow + (int64_t)gridDim.x * 128 < a.T) prefetch_l2(a.x + (grow + (int64_t)gridDim.x * 128) * C, C * 4);
            // ---- LayerNorm of the row -> split-bf16 A tile (each of the two threads of a row writes half of the chunks)
            float xr[C];
            if (ok) {
                const float4* xp = reinterpret_cast<const float4*>(a.x + grow * C);
#pragma unroll
                for (int i = 0; i < C / 4; ++i) {
                    const float4 v = __ldg(xp + i);
                    xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < C; ++i) xr[i] = 0.f;
            }
            float mean = 0.f;
#pragma unroll
            for (int i = 0; i < C; ++i) mean += xr[i];
            mean *= (1.f / C);
            float var = 0.f;
#pragma unroll
            for (int i = 0; i < C; ++i) { const float d = xr[i] - mean; var = fmaf(d, d, var); }
            const float rstd = rsqrtf(var * (1.f / C) + a.eps);
            constexpr int NCH = CP / 8;                          // chunks of the A tile row
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if ((c < (NCH + 1) / 2) == (half == 0)) {
                    float v8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int col = c * 8 + e;
                        v8[e] = col < C ? (xr[col < C ? col : 0] - mean) * rstd * sg[col < C ? col : 0] + sbt[col < C ? col : 0]
                                        : 0.f;          // (indices clamped only to keep the unrolled dead branch in range)
                    }
                    store_chunk(sA1h, sA1l, row, c, v8);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(a1_full);
            // ---- hidden = GELU(fc1 + b1): TMEM -> registers -> split-bf16 A tile of fc2
            bar_wait(hp_full, ph);
            fence_after();
            constexpr int HH = HID / 2;                          // columns per thread
#pragma unroll 1
            for (int g0 = 0; g0 < HH; g0 += 16) {
                const int col0 = half * HH + g0;
                float v[16];
                ld16(tmem + lane_base + (uint32_t)col0, v);
                ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = gelu_fast(v[e] + sb1[col0 + e]);
                uint8_t* ph_ = sA2h + (col0 >> 6) * 16384;
                uint8_t* pl_ = sA2l + (col0 >> 6) * 16384;
                store_chunk(ph_, pl_, row, (col0 & 63) >> 3, v);
                store_chunk(ph_, pl_, row, ((col0 & 63) >> 3) + 1, v + 8);
            }
            fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(h_full);
            // ---- y = x + rowscale * (fc2 + b2): the first thread of each row drains the C output columns
            bar_wait(y_full, ph);
            fence_after();
            if (half == 0) {
                float o[CP];
#pragma unroll
                for (int c0 = 0; c0 < CP; c0 += 16) ld16(t_y + lane_base + (uint32_t)c0, o + c0);
                ld_wait();
                if (ok) {
                    const float rs = a.rowscale ? a.rowscale[grow / a.rps] : 1.f;
                    float4* yp = reinterpret_cast<float4*>(a.y + grow * C);
#pragma unroll
                    for (int i = 0; i < C / 4; ++i) {
                        float4 r;
                        r.x = xr[4 * i] + rs * (o[4 * i] + sb2[4 * i]);
                        r.y = xr[4 * i + 1] + rs * (o[4 * i + 1] + sb2[4 * i + 1]);
                        r.z = xr[4 * i + 2] + rs * (o[4 * i + 2] + sb2[4 * i + 2]);
                        r.w = xr[4 * i + 3] + rs * (o[4 * i + 3] + sb2[4 * i + 3]);
                        yp[i] = r;
                    }
                }
            }
            fence_before();          // TMEM reads of this tile are ordered before the next tile's barrier arrivals
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        tmem_dealloc(tmem, K::TCOLS);
    }
}

template <int C>
static int launch_mlp_fwd(const MlpFwdArgs& a, cudaStream_t st) {
    using K = MlpCfg<C>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(mlp_block_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM) != cudaSuccess) {
            cudaGetLastError();
            return MIC_ERR_UNSUPPORTED;
        }
        attr = true;
    }
    int grid = num_sms();
    if (grid > a.ntiles) grid = a.ntiles;
    mic::launch(mlp_block_fwd_kernel<C>, dim3(grid), dim3(MLP_THREADS), (size_t)K::SMEM, st, a);
    return check_launch("mlp_block_fwd_kernel");
}

}  // namespace mic

using namespace mic;

extern "C" int mic_weight_images(const void* jobs, int n_jobs, int64_t max_chunks, void* stream) {
    MIC_REQUIRE(jobs && n_jobs > 0 && max_chunks > 0, "weight_images: bad arguments");
    int gx = (int)((max_chunks + 255) / 256);
    if (gx > 64) gx = 64;
    mic::launch(weight_image_kernel, dim3(gx, n_jobs), dim3(256), 0, (cudaStream_t)stream, (const ImgJob*)jobs);
    return check_launch("weight_image_kernel");
}

extern "C" int mic_conv_weight_layouts(const void* jobs, int n_jobs, int64_t max_elems, int max_co, void* stream) {
    MIC_REQUIRE(jobs && n_jobs > 0 && max_elems > 0 && max_co > 0 && max_co <= 64, "conv_weight_layouts: bad arguments");
    const int max_cin = (int)(max_elems / 27);                 // an upper bound of every job's Cin (Co >= 1)
    const size_t smem = (size_t)max_co * (32 * 27 + 1) * sizeof(float);
    static size_t attr = 0;
    if (smem > 48 * 1024 && attr < smem) {
        cudaFuncSetAttribute(conv_weight_layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    mic::launch(conv_weight_layout_kernel, dim3((max_cin + 31) / 32, n_jobs), dim3(256), smem, (cudaStream_t)stream, (const ConvWJob*)jobs);
    return check_launch("conv_weight_layout_kernel");
}

extern "C" int mic_mlp_block_smem(int C) {
    switch (C) {
        case 24: return MlpCfg<24>::SMEM;
        case 48: return MlpCfg<48>::SMEM;
        default: return -1;
    }
}

extern "C" int mic_mlp_block_fwd(const float* x, float* y, const float* gamma, const float* beta, const float* b1,
                                 const float* b2, const void* w1_hi, const void* w1_lo, const void* w2_hi, const void* w2_lo,
                                 const float* rowscale, int rows_per_sample, int T, int C, float eps, void* stream) {
    MIC_REQUIRE(x && y && gamma && beta && b1 && b2 && w1_hi && w1_lo && w2_hi && w2_lo && T > 0, "mlp_block_fwd: bad arguments");
    MIC_REQUIRE(!rowscale || rows_per_sample > 0, "mlp_block_fwd: rows_per_sample");
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w1_hi) |
          reinterpret_cast<uintptr_t>(w1_lo) | reinterpret_cast<uintptr_t>(w2_hi) | reinterpret_cast<uintptr_t>(w2_lo)) & 15) != 0)
        return fail(MIC_ERR_UNSUPPORTED, "mlp_block_fwd: pointers must be 16-byte aligned");
    MlpFwdArgs a;
    a.x = x; a.y = y; a.gamma = gamma; a.beta = beta; a.b1 = b1; a.b2 = b2;
    a.w1_hi = (const uint8_t*)w1_hi; a.w1_lo = (const uint8_t*)w1_lo; a.w2_hi = (const uint8_t*)w2_hi; a.w2_lo = (const uint8_t*)w2_lo;
    a.rowscale = rowscale; a.rps = rows_per_sample > 0 ? rows_per_sample : 1;
    a.T = T; a.ntiles = (T + 127) / 128; a.eps = eps;
    switch (C) {
        case 24: return launch_mlp_fwd<24>(a, (cudaStream_t)stream);
        case 48: return launch_mlp_fwd<48>(a, (cudaStream_t)stream);
        default: return fail(MIC_ERR_UNSUPPORTED, "mlp_block_fwd: C=%d is not built (24, 48)", C);
    }
}
