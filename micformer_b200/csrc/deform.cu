// Offset head (LN16 -> GELU -> 1x1 conv -> + reference points) and the deformable trilinear resampling of the
// other modality's token grid (== SpatialTransformer / grid_sample(bilinear, zeros, align_corners=False)).
// Both are HBM/L2-bound gather kernels: channels-last rows, float4 per thread, vector atomics for the scatter.
#include "common.cuh"

namespace mic {

constexpr int HC = 16;

__device__ __forceinline__ void ref_point(int z, int y, int x, int Dp, int Hp, int Wp, float r[3]) {
    // models/MICFormer_self.py:333-335 -- the z channel is normalised by H, y by W, x by D (sic)
    r[0] = __fsub_rn(__fmul_rn(__fdiv_rn((float)z + 0.5f, (float)Hp), 2.f), 1.f);
    r[1] = __fsub_rn(__fmul_rn(__fdiv_rn((float)y + 0.5f, (float)Wp), 2.f), 1.f);
    r[2] = __fsub_rn(__fmul_rn(__fdiv_rn((float)x + 0.5f, (float)Dp), 2.f), 1.f);
}

__global__ void __launch_bounds__(256) offset_head_fwd_kernel(const float* __restrict__ h,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ w3, float* __restrict__ pos,
                                                              int64_t P, int Dp, int Hp, int Wp, float eps) {
    pdl_sync();
    __shared__ float sg[HC], sb[HC], sw[3 * HC];
    if (threadIdx.x < HC) { sg[threadIdx.x] = gamma[threadIdx.x]; sb[threadIdx.x] = beta[threadIdx.x]; }
    if (threadIdx.x < 3 * HC) sw[threadIdx.x] = w3[threadIdx.x];
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        float v[HC];
#pragma unroll
        for (int c = 0; c < HC / 4; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(h + p * HC + c * 4);
            v[c * 4] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
        }
        float mu = 0.f;
#pragma unroll
        for (int c = 0; c < HC; ++c) mu += v[c];
        mu *= (1.f / HC);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < HC; ++c) { const float d = v[c] - mu; var += d * d; }
        const float rs = rsqrtf(var * (1.f / HC) + eps);
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
        for (int c = 0; c < HC; ++c) {
            const float a = gelu_erf((v[c] - mu) * rs * sg[c] + sb[c]);
            o0 = fmaf(a, sw[c], o0); o1 = fmaf(a, sw[HC + c], o1); o2 = fmaf(a, sw[2 * HC + c], o2);
        }
        int64_t t = p;
        int x; divmod(t, Wp, x);
        int y; divmod(t, Hp, y);
        int z; divmod(t, Dp, z);
        float r[3];
        ref_point(z, y, x, Dp, Hp, Wp, r);
        pos[p * 3 + 0] = o0 + r[0];
        pos[p * 3 + 1] = o1 + r[1];
        pos[p * 3 + 2] = o2 + r[2];
    }
}

__global__ void __launch_bounds__(256) offset_head_bwd_kernel(const float* __restrict__ dpos,
                                                              const float* __restrict__ h,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ w3, float* __restrict__ dh,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float* __restrict__ dw3, int64_t P, float eps) {
    pdl_sync();
    __shared__ float sg[HC], sb[HC], sw[3 * HC];
    __shared__ float red[8][5 * HC];   // per warp: dgamma | dbeta | dw3[3][HC]  (no shared-memory float atomics: CAS loops)
    if (threadIdx.x < HC) { sg[threadIdx.x] = gamma[threadIdx.x]; sb[threadIdx.x] = beta[threadIdx.x]; }
    if (threadIdx.x < 3 * HC) sw[threadIdx.x] = w3[threadIdx.x];
    __syncthreads();
    float ag[HC], ab[HC], aw0[HC], aw1[HC], aw2[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) { ag[c] = 0.f; ab[c] = 0.f; aw0[c] = 0.f; aw1[c] = 0.f; aw2[c] = 0.f; }
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        float v[HC];
#pragma unroll
        for (int c = 0; c < HC / 4; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(h + p * HC + c * 4);
            v[c * 4] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
        }
        const float d0 = dpos[p * 3], d1 = dpos[p * 3 + 1], d2 = dpos[p * 3 + 2];
        float mu = 0.f;
#pragma unroll
        for (int c = 0; c < HC; ++c) mu += v[c];
        mu *= (1.f / HC);
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < HC; ++c) { const float d = v[c] - mu; var += d * d; }
        const float rs = rsqrtf(var * (1.f / HC) + eps);
        float gsum = 0.f, gxsum = 0.f;
        float gg[HC];
#pragma unroll
        for (int c = 0; c < HC; ++c) {
            const float xh = (v[c] - mu) * rs;
            const float ln = xh * sg[c] + sb[c];
            const float a = gelu_erf(ln);
            aw0[c] = fmaf(d0, a, aw0[c]); aw1[c] = fmaf(d1, a, aw1[c]); aw2[c] = fmaf(d2, a, aw2[c]);
            const float da = d0 * sw[c] + d1 * sw[HC + c] + d2 * sw[2 * HC + c];
            const float dln = da * gelu_erf_grad(ln);
            ag[c] = fmaf(dln, xh, ag[c]);
            ab[c] += dln;
            gg[c] = dln * sg[c];
            gsum += gg[c];
            gxsum = fmaf(gg[c], xh, gxsum);
            v[c] = xh;
        }
        gsum *= (1.f / HC); gxsum *= (1.f / HC);
#pragma unroll
        for (int c = 0; c < HC / 4; ++c) {
            float4 t;
            t.x = rs * (gg[c * 4] - gsum - v[c * 4] * gxsum);
            t.y = rs * (gg[c * 4 + 1] - gsum - v[c * 4 + 1] * gxsum);
            t.z = rs * (gg[c * 4 + 2] - gsum - v[c * 4 + 2] * gxsum);
            t.w = rs * (gg[c * 4 + 3] - gsum - v[c * 4 + 3] * gxsum);
            *reinterpret_cast<float4*>(dh + p * HC + c * 4) = t;
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < HC; ++c) {
        const float r0 = warp_sum(ag[c]), r1 = warp_sum(ab[c]), r2 = warp_sum(aw0[c]), r3 = warp_sum(aw1[c]),
                    r4 = warp_sum(aw2[c]);
        if (lane == 0) {
            red[wid][c] = r0; red[wid][HC + c] = r1; red[wid][2 * HC + c] = r2;
            red[wid][3 * HC + c] = r3; red[wid][4 * HC + c] = r4;
        }
    }
    __syncthreads();
    if (threadIdx.x < 5 * HC) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        if (threadIdx.x < HC) atomicAdd(&dgamma[threadIdx.x], v);
        else if (threadIdx.x < 2 * HC) atomicAdd(&dbeta[threadIdx.x - HC], v);
        else atomicAdd(&dw3[threadIdx.x - 2 * HC], v);
    }
}

// ------------------------------------------------------------------------------------------ deformable gather
struct SampGeom {
    int B, D, H, W, Dp, Hp, Wp, C;
};

__device__ __forceinline__ float sample_coord(int idx, float off, int S) {
    // STN.py:20-23 then grid_sampler unnormalize (align_corners=False): ((2*(v/(S-1) - .5) + 1) * S - 1) / 2
    const float v = __fadd_rn((float)idx, off);
    const float nl = __fmul_rn(2.f, __fsub_rn(__fdiv_rn(v, (float)(S - 1)), 0.5f));
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(nl, 1.f), (float)S), 1.f), 0.5f);
}

__global__ void __launch_bounds__(256) deform_sample_fwd_kernel(const float* __restrict__ src,
                                                                const float* __restrict__ pos,
                                                                float* __restrict__ out, SampGeom g, int64_t total) {
    pdl_sync();
    const int C4 = g.C >> 2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t pq = idx;
        int c4; divmod(pq, C4, c4);
        const int64_t p = pq;
        int64_t t = p;
        int x; divmod(t, g.Wp, x);
        int y; divmod(t, g.Hp, y);
        int z; divmod(t, g.Dp, z);
        const int b = (int)t;
        const float cz = sample_coord(z, pos[p * 3 + 0], g.Dp);
        const float cy = sample_coord(y, pos[p * 3 + 1], g.Hp);
        const float cx = sample_coord(x, pos[p * 3 + 2], g.Wp);
        const float fz = floorf(cz), fy = floorf(cy), fx = floorf(cx);
        const float tz = cz - fz, ty = cy - fy, tx = cx - fx;
        const int iz = (int)fz, iy = (int)fy, ix = (int)fx;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
            const int zz = iz + dz, yy = iy + dy, xx = ix + dx;
            // zeros padding outside the padded grid; the pad region itself holds zeros (F.pad of xa)
            if (zz < 0 || zz >= g.D || yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) continue;
            const float w = (dz ? tz : 1.f - tz) * (dy ? ty : 1.f - ty) * (dx ? tx : 1.f - tx);
            const float4 v = *reinterpret_cast<const float4*>(
                src + ((((int64_t)b * g.D + zz) * g.H + yy) * g.W + xx) * g.C + c4 * 4);
            acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(out + p * g.C + c4 * 4) = acc;
    }
}

__global__ void __launch_bounds__(256) deform_sample_bwd_kernel(const float* __restrict__ dout,
                                                                const float* __restrict__ src,
                                                                const float* __restrict__ pos,
                                                                float* __restrict__ dsrc, float* __restrict__ dpos,
                                                                SampGeom g, int64_t total) {
    pdl_sync();
    const int C4 = g.C >> 2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t pq = idx;
        int c4; divmod(pq, C4, c4);
        const int64_t p = pq;
        int64_t t = p;
        int x; divmod(t, g.Wp, x);
        int y; divmod(t, g.Hp, y);
        int z; divmod(t, g.Dp, z);
        const int b = (int)t;
        const float cz = sample_coord(z, pos[p * 3 + 0], g.Dp);
        const float cy = sample_coord(y, pos[p * 3 + 1], g.Hp);
        const float cx = sample_coord(x, pos[p * 3 + 2], g.Wp);
        const float fz = floorf(cz), fy = floorf(cy), fx = floorf(cx);
        const float tz = cz - fz, ty = cy - fy, tx = cx - fx;
        const int iz = (int)fz, iy = (int)fy, ix = (int)fx;
        const float4 go = *reinterpret_cast<const float4*>(dout + p * g.C + c4 * 4);
        float gz = 0.f, gy = 0.f, gx = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
            const int zz = iz + dz, yy = iy + dy, xx = ix + dx;
            if (zz < 0 || zz >= g.D || yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) continue;
            const float wz = dz ? tz : 1.f - tz, wy = dy ? ty : 1.f - ty, wx = dx ? tx : 1.f - tx;
            const int64_t off = ((((int64_t)b * g.D + zz) * g.H + yy) * g.W + xx) * g.C + c4 * 4;
            const float4 v = *reinterpret_cast<const float4*>(src + off);
            const float w = wz * wy * wx;
            atomicAdd(reinterpret_cast<float4*>(dsrc + off), make_float4(w * go.x, w * go.y, w * go.z, w * go.w));
            const float dot = v.x * go.x + v.y * go.y + v.z * go.z + v.w * go.w;
            gz += (dz ? dot : -dot) * wy * wx;
            gy += (dy ? dot : -dot) * wz * wx;
            gx += (dx ? dot : -dot) * wz * wy;
        }
        atomicAdd(&dpos[p * 3 + 0], gz * ((float)g.Dp / (float)(g.Dp - 1)));
        atomicAdd(&dpos[p * 3 + 1], gy * ((float)g.Hp / (float)(g.Hp - 1)));
        atomicAdd(&dpos[p * 3 + 2], gx * ((float)g.Wp / (float)(g.Wp - 1)));
    }
}

static unsigned grid_for(int64_t total, int threads) {
    int64_t b = ceil_div64(total, threads);
    const int64_t cap = (int64_t)num_sms() * 32;
    return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace mic

using namespace mic;

extern "C" int mic_offset_head_fwd(const float* h, const float* gamma, const float* beta, const float* w3, float* pos,
                                   int B, int Dp, int Hp, int Wp, int hc, float eps, void* stream) {
    MIC_REQUIRE(h && gamma && beta && w3 && pos, "offset_head_fwd: null pointer");
    if (hc != HC) return fail(MIC_ERR_UNSUPPORTED, "offset_head: hidden_channels=%d (only 16 is built)", hc);
    const int64_t P = (int64_t)B * Dp * Hp * Wp;
    mic::launch(offset_head_fwd_kernel, dim3(grid_for(P, 256)), dim3(256), 0, (cudaStream_t)stream, h, gamma, beta, w3, pos, P, Dp, Hp, Wp, eps);
    return check_launch("offset_head_fwd_kernel");
}

extern "C" int mic_offset_head_bwd(const float* dpos, const float* h, const float* gamma, const float* beta,
                                   const float* w3, float* dh, float* dgamma, float* dbeta, float* dw3, int B, int Dp,
                                   int Hp, int Wp, int hc, float eps, void* stream) {
    MIC_REQUIRE(dpos && h && gamma && beta && w3 && dh && dgamma && dbeta && dw3, "offset_head_bwd: null pointer");
    if (hc != HC) return fail(MIC_ERR_UNSUPPORTED, "offset_head: hidden_channels=%d (only 16 is built)", hc);
    const int64_t P = (int64_t)B * Dp * Hp * Wp;
    int64_t blocks = ceil_div64(P, 256);
    const int64_t cap = (int64_t)num_sms() * 2;
    if (blocks > cap) blocks = cap;
    mic::launch(offset_head_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, dpos, h, gamma, beta, w3, dh, dgamma, dbeta,
                                                                             dw3, P, eps);
    return check_launch("offset_head_bwd_kernel");
}

extern "C" int mic_deform_sample_fwd(const float* src, const float* pos, float* out, int B, int D, int H, int W, int Dp,
                                     int Hp, int Wp, int C, void* stream) {
    MIC_REQUIRE(src && pos && out, "deform_sample_fwd: null pointer");
    MIC_REQUIRE(C > 0 && (C & 3) == 0, "deform_sample_fwd: C=%d must be a multiple of 4", C);
    MIC_REQUIRE(Dp >= D && Hp >= H && Wp >= W, "deform_sample_fwd: bad geometry");
    SampGeom g{B, D, H, W, Dp, Hp, Wp, C};
    const int64_t total = (int64_t)B * Dp * Hp * Wp * (C / 4);
    mic::launch(deform_sample_fwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, pos, out, g, total);
    return check_launch("deform_sample_fwd_kernel");
}

extern "C" int mic_deform_sample_bwd(const float* dout, const float* src, const float* pos, float* dsrc, float* dpos,
                                     int B, int D, int H, int W, int Dp, int Hp, int Wp, int C, void* stream) {
    MIC_REQUIRE(dout && src && pos && dsrc && dpos, "deform_sample_bwd: null pointer");
    MIC_REQUIRE(C > 0 && (C & 3) == 0, "deform_sample_bwd: C=%d must be a multiple of 4", C);
    SampGeom g{B, D, H, W, Dp, Hp, Wp, C};
    const int64_t P = (int64_t)B * Dp * Hp * Wp;
    cudaError_t e = cudaMemsetAsync(dpos, 0, P * 3 * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(MIC_ERR_CUDA, "deform_sample_bwd memset: %s", cudaGetErrorString(e));
    const int64_t total = P * (C / 4);
    mic::launch(deform_sample_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dout, src, pos, dsrc, dpos, g, total);
    return check_launch("deform_sample_bwd_kernel");
}
