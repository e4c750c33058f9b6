// fp32 CUDA-core GEMM family: C[i,j] = sum_r A(i,r) * B(r,j) with fused epilogues.  This is the exact-fp32
// (parity) path for every Linear / stride==kernel convolution of the model and the fallback for shapes the
// tcgen05 path (gemm_tc.cu) does not take.  128x64x16 tiles, 256 threads, 8x4 register tile per thread.
//
// One templated kernel covers forward (X W^T), backward-data (dY W) and backward-weight (dY^T X, split over
// the long row axis with atomic accumulation) -- they differ only in which operand axis is contiguous.
#include "common.cuh"

namespace mic {

constexpr int BM = 128, BN = 64, BK = 16, GT = 256;

struct GemmArgs {
    const float* A; int64_t sa_i, sa_r;
    const float* B; int64_t sb_r, sb_j;
    float* C; int64_t ldc;
    int I, J, R;
    int r_chunk;                 // rows of R per blockIdx.z
    // epilogue
    const float* bias;           // [J]
    int act;                     // 1: gelu(val), pre-activation stored to `pre`
    float* pre; int64_t ldpre;
    const float* mulgrad; int64_t ldmg;     // val *= gelu'(mulgrad[i,j])
    const float* res; int64_t ldres;        // val = res + rowscale*val
    const float* rowscale_i; int rps_i;     // per output row i
    const float* rowscale_r; int rps_r;     // per reduction index r (backward-weight)
    int accumulate;              // 1: C += val (plain RMW), 2: atomicAdd
};

template <bool A_RC, bool B_RC>
__global__ void __launch_bounds__(GT) gemm_kernel(GemmArgs p) {
    pdl_sync();
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    const int r_begin = blockIdx.z * p.r_chunk;
    const int r_end = min(p.R, r_begin + p.r_chunk);
    const int ty = tid >> 4, tx = tid & 15;

    const bool a_vec = aligned16(p.A) && ((A_RC ? p.sa_i : p.sa_r) % 4 == 0);
    const bool b_vec = aligned16(p.B) && ((B_RC ? p.sb_j : p.sb_r) % 4 == 0);

    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += BK) {
        // ---- A tile: BM x BK ----
        if (A_RC) {   // A[i*sa_i + r], r contiguous: float4 along r, transposed store
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                const int idx = tid + l * GT;            // 0..511
                const int i = idx >> 2, r4 = (idx & 3) * 4;
                const int gi = i0 + i, gr = r0 + r4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gi < p.I) {
                    const float* src = p.A + (int64_t)gi * p.sa_i + gr;
                    if (a_vec && gr + 3 < r_end) v = *reinterpret_cast<const float4*>(src);
                    else {
                        if (gr + 0 < r_end) v.x = src[0];
                        if (gr + 1 < r_end) v.y = src[1];
                        if (gr + 2 < r_end) v.z = src[2];
                        if (gr + 3 < r_end) v.w = src[3];
                    }
                }
                As[r4 + 0][i] = v.x; As[r4 + 1][i] = v.y; As[r4 + 2][i] = v.z; As[r4 + 3][i] = v.w;
            }
        } else {      // A[r*sa_r + i], i contiguous: float4 along i, direct store (optionally scaled per r)
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                const int idx = tid + l * GT;
                const int r = idx >> 5, i4 = (idx & 31) * 4;
                const int gr = r0 + r, gi = i0 + i4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < r_end) {
                    const float* src = p.A + (int64_t)gr * p.sa_r + gi;
                    if (a_vec && gi + 3 < p.I) v = *reinterpret_cast<const float4*>(src);
                    else {
                        if (gi + 0 < p.I) v.x = src[0];
                        if (gi + 1 < p.I) v.y = src[1];
                        if (gi + 2 < p.I) v.z = src[2];
                        if (gi + 3 < p.I) v.w = src[3];
                    }
                    if (p.rowscale_r) {
                        const float s = p.rowscale_r[gr / p.rps_r];
                        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
                    }
                }
                *reinterpret_cast<float4*>(&As[r][i4]) = v;
            }
        }
        // ---- B tile: BK x BN ----
        if (B_RC) {   // B[j*sb_j + r]
            const int j = tid >> 2, r4 = (tid & 3) * 4;
            const int gj = j0 + j, gr = r0 + r4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gj < p.J) {
                const float* src = p.B + (int64_t)gj * p.sb_j + gr;
                if (b_vec && gr + 3 < r_end) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (gr + 0 < r_end) v.x = src[0];
                    if (gr + 1 < r_end) v.y = src[1];
                    if (gr + 2 < r_end) v.z = src[2];
                    if (gr + 3 < r_end) v.w = src[3];
                }
            }
            Bs[r4 + 0][j] = v.x; Bs[r4 + 1][j] = v.y; Bs[r4 + 2][j] = v.z; Bs[r4 + 3][j] = v.w;
        } else {      // B[r*sb_r + j]
            const int r = tid >> 4, j4 = (tid & 15) * 4;
            const int gr = r0 + r, gj = j0 + j4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < r_end) {
                const float* src = p.B + (int64_t)gr * p.sb_r + gj;
                if (b_vec && gj + 3 < p.J) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (gj + 0 < p.J) v.x = src[0];
                    if (gj + 1 < p.J) v.y = src[1];
                    if (gj + 2 < p.J) v.z = src[2];
                    if (gj + 3 < p.J) v.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(&Bs[r][j4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
        }
        __syncthreads();
    }

    // ---- epilogue ----
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int gi = i0 + ty * 8 + a;
        if (gi >= p.I) continue;
        const float rs = p.rowscale_i ? p.rowscale_i[gi / p.rps_i] : 1.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gj = j0 + tx * 4 + c;
            if (gj >= p.J) continue;
            float v = acc[a][c];
            if (p.bias) v += p.bias[gj];
            if (p.act == 1) {
                if (p.pre) p.pre[(int64_t)gi * p.ldpre + gj] = v;
                v = gelu_erf(v);
            }
            if (p.mulgrad) v *= gelu_erf_grad(p.mulgrad[(int64_t)gi * p.ldmg + gj]);
            v *= rs;
            if (p.res) v += p.res[(int64_t)gi * p.ldres + gj];
            float* dst = p.C + (int64_t)gi * p.ldc + gj;
            if (p.accumulate == 2) atomicAdd(dst, v);
            else if (p.accumulate == 1) *dst += v;
            else *dst = v;
        }
    }
}

// db[j] += sum_m rowscale[m]*dY[m,j]
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ X, int64_t ldx, int M, int N,
                                                     const float* __restrict__ rowscale, int rps,
                                                     float* __restrict__ out, int rows_per_block) {
    pdl_sync();
    // block (32 x 32): x -> column, y -> row lane
    __shared__ float red[32][33];
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int m0 = blockIdx.y * rows_per_block;
    const int m1 = min(M, m0 + rows_per_block);
    float s = 0.f;
    if (j < N)
        for (int m = m0 + threadIdx.y; m < m1; m += 32) {
            float v = X[(int64_t)m * ldx + j];
            if (rowscale) v *= rowscale[m / rps];
            s += v;
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && j < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][threadIdx.x];
        atomicAdd(&out[j], t);
    }
}

int launch_gemm(const GemmArgs& p, bool a_rc, bool b_rc, int splits, cudaStream_t st) {
    dim3 grid(ceil_div(p.I, BM), ceil_div(p.J, BN), splits);
    if (a_rc && b_rc) mic::launch((gemm_kernel<true, true>), grid, dim3(GT), 0, st, p);
    else if (a_rc && !b_rc) mic::launch((gemm_kernel<true, false>), grid, dim3(GT), 0, st, p);
    else if (!a_rc && b_rc) mic::launch((gemm_kernel<false, true>), grid, dim3(GT), 0, st, p);
    else mic::launch((gemm_kernel<false, false>), grid, dim3(GT), 0, st, p);
    return check_launch("gemm_kernel");
}

int colsum(const float* X, int64_t ldx, int M, int N, const float* rowscale, int rps, float* out, cudaStream_t st) {
    int rows_per_block = M >= 16384 ? 512 : 128;
    dim3 grid(ceil_div(N, 32), ceil_div(M, rows_per_block));
    mic::launch(colsum_kernel, grid, dim3(32, 32), 0, st, X, ldx, M, N, rowscale, rps, out, rows_per_block);
    return check_launch("colsum_kernel");
}

// ---- SIMT entry points used by api.cu (which may route to the tensor-core path first) ----
int simt_linear_fwd(const float* X, int ldx, const float* W, int ldw, int w_is_kn, const float* bias, float* Y, int ldy,
                    int M, int N, int K, int act, float* pre, int ldpre, const float* res, int ldres,
                    const float* rowscale, int rps, int accumulate, cudaStream_t st) {
    GemmArgs p{};
    p.A = X; p.sa_i = ldx; p.sa_r = 1;
    p.B = W;
    if (w_is_kn) { p.sb_r = ldw; p.sb_j = 1; } else { p.sb_j = ldw; p.sb_r = 1; }
    p.C = Y; p.ldc = ldy; p.I = M; p.J = N; p.R = K; p.r_chunk = K;
    p.bias = bias; p.act = act; p.pre = pre; p.ldpre = ldpre; p.res = res; p.ldres = ldres;
    p.rowscale_i = rowscale; p.rps_i = rps > 0 ? rps : 1; p.accumulate = accumulate ? 1 : 0;
    return launch_gemm(p, true, !w_is_kn, 1, st);
}

int simt_linear_bwd_data(const float* dY, int lddy, const float* W, int ldw, int w_is_kn, float* dX, int lddx, int M,
                         int N, int K, const float* gelu_pre, int ldpre, const float* rowscale, int rps, int accumulate,
                         cudaStream_t st) {
    GemmArgs p{};
    p.A = dY; p.sa_i = lddy; p.sa_r = 1;
    p.B = W;   // B(r=n, j=k)
    if (w_is_kn) { p.sb_j = ldw; p.sb_r = 1; } else { p.sb_r = ldw; p.sb_j = 1; }
    p.C = dX; p.ldc = lddx; p.I = M; p.J = K; p.R = N; p.r_chunk = N;
    p.mulgrad = gelu_pre; p.ldmg = ldpre;
    p.rowscale_i = rowscale; p.rps_i = rps > 0 ? rps : 1; p.accumulate = accumulate ? 1 : 0;
    return launch_gemm(p, true, w_is_kn != 0, 1, st);
}

int simt_linear_bwd_weight(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, int w_is_kn,
                           float* db, int M, int N, int K, const float* rowscale, int rps, cudaStream_t st) {
    GemmArgs p{};
    if (!w_is_kn) {   // dW[n,k] = sum_m dY[m,n] X[m,k]
        p.A = dY; p.sa_r = lddy; p.sa_i = 1; p.B = X; p.sb_r = ldx; p.sb_j = 1; p.I = N; p.J = K;
    } else {          // dW[k,n] = sum_m X[m,k] dY[m,n]
        p.A = X; p.sa_r = ldx; p.sa_i = 1; p.B = dY; p.sb_r = lddy; p.sb_j = 1; p.I = K; p.J = N;
    }
    p.C = dW; p.ldc = lddw; p.R = M;
    p.rowscale_r = rowscale; p.rps_r = rps > 0 ? rps : 1;
    p.accumulate = 2;
    const int tiles = ceil_div(p.I, BM) * ceil_div(p.J, BN);
    int splits = (num_sms() * 2 + tiles - 1) / tiles;
    const int max_splits = ceil_div(M, 4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int chunk = ceil_div(M, splits);
    chunk = ceil_div(chunk, BK) * BK;
    splits = ceil_div(M, chunk);
    p.r_chunk = chunk;
    int rc = launch_gemm(p, false, false, splits, st);
    if (rc) return rc;
    if (db) return colsum(dY, lddy, M, N, rowscale, p.rps_r, db, st);
    return MIC_OK;
}

}  // namespace mic
