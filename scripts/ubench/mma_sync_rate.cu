// Legacy warp-level MMA issue rates on sm_100a: m16n8k8 TF32 vs m16n8k16 BF16 (fp32 accumulate), 16 warps per SM,
// 8 independent accumulators per warp.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void __launch_bounds__(512) k(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 11, b0 = 5, b1 = 9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 512);
    const int iters = 4096;
    for (int kind = 0; kind < 2; ++kind) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (kind == 0) k<0><<<sms, 512>>>(out, iters); else k<1><<<sms, 512>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double mmas = (double)sms * 16 * iters * 8;
        const double macs = mmas * 16 * 8 * (kind == 0 ? 8 : 16);
        printf("%s: %.3f ms  %.1f TFLOP/s  %.1f MAC/clk/SM @1.965GHz  %.2f clk per MMA per SM\n", kind == 0 ? "m16n8k8 tf32" : "m16n8k16 bf16", ms,
               2 * macs / ms / 1e9, macs / sms / (ms * 1e-3 * 1.965e9), (ms * 1e-3 * 1.965e9) / (mmas / sms));
    }
    return 0;
}
