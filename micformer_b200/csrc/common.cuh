// Shared helpers for the micformer_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/micformer_b200.h"

namespace mic {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MIC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return MIC_OK;
}

#define MIC_REQUIRE(cond, ...)                                      \
    do {                                                            \
        if (!(cond)) return ::mic::fail(MIC_ERR_INVALID, __VA_ARGS__); \
    } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact (erf) GELU and its derivative -- nn.GELU(approximate='none')
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// t -> (t / d, t % d).  Linear indices of this model fit 32 bits; 64-bit division is a ~100-instruction software routine,
// so take the hardware-friendly 32-bit path whenever the value allows it (warp-uniform in practice).
__device__ __forceinline__ void divmod(int64_t& t, int d, int& r) {
    if ((uint64_t)t <= 0xFFFFFFFFull) {
        const uint32_t q = (uint32_t)t / (uint32_t)d;
        r = (int)((uint32_t)t - q * (uint32_t)d);
        t = q;
    } else {
        r = (int)(t % d);
        t /= d;
    }
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- programmatic dependent launch (PDL).  Every kernel is launched with the programmatic-stream-serialization
// attribute and starts with pdl_sync(): griddepcontrol.wait blocks until the preceding kernel of the stream has
// completed and flushed (so all reads/writes below it are ordered exactly as without PDL), then launch_dependents lets
// the NEXT kernel's CTAs be scheduled and run their prologue (barrier init, TMEM allocation, descriptor prefetch, index
// math) while this kernel works.  The chains of 5-10 us kernels of the deep stages overlap launch latency + prologue.
// MICFORMER_PDL=0 disables the attribute (griddepcontrol is then a no-op).
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace mic
