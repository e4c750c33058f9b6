// Multi-tensor Adam step (torch.optim.Adam semantics: amsgrad=False, maximize=False, L2 weight decay folded into the
// gradient) over ALL parameter tensors in ONE launch.  The reference's optimizer is torch.optim.Adam(lr=1e-4, wd=0)
// stepped every iteration (train_mmwhs_noPad.py:114,201); SURVEY 8(f) rank 1.
//   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// The per-tensor step counts and the learning rate live in device memory so the launch is CUDA-graph capturable; the
// per-iteration cosine schedule (:148,206-207) only rewrites the device lr scalar.
// HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"

namespace mic {

constexpr int ADAM_CHUNK = 32768;      // elements per CTA

// per-tensor step counters (torch.optim.Adam keeps one per parameter and skips parameters without a gradient)
__global__ void adam_tick_kernel(float* __restrict__ steps, const float* const* __restrict__ grads, int n_tensors) {
    pdl_sync();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_tensors && grads[t] != nullptr) steps[t] += 1.f;
}

__global__ void __launch_bounds__(256) adam_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads,
                                                   float* const* __restrict__ exp_avg, float* const* __restrict__ exp_avg_sq,
                                                   const int64_t* __restrict__ sizes, const int* __restrict__ chunk_tensor,
                                                   const int* __restrict__ chunk_index, const float* __restrict__ steps,
                                                   const float* __restrict__ lr_dev, float beta1, float beta2, float eps,
                                                   float weight_decay) {
    pdl_sync();
    const int t = chunk_tensor[blockIdx.x];
    const int64_t off = (int64_t)chunk_index[blockIdx.x] * ADAM_CHUNK;
    const int64_t n = min((int64_t)ADAM_CHUNK, sizes[t] - off);
    float* p = params[t] + off;
    const float* g = grads[t];
    if (g == nullptr) return;                    // parameter without a gradient this step (e.g. concat_back_dim.0)
    g += off;
    float* m = exp_avg[t] + off;
    float* v = exp_avg_sq[t] + off;
    const float tstep = steps[t];
    const float step_size = lr_dev[0] / (1.f - powf(beta1, tstep));
    const float inv_sqrt_bc2 = rsqrtf(1.f - powf(beta2, tstep));
    const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = threadIdx.x; i < n4; i += 256) {
            float4 pp = reinterpret_cast<float4*>(p)[i];
            float4 gg = reinterpret_cast<const float4*>(g)[i];
            float4 mm = reinterpret_cast<float4*>(m)[i];
            float4 vv = reinterpret_cast<float4*>(v)[i];
            float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float gk = ga[k] + weight_decay * pa[k];
                ma[k] = beta1 * ma[k] + omb1 * gk;
                va[k] = beta2 * va[k] + omb2 * gk * gk;
                pa[k] -= step_size * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
            }
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
        }
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
            float gk = g[i] + weight_decay * p[i];
            m[i] = beta1 * m[i] + omb1 * gk;
            v[i] = beta2 * v[i] + omb2 * gk * gk;
            p[i] -= step_size * m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps);
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += 256) {
            float gk = g[i] + weight_decay * p[i];
            m[i] = beta1 * m[i] + omb1 * gk;
            v[i] = beta2 * v[i] + omb2 * gk * gk;
            p[i] -= step_size * m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps);
        }
    }
}

}  // namespace mic

using namespace mic;

extern "C" int mic_adam_chunk_elems(void) { return ADAM_CHUNK; }

extern "C" int mic_adam_step(void* params, void* grads, void* exp_avg, void* exp_avg_sq, const int64_t* sizes,
                             const int* chunk_tensor, const int* chunk_index, int n_chunks, int n_tensors, float* steps,
                             const float* lr, float beta1, float beta2, float eps, float weight_decay, void* stream) {
    MIC_REQUIRE(params && grads && exp_avg && exp_avg_sq && sizes && chunk_tensor && chunk_index && steps && lr &&
                    n_chunks > 0 && n_tensors > 0, "adam_step: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    mic::launch(adam_tick_kernel, dim3((n_tensors + 255) / 256), dim3(256), 0, st, steps, reinterpret_cast<const float* const*>(grads), n_tensors);
    int rc = check_launch("adam_tick_kernel");
    if (rc) return rc;
    mic::launch(adam_kernel, dim3(n_chunks), dim3(256), 0, st, reinterpret_cast<float* const*>(params), reinterpret_cast<const float* const*>(grads),
                                          reinterpret_cast<float* const*>(exp_avg), reinterpret_cast<float* const*>(exp_avg_sq),
                                          sizes, chunk_tensor, chunk_index, steps, lr, beta1, beta2, eps, weight_decay);
    return check_launch("adam_kernel");
}
