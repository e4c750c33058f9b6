// Micro-benchmarks for the open hardware questions of DESIGN.md section 7 (B200, sm_100a).  One CTA per measurement,
// clock64() around the measured region, results in cycles.  Build + run:  python scripts/ubench/run.py
//   1. tcgen05.ld / tcgen05.st (32x32b.x32): latency of a dependent ld+wait, throughput with 1 / 4 / 8 warps
//   2. tcgen05.mma kind::tf32, M=128, K=8, N in {16,32,64,96,128,192,256}: cycles per MMA when all accumulate into one
//      TMEM tile (the K-loop case) and when they rotate over independent tiles; SS (A from smem) and TS (A from TMEM)
//   3. cp.async 16 B gathers with a 192-byte stride (the conv3_tc producer pattern): latency of one group, cycles per
//      group with 5 groups in flight, and the same with 32 B per thread
//   4. fence.proxy.async.shared::cta after a shared-memory store
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tst32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot, int warp) {
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t tmem, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ---- 1. TMEM ld / st.  mode 0: ld + wait each (latency); 1: 8 lds then one wait (throughput); 2/3: same for st
__global__ void tmem_ldst_kernel(int mode, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t tmem = tmem_alloc_all(&slot, warp);
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    for (int c = 0; c < 8; ++c) tst32(base + c * 32, r);      // defined contents
    WAIT_ST();
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
            tld32(base + (it & 7) * 32, r); WAIT_LD(); acc += r[0];
        } else if (mode == 1) {
#pragma unroll
            for (int c = 0; c < 8; ++c) { tld32(base + c * 32, r); }
            WAIT_LD(); acc += r[0];
        } else if (mode == 2) {
            tst32(base + (it & 7) * 32, r); WAIT_ST();
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) tst32(base + c * 32, r);
            WAIT_ST();
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0xdeadbeef) sink[0] = acc;
    tmem_free(tmem, warp);
}

// ---- 2. MMA rate
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__global__ void mma_rate_kernel(int N, int nmma, int rotate, int a_from_tmem, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (4096 + 256 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t tmem = tmem_alloc_all(&slot, warp);
    if (threadIdx.x == 0) {
        // A: 128 rows x 8 tf32 (K-major, no swizzle): [kchunk 2][16 row groups][8 rows][16 B]; B: N rows likewise
        const uint32_t a_addr = s32(sm), b_addr = s32(sm + 4096);
        const uint64_t ad = desc_noswz(a_addr, 16 * 128, 128);
        const uint64_t bd = desc_noswz(b_addr, (N / 8) * 128, 128);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int ntile = rotate ? (N <= 128 ? 2 : 1) : 1;       // independent accumulators inside 256 columns (A tile sits above)
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint32_t d = tmem + (uint32_t)((i % ntile) * N);
            if (a_from_tmem) {
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                             ::"r"(d), "r"(tmem + 384u), "l"(bd), "r"(idesc), "r"(i >= ntile ? 1u : 0u) : "memory");
            } else {
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(i >= ntile ? 1u : 0u) : "memory");
            }
        }
        const long long t_issue = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n"
                     ::"r"(s32(&bar)), "r"(0) : "memory");
        const long long t1 = clock64();
        out[0] = t1 - t0;
        out[1] = t_issue - t0;
    }
    tmem_free(tmem, warp);
}


// ---- 2b. several issuing threads (one per warp), each accumulating into its own TMEM tile: is the ~100-clock floor of a
// K=8 TF32 MMA per issuing thread or per SM?  kind_f16 = 1 issues kind::f16 (bf16 operands, K=16 per instruction) instead.
__global__ void mma_multi_kernel(int N, int nmma, int issuers, int kind_f16, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (4096 + 256 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t tmem = tmem_alloc_all(&slot, warp);
    long long dt = 0;
    if (lane == 0 && warp < issuers) {
        const uint32_t a_addr = s32(sm), b_addr = s32(sm + 4096);
        const uint64_t ad = desc_noswz(a_addr, 16 * 128, 128);
        const uint64_t bd = desc_noswz(b_addr, (N / 8) * 128, 128);
        // kind::tf32: a/b format 2 (tf32) at bits 7/10; kind::f16: format 1 (bf16)
        const uint32_t fmt = kind_f16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t d = tmem + (uint32_t)(warp * 128);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            if (kind_f16)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[warp])) : "memory");
        asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n"
                     ::"r"(s32(&bar[warp])), "r"(0) : "memory");
        dt = clock64() - t0;
        out[warp] = dt;
    }
    tmem_free(tmem, warp);
}

// ---- 3b. the conv3_tc producer's exact pattern: lane pairs fetch the two 16-byte halves of one 32-byte sector (8 channels
// of a position); lanes_per_pos = 2 (32 B per position) or 4 (64 B per position)
__global__ void cpasync_pair_kernel(const float* __restrict__ src, int64_t stride_floats, int lanes_per_pos, int groups,
                                    long long* out) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int tid = threadIdx.x;
    const uint32_t sbase = s32(sm);
    auto issue = [&](int g) {
        for (int k = 0; k < 3; ++k) {
            const int e = (g * 3 + k) * blockDim.x + tid;
            const int pos = e / lanes_per_pos, part = e % lanes_per_pos;
            const float* p = src + (int64_t)pos * stride_floats + 4 * part;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, 16;" ::"r"(sbase + (uint32_t)((e % 2048) * 16)), "l"(p) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncthreads();
    const long long t0 = clock64();
    int issued = 0;
    for (; issued < 5 && issued < groups; ++issued) issue(issued);
    for (int g = 0; g < groups; ++g) {
        asm volatile("cp.async.wait_group 4;" ::: "memory");
        if (issued < groups) issue(issued++);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
}

// ---- 3. cp.async gathers.  bytes_per_thread 16 or 32; groups in flight: 1 (latency) or 5
__global__ void cpasync_kernel(const float* __restrict__ src, int64_t stride_floats, int per_thread16, int groups, int depth,
                               long long* out) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int tid = threadIdx.x;
    const uint32_t sbase = s32(sm);
    auto issue = [&](int g) {
        for (int k = 0; k < 3; ++k) {                        // three elements per thread per group (as in conv3_tc)
            const int e = (g * 3 + k) * blockDim.x + tid;
            const float* p = src + (int64_t)e * stride_floats;
            for (int h = 0; h < per_thread16; ++h)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, 16;" ::"r"(sbase + (uint32_t)(((e % 1024) * 2 + h) * 16)), "l"(p + 4 * h) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncthreads();
    const long long t0 = clock64();
    int issued = 0;
    for (; issued < depth && issued < groups; ++issued) issue(issued);
    for (int g = 0; g < groups; ++g) {
        if (depth >= 5) asm volatile("cp.async.wait_group 4;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (issued < groups) issue(issued++);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
}

// ---- 4. proxy fence
__global__ void fence_kernel(int iters, long long* out) {
    __shared__ float buf[1024];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        buf[(threadIdx.x + i) & 1023] = (float)i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (buf[threadIdx.x] == -1.f) out[1] = 1;
}

int main() {
    long long* out; uint32_t* sink;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 64));
    long long h[2];
    auto get = [&]() { CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost)); };
    printf("== tcgen05.ld / st 32x32b.x32 (4 KB per warp instruction)\n");
    const char* names[4] = {"ld + wait (dependent)", "8 x ld, one wait", "st + wait (dependent)", "8 x st, one wait"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8}) {
            const int iters = 200;
            tmem_ldst_kernel<<<1, warps * 32>>>(mode, iters, out, sink); get();
            const double per = (double)h[0] / (iters * ((mode & 1) ? 8 : 1));
            printf("  %-24s %d warps: %7.1f cycles per instruction per warp  -> %6.1f B/clk per SM\n", names[mode], warps, per,
                   warps * 4096.0 / per);
        }
    printf("== tcgen05.mma kind::tf32 M=128 K=8: cycles per MMA (issue..commit), issue-only cycles per MMA\n");
    CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int ts = 0; ts < 2; ++ts)
        for (int rot = 0; rot < 2; ++rot)
            for (int N : {16, 32, 64, 96, 128, 192, 256}) {
                if (ts && N > 128) continue;
                const int nm = 256;
                mma_rate_kernel<<<1, 128, 4096 + 256 * 32 + 1024>>>(N, nm, rot, ts, out); get();
                printf("  %s N=%3d %s: %7.1f cycles/MMA  (issue %5.1f)   = %6.0f MAC/clk\n", ts ? "TS" : "SS", N,
                       rot ? "rotating accumulators" : "one accumulator      ", (double)h[0] / nm, (double)h[1] / nm,
                       128.0 * N * 8 / ((double)h[0] / nm));
            }
    printf("== tcgen05.mma issue floor vs number of issuing threads (one per warp, own TMEM tile each): cycles per MMA per issuer\n");
    CK(cudaFuncSetAttribute(mma_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int f16 = 0; f16 < 2; ++f16)
        for (int N : {16, 32, 96})
            for (int issuers : {1, 2, 4}) {
                const int nm = 256;
                long long hh[4] = {0, 0, 0, 0};
                CK(cudaMemset(out, 0, 64));
                mma_multi_kernel<<<1, 128, 4096 + 256 * 32 + 1024>>>(N, nm, issuers, f16, out);
                CK(cudaDeviceSynchronize()); CK(cudaMemcpy(hh, out, 32, cudaMemcpyDeviceToHost));
                long long mx = 0; for (int i = 0; i < issuers; ++i) mx = hh[i] > mx ? hh[i] : mx;
                printf("  %s N=%3d, %d issuer(s): %7.1f cycles/MMA per issuer -> %6.0f MAC/clk per SM\n", f16 ? "kind::f16 (K=16)" : "kind::tf32 (K=8)",
                       N, issuers, (double)mx / nm, issuers * 128.0 * N * (f16 ? 16 : 8) / ((double)mx / nm));
            }
    printf("== cp.async 16 B gathers, 128 threads x 3 elements per group, 192 B stride (L2-resident source)\n");
    float* src; CK(cudaMalloc(&src, (size_t)64 << 20)); CK(cudaMemset(src, 0, (size_t)64 << 20));
    CK(cudaFuncSetAttribute(cpasync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int b16 : {1, 2})
        for (int depth : {1, 5}) {
            const int groups = 64;
            for (int rep = 0; rep < 2; ++rep) { cpasync_kernel<<<1, 128, 40960>>>(src, 48, b16, groups, depth, out); get(); }
            printf("  %2d B per element, %d group(s) in flight: %7.1f cycles per group (%d elements)\n", 16 * b16, depth,
                   (double)h[0] / groups, 3 * 128);
        }
    printf("== cp.async, conv3_tc producer pattern (lane groups share a position), 5 groups in flight\n");
    CK(cudaFuncSetAttribute(cpasync_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int lpp : {2, 4}) {
        const int groups = 64;
        for (int rep = 0; rep < 2; ++rep) { cpasync_pair_kernel<<<1, 128, 40960>>>(src, 48, lpp, groups, out); get(); }
        printf("  %d lanes (%2d B) per position: %7.1f cycles per group of %d positions = %5.1f B/clk per CTA\n", lpp, 16 * lpp,
               (double)h[0] / groups, 384 / lpp, 384.0 * 16 / ((double)h[0] / groups));
    }
    printf("== fence.proxy.async.shared::cta after a shared store\n");
    for (int threads : {32, 128, 256}) { fence_kernel<<<1, threads>>>(200, out); get(); printf("  %3d threads: %6.1f cycles per (store + fence)\n", threads, (double)h[0] / 200); }
    return 0;
}
