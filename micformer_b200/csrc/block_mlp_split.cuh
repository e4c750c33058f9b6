// Shared definitions of the hidden-split, operand-streaming fused MLP kernels (block_mlp_split.cu forward,
// block_mlp_split_bwd.cu backward).
#pragma once
#include <stdint.h>

namespace mic {

constexpr int MS_THREADS = 320;

struct MlpSplitFwdArgs {
    const float* x; float* y;
    const float* gamma; const float* beta; const float* b1; const float* b2;
    const uint8_t* w1_hi; const uint8_t* w1_lo;      // fc1 image: N = HID rows (n_pad1), K = C in 64-wide panels
    const uint8_t* w2_hi; const uint8_t* w2_lo;      // fc2 image: N = C rows (CP), K = HID: panel j = hidden chunk j
    const float* rowscale; int rps;
    int T, C, HID, CP, n_pad1;
    float eps;
};

struct MsF {                                          // forward shared-memory map (bytes)
    static constexpr int XN = 0;                      // 2 slots x (hi 16 KB + lo 16 KB)
    static constexpr int W1 = 65536;                  // 2 slots x (hi 8 KB + lo 8 KB)
    static constexpr int H = 98304;                   // hi 16 KB + lo 16 KB
    static constexpr int W2 = 131072;                 // 2 slots x (hi 16 KB + lo 16 KB)
    static constexpr int PART = 196608;               // LayerNorm partial sums [128][2]
    static constexpr int BAR = PART + 1024;
    static constexpr int SMEM = BAR + 256 + 1024;
};

}  // namespace mic
