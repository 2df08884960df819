// Implicit-GEMM convolution on tcgen05 (see conv_igemm.cu).
#pragma once
#include <cuda_runtime.h>

namespace fcuda {

struct IgemmProblem {
    const float* input;   // (N, IC, H, W) fp32 NCHW
    const float* w_hi;    // packed filters [tap][OC][IC], TF32-exact plane
    const float* w_lo;    // fp32 remainder plane (planes == 2) or null
    const float* bias;    // OC floats or null
    float* output;        // (N, OC, OH, OW)
    int N, IC, H, W, OC, OH, OW;
    int KH, KW, pad_top, pad_left;   // stride is 1
    int planes;           // 1 = TF32, 2 = 3xTF32
    int relu;
};

// stride 1, W % 4 == 0 and IC % 4 == 0 (TMA's 16-byte stride rule), 16-byte aligned input.
bool conv_igemm_supported(int IC, int W, int stride_h, int stride_w, const void* input);
// raw (OC, IC, KH, KW) -> [tap][OC][IC] hi (+ lo) planes.
int conv_igemm_pack_weights(const float* w, float* w_hi, float* w_lo, int OC, int IC, int taps, cudaStream_t s);
int conv_igemm_forward(const IgemmProblem& p, cudaStream_t stream);

}  // namespace fcuda
