// Packing kernels for TensorGEMM operands (see pack.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace fcuda {

struct PackGeom {
    int IC, H, W;
    int KH, KW;
    int OH, OW;
    int stride_h, stride_w;
    int pad_top, pad_left;
    int K;   // IC*KH*KW
    int Kp;  // K rounded up to a multiple of 4
};

// Rows [m0, m0+rows) of the im2col matrix (global pixel index m = img*OH*OW + oy*OW + ox) -> P[rows][Kp].
int im2col_pack(const float* in, float* P_hi, float* P_lo, const PackGeom& g, long long m0, int rows, cudaStream_t s);
// W[rows][K] -> hi/lo planes [rows][Kp] (zero padded).
int pack_weights(const float* w, float* W_hi, float* W_lo, int rows, int K, int Kp, cudaStream_t s);
// hi/lo split of n floats (lo may be null: only the TF32-rounded plane is written).
int split_tf32_planes(const float* x, float* hi, float* lo, size_t n, cudaStream_t s);

}  // namespace fcuda
