#pragma once
#include <cuda_runtime.h>
namespace fcuda {
struct DwGeom {
    int C, H, W, KH, KW, OH, OW, stride_h, stride_w, pad_top, pad_left;
};
// out[n][c] = act(in[n][c] (*) w[c] + bias[c]); w is (C, KH, KW).
int depthwise_forward(const float* in, const float* w, const float* bias, float* out, const DwGeom& g, int relu,
                      int batch, cudaStream_t s);
}  // namespace fcuda
