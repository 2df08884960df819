// GPU input staging: from_pixels (+resize) fused with substract_mean_normalize (see preprocess.cu).
#pragma once
#include <cuda_runtime.h>
namespace fcuda {
// Source / output channel counts of an ncnn pixel type (mat.h:126-146); -200 for an unknown type.
int pixel_channels(int type, int* src_c, int* out_c);
// pixels: DEVICE u8, `batch` images of w*h*src_c bytes; out: (batch, out_c, target_h, target_w) fp32.  target <= 0 keeps
// the source size.  mean_vals / norm_vals: HOST arrays of out_c floats or null.
int from_pixels(float* out, const unsigned char* pixels, int type, int w, int h, int target_w, int target_h,
                const float* mean_vals, const float* norm_vals, int batch, cudaStream_t s);
}  // namespace fcuda
