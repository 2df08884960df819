// Input staging on the GPU: the step BEFORE the hot path, fused into one pass.
//
// Replaces, on the device and for a whole batch:
//   ncnn::Mat::from_pixels          /root/reference/src/ncnn/mat_pixel.cpp:27-1131,1329-1367 (u8 interleaved -> fp32 planar,
//                                   with the RGB<->BGR / ->GRAY / RGBA-> conversions of mat.h:126-146)
//   ncnn::Mat::from_pixels_resize   mat_pixel.cpp:1369-1410 -> resize_bilinear_c1/c3/c4, mat_pixel_resize.cpp:26-278 (11-bit
//                                   fixed-point bilinear on the u8 image; restated bit-exactly: same float coefficient
//                                   computation, same >>4 / >>16 / +2 >>2 integer pipeline)
//   ncnn::Mat::substract_mean_normalize   mat.h:159-160, mat.cpp:30-107 (x - mean, x * norm, or x * norm + (-mean * norm))
// One thread per output pixel: 4 source pixels per channel -> resized u8 value -> channel map / gray weights -> mean / norm
// -> coalesced plane stores.  HBM-bound: reads <= 4 B / pixel, writes 4 B / pixel / channel.
#include "preprocess.cuh"

#include "common.cuh"

namespace fcuda {

namespace {

// mat.h:123-146
constexpr int PIXEL_CONVERT_SHIFT = 16;
constexpr int PIXEL_FORMAT_MASK = 0x0000ffff;
constexpr int PIXEL_RGB = 1, PIXEL_BGR = 1 << 1, PIXEL_GRAY = 1 << 2, PIXEL_RGBA = 1 << 3;

struct PixelArgs {
    const unsigned char* src;
    float* dst;
    int src_w, src_h, w, h;      // source and output geometry
    int src_c, out_c;
    int map[4];                  // out channel i <- source channel map[i] (plain copies)
    int gray;                    // 1: single output = (p[map0]*77 + p[map1]*150 + p[map2]*29) >> 8  (mat_pixel.cpp:545-548)
    int resize;
    double scale_x, scale_y;     // (double)src / dst, mat_pixel_resize.cpp:32-33
    float mul[4], add[4];        // y = x * mul + add ; has_mul / has_add select the reference's three variants
    int has_mul, has_add;
    unsigned total;              // batch * h * w
};

// (short) saturate_cast of X +- 0.5 (mat_pixel_resize.cpp:50)
__device__ __forceinline__ int sat_short(float x) {
    int v = static_cast<int>(__fadd_rn(x, x >= 0.f ? 0.5f : -0.5f));
    return max(-32768, min(32767, v));
}

// source offset and the two 11-bit coefficients of one output coordinate (mat_pixel_resize.cpp:52-74)
__device__ __forceinline__ void coef(int d, double scale, int src_n, int& s, int& c0, int& c1) {
    // exact IEEE double operations, no FMA contraction: the reference computes (d + 0.5) * scale - 0.5 in double
    float f = static_cast<float>(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), -0.5));
    s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src_n - 1) { s = src_n - 2; f = 1.f; }
    c0 = sat_short(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c1 = sat_short(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(256)
from_pixels_kernel(const PixelArgs a) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.total) return;
    const unsigned plane = static_cast<unsigned>(a.w) * a.h;
    const unsigned n = idx / plane, rem = idx - n * plane;
    const int y = static_cast<int>(rem / a.w), x = static_cast<int>(rem - static_cast<unsigned>(y) * a.w);
    const unsigned char* img = a.src + static_cast<size_t>(n) * a.src_w * a.src_h * a.src_c;
    int p[4] = {0, 0, 0, 0};
    if (!a.resize) {
        const unsigned char* s = img + (static_cast<size_t>(y) * a.src_w + x) * a.src_c;
        for (int c = 0; c < a.src_c; ++c) p[c] = s[c];
    } else {
        int sx, a0, a1, sy, b0, b1;
        coef(x, a.scale_x, a.src_w, sx, a0, a1);
        coef(y, a.scale_y, a.src_h, sy, b0, b1);
        const unsigned char* s0 = img + (static_cast<size_t>(sy) * a.src_w + sx) * a.src_c;
        const unsigned char* s1 = s0 + static_cast<size_t>(a.src_w) * a.src_c;
        for (int c = 0; c < a.src_c; ++c) {
            // hresize: (S[0]*a0 + S[1]*a1) >> 4 stored as short (:138,160-161); vresize (:273)
            const short r0 = static_cast<short>((s0[c] * a0 + s0[a.src_c + c] * a1) >> 4);
            const short r1 = static_cast<short>((s1[c] * a0 + s1[a.src_c + c] * a1) >> 4);
            const int v = (static_cast<short>((b0 * r0) >> 16) + static_cast<short>((b1 * r1) >> 16) + 2) >> 2;
            p[c] = static_cast<unsigned char>(v);
        }
    }
    float* out = a.dst + static_cast<size_t>(n) * a.out_c * plane + rem;
    if (a.gray) {
        float v = static_cast<float>((p[a.map[0]] * 77 + p[a.map[1]] * 150 + p[a.map[2]] * 29) >> 8);
        if (a.has_mul && a.has_add) v = v * a.mul[0] + a.add[0];
        else if (a.has_mul) v = v * a.mul[0];
        else if (a.has_add) v = v + a.add[0];
        out[0] = v;
    } else {
        for (int c = 0; c < a.out_c; ++c) {
            float v = static_cast<float>(p[a.map[c]]);
            if (a.has_mul && a.has_add) v = v * a.mul[c] + a.add[c];
            else if (a.has_mul) v = v * a.mul[c];
            else if (a.has_add) v = v + a.add[c];
            out[static_cast<size_t>(c) * plane] = v;
        }
    }
}

// mat_pixel.cpp:1329-1367: which source layout a type reads and which planes it produces
int plan_type(int type, PixelArgs* a) {
    const int from = type & PIXEL_FORMAT_MASK, to = type >> PIXEL_CONVERT_SHIFT;
    a->gray = 0;
    for (int i = 0; i < 4; ++i) a->map[i] = i;
    if (from == PIXEL_RGB || from == PIXEL_BGR) a->src_c = 3;
    else if (from == PIXEL_GRAY) a->src_c = 1;
    else if (from == PIXEL_RGBA) a->src_c = 4;
    else return -200;
    if (to == 0) {  // plain copy of the source channels
        a->out_c = a->src_c;
        return 0;
    }
    if ((from == PIXEL_RGB && to == PIXEL_BGR) || (from == PIXEL_BGR && to == PIXEL_RGB) || (from == PIXEL_RGBA && to == PIXEL_BGR)) {
        a->out_c = 3; a->map[0] = 2; a->map[1] = 1; a->map[2] = 0;
        return 0;
    }
    if (from == PIXEL_RGBA && to == PIXEL_RGB) { a->out_c = 3; return 0; }
    if (from == PIXEL_GRAY && (to == PIXEL_RGB || to == PIXEL_BGR)) { a->out_c = 3; a->map[0] = a->map[1] = a->map[2] = 0; return 0; }
    if (to == PIXEL_GRAY && (from == PIXEL_RGB || from == PIXEL_RGBA)) { a->out_c = 1; a->gray = 1; return 0; }           // r g b
    if (to == PIXEL_GRAY && from == PIXEL_BGR) { a->out_c = 1; a->gray = 1; a->map[0] = 2; a->map[1] = 1; a->map[2] = 0; return 0; }
    return -200;  // from_pixels returns an empty Mat for anything else
}

}  // namespace

int pixel_channels(int type, int* src_c, int* out_c) {
    PixelArgs a;
    const int rc = plan_type(type, &a);
    if (rc) return rc;
    if (src_c) *src_c = a.src_c;
    if (out_c) *out_c = a.out_c;
    return 0;
}

int from_pixels(float* out, const unsigned char* pixels, int type, int w, int h, int target_w, int target_h,
                const float* mean_vals, const float* norm_vals, int batch, cudaStream_t s) {
    PixelArgs a;
    int rc = plan_type(type, &a);
    if (rc) return rc;
    if (w <= 0 || h <= 0 || batch < 1) return -100;
    if (target_w <= 0) target_w = w;
    if (target_h <= 0) target_h = h;
    a.src = pixels; a.dst = out;
    a.src_w = w; a.src_h = h; a.w = target_w; a.h = target_h;
    a.resize = (target_w != w || target_h != h) ? 1 : 0;  // mat_pixel.cpp:1371-1372
    if (a.resize && (w < 2 || h < 2)) return -100;
    a.scale_x = static_cast<double>(w) / target_w;
    a.scale_y = static_cast<double>(h) / target_h;
    a.has_mul = norm_vals != nullptr;
    a.has_add = mean_vals != nullptr;
    for (int c = 0; c < 4; ++c) {
        a.mul[c] = 1.f; a.add[c] = 0.f;
        if (c >= a.out_c) continue;
        if (norm_vals) a.mul[c] = norm_vals[c];
        if (mean_vals) a.add[c] = norm_vals ? -mean_vals[c] * norm_vals[c] : -mean_vals[c];  // mat.cpp:48,89
    }
    const unsigned long long total = static_cast<unsigned long long>(batch) * target_w * target_h;
    if (total >= (1ull << 32)) return -100;
    a.total = static_cast<unsigned>(total);
    const int prof = prof_begin(s, PROF_ELEMENTWISE, 0, 0,
                                static_cast<double>(batch) * (static_cast<double>(w) * h * a.src_c +
                                                              4.0 * target_w * target_h * a.out_c));
    from_pixels_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(a);
    prof_end(prof, s);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

}  // namespace fcuda
