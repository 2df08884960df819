// TEST INFRASTRUCTURE ONLY.  extern "C" shim around the UNMODIFIED reference (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libfeather_ref.so).  This file is ours;
// it contains no reference code — it only calls the reference's public API:
//   booster::ConvBooster            /root/reference/src/booster/include/booster/booster.h:151-170
//   feather::Net                    /root/reference/src/net.h:30-70
//   ncnn::ModelBinFromMemory        /root/reference/src/ncnn/modelbin.h:55-65 (weight-blob decoder: fp32 / fp16 / LUT)
//   ncnn::Mat::from_pixels[_resize] /root/reference/src/ncnn/mat.h:149-152 (input staging)
// Used by tests/ (parity oracle), __graft_entry__.smoke() and bench.py's CPU baseline.
#include <booster/booster.h>
#include <ncnn/mat.h>
#include <ncnn/modelbin.h>
#include <net.h>

#include <fcntl.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace {

// The reference printf()s on its hot path (net.cpp:84,104,209,241,310; conv_layer.h:148).
// Silence fd 1/2 around calls so timings are not dominated by terminal I/O.
struct Quiet {
    int saved_out, saved_err;
    Quiet() {
        fflush(stdout);
        fflush(stderr);
        saved_out = dup(1);
        saved_err = dup(2);
        int dn = open("/dev/null", O_WRONLY);
        dup2(dn, 1);
        dup2(dn, 2);
        close(dn);
    }
    ~Quiet() {
        fflush(stdout);
        fflush(stderr);
        dup2(saved_out, 1);
        dup2(saved_err, 2);
        close(saved_out);
        close(saved_err);
    }
};

float* aligned_floats(size_t n) {
    void* p = nullptr;
    if (posix_memalign(&p, 128, (n ? n : 1) * sizeof(float)) != 0) return nullptr;
    memset(p, 0, (n ? n : 1) * sizeof(float));
    return static_cast<float*>(p);
}

void fill_param(booster::ConvParam& cp, const int* p) {
    // order: oc, ic, in_h, in_w, k_h, k_w, s_h, s_w, pad_l, pad_b, pad_r, pad_t, group, bias_term, activation
    cp.output_channels = p[0];
    cp.input_channels = p[1];
    cp.input_h = p[2];
    cp.input_w = p[3];
    cp.kernel_h = p[4];
    cp.kernel_w = p[5];
    cp.stride_h = p[6];
    cp.stride_w = p[7];
    cp.pad_left = p[8];
    cp.pad_bottom = p[9];
    cp.pad_right = p[10];
    cp.pad_top = p[11];
    cp.group = p[12];
    cp.bias_term = p[13] != 0;
    cp.activation = p[14] ? booster::ReLU : booster::None;
    cp.AssignOutputDim();
}

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

}  // namespace

extern "C" {

// One ConvBooster GetBufferSize -> Init -> Forward round, exactly the protocol of
// /root/reference/src/layers/conv_layer.h:92-172.  algo < 0 => SelectAlgo, else ForceSelectAlgo(algo).
// `repeat` > 1 re-runs Forward and reports the mean seconds per Forward in *seconds (may be NULL).
__attribute__((visibility("default")))
int ref_conv_forward(const int* params, int algo, const float* input, const float* weights, const float* bias,
                     float* output, int* out_dims, int repeat, double* seconds) {
    Quiet q;
    booster::ConvParam cp;
    fill_param(cp, params);
    booster::ConvBooster cb;
    int rc = algo < 0 ? cb.SelectAlgo(&cp) : cb.ForceSelectAlgo(static_cast<booster::ConvAlgo>(algo));
    if (rc != 0) return rc;
    int buffer_size = 0, kernel_size = 0;
    cb.GetBufferSize(&cp, &buffer_size, &kernel_size);
    const size_t in_n = static_cast<size_t>(cp.input_channels) * cp.input_h * cp.input_w;
    const size_t out_n = static_cast<size_t>(cp.output_channels) * cp.output_h * cp.output_w;
    const size_t w_n = static_cast<size_t>(cp.output_channels) * cp.input_channels * cp.kernel_h * cp.kernel_w /
                       (cp.group == cp.input_channels && cp.group > 1 ? cp.input_channels : 1);
    float* in = aligned_floats(in_n + 64);
    float* out = aligned_floats(out_n + 64);
    float* w = aligned_floats(w_n + 64);
    float* pk = aligned_floats(static_cast<size_t>(kernel_size) + 64);
    float* buf = aligned_floats(static_cast<size_t>(buffer_size) + 64);
    float* b = aligned_floats(static_cast<size_t>(cp.output_channels) + 64);
    memcpy(in, input, in_n * sizeof(float));
    memcpy(w, weights, w_n * sizeof(float));
    if (bias) memcpy(b, bias, cp.output_channels * sizeof(float));
    cb.Init(&cp, pk, w);
    if (repeat < 1) repeat = 1;
    cb.Forward(&cp, out, in, pk, buf, cp.bias_term ? b : NULL, 1);  // warm-up / result
    if (repeat > 1 || seconds) {
        double t0 = now_s();
        for (int i = 0; i < repeat; ++i) cb.Forward(&cp, out, in, pk, buf, cp.bias_term ? b : NULL, 1);
        if (seconds) *seconds = (now_s() - t0) / repeat;
    }
    memcpy(output, out, out_n * sizeof(float));
    if (out_dims) {
        out_dims[0] = cp.output_channels;
        out_dims[1] = cp.output_h;
        out_dims[2] = cp.output_w;
    }
    free(in); free(out); free(w); free(pk); free(buf); free(b);
    return 0;
}

// Which algorithm the reference's SelectAlgo picks cannot be read back (private member), so the
// tests restate the rule (avx/booster.cpp:283-310) and cross-check numerics with ForceSelectAlgo(NAIVE).

__attribute__((visibility("default")))
void* ref_net_create() {
    Quiet q;
    return new feather::Net();
}

__attribute__((visibility("default")))
void ref_net_destroy(void* h) {
    Quiet q;
    delete static_cast<feather::Net*>(h);
}

__attribute__((visibility("default")))
int ref_net_load(void* h, const char* param_path, const char* bin_path) {
    Quiet q;
    feather::Net* net = static_cast<feather::Net*>(h);
    int rc = net->LoadParam(param_path);
    if (rc != 0) return rc;
    return net->LoadWeights(bin_path);
}

// Feed one CHW image (dense, channel stride h*w) and run Forward().
__attribute__((visibility("default")))
int ref_net_forward(void* h, const char* input_name, const float* chw, int c, int hgt, int w) {
    Quiet q;
    feather::Net* net = static_cast<feather::Net*>(h);
    ncnn::Mat in(w, hgt, c);
    for (int ch = 0; ch < c; ++ch)
        memcpy(in.channel(ch), chw + static_cast<size_t>(ch) * hgt * w, sizeof(float) * hgt * w);
    int rc = net->FeedInput(input_name, in);
    if (rc != 0) return rc;
    return net->Forward();
}

// Raw-pointer Extract (net.cpp:260-278) — the ncnn::Mat overload is broken (net.cpp:291-294).
__attribute__((visibility("default")))
int ref_net_extract(void* h, const char* blob, const float** data, int* n, int* c, int* hgt, int* w) {
    Quiet q;
    feather::Net* net = static_cast<feather::Net*>(h);
    float* p = NULL;
    int rc = net->Extract(std::string(blob), &p, n, c, hgt, w);
    *data = p;
    return rc;
}

// Mean seconds per Forward over `iters` timed runs after one warm-up (includes lazy Init).
__attribute__((visibility("default")))
double ref_net_time_forward(void* h, const char* input_name, const float* chw, int c, int hgt, int w, int iters) {
    if (ref_net_forward(h, input_name, chw, c, hgt, w) != 0) return -1.0;
    Quiet q;
    feather::Net* net = static_cast<feather::Net*>(h);
    double t0 = now_s();
    for (int i = 0; i < iters; ++i) net->Forward();
    return (now_s() - t0) / (iters > 0 ? iters : 1);
}

// Decodes one weight blob of `w` floats from an in-memory .bin image with the reference's own loader
// (modelbin.cpp:204-293).  Returns the number of bytes consumed, or -1 when the loader returned an empty Mat.
long ref_modelbin_load_mem(const unsigned char* buf, int w, int type, float* out) {
    Quiet q;
    const unsigned char* mem = buf;
    ncnn::ModelBinFromMemory mb(mem);
    ncnn::Mat m = mb.load(w, type);
    if (m.empty()) return -1;
    memcpy(out, m.data, sizeof(float) * static_cast<size_t>(w));
    return static_cast<long>(mem - buf);
}

// ncnn::Mat::from_pixels / from_pixels_resize of the reference (mat_pixel.cpp:1329-1410, mat_pixel_resize.cpp) -> dense
// planar floats.  Returns the number of output channels, or -1 when the reference returned an empty Mat.
int ref_from_pixels(const unsigned char* pixels, int type, int w, int h, int target_w, int target_h, float* out) {
    Quiet q;
    ncnn::Mat m = (target_w == w && target_h == h) ? ncnn::Mat::from_pixels(pixels, type, w, h)
                                                   : ncnn::Mat::from_pixels_resize(pixels, type, w, h, target_w, target_h);
    if (m.empty()) return -1;
    const size_t plane = static_cast<size_t>(m.w) * m.h;
    for (int c = 0; c < m.c; ++c) memcpy(out + plane * c, m.channel(c), sizeof(float) * plane);
    return m.c;
}

}  // extern "C"
