/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's CPU algorithms for the
 * convolution / dense / pooling / activation hot path.  Never linked into,
 * imported by or executed from the product path (feathercnn_b200/): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Parity pin: tests/test_oracle.py checks every function here against the UNMODIFIED reference
 * compiled into oracle/_ref/libfeather_ref.so (built from /root/reference by oracle/Makefile) and
 * against the committed fixtures in tests/golden/ (generated from that build by
 * tests/golden/make_golden.py).  The reference itself ships no golden vectors.
 */
#ifndef FEATHER_ORACLE_H
#define FEATHER_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Field-for-field mirror of booster::ConvParam's integer geometry
 * (/root/reference/src/booster/include/booster/booster.h:59-77). */
typedef struct {
    int output_channels, input_channels, input_h, input_w;
    int kernel_h, kernel_w, output_h, output_w;
    int stride_h, stride_w;
    int pad_left, pad_bottom, pad_right, pad_top;
    int group;
    int bias_term;
    int activation; /* 0 None, 1 ReLU */
} OracleConvParam;

enum { ORACLE_NAIVE = 0, ORACLE_IM2COL = 1, ORACLE_SGECONV = 2, ORACLE_DEPTHWISE = 3, ORACLE_WINOGRADF63 = 4,
       ORACLE_WINOGRADF63FUSED = 5, ORACLE_WINOGRADF23 = 6 };

/* booster.h:113-125 */
void oracle_assign_output_dim(OracleConvParam* p);
/* avx/booster.cpp:283-310; returns the ConvAlgo enum value or -1 (partial groups). */
int oracle_select_algo(const OracleConvParam* p);

/* Direct convolution with the im2col index convention of generic_kernels.cpp:50-85
 * (row = u - pad_top + i*stride_h, col = v - pad_left + j*stride_w), group == 1.
 * accumulate_double != 0 uses fp64 accumulators (second opinion); else fp32 in (ic,u,v) order like naive_sgemm. */
void oracle_conv_direct(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                        float* output, int accumulate_double);

/* Winograd F(6,3), restating winograd_kernels_F63.cpp: G table (191-201), B^T (272-325), A^T (1040-1045),
 * tiling nRow=(Wp+3)/6, nCol=(Hp+3)/6 with zero-extended edge tiles (378-414) and clipped stores (1200-1231). */
void oracle_conv_winograd_f63(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                              float* output);

/* avx/depthwise.cpp:161-207 (incl. the stride_w/stride_h swap at :184) and 30-55 (global case),
 * after pad_input (generic_kernels.cpp:31-48).  weights: (C, kh, kw). */
void oracle_conv_depthwise(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                           float* output);

/* Dispatch exactly like ConvBooster::SelectAlgo + Forward. Returns algo or -1. */
int oracle_conv_forward(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                        float* output);

/* pooling_layer.h:38-134.  type 0 = max, else average over in-bounds count.  Output dims via
 * oracle_pool_out_dim (ceil mode, :129-130).  Window start double-subtracts pad (:56,:67). */
int oracle_pool_out_dim(int in, int pad_a, int pad_b, int kernel, int stride);
void oracle_pooling(const float* input, int channels, int in_h, int in_w, int type, int kernel_h, int kernel_w,
                    int stride_h, int stride_w, int pad_left, int pad_right, int pad_top, int pad_bottom,
                    int global_pooling, float* output);

/* inner_product_layer.h:33-170 -> sgemv.cpp:317-395:  z = W x (+ b) (ReLU). W is (out, in) row-major. */
void oracle_inner_product(const float* x, const float* w, const float* bias, int in_size, int out_size, int relu,
                          float* z);

/* batchnorm_layer.h:43-77 fold + generic_kernels.cpp:237-279 apply: y = beta*x + alpha. */
void oracle_batchnorm(const float* input, int channels, int stride, const float* slope, const float* mean,
                      const float* var, const float* bias, float eps, float* output);
/* scale_layer.h / generic_kernels.cpp:203-233 */
void oracle_scale(const float* input, int channels, int stride, const float* scale, const float* bias, float* output);
/* eltwise_layer.h:68-82 (SUM only) */
void oracle_eltwise_add(const float* a, const float* b, long n, int relu, float* out);
/* relu_layer.h:29-41 */
void oracle_relu(const float* in, long n, float* out);
/* softmax_layer.h:32-55 — over the whole blob */
void oracle_softmax(const float* in, long n, float* out);
/* dropout_layer.h:36-57 */
void oracle_dropout(const float* in, long n, float scale, float* out);

#ifdef __cplusplus
}
#endif
#endif
