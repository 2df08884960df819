/* TEST INFRASTRUCTURE ONLY — see feather_oracle.h.  Plain C restatement of the reference algorithms;
 * every function cites the reference file:line it follows.  Scalar fp32 arithmetic in the same
 * operation order as the reference where that order is observable (Winograd transforms). */
#include "feather_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * ConvParam helpers
 * ---------------------------------------------------------------------------------------------- */
void oracle_assign_output_dim(OracleConvParam* p) {
    /* booster.h:113-125 */
    if (p->group == 0) p->group = 1;
    if (p->stride_h == 0) p->stride_h = 1;
    if (p->stride_w == 0) p->stride_w = 1;
    p->output_h = (p->input_h + p->pad_top + p->pad_bottom - p->kernel_h) / p->stride_h + 1;
    p->output_w = (p->input_w + p->pad_left + p->pad_right - p->kernel_w) / p->stride_w + 1;
    if (p->group == p->input_channels) p->output_channels = p->input_channels;
}

int oracle_select_algo(const OracleConvParam* p) {
    /* avx/booster.cpp:283-310 */
    if (p->group == p->input_channels) return ORACLE_DEPTHWISE;
    if (p->group == 1 && p->kernel_h == 3 && p->kernel_w == 3 && p->stride_h == 1 && p->stride_w == 1 &&
        p->input_h > 8 && p->input_w > 8 && p->output_channels % 4 == 0 && p->input_channels % 4 == 0)
        return ORACLE_WINOGRADF63;
    if (p->group == 1) return ORACLE_IM2COL;
    return -1;
}

static void pad_input(float* padded, const float* input, int channels, int w, int h, int pl, int pt, int pr, int pb) {
    /* generic_kernels.cpp:31-48 */
    const int pw = w + pl + pr, ph = h + pt + pb;
    memset(padded, 0, sizeof(float) * (size_t)pw * ph * channels);
    for (int c = 0; c < channels; ++c)
        for (int i = 0; i < h; ++i)
            memcpy(padded + ((size_t)c * ph + pt + i) * pw + pl, input + ((size_t)c * h + i) * w, sizeof(float) * w);
}

/* ------------------------------------------------------------------------------------------------
 * Direct convolution (NAIVE / IM2COL semantics)
 * ---------------------------------------------------------------------------------------------- */
void oracle_conv_direct(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                        float* output, int accumulate_double) {
    const int OC = p->output_channels, IC = p->input_channels, OH = p->output_h, OW = p->output_w;
    const int KH = p->kernel_h, KW = p->kernel_w, IH = p->input_h, IW = p->input_w;
#pragma omp parallel for collapse(2) schedule(static)
    for (int oc = 0; oc < OC; ++oc) {
        for (int i = 0; i < OH; ++i) {
            for (int j = 0; j < OW; ++j) {
                double accd = 0.0;
                float accf = 0.f;
                for (int k = 0; k < IC; ++k)
                    for (int u = 0; u < KH; ++u) {
                        /* generic_kernels.cpp:66-67 */
                        const int row = u - p->pad_top + i * p->stride_h;
                        if (row < 0 || row >= IH) continue;
                        for (int v = 0; v < KW; ++v) {
                            const int col = v - p->pad_left + j * p->stride_w;
                            if (col < 0 || col >= IW) continue;
                            const float x = input[((size_t)k * IH + row) * IW + col];
                            const float w = weights[(((size_t)oc * IC + k) * KH + u) * KW + v];
                            if (accumulate_double) accd += (double)x * (double)w;
                            else accf += x * w;
                        }
                    }
                float r = accumulate_double ? (float)accd : accf;
                if (p->bias_term && bias) r += bias[oc];
                if (p->activation) r = r > 0.f ? r : 0.f;
                output[((size_t)oc * OH + i) * OW + j] = r;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Winograd F(6,3)
 * ---------------------------------------------------------------------------------------------- */
/* G ("ktm"), winograd_kernels_F63.cpp:191-201 — rows 5/6 are the textbook rows divided by 32. */
static const float kG[8][3] = {
    {1.0f, 0.0f, 0.0f},
    {-2.0f / 9, -2.0f / 9, -2.0f / 9},
    {-2.0f / 9, 2.0f / 9, -2.0f / 9},
    {1.0f / 90, 1.0f / 45, 2.0f / 45},
    {1.0f / 90, -1.0f / 45, 2.0f / 45},
    {1.0f / 45, 1.0f / 90, 1.0f / 180},
    {1.0f / 45, -1.0f / 90, 1.0f / 180},
    {0.0f, 0.0f, 1.0f}};

/* U = G g G^T.  winograd_kernels_F63.cpp:222-254 computes mid = G g (naive_gemm_temp 8x3x3), transposes it
 * and multiplies by G again, i.e. it stores bigBlock = G (G g)^T = U^T; its SSE input transform produces V^T to
 * match.  Those storage orientations are private to the AVX backend (SURVEY.md §8 preamble); this restatement keeps
 * the same arithmetic (same products, same fp32 summation order over k) in the textbook orientation. */
static void kernel_transform(const float* g, float* U /*8x8*/) {
    float mid[8][3];
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += kG[i][k] * g[k * 3 + j];
            mid[i][j] = s;
        }
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            float s = 0.f;
            for (int k = 0; k < 3; ++k) s += kG[i][k] * mid[j][k]; /* bigBlock[i][j] of the reference */
            U[j * 8 + i] = s;                                       /* == (G g G^T)[j][i] */
        }
}

/* 1-D B^T, winograd_kernels_F63.cpp:272-325, same operation order. */
static void input_transform_1d(const float r[8], float o[8]) {
    const float r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4], r5 = r[5], r6 = r[6], r7 = r[7];
    o[0] = (r0 - r6) + (r4 - r2) * 5.25f;
    o[7] = (r7 - r1) + (r3 - r5) * 5.25f;
    const float t1 = (r2 + r6) - r4 * 4.25f;
    const float t2 = (r1 + r5) - r3 * 4.25f;
    const float s1 = r4 * 1.25f;
    const float s2 = r3 * 2.5f;
    float p1 = r6 + (r2 * 0.25f - s1);
    float p2 = (r1 * 0.5f - s2) + r5 * 2.f;
    o[3] = p1 + p2;
    o[4] = p1 - p2;
    p1 = r6 + (r2 - s1) * 4.f;
    p2 = (r1 * 2.f - s2) + r5 * 0.5f;
    o[5] = p1 + p2;
    o[6] = p1 - p2;
    o[1] = t1 + t2;
    o[2] = t1 - t2;
}

/* 1-D A^T, winograd_kernels_F63.cpp:1040-1045 */
static void output_transform_1d(const float m[8], float s[6]) {
    const float a12 = m[1] + m[2], s12 = m[1] - m[2];
    const float a34 = m[3] + m[4], s34 = m[3] - m[4];
    const float a56 = m[5] + m[6], s56 = m[5] - m[6];
    s[0] = m[0] + a12 + a34 + 32.f * a56;
    s[1] = s12 + 2.f * s34 + 16.f * s56;
    s[2] = a12 + 4.f * a34 + 8.f * a56;
    s[3] = s12 + 8.f * s34 + 4.f * s56;
    s[4] = a12 + 16.f * a34 + 2.f * a56;
    s[5] = s12 + 32.f * s34 + s56 + m[7];
}

void oracle_conv_winograd_f63(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                              float* output) {
    const int IC = p->input_channels, OC = p->output_channels;
    const int Hp = p->input_h + p->pad_top + p->pad_bottom, Wp = p->input_w + p->pad_left + p->pad_right;
    const int OH = Hp - 2, OW = Wp - 2;
    const int nRow = (Wp + 3) / 6; /* tiles along x, winograd_kernels_F63.cpp:2320 */
    const int nCol = (Hp + 3) / 6; /* tiles along y */
    const int nB = nRow * nCol;
    float* padded = (float*)malloc(sizeof(float) * (size_t)IC * Hp * Wp);
    pad_input(padded, input, IC, p->input_w, p->input_h, p->pad_left, p->pad_top, p->pad_right, p->pad_bottom);
    float* U = (float*)malloc(sizeof(float) * 64 * (size_t)IC * OC); /* [oc][ic][64] */
    float* V = (float*)malloc(sizeof(float) * 64 * (size_t)IC * nB); /* [ic][tile][64] */
    for (int oc = 0; oc < OC; ++oc)
        for (int ic = 0; ic < IC; ++ic) kernel_transform(weights + ((size_t)oc * IC + ic) * 9, U + ((size_t)oc * IC + ic) * 64);
#pragma omp parallel for schedule(static)
    for (int ic = 0; ic < IC; ++ic)
        for (int ty = 0; ty < nCol; ++ty)
            for (int tx = 0; tx < nRow; ++tx) {
                float d[8][8], tmp[8][8];
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x) {
                        const int yy = ty * 6 + y, xx = tx * 6 + x;
                        d[y][x] = (yy < Hp && xx < Wp) ? padded[((size_t)ic * Hp + yy) * Wp + xx] : 0.f; /* :378-414 */
                    }
                /* rows (along y index: combine the 8 row-vectors), then columns */
                for (int x = 0; x < 8; ++x) {
                    float col[8], o[8];
                    for (int y = 0; y < 8; ++y) col[y] = d[y][x];
                    input_transform_1d(col, o);
                    for (int y = 0; y < 8; ++y) tmp[y][x] = o[y];
                }
                float* v = V + ((size_t)ic * nB + ty * nRow + tx) * 64;
                for (int y = 0; y < 8; ++y) {
                    float o[8];
                    input_transform_1d(tmp[y], o);
                    for (int x = 0; x < 8; ++x) v[y * 8 + x] = o[x];
                }
            }
#pragma omp parallel for schedule(static)
    for (int oc = 0; oc < OC; ++oc) {
        for (int t = 0; t < nB; ++t) {
            float M[64];
            for (int e = 0; e < 64; ++e) M[e] = 0.f;
            /* TensorGEMM, :518-757: M_e[oc,tile] = sum_ic U_e[oc,ic] * V_e[ic,tile], fp32, ic ascending. */
            for (int ic = 0; ic < IC; ++ic) {
                const float* u = U + ((size_t)oc * IC + ic) * 64;
                const float* v = V + ((size_t)ic * nB + t) * 64;
                for (int e = 0; e < 64; ++e) M[e] += u[e] * v[e];
            }
            /* output transform: Y = A^T M A, then bias, ReLU, clipped store (:1088-1269) */
            float tmp[6][8];
            for (int x = 0; x < 8; ++x) {
                float col[8], s[6];
                for (int y = 0; y < 8; ++y) col[y] = M[y * 8 + x];
                output_transform_1d(col, s);
                for (int y = 0; y < 6; ++y) tmp[y][x] = s[y];
            }
            const int ty = t / nRow, tx = t % nRow;
            for (int y = 0; y < 6; ++y) {
                float s[6];
                output_transform_1d(tmp[y], s);
                for (int x = 0; x < 6; ++x) {
                    const int oy = ty * 6 + y, ox = tx * 6 + x;
                    if (oy >= OH || ox >= OW) continue;
                    float r = s[x];
                    if (p->bias_term && bias) r += bias[oc];
                    if (p->activation) r = r > 0.f ? r : 0.f;
                    output[((size_t)oc * OH + oy) * OW + ox] = r;
                }
            }
        }
    }
    free(padded);
    free(U);
    free(V);
}

/* ------------------------------------------------------------------------------------------------
 * Depthwise
 * ---------------------------------------------------------------------------------------------- */
void oracle_conv_depthwise(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                           float* output) {
    const int C = p->input_channels;
    const int inh = p->input_h + p->pad_top + p->pad_bottom, inw = p->input_w + p->pad_left + p->pad_right;
    float* padded = (float*)malloc(sizeof(float) * (size_t)C * inh * inw);
    pad_input(padded, input, C, p->input_w, p->input_h, p->pad_left, p->pad_top, p->pad_right, p->pad_bottom);
    const int kw = p->kernel_w, kh = p->kernel_h;
    if (kw == inw && kh == inh) {
        /* globalDwConv, depthwise.cpp:30-55 (group == channels => k = 0-based block arithmetic reduces to i) */
        const int step = inw * inh;
        for (int i = 0; i < C; ++i) {
            float s = 0.f;
            for (int j = 0; j < step; ++j) s += padded[(size_t)i * step + j] * weights[(size_t)i * step + j];
            if (p->bias_term && bias) s += bias[i];
            if (p->activation) s = s > 0.f ? s : 0.f;
            output[i] = s;
        }
    } else {
        const int outw = (inw - kw) / p->stride_w + 1, outh = (inh - kh) / p->stride_h + 1;
        for (int g = 0; g < C; ++g)
            for (int i = 0; i < outh; ++i)
                for (int j = 0; j < outw; ++j) {
                    /* depthwise.cpp:184 — stride_w multiplies the row index and stride_h the column index */
                    const float* inp = padded + (size_t)g * inw * inh + (size_t)inw * (i * p->stride_w) + (j * p->stride_h);
                    float s = 0.f;
                    for (int m = 0; m < kh; ++m)
                        for (int n = 0; n < kw; ++n) s += inp[m * inw + n] * weights[(size_t)g * kw * kh + m * kw + n];
                    if (p->bias_term && bias) s += bias[g];
                    if (p->activation) s = s > 0.f ? s : 0.f;
                    output[((size_t)g * outh + i) * outw + j] = s;
                }
    }
    free(padded);
}

int oracle_conv_forward(const OracleConvParam* p, const float* input, const float* weights, const float* bias,
                        float* output) {
    const int algo = oracle_select_algo(p);
    switch (algo) {
        case ORACLE_DEPTHWISE: oracle_conv_depthwise(p, input, weights, bias, output); break;
        case ORACLE_WINOGRADF63: oracle_conv_winograd_f63(p, input, weights, bias, output); break;
        case ORACLE_IM2COL: oracle_conv_direct(p, input, weights, bias, output, 0); break;
        default: return -1;
    }
    return algo;
}

/* ------------------------------------------------------------------------------------------------
 * Pooling
 * ---------------------------------------------------------------------------------------------- */
int oracle_pool_out_dim(int in, int pad_a, int pad_b, int kernel, int stride) {
    /* pooling_layer.h:129-130 */
    return (int)ceilf((float)(in + pad_a + pad_b - kernel) / (float)stride) + 1;
}

void oracle_pooling(const float* input, int channels, int in_h, int in_w, int type, int kernel_h, int kernel_w,
                    int stride_h, int stride_w, int pad_left, int pad_right, int pad_top, int pad_bottom,
                    int global_pooling, float* output) {
    int out_h, out_w;
    if (global_pooling) {
        kernel_h = in_h; kernel_w = in_w; out_h = 1; out_w = 1; /* :114-121 */
    } else {
        out_h = oracle_pool_out_dim(in_h, pad_top, pad_bottom, kernel_h, stride_h);
        out_w = oracle_pool_out_dim(in_w, pad_left, pad_right, kernel_w, stride_w);
    }
    for (int c = 0; c < channels; ++c)
        for (int j = 0; j < out_h; ++j) {
            const int tmp_pos = j * stride_h - pad_top - pad_bottom; /* :56 double pad subtraction */
            const int x_min = tmp_pos > 0 ? tmp_pos : 0;
            const int x_max = (tmp_pos + kernel_h) < in_h ? (tmp_pos + kernel_h) : in_h;
            for (int k = 0; k < out_w; ++k) {
                const int local_pos = k * stride_w - pad_left - pad_right; /* :67 */
                const int y_min = local_pos > 0 ? local_pos : 0;
                const int y_max = (local_pos + kernel_w) < in_w ? (local_pos + kernel_w) : in_w;
                int counter = 0;
                float total = type != 0 ? 0.f : -FLT_MAX;
                for (int x = x_min; x < x_max; ++x)
                    for (int y = y_min; y < y_max; ++y) {
                        const float v = input[((size_t)c * in_h + x) * in_w + y];
                        if (type != 0) { total += v; counter++; }
                        else total = total > v ? total : v;
                    }
                float* o = output + ((size_t)c * out_h + j) * out_w + k;
                if (type != 0) *o = 0.f + total / counter; /* :84 */
                else *o = (-FLT_MAX > total) ? -FLT_MAX : total;
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * Dense and element-wise layers
 * ---------------------------------------------------------------------------------------------- */
void oracle_inner_product(const float* x, const float* w, const float* bias, int in_size, int out_size, int relu,
                          float* z) {
    /* sgemv.cpp:317-332 */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < out_size; ++i) {
        float sum = 0.f;
        for (int j = 0; j < in_size; ++j) sum += x[j] * w[(size_t)i * in_size + j];
        if (bias) sum += bias[i];
        if (relu) sum = sum > 0.f ? sum : 0.f;
        z[i] = sum;
    }
}

void oracle_batchnorm(const float* input, int channels, int stride, const float* slope, const float* mean,
                      const float* var, const float* bias, float eps, float* output) {
    for (int i = 0; i < channels; ++i) {
        /* batchnorm_layer.h:70-75 */
        const float sqrt_var = sqrtf(var[i] + eps);
        const float alpha = bias[i] - slope[i] * mean[i] / sqrt_var;
        const float beta = slope[i] / sqrt_var;
        for (int j = 0; j < stride; ++j) /* generic_kernels.cpp:266 */
            output[(size_t)i * stride + j] = beta * input[(size_t)i * stride + j] + alpha;
    }
}

void oracle_scale(const float* input, int channels, int stride, const float* scale, const float* bias, float* output) {
    for (int i = 0; i < channels; ++i)
        for (int j = 0; j < stride; ++j) {
            float v = input[(size_t)i * stride + j] * scale[i]; /* generic_kernels.cpp:224-228 */
            if (bias) v = v + bias[i];
            output[(size_t)i * stride + j] = v;
        }
}

void oracle_eltwise_add(const float* a, const float* b, long n, int relu, float* out) {
    for (long i = 0; i < n; ++i) {
        float s = a[i] + b[i];
        if (relu) s = s > 0.f ? s : 0.f;
        out[i] = s;
    }
}

void oracle_relu(const float* in, long n, float* out) {
    for (long i = 0; i < n; ++i) out[i] = in[i] > 0 ? in[i] : 0;
}

void oracle_softmax(const float* in, long n, float* out) {
    /* softmax_layer.h:40-53 */
    float sum = 0.f, mx = -FLT_MAX;
    for (long i = 0; i < n; ++i) mx = mx > in[i] ? mx : in[i];
    for (long i = 0; i < n; ++i) {
        out[i] = (float)exp(in[i] - mx);
        sum += out[i];
    }
    for (long i = 0; i < n; ++i) out[i] = out[i] / sum;
}

void oracle_dropout(const float* in, long n, float scale, float* out) {
    for (long i = 0; i < n; ++i) out[i] = scale == 1.f ? in[i] : in[i] * scale;
}
