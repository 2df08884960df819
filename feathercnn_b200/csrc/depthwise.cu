// Depthwise convolution (group == channels).
//
// Replaces DEPTHWISE_Forward (/root/reference/src/booster/avx/booster.cpp:136-160) -> dwConv_template /
// globalDwConv (avx/depthwise.cpp:161-207, 30-55) and the hand-unrolled NEON dwConvs1/dwConvs2
// (arm/depthwise.cpp:166,803).  The op is HBM-bound (18 FLOP per 8 bytes): the 3x3 kernels keep a 3-row
// sliding window in registers, read every input element once per warp with coalesced row loads and fetch
// the horizontal neighbours with warp shuffles (only the two edge lanes issue an extra halo load).
// Zero padding comes from the bounds checks — no padded copy (the reference's pad_input) is made.
// Stride semantics are the geometric ones; the reference swaps stride_w/stride_h in its index
// (depthwise.cpp:184), which is identical whenever stride_h == stride_w (all models in scope).
#include "depthwise.cuh"
#include "common.cuh"

namespace fcuda {

constexpr int kDwRows = 16;  // output rows marched by one warp

template <int STRIDE>
__global__ void __launch_bounds__(128)
dw3x3_shuffle_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                     float* __restrict__ out, int C, int H, int W, int OH, int OW, int relu, long long items,
                     int xstrips, int ystrips) {
    const long long item = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
    if (item >= items) return;
    const int lane = threadIdx.x & 31;
    const int xs = static_cast<int>(item % xstrips);
    const long long t = item / xstrips;
    const int ys = static_cast<int>(t % ystrips);
    const long long plane = t / ystrips;  // n * C + c
    const int c = static_cast<int>(plane % C);
    const float* ip = in + plane * H * W;
    float* op = out + plane * OH * OW;

    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
    const float bv = bias ? __ldg(bias + c) : 0.f;

    const int ox = xs * 32 + lane;
    const int oy0 = ys * kDwRows;
    const int oy1 = min(oy0 + kDwRows, OH);

    // The op is latency-bound unless many row loads are in flight per warp (measured: one row at a time = 29 % of HBM
    // bandwidth with 40 resident warps).  So input rows are fetched in batches — all global loads of a batch are issued
    // before the first shuffle (a shuffle waits for its load and would serialise the loads behind it) — and the
    // horizontal neighbours come from warp shuffles; only lanes 0 / 31 issue one extra (halo) load per row.
    auto fma3 = [&](float v, int kr, float l, float m, float r) {
        v = fmaf(k[kr * 3 + 0], l, v);
        v = fmaf(k[kr * 3 + 1], m, v);
        return fmaf(k[kr * 3 + 2], r, v);
    };
    if (STRIDE == 1) {
        constexpr int NB = 4;  // output rows per batch == new input rows per batch
        const int ix = ox;     // centre column (pad_left == 1)
        const int hx = lane == 0 ? ix - 1 : ix + 1;  // halo column of the edge lanes
        const bool m_ok = ix < W, h_ok = (lane == 0 || lane == 31) && hx >= 0 && hx < W;
        auto fetch = [&](int iy, float& m, float& h) {
            const bool row_ok = iy >= 0 && iy < H;
            const float* rp = ip + static_cast<long long>(iy) * W;
            m = (row_ok && m_ok) ? __ldg(rp + ix) : 0.f;
            h = (row_ok && h_ok) ? __ldg(rp + hx) : 0.f;
        };
        auto spread = [&](float m, float h, float& l, float& r) {
            l = __shfl_up_sync(0xffffffffu, m, 1);
            r = __shfl_down_sync(0xffffffffu, m, 1);
            if (lane == 0) l = h;
            if (lane == 31) r = h;
        };
        float am, ah, bm, bh, al, ar, bl, br;
        fetch(oy0 - 1, am, ah);
        fetch(oy0, bm, bh);
        spread(am, ah, al, ar);
        spread(bm, bh, bl, br);
        for (int oy = oy0; oy < oy1; oy += NB) {
            float cm[NB], ch[NB], cl[NB], cr[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) fetch(oy + 1 + i, cm[i], ch[i]);
#pragma unroll
            for (int i = 0; i < NB; ++i) spread(cm[i], ch[i], cl[i], cr[i]);
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                // window rows: (a, b, c0), (b, c0, c1), (c0, c1, c2), (c1, c2, c3)
                const float tl = i == 0 ? al : i == 1 ? bl : cl[i - 2], tm = i == 0 ? am : i == 1 ? bm : cm[i - 2],
                            tr = i == 0 ? ar : i == 1 ? br : cr[i - 2];
                const float ml = i == 0 ? bl : cl[i - 1], mm = i == 0 ? bm : cm[i - 1], mr = i == 0 ? br : cr[i - 1];
                float v = fma3(bv, 0, tl, tm, tr);
                v = fma3(v, 1, ml, mm, mr);
                v = fma3(v, 2, cl[i], cm[i], cr[i]);
                if (relu) v = fmaxf(v, 0.f);
                if (ox < OW && oy + i < oy1) op[static_cast<long long>(oy + i) * OW + ox] = v;
            }
            al = cl[NB - 2]; am = cm[NB - 2]; ar = cr[NB - 2];
            bl = cl[NB - 1]; bm = cm[NB - 1]; br = cr[NB - 1];
        }
    } else {
        constexpr int NB = 2;  // output rows per batch (4 new input rows)
        const int ix = 2 * ox;  // centre column 2*ox - 1 + 1
        const bool m_ok = ix < W, r_ok = ix + 1 < W, h_ok = lane == 0 && ix - 1 >= 0 && ix - 1 < W;
        auto fetch = [&](int iy, float& m, float& r, float& h) {
            const bool row_ok = iy >= 0 && iy < H;
            const float* rp = ip + static_cast<long long>(iy) * W;
            m = (row_ok && m_ok) ? __ldg(rp + ix) : 0.f;
            r = (row_ok && r_ok) ? __ldg(rp + ix + 1) : 0.f;
            h = (row_ok && h_ok) ? __ldg(rp + ix - 1) : 0.f;
        };
        auto spread = [&](float r, float h, float& l) {
            l = __shfl_up_sync(0xffffffffu, r, 1);
            if (lane == 0) l = h;
        };
        float am, ar, ah, al;
        fetch(2 * oy0 - 1, am, ar, ah);
        spread(ar, ah, al);
        for (int oy = oy0; oy < oy1; oy += NB) {
            float cm[2 * NB], cr[2 * NB], ch[2 * NB], cl[2 * NB];
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) fetch(2 * oy + i, cm[i], cr[i], ch[i]);
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) spread(cr[i], ch[i], cl[i]);
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                // output row oy+i reads input rows 2(oy+i)-1, 2(oy+i), 2(oy+i)+1 = (a | c[2i-1]), c[2i], c[2i+1]
                const float tl = i == 0 ? al : cl[2 * i - 1], tm = i == 0 ? am : cm[2 * i - 1], tr = i == 0 ? ar : cr[2 * i - 1];
                float v = fma3(bv, 0, tl, tm, tr);
                v = fma3(v, 1, cl[2 * i], cm[2 * i], cr[2 * i]);
                v = fma3(v, 2, cl[2 * i + 1], cm[2 * i + 1], cr[2 * i + 1]);
                if (relu) v = fmaxf(v, 0.f);
                if (ox < OW && oy + i < oy1) op[static_cast<long long>(oy + i) * OW + ox] = v;
            }
            al = cl[2 * NB - 1]; am = cm[2 * NB - 1]; ar = cr[2 * NB - 1];
        }
    }
}

// 3x3 depthwise on small planes (MobileNet's 14x14 and 7x7 stages, where a 32-wide strip would idle most lanes): one warp
// per (image, channel) plane.  The plane goes to shared memory with a zero halo of one pixel, then every lane computes
// output pixels lane, lane+32, ...; row/column come from a 16-bit reciprocal multiply (exact for planes <= 34x34).
template <int STRIDE>
__global__ void __launch_bounds__(128)
dw3x3_plane_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, int C, int H, int W, int OH, int OW, int pad_top, int pad_left, int relu,
                   long long planes) {
    extern __shared__ float dw_smem[];
    const int PW = W + 2, PH = H + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long plane = static_cast<long long>(blockIdx.x) * 4 + warp;
    if (plane >= planes) return;
    float* tile = dw_smem + warp * PH * PW;
    const int c = static_cast<int>(plane % C);
    const float* ip = in + plane * H * W;
    const unsigned rcp_pw = (65536u + PW - 1) / PW, rcp_ow = (65536u + OW - 1) / OW;
    // eight loads in flight per lane, then the eight shared-memory stores (a store right behind its load would make the
    // next load wait for it)
    const unsigned n_tile = static_cast<unsigned>(PH * PW);
    for (unsigned base = 0; base < n_tile; base += 256) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned i = base + j * 32 + lane;
            const unsigned y = (i * rcp_pw) >> 16, x = i - y * PW;
            const int iy = static_cast<int>(y) - 1, ix = static_cast<int>(x) - 1;
            v[j] = (i < n_tile && iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(ip + iy * W + ix) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned i = base + j * 32 + lane;
            if (i < n_tile) tile[i] = v[j];
        }
    }
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
    const float bv = bias ? __ldg(bias + c) : 0.f;
    __syncwarp();
    float* op = out + plane * OH * OW;
    for (unsigned o = lane; o < static_cast<unsigned>(OH * OW); o += 32) {
        const unsigned oy = (o * rcp_ow) >> 16, ox = o - oy * OW;
        // window start in image coordinates is (oy*S - pad_top, ox*S - pad_left); +1 for the halo
        const float* t = tile + (oy * STRIDE + 1 - pad_top) * PW + (ox * STRIDE + 1 - pad_left);
        float v = bv;
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int x = 0; x < 3; ++x) v = fmaf(k[u * 3 + x], t[u * PW + x], v);
        if (relu) v = fmaxf(v, 0.f);
        op[o] = v;
    }
}

// Generic k x k / any stride / any padding depthwise: one thread per output element.
__global__ void __launch_bounds__(256)
dw_generic_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, DwGeom g, int relu, long long total) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox = static_cast<int>(idx % g.OW);
    long long t = idx / g.OW;
    const int oy = static_cast<int>(t % g.OH);
    const long long plane = t / g.OH;
    const int c = static_cast<int>(plane % g.C);
    const float* ip = in + plane * g.H * g.W;
    const float* kp = w + static_cast<size_t>(c) * g.KH * g.KW;
    float v = 0.f;
    for (int u = 0; u < g.KH; ++u) {
        const int iy = oy * g.stride_h - g.pad_top + u;
        if (iy < 0 || iy >= g.H) continue;
        for (int x = 0; x < g.KW; ++x) {
            const int ix = ox * g.stride_w - g.pad_left + x;
            if (ix < 0 || ix >= g.W) continue;
            v = fmaf(__ldg(ip + static_cast<long long>(iy) * g.W + ix), __ldg(kp + u * g.KW + x), v);
        }
    }
    if (bias) v += __ldg(bias + c);
    if (relu) v = fmaxf(v, 0.f);
    out[idx] = v;
}

int depthwise_forward(const float* in, const float* w, const float* bias, float* out, const DwGeom& g, int relu,
                      int batch, cudaStream_t s) {
    const bool k3 = g.KH == 3 && g.KW == 3 && g.pad_top == 1 && g.pad_left == 1 && g.stride_h == g.stride_w &&
                    (g.stride_h == 1 || g.stride_h == 2) && g.OW >= 24;
    const double planes = static_cast<double>(batch) * g.C;
    const int prof = prof_begin(s, PROF_DEPTHWISE, 2.0 * planes * g.OH * g.OW * g.KH * g.KW, 0,
                                4.0 * planes * (static_cast<double>(g.H) * g.W + static_cast<double>(g.OH) * g.OW));
    if (k3) {
        const int xstrips = ceil_div(g.OW, 32), ystrips = ceil_div(g.OH, kDwRows);
        const long long items = static_cast<long long>(batch) * g.C * xstrips * ystrips;
        const unsigned blocks = static_cast<unsigned>((items + 3) / 4);
        if (g.stride_h == 1)
            dw3x3_shuffle_kernel<1><<<blocks, 128, 0, s>>>(in, w, bias, out, g.C, g.H, g.W, g.OH, g.OW, relu, items,
                                                            xstrips, ystrips);
        else
            dw3x3_shuffle_kernel<2><<<blocks, 128, 0, s>>>(in, w, bias, out, g.C, g.H, g.W, g.OH, g.OW, relu, items,
                                                            xstrips, ystrips);
    } else if (g.KH == 3 && g.KW == 3 && (g.pad_top == 0 || g.pad_top == 1) && (g.pad_left == 0 || g.pad_left == 1) &&
               g.stride_h == g.stride_w && (g.stride_h == 1 || g.stride_h == 2) && g.H <= 34 && g.W <= 34 &&
               (g.OH - 1) * g.stride_h - g.pad_top + 2 <= g.H && (g.OW - 1) * g.stride_w - g.pad_left + 2 <= g.W) {
        // (the last window may reach one pixel past the image: that is the zero halo)
        const long long planes_ll = static_cast<long long>(batch) * g.C;
        const unsigned blocks = static_cast<unsigned>((planes_ll + 3) / 4);
        const int smem = 4 * (g.H + 2) * (g.W + 2) * static_cast<int>(sizeof(float));
        if (g.stride_h == 1)
            dw3x3_plane_kernel<1><<<blocks, 128, smem, s>>>(in, w, bias, out, g.C, g.H, g.W, g.OH, g.OW, g.pad_top,
                                                             g.pad_left, relu, planes_ll);
        else
            dw3x3_plane_kernel<2><<<blocks, 128, smem, s>>>(in, w, bias, out, g.C, g.H, g.W, g.OH, g.OW, g.pad_top,
                                                             g.pad_left, relu, planes_ll);
    } else {
        const long long total = static_cast<long long>(batch) * g.C * g.OH * g.OW;
        dw_generic_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, w, bias, out, g, relu, total);
    }
    prof_end(prof, s);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

}  // namespace fcuda
