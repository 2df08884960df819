// SGECONV — implicit-GEMM convolution straight from the NCHW blob on tcgen05 (sm_100a).
//
// Replaces the idea of the reference's SGECONV algorithm (/root/reference/src/booster/arm/sgeconv.cpp:1311-1856,
// "packs B directly from the padded input, no materialised im2col"; a stub in the AVX dispatcher,
// avx/booster.cpp:105-118) with a Blackwell formulation that has NO intermediate in HBM at all:
//
//   D[pixel][oc] = sum_{tap=(u,v)} sum_{ic}  X[n][ic][oy*s+u-pad][ox*s+v-pad] * W[oc][ic][u][v]
//
//   M tile     = 128 output pixels = four 32-pixel row segments ("boxes", consecutive in (n, oy, ox/32) order).
//   A operand  = gathered by 16 producer warps directly from the input (coalesced along x, zero padding and stride by
//                index arithmetic — the im2col rule of generic_kernels.cpp:66-67), split on the fly into TF32 hi and
//                fp32 lo and stored as 16-byte vectors into shared memory in the canonical MN-major UMMA layout for
//                32-bit operands (128-byte rows of 32 pixels per channel, 32-byte-chunk swizzle, layout type
//                SWIZZLE_128B_BASE32B).  TMA cannot do this gather: its innermost box coordinate must be 16-byte
//                aligned, and a 3x3 tap shifts x by one float (measured: illegal-instruction trap).
//   B operand  = filters re-packed once at Init to Wp[tap][oc][ic] (K-major rows), TF32 hi / fp32 lo planes, by TMA.
//   MMA        = one thread issues tcgen05.mma kind::tf32 (A MN-major, B K-major), 3 MMAs per k-step in 3xTF32 mode,
//                fp32 accumulators in a 2-deep TMEM ring.
//   epilogue   = tcgen05.ld -> +bias -> ReLU -> NCHW store; lanes hold consecutive pixels -> 128-byte coalesced rows.
//
// Used where non-fused Winograd is bandwidth-bound (large images, <= 128 channels) and for the layers the reference
// sends to im2col + SGEMM (1x1, strided, 5x5/7x7).
#include "conv_igemm.cuh"

#include "common.cuh"
#include "tcgen05.cuh"

namespace fcuda {

namespace {

constexpr int kProducerWarps = 16;  // two groups of 8: group g gathers the k-blocks whose running index has parity g
constexpr int kThreadsIg = (8 + kProducerWarps) * 32;  // warp 0 TMA(B), warp 1 MMA, warps 4-7 epilogue, 8.. producers
constexpr int kABox = 32 * 32 * 4;      // one 32-pixel box: 32 channel rows of 128 bytes = 4 KB
constexpr int kATileBytes = 4 * kABox;  // 128 pixels x 32 channels = 16 KB

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct IgemmArgs {
    const float* in;
    float* out;
    const float* bias;
    int N, IC, H, W, OC, OH, OW;
    int KH, KW, pad_top, pad_left, stride_h, stride_w;
    int bpr;                 // 32-pixel boxes per output row
    long long total_boxes;   // N * OH * bpr
    int cblocks;             // ceil(IC / 32)
    int num_n;               // ceil(OC / BN)
    long long pixel_tiles;   // ceil(total_boxes / 4)
    int relu;
};

// MN-major 32-bit operand.  tcgen05 accepts only one shared-memory layout for MN-major tf32: 128-byte rows of 32 MN
// elements swizzled in 32-byte chunks ("SWIZZLE_128B_BASE32B", layout type 1, Swizzle<2,5,2>: byte-offset bits [5,7)
// ^= bits [7,9)), atoms of 32 MN x 4 K rows (512 B).  LBO = byte stride between atoms along MN (one 4 KB box),
// SBO = byte stride between 4-row K groups (512 B).  Verified on B200 against TMA's SWIZZLE_128B_ATOM_32B tiles.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(1) << 61;  // SWIZZLE_128B_BASE32B
    return d;
}

struct BoxCoord { int n, oy, ox0; bool valid; };

__device__ __forceinline__ BoxCoord decode_box(long long b, const IgemmArgs& a) {
    BoxCoord c;
    c.valid = b < a.total_boxes;
    const long long per_img = static_cast<long long>(a.OH) * a.bpr;
    const long long n = b / per_img;
    const int rem = static_cast<int>(b - n * per_img);
    c.n = static_cast<int>(n);
    c.oy = rem / a.bpr;
    c.ox0 = (rem - c.oy * a.bpr) * 32;
    return c;
}

template <int BN, int PLANES, int STAGES>
__global__ void __launch_bounds__(kThreadsIg, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
                  const IgemmArgs args) {
    constexpr int kBTile = BN * 32 * 4;
    constexpr int kStage = PLANES * (kATileBytes + kBTile);
    // stage layout: [A_hi][B_hi][A_lo][B_lo]
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ uint64_t b_full_bar[STAGES];   // filters landed (TMA transaction bytes)
    __shared__ uint64_t a_ready_bar[STAGES];  // 8 producer warps finished the gathered A tile
    __shared__ uint64_t empty_bar[STAGES];    // MMAs that read the stage have retired
    __shared__ uint64_t tmem_full_bar[2];
    __shared__ uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long total_tiles = args.pixel_tiles * args.num_n;
    const int taps = args.KH * args.KW;
    const int kblocks = args.cblocks * taps;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&b_full_bar[s], 1);
            ptx::mbar_init(&a_ready_bar[s], 8);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full_bar[s], 1);
            ptx::mbar_init(&tmem_empty_bar[s], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmW);
        if (PLANES == 2) ptx::prefetch_tensormap(&tmWlo);
    }
    constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    if (warp == 1) {
        ptx::tmem_alloc(&tmem_base_smem, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer for the filter tiles =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_blk = static_cast<int>(tile % args.num_n);
                for (int cb = 0; cb < args.cblocks; ++cb) {
                    for (int tap = 0; tap < taps; ++tap) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* st = smem + stage * kStage;
                        ptx::mbar_arrive_expect_tx(&b_full_bar[stage], PLANES * kBTile);
                        ptx::tma_load_3d(st + kATileBytes, &tmW, &b_full_bar[stage], cb * 32, n_blk * BN, tap);
                        if (PLANES == 2)
                            ptx::tma_load_3d(st + 2 * kATileBytes + kBTile, &tmWlo, &b_full_bar[stage], cb * 32, n_blk * BN, tap);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(BN) | (1u << 15);  // A is MN-major
            int stage = 0;
            uint32_t phase = 0;
            long long it = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int as = static_cast<int>(it & 1);
                const uint32_t aphase = static_cast<uint32_t>((it >> 1) & 1);
                ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    ptx::mbar_wait(&a_ready_bar[stage], phase);
                    ptx::mbar_wait(&b_full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + stage * kStage);
                    const uint64_t dB = make_smem_desc_sw128(st + kATileBytes);
                    const uint64_t dBlo = make_smem_desc_sw128(st + 2 * kATileBytes + kBTile);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 8 channels per MMA = two 4-row K groups = 1 KB
                        const uint64_t dA = make_smem_desc_mn_sw128(st + k * 1024, kABox, 512);
                        const uint64_t dAlo = make_smem_desc_mn_sw128(st + kATileBytes + kBTile + k * 1024, kABox, 512);
                        const uint64_t koff = static_cast<uint64_t>(k * 2);  // B: +32 bytes inside the swizzle atom
                        const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
                        if (PLANES == 2) {
                            ptx::umma_tf32(tmem_d, dAlo, dB + koff, idesc, first);
                            ptx::umma_tf32(tmem_d, dA, dBlo + koff, idesc, 1u);
                            ptx::umma_tf32(tmem_d, dA, dB + koff, idesc, 1u);
                        } else {
                            ptx::umma_tf32(tmem_d, dA, dB + koff, idesc, first);
                        }
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full_bar[as]);
            }
        }
    } else if (warp >= 8) {
        // ===================== A producers: im2col gather + 3xTF32 split into UMMA MN-major tiles =====================
        const int pw = warp - 8;
        const int group = pw >> 3;  // 0 / 1: which k-block parity this warp serves
        const int wsub = pw & 7;    // channels wsub*4 .. wsub*4+3 of the 32-channel block
        const int q = lane >> 3;    // box (32-pixel segment) of this lane's 4 pixels
        const int xb = (lane & 7) * 4;  // first of the lane's 4 consecutive pixels inside the box
        const size_t plane = static_cast<size_t>(args.H) * args.W;
        long long g = 0;  // running k-block index over all tiles of this CTA
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const long long ptile = tile / args.num_n;
            const BoxCoord bx = decode_box(ptile * 4 + q, args);
            const int iy0 = bx.oy * args.stride_h - args.pad_top;
            const int ix0 = (bx.ox0 + xb) * args.stride_w - args.pad_left;
            const float* img = args.in + static_cast<size_t>(bx.valid ? bx.n : 0) * args.IC * plane;
            for (int cb = 0; cb < args.cblocks; ++cb) {
                for (int tap = 0; tap < taps; ++tap, ++g) {
                    if ((g & 1) != group) continue;
                    const int stage = static_cast<int>(g % STAGES);
                    const uint32_t phase = static_cast<uint32_t>((g / STAGES) & 1);
                    const int u = tap / args.KW, v = tap - u * args.KW;
                    const int iy = iy0 + u;
                    const bool row_ok = bx.valid && iy >= 0 && iy < args.H;
                    float x[4][4];
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int c = cb * 32 + wsub * 4 + ci;
                        const float* rp = img + static_cast<size_t>(c) * plane + static_cast<size_t>(iy) * args.W;
                        const bool c_ok = row_ok && c < args.IC;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ix = ix0 + j * args.stride_w + v;
                            x[ci][j] = (c_ok && ix >= 0 && ix < args.W) ? __ldg(rp + ix) : 0.f;
                        }
                    }
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);  // loads are in flight while the slot drains
                    uint8_t* st = smem + stage * kStage;
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int cl = wsub * 4 + ci;  // channel row inside the 32-channel block
                        // row cl of box q: 128 bytes; 32-byte chunk index swizzled with (row & 3)
                        const uint32_t off = static_cast<uint32_t>(q) * kABox + static_cast<uint32_t>(cl) * 128 +
                                             ((static_cast<uint32_t>(xb >> 3) ^ (cl & 3)) << 5) + ((xb & 7) << 2);
                        if (PLANES == 2) {
                            float4 h, l;
                            split_tf32(x[ci][0], h.x, l.x);
                            split_tf32(x[ci][1], h.y, l.y);
                            split_tf32(x[ci][2], h.z, l.z);
                            split_tf32(x[ci][3], h.w, l.w);
                            *reinterpret_cast<float4*>(st + off) = h;
                            *reinterpret_cast<float4*>(st + kATileBytes + kBTile + off) = l;
                        } else {
                            *reinterpret_cast<float4*>(st + off) = make_float4(x[ci][0], x[ci][1], x[ci][2], x[ci][3]);
                        }
                    }
                    // make the generic-proxy writes visible to the tensor core's async-proxy reads
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&a_ready_bar[stage]);
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (warps 4..7) =====================
        const int q = warp & 3;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const long long ptile = tile / args.num_n;
            const int n_blk = static_cast<int>(tile - ptile * args.num_n);
            const int as = static_cast<int>(it & 1);
            const uint32_t aphase = static_cast<uint32_t>((it >> 1) & 1);
            ptx::mbar_wait(&tmem_full_bar[as], aphase);
            ptx::tc_fence_after();
            const BoxCoord bx = decode_box(ptile * 4 + q, args);  // this warp's 32 TMEM lanes are box q
            const int ox = bx.ox0 + lane;
            const bool ok = bx.valid && ox < args.OW;
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            const size_t oplane = static_cast<size_t>(args.OH) * args.OW;
            float* dst0 = args.out + (static_cast<size_t>(ok ? bx.n : 0) * args.OC) * oplane +
                          static_cast<size_t>(bx.oy) * args.OW + ox;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int oc0 = n_blk * BN + c0;
                if (oc0 >= args.OC) break;
                uint32_t r[32];
                ptx::tmem_ld_32x32(taddr0 + c0, r);
                ptx::tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (oc0 + j < args.OC) {
                            float v = __uint_as_float(r[j]);
                            if (args.bias) v += __ldg(args.bias + oc0 + j);
                            if (args.relu) v = fmaxf(v, 0.f);
                            dst0[static_cast<size_t>(oc0 + j) * oplane] = v;
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// W[oc][ic][tap] -> Wp[tap][oc][icp] hi/lo planes (icp = IC rounded up to 4 for TMA's 16-byte row rule)
__global__ void __launch_bounds__(256)
igemm_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int OC, int IC,
                          int ICp, int taps) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t total = static_cast<size_t>(OC) * ICp * taps;
    if (idx >= total) return;
    const int ic = static_cast<int>(idx % ICp);
    const size_t t = idx / ICp;
    const int oc = static_cast<int>(t % OC);
    const int tap = static_cast<int>(t / OC);
    const float v = ic < IC ? w[(static_cast<size_t>(oc) * IC + ic) * taps + tap] : 0.f;
    if (lo) {
        float h, l;
        split_tf32(v, h, l);
        hi[idx] = h;
        lo[idx] = l;
    } else {
        hi[idx] = v;
    }
}

template <int BN, int PLANES, int STAGES>
int launch_igemm(const IgemmProblem& p, cudaStream_t stream) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return FCUDA_ERR_CUDA;
    CUtensorMap tmW, tmWlo;
    const int taps = p.KH * p.KW;
    const int ICp = (p.IC + 3) & ~3;
    for (int pl = 0; pl < PLANES; ++pl) {
        cuuint64_t dims[3] = {(cuuint64_t)ICp, (cuuint64_t)p.OC, (cuuint64_t)taps};
        cuuint64_t strides[2] = {(cuuint64_t)ICp * 4, (cuuint64_t)ICp * p.OC * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        const float* base = pl == 0 ? p.w_hi : p.w_lo;
        CUresult r = enc(pl == 0 ? &tmW : &tmWlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            fprintf(stderr, "fcuda: igemm weight tensor map failed (%d)\n", (int)r);
            return FCUDA_ERR_CUDA;
        }
    }
    if (PLANES == 1) tmWlo = tmW;
    IgemmArgs a;
    a.in = p.input; a.out = p.output; a.bias = p.bias;
    a.N = p.N; a.IC = p.IC; a.H = p.H; a.W = p.W; a.OC = p.OC; a.OH = p.OH; a.OW = p.OW;
    a.KH = p.KH; a.KW = p.KW; a.pad_top = p.pad_top; a.pad_left = p.pad_left;
    a.stride_h = p.stride_h; a.stride_w = p.stride_w;
    a.bpr = ceil_div(p.OW, 32);
    a.total_boxes = static_cast<long long>(p.N) * p.OH * a.bpr;
    a.cblocks = ceil_div(p.IC, 32);
    a.num_n = ceil_div(p.OC, BN);
    a.pixel_tiles = (a.total_boxes + 3) / 4;
    a.relu = p.relu;
    const long long total = a.pixel_tiles * a.num_n;
    const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
    constexpr int kStage = PLANES * (kATileBytes + BN * 32 * 4);
    static_assert(STAGES * kStage + 1024 <= 227 * 1024, "smem budget");
    const int smem = STAGES * kStage + 1024;
    auto kern = conv_igemm_kernel<BN, PLANES, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        FCUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    kern<<<grid, kThreadsIg, smem, stream>>>(tmW, tmWlo, a);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

}  // namespace

size_t conv_igemm_packed_floats(int OC, int IC, int taps, int planes) {
    return static_cast<size_t>(planes) * taps * OC * ((IC + 3) & ~3);
}

int conv_igemm_pack_weights(const float* w, float* w_hi, float* w_lo, int OC, int IC, int taps, cudaStream_t s) {
    const int ICp = (IC + 3) & ~3;
    const size_t total = static_cast<size_t>(OC) * ICp * taps;
    igemm_pack_weights_kernel<<<static_cast<unsigned>(ceil_div_sz(total, 256)), 256, 0, s>>>(w, w_hi, w_lo, OC, IC, ICp, taps);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

int conv_igemm_forward(const IgemmProblem& p, cudaStream_t stream) {
    const bool x3 = p.planes == 2;
    if (p.OC <= 32) return x3 ? launch_igemm<32, 2, 4>(p, stream) : launch_igemm<32, 1, 8>(p, stream);
    if (p.OC <= 64) return x3 ? launch_igemm<64, 2, 4>(p, stream) : launch_igemm<64, 1, 8>(p, stream);
    return x3 ? launch_igemm<128, 2, 3>(p, stream) : launch_igemm<128, 1, 6>(p, stream);
}

}  // namespace fcuda
