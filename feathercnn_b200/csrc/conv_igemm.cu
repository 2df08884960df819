// SGECONV — implicit-GEMM convolution straight from the NCHW blob on tcgen05 (sm_100a).
//
// Replaces the idea of the reference's SGECONV algorithm (/root/reference/src/booster/arm/sgeconv.cpp:1311-1856,
// "packs B directly from the padded input, no materialised im2col"; a stub in the AVX dispatcher,
// avx/booster.cpp:105-118) with a Blackwell formulation that has NO intermediate in HBM at all:
//
//   D[pixel][oc] = sum_{tap=(u,v)} sum_{ic}  X[n][ic][oy+u-pad][ox+v-pad] * W[oc][ic][u][v]
//
//   A operand  = the input itself.  A 4-D TMA box (32 pixels along W) x (32 channels) lands in shared memory as
//                [channel][pixel] rows of 128 bytes with the 128B swizzle, which is exactly the canonical MN-major
//                UMMA layout (M = pixels contiguous, K = channels).  Zero padding is TMA's out-of-bounds fill:
//                the box is simply addressed at (ox+v-pad, oy+u-pad).  Four boxes (32 pixels each, consecutive
//                row segments, possibly from different rows/images) form the 128-row M tile.
//   B operand  = filters re-packed once at Init to Wp[tap][oc][ic] (K-major rows), TF32 hi / fp32 lo planes.
//   3xTF32     = four "splitter" warps turn each landed raw A tile into hi (in place) + lo (second buffer) in shared
//                memory (the transform is element-wise, hence indifferent to the swizzle), then the MMA thread issues
//                A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.
//   epilogue   = tcgen05.ld -> +bias -> ReLU -> NCHW store; lanes hold consecutive pixels -> 128-byte coalesced rows.
//
// Used where non-fused Winograd is bandwidth-bound (large images, <= 128 channels) and for stride-1 1x1 convolutions.
#include "conv_igemm.cuh"

#include "common.cuh"
#include "tcgen05.cuh"

namespace fcuda {

namespace {

constexpr int kThreadsIg = 384;  // warp 0 TMA, warp 1 MMA, warps 4-7 epilogue, warps 8-11 splitters
constexpr int kABox = 32 * 32 * 4;   // one TMA box: 32 channels x 32 pixels fp32 = 4 KB
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
    float* out;
    const float* bias;
    int N, OC, OH, OW;
    int KH, KW, pad_top, pad_left;
    int bpr;                 // 32-pixel boxes per output row
    long long total_boxes;   // N * OH * bpr
    int cblocks;             // ceil(IC / 32)
    int num_n;               // ceil(OC / BN)
    long long pixel_tiles;   // ceil(total_boxes / 4)
    int relu;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
        : "memory");
}

// MN-major operand, 128B swizzle: atoms of (32 MN elements = 128 B) x (8 K rows); LBO = byte stride between atoms
// along MN (here: one 4 KB TMA box), SBO = byte stride between 8-row K groups (1 KB).
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
    return d;
}

struct BoxCoord { int n, oy, ox0; bool valid; };

__device__ __forceinline__ BoxCoord decode_box(long long b, const IgemmArgs& a) {
    BoxCoord c;
    c.valid = b < a.total_boxes;
    const long long per_img = static_cast<long long>(a.OH) * a.bpr;
    const long long n = b / per_img;
    const int rem = static_cast<int>(b - n * per_img);
    c.n = c.valid ? static_cast<int>(n) : a.N;  // n == N is out of bounds for TMA -> zero fill
    c.oy = rem / a.bpr;
    c.ox0 = (rem - c.oy * a.bpr) * 32;
    return c;
}

template <int BN, int PLANES, int STAGES>
__global__ void __launch_bounds__(kThreadsIg, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmWlo, const IgemmArgs args) {
    constexpr int kBTile = BN * 32 * 4;
    constexpr int kStage = PLANES * (kATileBytes + kBTile);
    // stage layout: [A (raw -> hi)][B_hi][A_lo][B_lo]
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint64_t split_bar[STAGES];
    __shared__ uint64_t empty_bar[STAGES];
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
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&split_bar[s], 4);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full_bar[s], 1);
            ptx::mbar_init(&tmem_empty_bar[s], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmX);
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const long long ptile = tile / args.num_n;
                const int n_blk = static_cast<int>(tile - ptile * args.num_n);
                BoxCoord box[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) box[j] = decode_box(ptile * 4 + j, args);
                for (int cb = 0; cb < args.cblocks; ++cb) {
                    for (int tap = 0; tap < taps; ++tap) {
                        const int u = tap / args.KW, v = tap - u * args.KW;
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* st = smem + stage * kStage;
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], (PLANES == 2 ? 2 : 1) * kBTile + kATileBytes);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            tma_load_4d(st + j * kABox, &tmX, &full_bar[stage], box[j].ox0 + v - args.pad_left,
                                        box[j].oy + u - args.pad_top, cb * 32, box[j].n);
                        ptx::tma_load_3d(st + kATileBytes, &tmW, &full_bar[stage], cb * 32, n_blk * BN, tap);
                        if (PLANES == 2)
                            ptx::tma_load_3d(st + 2 * kATileBytes + kBTile, &tmWlo, &full_bar[stage], cb * 32, n_blk * BN, tap);
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
                    ptx::mbar_wait(PLANES == 2 ? &split_bar[stage] : &full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + stage * kStage);
                    const uint64_t dB = make_smem_desc_sw128(st + kATileBytes);
                    const uint64_t dBlo = make_smem_desc_sw128(st + 2 * kATileBytes + kBTile);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 K-atoms of 8 channels
                        const uint64_t dA = make_smem_desc_mn_sw128(st + k * 1024, kABox, 1024);
                        const uint64_t dAlo = make_smem_desc_mn_sw128(st + kATileBytes + kBTile + k * 1024, kABox, 1024);
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
        // ===================== hi/lo splitters (3xTF32 only) =====================
        if (PLANES == 2) {
            const int t = threadIdx.x - 8 * 32;  // 0..127
            int stage = 0;
            uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    float4* a = reinterpret_cast<float4*>(smem + stage * kStage);
                    float4* alo = reinterpret_cast<float4*>(smem + stage * kStage + kATileBytes + kBTile);
#pragma unroll
                    for (int i = 0; i < kATileBytes / 16 / 128; ++i) {
                        const int idx = i * 128 + t;
                        const float4 x = a[idx];
                        float4 h, l;
                        split_tf32(x.x, h.x, l.x);
                        split_tf32(x.y, h.y, l.y);
                        split_tf32(x.z, h.z, l.z);
                        split_tf32(x.w, h.w, l.w);
                        a[idx] = h;
                        alo[idx] = l;
                    }
                    // make the generic-proxy writes visible to the tensor core's async-proxy reads
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&split_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
            const BoxCoord bx = decode_box(ptile * 4 + q, args);  // this warp's 32 rows are box q
            const int ox = bx.ox0 + lane;
            const bool ok = bx.valid && ox < args.OW;
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            const size_t plane = static_cast<size_t>(args.OH) * args.OW;
            float* dst0 = args.out + (static_cast<size_t>(ok ? bx.n : 0) * args.OC) * plane +
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
                            dst0[static_cast<size_t>(oc0 + j) * plane] = v;
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

// W[oc][ic][tap] -> Wp[tap][oc][ic] hi/lo planes
__global__ void __launch_bounds__(256)
igemm_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int OC, int IC,
                          int taps) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t total = static_cast<size_t>(OC) * IC * taps;
    if (idx >= total) return;
    const int ic = static_cast<int>(idx % IC);
    const size_t t = idx / IC;
    const int oc = static_cast<int>(t % OC);
    const int tap = static_cast<int>(t / OC);
    const float v = w[(static_cast<size_t>(oc) * IC + ic) * taps + tap];
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
    CUtensorMap tmX, tmW, tmWlo;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.IC, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.W * p.H * 4, (cuuint64_t)p.W * p.H * p.IC * 4};
        cuuint32_t box[4] = {32, 1, 32, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.input), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            fprintf(stderr, "fcuda: igemm input tensor map failed (%d)\n", (int)r);
            return FCUDA_ERR_CUDA;
        }
    }
    const int taps = p.KH * p.KW;
    for (int pl = 0; pl < PLANES; ++pl) {
        cuuint64_t dims[3] = {(cuuint64_t)p.IC, (cuuint64_t)p.OC, (cuuint64_t)taps};
        cuuint64_t strides[2] = {(cuuint64_t)p.IC * 4, (cuuint64_t)p.IC * p.OC * 4};
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
    a.out = p.output; a.bias = p.bias;
    a.N = p.N; a.OC = p.OC; a.OH = p.OH; a.OW = p.OW;
    a.KH = p.KH; a.KW = p.KW; a.pad_top = p.pad_top; a.pad_left = p.pad_left;
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
    kern<<<grid, kThreadsIg, smem, stream>>>(tmX, tmW, tmWlo, a);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

}  // namespace

bool conv_igemm_supported(int IC, int W, int stride_h, int stride_w, const void* input) {
    return stride_h == 1 && stride_w == 1 && W % 4 == 0 && IC % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(input) & 15) == 0;
}

int conv_igemm_pack_weights(const float* w, float* w_hi, float* w_lo, int OC, int IC, int taps, cudaStream_t s) {
    const size_t total = static_cast<size_t>(OC) * IC * taps;
    igemm_pack_weights_kernel<<<static_cast<unsigned>(ceil_div_sz(total, 256)), 256, 0, s>>>(w, w_hi, w_lo, OC, IC, taps);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

int conv_igemm_forward(const IgemmProblem& p, cudaStream_t stream) {
    if (!conv_igemm_supported(p.IC, p.W, 1, 1, p.input)) return -1;
    const bool x3 = p.planes == 2;
    if (p.OC <= 32) return x3 ? launch_igemm<32, 2, 4>(p, stream) : launch_igemm<32, 1, 8>(p, stream);
    if (p.OC <= 64) return x3 ? launch_igemm<64, 2, 4>(p, stream) : launch_igemm<64, 1, 8>(p, stream);
    return x3 ? launch_igemm<128, 2, 3>(p, stream) : launch_igemm<128, 1, 6>(p, stream);
}

}  // namespace fcuda
