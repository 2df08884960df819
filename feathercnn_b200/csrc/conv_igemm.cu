// SGECONV — implicit-GEMM convolution straight from the NCHW blob on tcgen05 (sm_100a).
//
// Replaces the idea of the reference's SGECONV algorithm (/root/reference/src/booster/arm/sgeconv.cpp:1311-1856,
// "packs B directly from the padded input, no materialised im2col"; a stub in the AVX dispatcher,
// avx/booster.cpp:105-118) and of im2col + packed SGEMM (avx/booster.cpp:83-102) with a Blackwell formulation that
// has NO intermediate in HBM and not even an A tile in shared memory:
//
//   D[pixel][oc] = sum_k  A[pixel][k] * Wp[oc][k],    k = (u*KW + v)*IC + ic,
//   A[pixel][k]  = X[n][ic][oy*s+u-pad][ox*s+v-pad]   (0 outside the image: the im2col rule, generic_kernels.cpp:66-67)
//
//   M tile     = 128 output pixels = four 32-pixel row segments ("boxes", consecutive in (n, oy, ox/32) order);
//                pixel <-> TMEM lane.
//   A operand  = lives in TENSOR MEMORY.  Twelve producer warps (3 groups x 4; group g serves every third k-block) each
//                own 32 pixels (= their TMEM lane quadrant): a thread gathers the 32 k-values of its pixel (loads
//                coalesced across lanes along x; offsets and tap coordinates come from a small shared-memory table,
//                validity from per-pixel row/column bit masks), splits them into TF32 hi + fp32 lo and writes them with
//                tcgen05.st.  (TMA cannot gather: its innermost box coordinate must be 16-byte aligned — a +-1 pixel
//                tap shift traps; an A tile in shared memory costs 3x the instructions and all of the smem bandwidth.)
//   B operand  = filters re-packed once at Init to Wp[oc][k] (K-major), TF32 hi / fp32 lo planes, loaded by TMA.
//   MMA        = one thread issues tcgen05.mma kind::tf32 with A from TMEM ("TS" form), 3 MMAs per k-step in 3xTF32
//                mode, fp32 accumulators in a 2-deep TMEM ring next to the 4-deep A ring.
//   epilogue   = tcgen05.ld -> +bias -> ReLU -> NCHW store; lanes hold consecutive pixels -> 128-byte coalesced rows.
//
// Used where non-fused Winograd is bandwidth-bound (large images, <= 128 channels) and for every layer the reference
// sends to im2col + SGEMM (1x1, strided, 7x7, IC = 3).
#include "conv_igemm.cuh"

#include "common.cuh"
#include "tcgen05.cuh"

namespace fcuda {

// Optional timeline of CTA 0 (tests/cuda/igemm_trace.cu builds this file with -DFCUDA_IGEMM_TRACE): clock64() of the
// hand-offs of the first 64 k-blocks, one row per event kind.
#ifdef FCUDA_IGEMM_TRACE
__device__ long long g_igemm_trace[16 * 64];
#define IG_TRACE(slot, idx)                                                                              \
    do {                                                                                                 \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (idx) < 64u) g_igemm_trace[(slot) * 64 + (idx)] = clock64(); \
    } while (0)
int igemm_trace_read(long long* host) {
    return cudaMemcpyFromSymbol(host, g_igemm_trace, sizeof(long long) * 16 * 64) == cudaSuccess ? 0 : -1;
}
#else
#define IG_TRACE(slot, idx) do {} while (0)
#endif

namespace {

constexpr int kGroups = 3;                                // producer groups, 4 warps (128 pixels) each
constexpr int kWarpTma = 4 + 4 * kGroups;                  // warps 0-3 epilogue, 4.. producers (quadrant = warp % 4),
constexpr int kWarpMma = kWarpTma + 1;                    // then one TMA(B) warp and one MMA warp: no idle warps, so
constexpr int kThreadsIg = (kWarpMma + 1) * 32;           // 576 threads leave 112 registers for the producers
constexpr int kStagesIg = 4;                              // ring depth shared by the smem B tiles and the TMEM A tiles
constexpr int kMaxBStages = 8;                            // filter-tile ring: deeper than the A ring, as smem allows
constexpr int kMaxTableK = 8192;                          // k-table entries that fit beside the B ring

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
    int K;                   // KH*KW*IC
    int kblocks;             // ceil(K / 32)
    int use_table;           // 0 => IC % 32 == 0: every k-block is 32 channels of ONE tap, offsets are arithmetic
    int bpr;                 // 32-pixel boxes per output row
    long long total_boxes;   // N * OH * bpr
    int num_n;               // ceil(OC / BN)
    long long pixel_tiles;   // ceil(total_boxes / 4)
    int relu;
    int bstages;             // depth of the filter-tile ring in shared memory (<= kMaxBStages)
};

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

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int BN, int PLANES>
__global__ void __launch_bounds__(kThreadsIg, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
                  const IgemmArgs args) {
    constexpr int STAGES = kStagesIg;
    constexpr int kBTile = BN * 32 * 4;
    constexpr int kStage = PLANES * kBTile;                 // smem per stage: [B_hi][B_lo]
    constexpr uint32_t kAccCols = 2 * BN;                   // accumulator ring
    constexpr uint32_t kAStageCols = 32 * PLANES;           // [A_hi (32 cols)][A_lo (32 cols)]
    constexpr uint32_t kNeedCols = kAccCols + STAGES * kAStageCols;
    constexpr uint32_t kTmemCols = kNeedCols <= 32 ? 32 : kNeedCols <= 64 ? 64 : kNeedCols <= 128 ? 128 : kNeedCols <= 256 ? 256 : 512;
    static_assert(kNeedCols <= 512, "TMEM budget");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    int2* ktab = reinterpret_cast<int2*>(smem + args.bstages * kStage);  // {element offset, tap index} per k

    // A ring (tensor memory, STAGES slots): full = 4 arrivals of the owning producer group, empty = MMAs retired.
    // B ring (shared memory, args.bstages slots, deeper: a filter tile is ~1.5k cycles away — relaxed wake-up of the
    // TMA warp + L2 latency — while an A slot turns around in 3 k-blocks of MMA time).
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint64_t b_full_bar[kMaxBStages];
    __shared__ uint64_t b_empty_bar[kMaxBStages];
    __shared__ uint64_t empty_bar[STAGES];    // MMAs that read the stage have retired
    __shared__ uint64_t tmem_full_bar[2];
    __shared__ uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long total_tiles = args.pixel_tiles * args.num_n;
    const int kblocks = args.kblocks;
    const int plane = args.H * args.W;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 4);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < kMaxBStages; ++s) {
            ptx::mbar_init(&b_full_bar[s], 1);
            ptx::mbar_init(&b_empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full_bar[s], 1);
            ptx::mbar_init(&tmem_empty_bar[s], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kWarpTma && lane == 0) {
        ptx::prefetch_tensormap(&tmW);
        if (PLANES == 2) ptx::prefetch_tensormap(&tmWlo);
    }
    if (args.use_table) {
        for (int k = threadIdx.x; k < kblocks * 32; k += kThreadsIg) {
            int2 e = make_int2(0, 63);  // padding rows: tap 63 is never valid
            if (k < args.K) {
                const int tap = k / args.IC, ic = k - tap * args.IC;
                const int u = tap / args.KW, v = tap - u * args.KW;
                e = make_int2(ic * plane + u * args.W + v, tap);
            }
            ktab[k] = e;
        }
    }
    if (warp == kWarpMma) {
        ptx::tmem_alloc(&tmem_base_smem, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t tmem_a0 = tmem_base + kAccCols;

    if (warp == kWarpTma) {
        // ===================== TMA producer for the filter tiles =====================
        const bool leader = ptx::elect_one();
        int bs = 0;
        uint32_t bphase = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n_blk = static_cast<int>(tile % args.num_n);
            for (int kb = 0; kb < kblocks; ++kb) {
                ptx::mbar_wait(&b_empty_bar[bs], bphase ^ 1);
                if (leader) {
                    uint8_t* st = smem + bs * kStage;
                    ptx::mbar_arrive_expect_tx(&b_full_bar[bs], PLANES * kBTile);
                    ptx::tma_load_3d(st, &tmW, &b_full_bar[bs], kb * 32, n_blk * BN, 0);
                    if (PLANES == 2) ptx::tma_load_3d(st + kBTile, &tmWlo, &b_full_bar[bs], kb * 32, n_blk * BN, 0);
                }
                __syncwarp();
                if (++bs == args.bstages) { bs = 0; bphase ^= 1; }
            }
        }
    } else if (warp == kWarpMma) {
        // ===================== MMA issuer =====================
        // The whole warp walks the loop (warp-uniform control flow); one elected lane issues.  Descriptors are
        // formed once: per k-block only the stage offset (in 16-byte units / TMEM columns) is added.
        // The tensor pipe's queue is shallow (measured: issuing 12 MMAs takes as long as executing them, and anything
        // the warp does after the last issue runs with ~1.5 MMAs of work left), so nothing slow may sit between two
        // k-blocks: the barrier of the NEXT slot is probed half-way through the current slot's MMAs and the blocking
        // wait is only the fallback.
        constexpr uint32_t idesc = make_idesc_tf32(BN);
        const bool leader = ptx::elect_one();
        const uint64_t dB0 = make_smem_desc_sw128(ptx::smem_u32(smem));
        int stage = 0, bs = 0;
        uint32_t phase = 0, bphase = 0;
        long long it = 0;
        bool ready = false;  // full_bar[stage] and b_full_bar[bs] already observed complete
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = static_cast<int>(it & 1);
            const uint32_t aphase = static_cast<uint32_t>((it >> 1) & 1);
            ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int kb = 0; kb < kblocks; ++kb) {
                if (!ready) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::mbar_wait(&b_full_bar[bs], bphase);
                }
                IG_TRACE(5, static_cast<uint32_t>(it * kblocks + kb));
                ptx::tc_fence_after();
                const int nstage = (stage + 1) & (STAGES - 1);
                const uint32_t nphase = nstage == 0 ? phase ^ 1u : phase;
                const int nbs = bs + 1 == args.bstages ? 0 : bs + 1;
                const uint32_t nbphase = nbs == 0 ? bphase ^ 1u : bphase;
                const uint64_t dB = dB0 + static_cast<uint64_t>(bs * (kStage >> 4));
                const uint64_t dBlo = dB + static_cast<uint64_t>(kBTile >> 4);
                const uint32_t ta = tmem_a0 + stage * kAStageCols;
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // 8 k-values (TMEM columns / 32 smem bytes) per MMA
                    const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
                    if (leader) {
                        if (PLANES == 2) {
                            umma_tf32_ts(tmem_d, ta + 32 + k * 8, dB + 2 * k, idesc, first);   // A_lo * B_hi
                            umma_tf32_ts(tmem_d, ta + k * 8, dBlo + 2 * k, idesc, 1u);         // A_hi * B_lo
                            umma_tf32_ts(tmem_d, ta + k * 8, dB + 2 * k, idesc, 1u);           // A_hi * B_hi
                        } else {
                            umma_tf32_ts(tmem_d, ta + k * 8, dB + 2 * k, idesc, first);
                        }
                    }
                    if (k == 1) {  // results needed after the last MMA
                        const bool a_ok = ptx::mbar_test(&full_bar[nstage], nphase);
                        const bool b_ok = ptx::mbar_test(&b_full_bar[nbs], nbphase);
                        ready = a_ok && b_ok;
                    }
                }
                if (leader) {
                    ptx::umma_commit(&empty_bar[stage]);
                    ptx::umma_commit(&b_empty_bar[bs]);
                }
                __syncwarp();
                IG_TRACE(7, static_cast<uint32_t>(it * kblocks + kb));
                stage = nstage; phase = nphase;
                bs = nbs; bphase = nbphase;
            }
            if (leader) ptx::umma_commit(&tmem_full_bar[as]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== A producers: gather + 3xTF32 split -> tensor memory =====================
        // k-block number g (running over all tiles of this CTA) belongs to group g % kGroups and ring slot
        // g % STAGES; a group visits only its own k-blocks.  The group is latency-bound, not issue-bound (a gather
        // is ~800 cycles from L2), so the loop is software-pipelined: the 32 loads of the group's NEXT k-block — which
        // may belong to the next tile — are issued before it waits for the tensor-memory stores of the current one.
        static_assert((STAGES & (STAGES - 1)) == 0, "ring slot = g & (STAGES-1)");
        const int group = (warp - 4) >> 2;
        const int q = warp & 3;  // TMEM lane quadrant this warp may write == box index inside the tile
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const int cblocks = args.IC >> 5;                              // fast path only
        const uint32_t plane_bytes = static_cast<uint32_t>(plane) * 4u;  // host guarantees H*W < 2^30
        const int kb_mod = kblocks % kGroups;

        // cursor over this group's k-blocks
        long long tile = blockIdx.x;
        uint32_t g0 = 0;   // running index of the current tile's k-block 0
        int g0_mod = 0;    // g0 % kGroups
        int kb = 0;
        int tap = 0, tu = 0, tv = 0, cb = 0;  // fast path: k-block kb = tap * cblocks + cb, tap = tu*KW + tv
        const float* base = args.in;
        unsigned long long tapmask = 0;  // bit (u*KW+v) set <=> that tap of this lane's pixel lies inside the image
        bool have = false;

        // position the cursor on the first owned k-block of tile `tile` or a later tile
        auto enter_tile = [&]() {
            have = false;
            while (tile < total_tiles) {
                kb = group - g0_mod;
                if (kb < 0) kb += kGroups;
                if (kb < kblocks) { have = true; break; }
                tile += gridDim.x;
                g0 += static_cast<uint32_t>(kblocks);
                g0_mod += kb_mod;
                if (g0_mod >= kGroups) g0_mod -= kGroups;
            }
            if (!have) return;
            const long long ptile = tile / args.num_n;
            const BoxCoord bx = decode_box(ptile * 4 + q, args);
            const int ox = bx.ox0 + lane;
            const bool pix_ok = bx.valid && ox < args.OW;
            const int iy0 = bx.oy * args.stride_h - args.pad_top;
            const int ix0 = ox * args.stride_w - args.pad_left;
            tapmask = 0;
            if (pix_ok) {
                for (int u = 0; u < args.KH; ++u) {
                    if (iy0 + u < 0 || iy0 + u >= args.H) continue;
                    for (int v = 0; v < args.KW; ++v)
                        if (ix0 + v >= 0 && ix0 + v < args.W) tapmask |= 1ull << (u * args.KW + v);
                }
            }
            base = args.in + (static_cast<long long>(bx.valid ? bx.n : 0) * args.IC) * plane +
                   static_cast<long long>(iy0) * args.W + ix0;
            asm volatile("" : "+l"(base));  // opaque: see gather()
            tap = 0; tu = 0; tv = 0; cb = kb;
        };
        auto advance = [&]() {
            kb += kGroups;
            cb += kGroups;
            if (kb >= kblocks) {
                tile += gridDim.x;
                g0 += static_cast<uint32_t>(kblocks);
                g0_mod += kb_mod;
                if (g0_mod >= kGroups) g0_mod -= kGroups;
                enter_tile();
            }
        };
        // issue the 32 loads of the cursor's k-block
        auto gather = [&](float (&x)[32]) {
            if (!args.use_table) {
                // IC % 32 == 0: the whole k-block is one tap -> one predicate; the 32 channel addresses are
                // kp + r*plane_bytes, one IMAD.WIDE each off an opaque pointer (otherwise nvcc re-derives every
                // address from args.in with ~8 integer instructions per load)
                while (cb >= cblocks) {
                    cb -= cblocks;
                    ++tap;
                    if (++tv == args.KW) { tv = 0; ++tu; }
                }
                const char* kp = reinterpret_cast<const char*>(base) +
                                 (static_cast<long long>(cb * 32) * plane + tu * args.W + tv) * 4;
                asm volatile("" : "+l"(kp));
                const bool kb_ok = ((tapmask >> tap) & 1ull) != 0;
                if (__all_sync(0xffffffffu, kb_ok)) {  // interior: no predication, no zero fill
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        x[r] = __ldg(reinterpret_cast<const float*>(
                            kp + static_cast<unsigned long long>(plane_bytes) * static_cast<uint32_t>(r)));
                } else {
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        x[r] = kb_ok ? __ldg(reinterpret_cast<const float*>(
                                           kp + static_cast<unsigned long long>(plane_bytes) * static_cast<uint32_t>(r)))
                                     : 0.f;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const int2 e = ktab[kb * 32 + r];
                    x[r] = ((tapmask >> e.y) & 1ull) ? __ldg(base + e.x) : 0.f;
                }
            }
        };

        float x[32];
        enter_tile();
        if (have) gather(x);
        while (have) {
            const uint32_t g = g0 + static_cast<uint32_t>(kb);
            const int my_stage = static_cast<int>(g & (STAGES - 1));
            const uint32_t my_phase = (g / STAGES) & 1u;
            const uint32_t ta = tmem_a0 + lane_base + my_stage * kAStageCols;
            if (q == 0) IG_TRACE(0, g);
            ptx::mbar_wait(&empty_bar[my_stage], my_phase ^ 1);
            if (q == 0) IG_TRACE(1, g);
            ptx::tc_fence_after();
#pragma unroll
            for (int part = 0; part < 4; ++part) {  // 8 k-values at a time keeps the live set inside 96 registers
                uint32_t hi[8];
                if (PLANES == 2) {
                    uint32_t lo[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        // hi = x with the 13 sub-TF32 mantissa bits cleared (what the tensor core would read
                        // anyway), lo = x - hi exactly: 2 instructions per element instead of 5 for round-to-nearest;
                        // the residual after the hardware truncates lo is <= 2^-21 |x| either way
                        hi[r] = __float_as_uint(x[part * 8 + r]) & 0xFFFFE000u;
                        lo[r] = __float_as_uint(x[part * 8 + r] - __uint_as_float(hi[r]));
                    }
                    tmem_st_32x8(ta + part * 8, hi);
                    tmem_st_32x8(ta + 32 + part * 8, lo);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) hi[r] = __float_as_uint(x[part * 8 + r]);
                    tmem_st_32x8(ta + part * 8, hi);
                }
            }
            if (q == 0) IG_TRACE(2, g);
            advance();
            if (have) gather(x);  // in flight across the store drain, the arrive and the next slot wait
            if (q == 0) IG_TRACE(3, g);
            tmem_st_wait();
            if (q == 0) IG_TRACE(4, g);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&full_bar[my_stage]);
        }
    } else {
        // ===================== epilogue (warps 0..3) =====================
        const int q = warp & 3;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const long long ptile = tile / args.num_n;
            const int n_blk = static_cast<int>(tile - ptile * args.num_n);
            const int as = static_cast<int>(it & 1);
            const uint32_t aphase = static_cast<uint32_t>((it >> 1) & 1);
            ptx::mbar_wait_relaxed(&tmem_full_bar[as], aphase);
            if (q == 0) IG_TRACE(9, static_cast<uint32_t>(it));
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
            if (q == 0) IG_TRACE(10, static_cast<uint32_t>(it));
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// W[oc][ic][tap] -> Wp[oc][Kf] hi/lo planes, k = tap*IC + ic, Kf = K rounded up to 4 (TMA's 16-byte row rule)
__global__ void __launch_bounds__(256)
igemm_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int OC, int IC,
                          int taps, int Kf) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t total = static_cast<size_t>(OC) * Kf;
    if (idx >= total) return;
    const int k = static_cast<int>(idx % Kf);
    const int oc = static_cast<int>(idx / Kf);
    float v = 0.f;
    if (k < IC * taps) {
        const int tap = k / IC, ic = k - tap * IC;
        v = w[(static_cast<size_t>(oc) * IC + ic) * taps + tap];
    }
    if (lo) {
        float h, l;
        split_tf32(v, h, l);
        hi[idx] = h;
        lo[idx] = l;
    } else {
        hi[idx] = v;
    }
}

template <int BN, int PLANES>
int launch_igemm(const IgemmProblem& p, cudaStream_t stream) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return FCUDA_ERR_CUDA;
    CUtensorMap tmW, tmWlo;
    const int taps = p.KH * p.KW;
    const int K = taps * p.IC;
    const int Kf = (K + 3) & ~3;
    for (int pl = 0; pl < PLANES; ++pl) {
        cuuint64_t dims[3] = {(cuuint64_t)Kf, (cuuint64_t)p.OC, 1};
        cuuint64_t strides[2] = {(cuuint64_t)Kf * 4, (cuuint64_t)Kf * p.OC * 4};
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
    a.K = K;
    a.kblocks = ceil_div(K, 32);
    a.use_table = (p.IC % 32 == 0) ? 0 : 1;
    a.bpr = ceil_div(p.OW, 32);
    a.total_boxes = static_cast<long long>(p.N) * p.OH * a.bpr;
    a.num_n = ceil_div(p.OC, BN);
    a.pixel_tiles = (a.total_boxes + 3) / 4;
    a.relu = p.relu;
    const long long total = a.pixel_tiles * a.num_n;
    const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
    constexpr int kStage = PLANES * BN * 32 * 4;
    const int table_bytes = a.use_table ? a.kblocks * 32 * 8 : 0;
    int bstages = (227 * 1024 - 2048 - table_bytes) / kStage;
    if (bstages > kMaxBStages) bstages = kMaxBStages;
    if (bstages < 2) return -1;
    a.bstages = bstages;
    const int smem = bstages * kStage + table_bytes + 1024;
    auto kern = conv_igemm_kernel<BN, PLANES>;
    static int attr_smem = 0;
    if (smem > attr_smem) {
        FCUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    kern<<<grid, kThreadsIg, smem, stream>>>(tmW, tmWlo, a);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

}  // namespace

bool conv_igemm_supported(int IC, int KH, int KW) {
    if (KH * KW > 63) return false;                               // per-pixel tap validity mask is 64-bit
    if (IC % 32 != 0 && KH * KW * IC > kMaxTableK) return false;  // k-table must fit in shared memory
    return true;
}

size_t conv_igemm_packed_floats(int OC, int IC, int taps, int planes) {
    return static_cast<size_t>(planes) * OC * ((taps * IC + 3) & ~3);
}

int conv_igemm_pack_weights(const float* w, float* w_hi, float* w_lo, int OC, int IC, int taps, cudaStream_t s) {
    const int Kf = (taps * IC + 3) & ~3;
    const size_t total = static_cast<size_t>(OC) * Kf;
    igemm_pack_weights_kernel<<<static_cast<unsigned>(ceil_div_sz(total, 256)), 256, 0, s>>>(w, w_hi, w_lo, OC, IC, taps, Kf);
    FCUDA_CHECK_LAUNCH();
    count_launch();
    return 0;
}

int conv_igemm_forward(const IgemmProblem& p, cudaStream_t stream) {
    if (!conv_igemm_supported(p.IC, p.KH, p.KW)) return -1;
    const bool x3 = p.planes == 2;
    if (p.OC <= 32) return x3 ? launch_igemm<32, 2>(p, stream) : launch_igemm<32, 1>(p, stream);
    if (p.OC <= 64) return x3 ? launch_igemm<64, 2>(p, stream) : launch_igemm<64, 1>(p, stream);
    return x3 ? launch_igemm<128, 2>(p, stream) : launch_igemm<128, 1>(p, stream);
}

}  // namespace fcuda
