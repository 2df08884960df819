// Micro-benchmark (needs a B200): how fast can ONE CTA per SM write 128-pixel x 64-channel fp32 tiles into an NCHW tensor?
//   build: make -C feathercnn_b200/csrc store_rate      run: build/store_rate
// The implicit-GEMM conv's epilogue is the limiter of every short-K layer (igemm trace of VGG conv1_1: ~3,000 cycles per
// 128 x 64 tile = 11 B/clk per SM, HBM needs ~23).  Variants, all writing the same bytes in the same tile order:
//   stg   W warps, lane = pixel, one 4-byte store per (pixel, channel): a warp instruction writes one 128-byte line
//   stg4  W warps, lane = 4 consecutive pixels (st.global.v4): a warp instruction writes four channels x 128 bytes
//   tma32 4 warps, [32 channels][32 pixels] staged in shared memory, one 4-D TMA store per 4 KB chunk (the current epilogue)
//   tma64 4 warps, [64 channels][32 pixels] per store (8 KB)
//   tma128 one elected thread, [64 channels][128 pixels] per store (32 KB, rows of 512 bytes; needs OW % 128 == 0)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int OC = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// tile t = 128 consecutive pixels of the (n, y, x) order (OW % 128 == 0 here, so a tile is one row segment)
__global__ void __launch_bounds__(512, 1) stg_kernel(float* out, int N, int OH, int OW, int warps, int vec, long long tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= warps) return;
    const size_t plane = static_cast<size_t>(OH) * OW;
    const int tiles_per_row = OW / 128;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long row = t / tiles_per_row;
        const int x0 = static_cast<int>(t % tiles_per_row) * 128;
        const int n = static_cast<int>(row / OH), y = static_cast<int>(row % OH);
        float* base = out + (static_cast<size_t>(n) * OC) * plane + static_cast<size_t>(y) * OW + x0;
        if (!vec) {
            // warp w handles box (w & 3) = 32 pixels and the channels c = (w >> 2), (w >> 2) + warps/4, ...
            const int q = warp & 3, cstep = warps / 4;
            float* p = base + q * 32 + lane;
#pragma unroll 8
            for (int c = warp >> 2; c < OC; c += cstep) p[static_cast<size_t>(c) * plane] = static_cast<float>(c + lane);
        } else {
            // lane = 4 pixels of the 128; lanes 0-7 one line... a warp covers 128 pixels of ONE channel per instruction
            float4 v = make_float4(lane, 1.f, 2.f, 3.f);
            float* p = base + lane * 4;
#pragma unroll 8
            for (int c = warp; c < OC; c += warps) *reinterpret_cast<float4*>(p + static_cast<size_t>(c) * plane) = v;
        }
    }
}

__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int OH, int OW, int mode, long long tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_row = OW / 128;
    for (int i = threadIdx.x; i < 2 * 32768 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = static_cast<float>(i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const long long row = t / tiles_per_row;
        const int x0 = static_cast<int>(t % tiles_per_row) * 128;
        const int n = static_cast<int>(row / OH), y = static_cast<int>(row % OH);
        uint8_t* buf = smem + (it & 1) * 32768;
        if (mode == 0 || mode == 1) {
            const int chunk = mode == 0 ? 32 : 64;  // channels per store
            if (lane == 0) {
                for (int c0 = 0; c0 < OC; c0 += chunk) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(buf + warp * 8192 + (c0 ? 4096 : 0))), "r"(x0 + warp * 32), "r"(y), "r"(c0), "r"(n)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else {
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(buf)), "r"(x0), "r"(y), "r"(0), "r"(n)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// TMA LOAD rate: one thread streams boxes {bx, by, 32 channels} of the NCHW tensor into a 3-deep shared-memory ring (the
// implicit GEMM's 3x3 slab: {48, 6, 32} = 36 KB per box; nobody reads the data here).
__global__ void __launch_bounds__(128, 1) tma_load_kernel(const __grid_constant__ CUtensorMap tm, int OH, int OW, int bx, int by,
                                                          int box_bytes, long long boxes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[3];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int per_row = OW / 32, rows = OH / 4;
    uint32_t it = 0;
    for (long long b = blockIdx.x; b < boxes; b += gridDim.x, ++it) {
        const int st = it % 3;
        if (it >= 3) {  // the load that used this stage three boxes ago has landed
            const uint32_t parity = ((it / 3) - 1) & 1u;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[st])), "r"(parity) : "memory");
        }
        const long long t = b / 2;          // two channel blocks per tile
        const int cb = static_cast<int>(b & 1);
        const int n = static_cast<int>(t / (static_cast<long long>(rows) * per_row));
        const int rem = static_cast<int>(t % (static_cast<long long>(rows) * per_row));
        const int y0 = (rem / per_row) * 4 - 1, x0 = (rem % per_row) * 32 - 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[st])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(smem + st * 49152)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar[st])), "r"(x0), "r"(y0),
                     "r"(cb * 32), "r"(n)
                     : "memory");
    }
    // drain
    for (int k = 0; k < 3 && it > 0; ++k) {
        const uint32_t j = it - 1 - k;
        if (static_cast<int>(j) < 0) break;
        const int st = j % 3;
        const uint32_t parity = (j / 3) & 1u;
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[st])), "r"(parity) : "memory");
    }
}

// Small 1-D bulk copies from an L2-resident buffer (the implicit GEMM's filter tiles): `issuers` threads (one per warp) each
// stream `copies` copies of `bytes` into their own ring of `ring` stages.  Cycles per copy -> per-instruction cost of the TMA unit.
__global__ void __launch_bounds__(128, 1) bulk_copy_kernel(const unsigned char* src, int src_bytes, int bytes, int ring, int issuers,
                                                           int copies, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[4][16];
    if (threadIdx.x == 0) {
        for (int w = 0; w < 4; ++w)
            for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[w][i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) != 0 || warp >= issuers) return;
    uint8_t* my = smem + warp * ring * bytes;
    const long long t0 = clock64();
    for (int it = 0; it < copies; ++it) {
        const int st = it % ring;
        if (it >= ring) {
            const uint32_t parity = ((it / ring) - 1) & 1u;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[warp][st])), "r"(parity) : "memory");
        }
        const unsigned char* g = src + ((static_cast<unsigned>(it * 4 + warp + blockIdx.x * 7) * 4096u) & static_cast<unsigned>(src_bytes / 2 - 1));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp][st])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(my + st * bytes)), "l"(g), "r"(bytes), "r"(smem_u32(&bar[warp][st]))
                     : "memory");
    }
    for (int k = 0; k < ring && k < copies; ++k) {
        const int j = copies - 1 - k;
        const int st = j % ring;
        const uint32_t parity = (j / ring) & 1u;
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[warp][st])), "r"(parity) : "memory");
    }
    if (blockIdx.x == 0 && warp == 0) *out = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int N = 16, OH = 256, OW = 256;
    const size_t floats = static_cast<size_t>(N) * OC * OH * OW;
    float* out;
    CK(cudaMalloc(&out, floats * 4));
    const long long tiles = static_cast<long long>(N) * OH * (OW / 128);
    const double bytes = static_cast<double>(floats) * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    auto report = [&](const char* name, float ms) {
        printf("%-22s %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM at %.2f GHz (nominal)  %7.0f cycles per 32 KB tile\n", name, ms, bytes / ms / 1e6,
               bytes / (ms * 1e-3) / 148 / (clock_khz * 1e3), clock_khz / 1e6, ms * 1e-3 * clock_khz * 1e3 / (tiles / 148.0));
    };
    for (int vec = 0; vec <= 1; ++vec)
        for (int warps : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                stg_kernel<<<148, 512>>>(out, N, OH, OW, warps, vec, tiles);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            char name[64]; snprintf(name, sizeof name, "%s warps=%d", vec ? "stg4" : "stg", warps);
            report(name, ms);
        }
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768 + 1024));
    for (int mode = 0; mode < 3; ++mode) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)OC, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)OW * 4, (cuuint64_t)OW * OH * 4, (cuuint64_t)OC * OW * OH * 4};
        cuuint32_t box[4] = {mode == 2 ? 128u : 32u, 1, mode == 0 ? 32u : 64u, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("tensor map failed\n");
            return 1;
        }
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            tma_kernel<<<148, 128, 2 * 32768 + 1024>>>(tm, OH, OW, mode, tiles);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        report(mode == 0 ? "tma 32ch x 32px (4 KB)" : mode == 1 ? "tma 64ch x 32px (8 KB)" : "tma 64ch x 128px (32KB)", ms);
    }
    // ---- small bulk copies from L2
    {
        unsigned char* src;
        const int src_bytes = 1 << 20;
        CK(cudaMalloc(&src, src_bytes));
        CK(cudaMemset(src, 1, src_bytes));
        long long* d_out;
        CK(cudaMalloc(&d_out, 8));
        CK(cudaFuncSetAttribute(bulk_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int bytes : {2048, 4096, 8192, 16384})
            for (int issuers : {1, 2, 4})
                for (int ring : {2, 8}) {
                    if (issuers * ring * bytes > 196 * 1024 || ring > 16) continue;
                    const int copies = 2000;
                    for (int rep = 0; rep < 2; ++rep) {
                        bulk_copy_kernel<<<148, 128, 200 * 1024>>>(src, src_bytes, bytes, ring, issuers, copies, d_out);
                        CK(cudaDeviceSynchronize());
                    }
                    long long c;
                    CK(cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost));
                    printf("bulk copy %5d B x %d issuers, ring %d: %7.1f cycles per copy per issuer, %6.1f B/clk/SM\n", bytes, issuers, ring,
                           static_cast<double>(c) / copies, static_cast<double>(bytes) * issuers * copies / c);
                }
    }
    // ---- TMA loads of slab-shaped boxes
    CK(cudaFuncSetAttribute(tma_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 49152 + 1024));
    const int shapes[4][2] = {{48, 6}, {64, 6}, {32, 8}, {128, 3}};
    for (int sidx = 0; sidx < 4; ++sidx) {
        const int bx = shapes[sidx][0], by = shapes[sidx][1];
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)OC, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)OW * 4, (cuuint64_t)OW * OH * 4, (cuuint64_t)OC * OW * OH * 4};
        cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 32, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("tensor map failed\n");
            return 1;
        }
        const int box_bytes = bx * by * 32 * 4;
        const long long boxes = 2ll * N * (OH / 4) * (OW / 32);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            tma_load_kernel<<<148, 128, 3 * 49152 + 1024>>>(tm, OH, OW, bx, by, box_bytes, boxes);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        const double lb = static_cast<double>(boxes) * box_bytes;
        printf("tma load box {%3d,%d,32} %6d B: %8.3f ms  %7.1f GB/s  %6.1f B/clk/SM  %7.0f cycles per box\n", bx, by, box_bytes, ms,
               lb / ms / 1e6, lb / (ms * 1e-3) / 148 / (clock_khz * 1e3), ms * 1e-3 * clock_khz * 1e3 / (boxes / 148.0));
    }
    return 0;
}
