// Micro-benchmark (needs a B200): cycles per tcgen05.mma for the operand forms and shapes the conv kernels use.
//   build: make -C feathercnn_b200/csrc mma_rate      run: build/mma_rate
// One CTA per SM, one elected thread issues a long stream of MMAs on resident (garbage) operands and commits once;
// the clock is read around the stream.  Answers: is the 3xTF32 pipe bound by math (cycles ~ N) or by a per-instruction
// cost (operand fetch), and does it depend on A coming from shared (SS) or tensor memory (TS)?
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../feathercnn_b200/csrc/tcgen05.cuh"

using namespace fcuda;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, int f16) {
    if (f16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "r"(tmem_a), "l"(desc_b), "r"(idesc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "r"(tmem_a), "l"(desc_b), "r"(idesc)
                     : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, int f16) {
    if (f16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "l"(desc_a), "l"(desc_b), "r"(idesc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "l"(desc_a), "l"(desc_b), "r"(idesc)
                     : "memory");
}

struct Variant {
    int ts;       // A from tensor memory
    int n;        // MMA N
    int f16;      // kind::f16 (bf16 operands) instead of kind::tf32
    int d_ring;   // number of distinct accumulators the stream rotates over
    int k_steps;  // distinct k-slices walked inside the operand tiles (4 = one 128-byte swizzle row)
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Variant v, int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        ptx::tmem_alloc(&tmem_base_smem, 512);
        ptx::tmem_relinquish();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    if (threadIdx.x < 32) {
        const bool leader = ptx::elect_one();
        // idesc: D=f32; A/B format tf32 (2) or bf16 (1); K-major; N, M=128
        uint32_t idesc = (1u << 4) | ((v.f16 ? 1u : 2u) << 7) | ((v.f16 ? 1u : 2u) << 10) |
                         (static_cast<uint32_t>(v.n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        const uint64_t dA = make_smem_desc_sw128(ptx::smem_u32(smem));
        const uint64_t dB = make_smem_desc_sw128(ptx::smem_u32(smem + 16384));
        const uint32_t a_cols = tmem + 256;  // accumulators live in columns [0, 256)
        long long t0 = 0, t1 = 0;
        __syncwarp();
        if (leader) {
            t0 = clock64();
            // straight-line groups of 8 MMAs with compile-time operand offsets, like the conv kernels' issue loops
            const uint32_t dstep = v.d_ring > 1 ? static_cast<uint32_t>(v.n) : 0u;
            const uint32_t kmask = static_cast<uint32_t>(v.k_steps - 1);
            if (v.ts) {
                for (int i = 0; i < iters; i += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        umma_ts(tmem + (j & (v.d_ring - 1)) * dstep, a_cols + (j & kmask) * 8, dB + 2 * (j & kmask), idesc,
                                v.f16);
                }
            } else {
                for (int i = 0; i < iters; i += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        umma_ss(tmem + (j & (v.d_ring - 1)) * dstep, dA + 2 * (j & kmask), dB + 2 * (j & kmask), idesc,
                                v.f16);
                }
            }
            ptx::umma_commit(&bar);
        }
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        t1 = clock64();
        if (leader && blockIdx.x == 0) *cycles = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// Same MMA stream with 12 other warps storing into tensor memory at the same time (what the implicit-GEMM conv's
// producers do: 32 KB of A_hi/A_lo per k-block).  mode 0: stores only (fixed count), 1: MMAs + stores until the MMAs
// are done.  Reports cycles/MMA and bytes/cycle of tcgen05.st.
template <int X>
__device__ __forceinline__ void tmem_st(uint32_t taddr, uint32_t v) {
    if (X == 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
    else if (X == 16)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
    else
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
}

template <int X>
__global__ void __launch_bounds__(512, 1)
mma_st_kernel(int n, int with_mma, int store_warps, int iters, long long* out /* [0]=mma cycles [1]=st cycles [2]=stores */) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ volatile int done;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        done = 0;
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        ptx::tmem_alloc(&tmem_base_smem, 512);
        ptx::tmem_relinquish();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (with_mma) {
            const bool leader = ptx::elect_one();
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
                                   (static_cast<uint32_t>(128 >> 4) << 24);
            const uint64_t dB = make_smem_desc_sw128(ptx::smem_u32(smem + 16384));
            const uint32_t a_cols = tmem + 256;
            long long t0 = clock64();
            if (leader) {
                for (int i = 0; i < iters; i += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) umma_ts(tmem, a_cols + (j & 3) * 8, dB + 2 * (j & 3), idesc, 0);
                }
                ptx::umma_commit(&bar);
            }
            __syncwarp();
            ptx::mbar_wait(&bar, 0);
            long long t1 = clock64();
            if (leader && blockIdx.x == 0) out[0] = t1 - t0;
            done = 1;
        }
    } else if (warp >= 4 && warp < 4 + store_warps) {
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t col0 = tmem + lane_base + 384 + ((warp - 4) >> 2) * 32;  // away from the MMA's A columns
        long long t0 = clock64();
        long long n_st = 0;
        for (int i = 0; with_mma ? !done : i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 32 / X; ++j) tmem_st<X>(col0 + j * X, i);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            n_st += 32 / X;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 128) { out[1] = t1 - t0; out[2] = n_st; }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

template <int X>
static void run_st(int n, int with_mma, int store_warps, long long* d_out, int smem) {
    CK(cudaFuncSetAttribute(mma_st_kernel<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(d_out, 0, 24));
    const int iters = 4096;
    mma_st_kernel<X><<<148, 512, smem>>>(n, with_mma, store_warps, iters, d_out);
    CK(cudaDeviceSynchronize());
    long long o[3];
    CK(cudaMemcpy(o, d_out, 24, cudaMemcpyDeviceToHost));
    // every store instruction of one warp writes 32 lanes x X columns x 4 bytes
    const double bytes = static_cast<double>(o[2]) * 32 * X * 4 * store_warps;
    printf("N=%3d mma=%d store_warps=%2d x%-2d : cycles/MMA %7.1f   tcgen05.st %8.1f B/clk/SM (%lld stores/warp in %lld clk)\n", n,
           with_mma, store_warps, X, with_mma ? static_cast<double>(o[0]) / iters : 0.0, o[1] ? bytes / o[1] : 0.0, o[2], o[1]);
}

// ------------------------------------------------------------------------------------------------
// How far can the issuing thread run ahead of the tensor pipe?  Issues `iters` MMAs (TS form, N columns), reads the
// clock when the last one has been ISSUED and again when the commit barrier reports them COMPLETE.  With `vary` the
// operand descriptors are recomputed from a running stage index for every group of 12 (as a real k-loop does).
__global__ void __launch_bounds__(128, 1) mma_queue_kernel(int n, int iters, int vary, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        ptx::tmem_alloc(&tmem_base_smem, 512);
        ptx::tmem_relinquish();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    if (threadIdx.x < 32) {
        const bool leader = ptx::elect_one();
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
                               (static_cast<uint32_t>(128 >> 4) << 24);
        const uint64_t dB0 = make_smem_desc_sw128(ptx::smem_u32(smem + 16384));
        long long t0 = clock64(), t1 = 0, t2 = 0;
        int stage = 0;
        for (int i = 0; i < iters; i += 12) {
            const uint64_t dB = dB0 + static_cast<uint64_t>(vary ? stage * 64 : 0);
            const uint32_t ta = tmem + 256 + (vary ? stage * 64 : 0);
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_ts(tmem, ta + 32 + k * 8, dB + 2 * k, idesc, 0);
                    umma_ts(tmem, ta + k * 8, dB + 32 + 2 * k, idesc, 0);
                    umma_ts(tmem, ta + k * 8, dB + 2 * k, idesc, 0);
                }
            }
            __syncwarp();
            if (++stage == 4) stage = 0;
        }
        t1 = clock64();
        if (leader) ptx::umma_commit(&bar);
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        t2 = clock64();
        if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

int main() {
    long long* d_cycles;
    CK(cudaMalloc(&d_cycles, 8));
    const int smem = 16384 + 32768 + 1024;
    CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<Variant> vs;
    for (int f16 = 0; f16 <= 1; ++f16)
        for (int ts = 0; ts <= 1; ++ts)
            for (int n : {32, 64, 128, 256}) vs.push_back({ts, n, f16, 1, 4});
    vs.push_back({1, 64, 0, 2, 4});
    vs.push_back({1, 64, 0, 4, 4});
    vs.push_back({1, 128, 0, 2, 4});
    vs.push_back({0, 64, 0, 2, 4});
    vs.push_back({1, 64, 0, 1, 1});
    vs.push_back({1, 128, 0, 1, 1});
    const int iters = 4096;
    printf("kind  form   N  d_ring k_steps   cycles/MMA   ideal(tf32 4096, bf16 8192 FLOP/clk)  eff\n");
    for (const Variant& v : vs) {
        mma_rate_kernel<<<148, 128, smem>>>(v, iters, d_cycles);
        CK(cudaDeviceSynchronize());
        mma_rate_kernel<<<148, 128, smem>>>(v, iters, d_cycles);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost));
        const double per = static_cast<double>(c) / iters;
        const double flop = 2.0 * 128 * v.n * (v.f16 ? 16 : 8);
        const double ideal = flop / (v.f16 ? 8192.0 : 4096.0);
        printf("%s  %s  %4d  %5d  %5d   %10.1f   %8.1f   %5.1f%%\n", v.f16 ? "bf16" : "tf32", v.ts ? "TS" : "SS", v.n,
               v.d_ring, v.k_steps, per, ideal, 100.0 * ideal / per);
    }
    long long* d_out;
    CK(cudaMalloc(&d_out, 24));
    printf("---- tensor-memory stores alone\n");
    for (int w : {4, 8, 12}) { run_st<8>(64, 0, w, d_out, smem); run_st<16>(64, 0, w, d_out, smem); run_st<32>(64, 0, w, d_out, smem); }
    printf("---- TS MMA stream with concurrent stores\n");
    for (int n : {64, 128, 256})
        for (int w : {4, 12}) { run_st<8>(n, 1, w, d_out, smem); run_st<32>(n, 1, w, d_out, smem); }
    printf("---- run-ahead of the issuing thread (TS, 3 MMAs per k-step): cycles until last ISSUE vs until COMPLETE\n");
    CK(cudaFuncSetAttribute(mma_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int n : {64, 128})
        for (int vary = 0; vary <= 1; ++vary)
            for (int iters : {12, 24, 48, 96, 192, 1536}) {
                mma_queue_kernel<<<148, 128, smem>>>(n, iters, vary, d_out);
                CK(cudaDeviceSynchronize());
                mma_queue_kernel<<<148, 128, smem>>>(n, iters, vary, d_out);
                CK(cudaDeviceSynchronize());
                long long o[2];
                CK(cudaMemcpy(o, d_out, 16, cudaMemcpyDeviceToHost));
                printf("N=%3d vary=%d MMAs=%5d : issued after %7lld clk (%.1f/MMA), complete after %7lld clk (%.1f/MMA)\n", n, vary,
                       iters, o[0], (double)o[0] / iters, o[1], (double)o[1] / iters);
            }
    return 0;
}
