// TensorGEMM on tcgen05 — the B200 replacement for the reference's 64-way batched
// "TensorGEMM" (/root/reference/src/booster/avx/winograd_kernels_F63.cpp:518-757) and for its
// packed SGEMM (/root/reference/src/booster/avx/sgemm.cpp:377-433).
//
//   for g in [0,G):   D_g[M x N] = A_g[M x K] * B_g[N x K]^T        (fp32 in, fp32 out)
//
// Both operands are K-major ("TN").  The contraction runs on the 5th-gen tensor cores as kind::tf32 UMMA with
// fp32 accumulation in TMEM.  fp32 semantics are recovered with the 3xTF32 split (x = hi + lo exactly, hi
// TF32-representable): each k-step issues  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.
//   B (the small, reused operand: transformed filters / packed weights / FC activations) carries its hi and lo
//     planes in global memory (made once by the producer kernel);
//   A (the big streamed operand: Winograd V, im2col rows, FC weights) is stored ONCE as plain fp32: the kernel
//     lands the raw tile in shared memory by TMA, four splitter warps turn each row into hi/lo and park it in
//     TENSOR MEMORY (tcgen05.st), and the MMAs read A from TMEM ("TS" form).  That halves A's HBM traffic and takes
//     the A reads off the shared-memory port.
// planes == 1 runs plain TF32 (one MMA per k-step, operands straight from shared memory).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fcuda {

enum GemmEpilogue : int {
    // D[g][m][n], n contiguous, row stride = ldd floats, batch stride = M*ldd.
    // Used for the Winograd product buffer M_e[tile][oc].
    EPI_ROWMAJOR = 0,
    // m = img * P + pix ;  out[(img * N + n) * P + pix] = act(acc + bias[n])      (NCHW conv output)
    EPI_NCHW = 1,
    // part[ks][n * ldd + m] = acc  — m = output feature, n = batch row (InnerProduct), ks = k-split index; plane stride
    // = split_stride floats.  Plain stores: the partial planes are summed in a fixed order by fc_reduce (layers.cu), so
    // the result does not depend on scheduling (round 1 used fp32 atomics here and was not run-to-run deterministic).
    EPI_COLMAJOR_PARTIAL = 2,
};

struct GemmProblem {
    // operands (device pointers)
    const float* A;                         // [G][M][K] plain fp32
    const float* B_hi; const float* B_lo;   // [G][N][K]; B_lo may be null when planes == 1
    float* D;
    int M, N, K, G;
    int planes;        // 1 = TF32, 2 = 3xTF32 (A split in-kernel, B_hi/B_lo planes)
    int epilogue;      // GemmEpilogue
    int ldd;           // EPI_ROWMAJOR / EPI_COLMAJOR_PARTIAL leading dimension (floats)
    int P;             // EPI_NCHW: pixels per image
    long long m_offset;  // EPI_NCHW: global pixel index of row 0 (chunked im2col); image = (m_offset + m) / P
    const float* bias; // EPI_NCHW: per-n bias or null
    int relu;          // EPI_NCHW: fuse max(0, .)
    int split_k;       // >=1; only with EPI_COLMAJOR_PARTIAL
    long long split_stride;  // EPI_COLMAJOR_PARTIAL: floats between the partial planes of consecutive k-splits
    long long a_batch_stride;  // floats between consecutive g in A; 0 => dense (M*K)
    long long b_batch_stride;  // floats between consecutive g in B; 0 => dense (N*K)
    double algo_flops;         // algorithmic FLOPs this launch stands for (profiling only; 0 => 2*M*N*K*G)
};

// Launches the persistent tcgen05 kernel on `stream`.  Returns 0 or a negative fcuda error.
// Requirements: K % 4 == 0, 16-byte aligned operand pointers.  M/N/K tails are zero-filled by TMA.
int tensor_gemm(const GemmProblem& p, cudaStream_t stream);

// Per-launch profiling of the TensorGEMM (bench.py's roofline leg): when enabled every tensor_gemm launch is
// bracketed by CUDA events on its own stream.  collect() synchronises and returns totals since enable.
void gemm_profile_enable(bool on);
void gemm_profile_collect(double* total_ms, double* algo_flops, double* mma_flops, long long* launches);

// Number of k-splits tensor_gemm will really use for (K, requested split): every split owns at least one 32-wide k-block.
int tensor_gemm_effective_split(int K, int split_k);

// True when the problem satisfies the TMA alignment rules above.
bool tensor_gemm_supported(const GemmProblem& p);

// Reference CUDA-core fp32 GEMM with identical semantics (used by the self-test and as
// the fallback for shapes TMA cannot describe).
int simt_gemm(const GemmProblem& p, cudaStream_t stream);

}  // namespace fcuda
