// Timeline of the implicit-GEMM conv's hand-offs (needs a B200; diagnostic, not a test).
//   build: make -C feathercnn_b200/csrc igemm_trace      run: build/igemm_trace [IC OC HW batch [planes pool]]
//   planes: 2 = 3xTF32 (default), 3 = BF16x3; pool: 1 = fused 2x2 max pooling epilogue
// Compiles conv_igemm.cu with -DFCUDA_IGEMM_TRACE: CTA 0 records clock64() at every producer / MMA / epilogue hand-off
// of its first 64 k-blocks; printed relative to the first event, one line per k-block.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../feathercnn_b200/csrc/conv_igemm.cuh"

namespace fcuda { int igemm_trace_read(long long* host); }
using namespace fcuda;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

int main(int argc, char** argv) {
    const int IC = argc > 1 ? atoi(argv[1]) : 64, OC = argc > 2 ? atoi(argv[2]) : 64;
    const int HW = argc > 3 ? atoi(argv[3]) : 224, N = argc > 4 ? atoi(argv[4]) : 16;
    const int planes = argc > 5 ? atoi(argv[5]) : 2, pool = argc > 6 ? atoi(argv[6]) : 0;
    const size_t nin = (size_t)N * IC * HW * HW, nout = (size_t)N * OC * HW * HW, nw = (size_t)OC * IC * 9;
    float *in, *out, *w, *whi, *wlo, *bias;
    CK(cudaMalloc(&in, nin * 4)); CK(cudaMalloc(&out, nout * 4)); CK(cudaMalloc(&w, nw * 4));
    CK(cudaMalloc(&whi, nw * 4)); CK(cudaMalloc(&wlo, nw * 4)); CK(cudaMalloc(&bias, OC * 4));
    std::vector<float> h(nin);
    for (size_t i = 0; i < nin; ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f;
    CK(cudaMemcpy(in, h.data(), nin * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w, h.data(), nw * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(bias, 0, OC * 4));
    if (conv_igemm_pack_weights(w, whi, wlo, OC, IC, 9, 0, planes == 3)) return 1;
    IgemmProblem p{in, whi, wlo, bias, out, N, IC, HW, HW, OC, HW, HW, 3, 3, 1, 1, 1, 1, planes, 1};
    p.pool = pool;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) if (conv_igemm_forward(p, 0)) return 1;
    CK(cudaEventRecord(e0));
    if (conv_igemm_forward(p, 0)) return 1;
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const int kblocks = 9 * IC / 32;
    printf("IC=%d OC=%d %dx%d batch %d planes %d pool %d: %.3f ms (trace build), %d k-blocks per tile\n", IC, OC, HW, HW, N, planes,
           pool, ms, kblocks);
    std::vector<long long> t(16 * 64);
    if (igemm_trace_read(t.data())) return 1;
    long long t0 = t[0];
    for (int i = 0; i < 12 * 64; ++i) if (t[i] && t[i] < t0) t0 = t[i];
    printf("  g grp | P:wait_start  empty_ok  st_issued  gather_issued  st_done | M:full_ok  probe(next ready?)  issued | T:B issued | period\n");
    for (int g = 0; g < 64; ++g) {
        auto r = [&](int slot) { return t[slot * 64 + g] ? (long long)(t[slot * 64 + g] - t0) : -1LL; };
        printf("%3d  %d  | %10lld %9lld %10lld %14lld %8lld | %9lld %7lld(%lld) %7lld | %8lld | %lld\n", g, g % 3, r(0), r(1), r(2),
               r(3), r(4), r(5), r(6), 0LL, r(7), r(8), g ? r(5) - (t[5 * 64 + g - 1] - t0) : 0LL);
    }
    printf("  g | M: loop_top  full_ok  mma_issued  committed | P(q3) st_done\n");
    for (int g = 24; g < 44; ++g) {
        auto r = [&](int slot) { return t[slot * 64 + g] ? (long long)(t[slot * 64 + g] - t0) : -1LL; };
        printf("%3d | %10lld %10lld %10lld %10lld | %10lld\n", g, r(6), r(5), r(14), r(7), r(3));
    }
    printf("  g | T: loop_top  slot_free  copy_issued  after_syncwarp\n");
    for (int g = 24; g < 44; ++g) {
        auto r = [&](int slot) { return t[slot * 64 + g] ? (long long)(t[slot * 64 + g] - t0) : -1LL; };
        printf("%3d | %10lld %10lld %10lld %10lld\n", g, r(11), r(8), r(12), r(13));
    }
    printf("tile | E:tmem_full  done\n");
    for (int i = 0; i < 8; ++i) printf("%3d  | %10lld %8lld\n", i, t[9 * 64 + i] - t0, t[10 * 64 + i] - t0);
    return 0;
}
