// Device scratch pool (see include/feather/mempool.h; mirrors /root/reference/src/mempool.cpp:32-109).
#include <feather/mempool.h>

#include <cuda_runtime.h>

template <typename PTR_TYPE>
bool CommonMemPool<PTR_TYPE>::Request(size_t size_byte) {
    if (size_byte > common_size) common_size = size_byte;  // keep the maximum, mempool.cpp:61-68
    return true;
}

template <typename PTR_TYPE>
bool CommonMemPool<PTR_TYPE>::Alloc() {
    if (allocated_size >= common_size && (common_memory || common_size == 0)) return true;
    if (common_memory && cudaFree(common_memory) != cudaSuccess) {
        // e.g. called inside a stream capture: keep the old pool instead of leaking it
        fprintf(stderr, "CommonMemPool: cudaFree failed (%s); pool not resized\n", cudaGetErrorString(cudaGetLastError()));
        return false;
    }
    common_memory = nullptr;
    allocated_size = 0;
    if (common_size == 0) return true;
    void* p = nullptr;
    if (cudaMalloc(&p, common_size) != cudaSuccess) {
        fprintf(stderr, "CommonMemPool: cudaMalloc of %zu bytes failed\n", common_size);
        return false;
    }
    common_memory = static_cast<PTR_TYPE*>(p);
    allocated_size = common_size;
    return true;
}

template <typename PTR_TYPE>
bool CommonMemPool<PTR_TYPE>::GetPtr(PTR_TYPE** ptr) {
    if (!Alloc()) return false;  // lazy allocation, mempool.cpp:95-109
    *ptr = common_memory;
    return true;
}

template <typename PTR_TYPE>
bool CommonMemPool<PTR_TYPE>::Reset() {
    common_size = 0;
    return true;
}

template <typename PTR_TYPE>
bool CommonMemPool<PTR_TYPE>::Free() {
    if (common_memory) cudaFree(common_memory);
    common_memory = nullptr;
    allocated_size = 0;
    return true;
}

template class CommonMemPool<float>;
