// RuntimeParameter — per-Net runtime context handed to every layer
// (mirrors /root/reference/src/rt_param.h:28-55: {common mempool, num_threads}) plus the CUDA stream
// and device the Net runs on.
#pragma once

#include "blob.h"
#include "mempool.h"
#include "utils.h"

inline namespace feather_b200 {  // ABI tag (the reference declares the same global template)

template <typename Dtype>
class RuntimeParameter {
public:
    RuntimeParameter() : _common_mempool(nullptr), _num_threads(1), _stream(nullptr), _device(0) {}
    RuntimeParameter(CommonMemPool<Dtype>* common_mempool, size_t num_threads)
        : _common_mempool(common_mempool), _num_threads(num_threads), _stream(nullptr), _device(0) {}
    CommonMemPool<Dtype>* common_mempool() const { return _common_mempool; }
    size_t num_threads() const { return _num_threads; }  // kept for API compatibility; unused on the GPU
    void* stream() const { return _stream; }             // cudaStream_t
    void set_stream(void* s) { _stream = s; }
    int device() const { return _device; }
    void set_device(int d) { _device = d; }

private:
    CommonMemPool<Dtype>* _common_mempool;
    size_t _num_threads;
    void* _stream;
    int _device;
};

}  // inline namespace feather_b200
