// feather::NetGroup — one model replicated on the GPUs of one box from ONE host process, the C++ side of the north
// star's multi-GPU plan: "images of one batch shard across the 8 GPUs of one box with a single NCCL broadcast of weights at
// InitFromPath and no collective in Forward".
//
// The reference has no counterpart (it is a single-device CPU library: /root/reference/src/net.h:30-70); the member calls
// keep feather::Net's names and meaning.  InitFromPath loads the model ONCE (device 0 of the group: file I/O, weight
// decoding, one host->device upload of the weight arena), lays the same arena out on every other device without touching
// the file (Net::PrepareWeightArena) and fills it with one ncclBroadcast over NVLink (NCCL is dlopen'ed: libnccl.so.2; when
// it is absent the arena travels by cudaMemcpyPeerAsync); every device then runs its own Init (filter transforms) lazily.
// ForwardBatch shards the images of one host batch contiguously over the devices (sizes differ by at most one image),
// and drives every device's pipelined Net::SubmitBatch from the calling thread — the copies and kernels of all devices
// overlap — then waits for all of them.  No collective, results bit-identical to a single device (same kernels per image).
//
// The one-process-per-GPU launch (torchrun + feathercnn_b200/dist.py) remains the way bench.py scales; this class is the
// drop-in for a C++ application that owns the whole box.
#pragma once

#include <string>
#include <vector>

#include "net.h"

namespace feather {
inline namespace b200 {

class NetGroup {
public:
    NetGroup();
    ~NetGroup();
    NetGroup(const NetGroup&) = delete;
    NetGroup& operator=(const NetGroup&) = delete;

    // Options for every member Net; call before InitFromPath.
    void SetFusion(bool enable) { fusion_ = enable; }
    void SetCudaGraph(bool enable) { graph_ = enable; }

    // devices == NULL or count <= 0: every visible device.  model_path as for Net::InitFromPath (a FTHRB200 container, or
    // the stem of <stem>.param / <stem>.bin).  Returns 0 or a negative code (the first member's failure).
    int InitFromPath(const char* model_path, const int* devices, int count);

    int Size() const { return static_cast<int>(nets_.size()); }
    int Device(int i) const { return devices_[i]; }
    Net* Member(int i) { return nets_[i]; }
    // "nccl" or "cudaMemcpyPeer" (how the weight arena reached the non-root devices), "" before InitFromPath / for one device
    const char* BroadcastTransport() const { return transport_.c_str(); }

    // Forward `batch` images (host, NCHW, shaped like the model's Input layer) sharded over the devices and gather
    // `blob_name` of every image, in input order, into `host_out` (may be NULL).  Pinned host memory makes the copies
    // asynchronous.  Rows [*lo, *hi) of the batch ran on member i: ShardRange.
    int ForwardBatch(const float* host_nchw, int batch, const char* blob_name, float* host_out);
    static void ShardRange(int batch, int members, int i, int* lo, int* hi);
    int Synchronize();

private:
    int BroadcastArena();
    void Clear();
    std::vector<Net*> nets_;
    std::vector<int> devices_;
    std::string transport_;
    bool fusion_ = true, graph_ = true;
};

}  // inline namespace b200
}  // namespace feather
