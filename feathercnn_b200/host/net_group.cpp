// feather::NetGroup — see include/feather/net_group.h.
#include <feather/net_group.h>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <feather/utils.h>

namespace feather {
inline namespace b200 {

namespace {

// The five NCCL entry points this file needs, bound at run time (nccl.h:  ncclCommInitAll, ncclBroadcast, ncclGroupStart,
// ncclGroupEnd, ncclCommDestroy; ncclFloat32 == 7, ncclSuccess == 0).
struct Nccl {
    typedef void* Comm;
    int (*CommInitAll)(Comm*, int, const int*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    void* handle = nullptr;
    bool ok() const { return CommInitAll && Broadcast && GroupStart && GroupEnd && CommDestroy; }
};

Nccl load_nccl() {
    Nccl n;
    const char* names[] = {getenv("FEATHER_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
        if (!name || !*name) continue;
        n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (n.handle) break;
    }
    if (!n.handle) return n;
    n.CommInitAll = reinterpret_cast<int (*)(Nccl::Comm*, int, const int*)>(dlsym(n.handle, "ncclCommInitAll"));
    n.Broadcast = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, Nccl::Comm, cudaStream_t)>(dlsym(n.handle, "ncclBroadcast"));
    n.GroupStart = reinterpret_cast<int (*)()>(dlsym(n.handle, "ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<int (*)()>(dlsym(n.handle, "ncclGroupEnd"));
    n.CommDestroy = reinterpret_cast<int (*)(Nccl::Comm)>(dlsym(n.handle, "ncclCommDestroy"));
    return n;
}

struct DeviceGuard {
    int saved = 0;
    DeviceGuard() { cudaGetDevice(&saved); }
    ~DeviceGuard() { cudaSetDevice(saved); }
};

}  // namespace

NetGroup::NetGroup() {}
NetGroup::~NetGroup() { Clear(); }

void NetGroup::Clear() {
    DeviceGuard guard;
    for (size_t i = 0; i < nets_.size(); ++i) {
        cudaSetDevice(devices_[i]);
        delete nets_[i];
    }
    nets_.clear();
    devices_.clear();
    transport_.clear();
}

void NetGroup::ShardRange(int batch, int members, int i, int* lo, int* hi) {
    // contiguous; the first batch % members members get one extra image (the rule of feathercnn_b200/dist.py: shard_range)
    const int m = members > 0 ? members : 1;
    const int base = batch / m, extra = batch % m;
    *lo = i * base + (i < extra ? i : extra);
    *hi = *lo + base + (i < extra ? 1 : 0);
}

int NetGroup::InitFromPath(const char* model_path, const int* devices, int count) {
    if (!model_path) return -1;
    Clear();
    DeviceGuard guard;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
        LOGE("NetGroup: no CUDA device");
        return FEATHER_ERR_CUDA;
    }
    if (!devices || count <= 0) {
        for (int d = 0; d < visible; ++d) devices_.push_back(d);
    } else {
        for (int i = 0; i < count; ++i) {
            if (devices[i] < 0 || devices[i] >= visible) {
                LOGE("NetGroup: device %d is not visible (%d devices)", devices[i], visible);
                devices_.clear();
                return -1;
            }
            for (int j = 0; j < i; ++j)
                if (devices[j] == devices[i]) {
                    LOGE("NetGroup: device %d listed twice", devices[i]);
                    devices_.clear();
                    return -1;
                }
            devices_.push_back(devices[i]);
        }
    }
    // the model file is read by the root member only; the others get the graph from the same bytes and the weights over NVLink
    std::vector<unsigned char> container;
    bool is_container = false;
    if (FILE* fp = fopen(model_path, "rb")) {
        char magic[8] = {0};
        if (fread(magic, 1, 8, fp) == 8 && memcmp(magic, "FTHRB200", 8) == 0) {
            fseek(fp, 0, SEEK_END);
            const long len = ftell(fp);
            fseek(fp, 0, SEEK_SET);
            if (len > 0) {
                container.resize(static_cast<size_t>(len));
                is_container = fread(container.data(), 1, container.size(), fp) == container.size();
            }
        }
        fclose(fp);
    }
    const std::string stem(model_path);
    for (size_t i = 0; i < devices_.size(); ++i) {
        if (cudaSetDevice(devices_[i]) != cudaSuccess) {
            Clear();
            return FEATHER_ERR_CUDA;
        }
        Net* net = new Net();
        nets_.push_back(net);
        net->SetFusion(fusion_);
        net->SetCudaGraph(graph_);
        int rc;
        if (i == 0) {
            rc = is_container ? net->InitFromBuffer(container.data(), container.size()) : net->InitFromPath(model_path);
        } else {
            rc = is_container ? net->InitGraphFromBuffer(container.data(), container.size()) : net->LoadParam((stem + ".param").c_str());
            if (rc == 0) rc = net->PrepareWeightArena();
        }
        if (rc) {
            LOGE("NetGroup: member %d (device %d) failed to load: %d", (int)i, devices_[i], rc);
            Clear();
            return rc;
        }
        if (i > 0 && net->WeightArenaFloats() != nets_[0]->WeightArenaFloats()) {
            LOGE("NetGroup: weight arena layout differs between members");
            Clear();
            return -1;
        }
    }
    if (nets_.size() > 1) {
        const int rc = BroadcastArena();
        if (rc) {
            Clear();
            return rc;
        }
        for (size_t i = 1; i < nets_.size(); ++i)
            if (nets_[i]->AttachWeights()) {
                Clear();
                return -1;
            }
    }
    return 0;
}

// One broadcast of the root's weight arena to every other member.
int NetGroup::BroadcastArena() {
    const size_t floats = nets_[0]->WeightArenaFloats();
    if (floats == 0) return 0;
    const int n = static_cast<int>(nets_.size());
    static Nccl nccl = load_nccl();
    if (nccl.ok() && !getenv("FEATHER_NO_NCCL")) {
        std::vector<Nccl::Comm> comms(n, nullptr);
        if (nccl.CommInitAll(comms.data(), n, devices_.data()) == 0) {
            std::vector<cudaStream_t> streams(n, nullptr);
            int rc = 0;
            for (int i = 0; i < n && !rc; ++i) {
                cudaSetDevice(devices_[i]);
                if (cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking) != cudaSuccess) rc = FEATHER_ERR_CUDA;
            }
            if (!rc) {
                nccl.GroupStart();
                for (int i = 0; i < n; ++i) {
                    cudaSetDevice(devices_[i]);
                    if (nccl.Broadcast(nets_[0]->WeightArena(), nets_[i]->WeightArena(), floats, /*ncclFloat32*/ 7, /*root*/ 0,
                                       comms[i], streams[i]) != 0)
                        rc = -1;
                }
                if (nccl.GroupEnd() != 0) rc = -1;
            }
            for (int i = 0; i < n; ++i) {
                cudaSetDevice(devices_[i]);
                if (streams[i]) {
                    if (cudaStreamSynchronize(streams[i]) != cudaSuccess) rc = FEATHER_ERR_CUDA;
                    cudaStreamDestroy(streams[i]);
                }
                if (comms[i]) nccl.CommDestroy(comms[i]);
            }
            if (rc == 0) {
                transport_ = "nccl";
                return 0;
            }
            LOGE("NetGroup: NCCL broadcast failed (%d), falling back to peer copies", rc);
        }
    }
    // no NCCL: the arena travels device to device (over NVLink when peer access is available, else through the host)
    for (int i = 1; i < n; ++i) {
        cudaSetDevice(devices_[i]);
        if (cudaMemcpyPeer(nets_[i]->WeightArena(), devices_[i], nets_[0]->WeightArena(), devices_[0], floats * sizeof(float)) !=
            cudaSuccess) {
            LOGE("NetGroup: cudaMemcpyPeer to device %d failed", devices_[i]);
            return FEATHER_ERR_CUDA;
        }
    }
    transport_ = "cudaMemcpyPeer";
    return 0;
}

int NetGroup::ForwardBatch(const float* host_nchw, int batch, const char* blob_name, float* host_out) {
    if (nets_.empty() || !host_nchw || batch < 1) return -1;
    DeviceGuard guard;
    int c = 0, h = 0, w = 0;
    nets_[0]->InputShape(&c, &h, &w);
    const size_t in_img = static_cast<size_t>(c) * h * w;
    const int n = static_cast<int>(nets_.size());
    std::vector<int> ticket(n, -1);
    size_t out_img = 0;
    for (int i = 0; i < n; ++i) {
        int lo, hi;
        ShardRange(batch, n, i, &lo, &hi);
        if (hi <= lo) continue;
        if (cudaSetDevice(devices_[i]) != cudaSuccess) return FEATHER_ERR_CUDA;
        float* out_i = nullptr;
        if (blob_name && host_out) {
            if (out_img == 0 && lo > 0) return -1;  // (the first non-empty shard starts at row 0)
            out_i = host_out + static_cast<size_t>(lo) * out_img;
        }
        const int t = nets_[i]->SubmitBatch(host_nchw + static_cast<size_t>(lo) * in_img, hi - lo, blob_name, out_i);
        if (t < 0) return t;
        ticket[i] = t;
        if (blob_name && host_out && out_img == 0) {  // the blob's per-image size is known once a member has reshaped
            size_t total = 0;
            if (nets_[i]->GetBlobDataSize(&total, blob_name) != 0 || total == 0) return -1;
            out_img = total / static_cast<size_t>(hi - lo);
        }
    }
    int rc = 0;
    for (int i = 0; i < n; ++i) {
        if (ticket[i] < 0) continue;
        cudaSetDevice(devices_[i]);
        const int r = nets_[i]->WaitBatch(ticket[i]);
        if (r && !rc) rc = r;
    }
    return rc;
}

int NetGroup::Synchronize() {
    DeviceGuard guard;
    int rc = 0;
    for (size_t i = 0; i < nets_.size(); ++i) {
        cudaSetDevice(devices_[i]);
        const int r = nets_[i]->Synchronize();
        if (r && !rc) rc = r;
    }
    return rc;
}

}  // inline namespace b200
}  // namespace feather
