// feather::Layer — operator plugin interface, same virtuals and public members as the reference
// (/root/reference/src/layer.h:28-89).
#pragma once

#include <string>
#include <vector>

#include "blob.h"
#include "mempool.h"
#include "ncnn/modelbin.h"
#include "ncnn/paramdict.h"
#include "rt_param.h"
#include "utils.h"

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class Layer {
public:
    explicit Layer(RuntimeParameter<float>* rt_param);
    virtual ~Layer();

    // Load layer-specific parameters and weights from ncnn model files (layer.h:40-46).
    virtual int LoadParam(const ncnn::ParamDict& pd);
    virtual int LoadWeights(const ncnn::ModelBin& mb);
    // Infer top shapes (batch included) and allocate them (layer.h:48-52).
    virtual int Reshape();
    // One-time weight preprocessing (layer.h:54-59).
    virtual int Init();
    virtual int Forward();
    virtual int ForwardReshape();

    // Fusion hooks (layer.h:61-68).  The reference never calls TryFuse; here
    // Net::SetFusion(true) makes Net::LoadParam run it over consecutive layers.
    virtual int Fuse(Layer* next_layer);
    int TryFuse(Layer* next_layer);
    bool fusible() const;

    int FindBottomIDByName(std::string name);
    int FindTopIDByName(std::string name);

public:  // "We just make everything public" (layer.h:74)
    std::string name;
    std::string type;
    std::vector<Blob<float>*> bottoms;
    std::vector<Blob<float>*> tops;
    std::vector<Blob<float>*> weights;
    bool _fusible;
    bool _inplace;
    bool _fused_away = false;  // set on a layer absorbed into its producer; Net skips it
    CommonMemPool<float>* common_mempool;
    RuntimeParameter<float>* rt_param;

protected:
    void* stream() const { return rt_param->stream(); }
    // New weight blob staged on the host until Net::LoadWeights binds it into the device arena.
    Blob<float>* NewWeightBlob(const std::string& blob_name, int n, int c, int h, int w);
};

}  // inline namespace b200
}  // namespace feather
