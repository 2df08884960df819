// Built-in layer registrations: the 13 ncnn type names the reference serves (/root/reference/src/layer_factory.cpp:53-68)
// mapped onto this host's layer classes.  A table of {type name, factory} pairs instead of one macro line per layer; the
// DEFINE_LAYER_CREATOR / REGISTER_LAYER_CREATOR macros of <feather/layer_factory.h> remain the plugin API for user layers.
#include <feather/layer_factory.h>

#include <stdio.h>

#include "layers/batchnorm_layer.h"
#include "layers/conv_layer.h"
#include "layers/eltwise_layer.h"
#include "layers/inner_product_layer.h"
#include "layers/misc_layers.h"
#include "layers/pooling_layer.h"
#include "layers/relu_layer.h"
#include "layers/scale_layer.h"

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

namespace {

template <class L>
Layer* make_layer(RuntimeParameter<float>* rt_param) {
    return new L(rt_param);
}

struct BuiltIn {
    const char* ncnn_type;
    LayerRegistry::Creator create;
};

// "Convolution" and "ConvolutionDepthWise" share one class: group == channels selects the depthwise algorithm.
const BuiltIn kBuiltIns[] = {
    {"Input", make_layer<InputLayer>},
    {"Convolution", make_layer<ConvLayer>},
    {"ConvolutionDepthWise", make_layer<ConvLayer>},
    {"InnerProduct", make_layer<InnerProductLayer>},
    {"Pooling", make_layer<PoolingLayer>},
    {"BatchNorm", make_layer<BatchNormLayer>},
    {"Scale", make_layer<ScaleLayer>},
    {"Eltwise", make_layer<EltwiseLayer>},
    {"ReLU", make_layer<ReluLayer>},
    {"Softmax", make_layer<SoftmaxLayer>},
    {"Dropout", make_layer<DropoutLayer>},
    {"Split", make_layer<SplitLayer>},
    {"Concat", make_layer<ConcatLayer>},
};

}  // namespace

LayerRegistry::CreatorRegistry& LayerRegistry::Registry() {
    static CreatorRegistry* table = new CreatorRegistry();  // never destroyed: layers may be created during exit
    return *table;
}

void LayerRegistry::AddCreator(const std::string& type, Creator creator) { Registry()[type] = creator; }

Layer* LayerRegistry::CreateLayer(std::string type, RuntimeParameter<float>* rt_param) {
    const CreatorRegistry& table = Registry();
    const CreatorRegistry::const_iterator hit = table.find(type);
    if (hit == table.end()) {
        fprintf(stderr, "Layer type %s is not supported in FeatherCNN...Aborting\n", type.c_str());
        return NULL;
    }
    return hit->second(rt_param);
}

void register_layer_creators() {
    for (const BuiltIn& b : kBuiltIns) LayerRegistry::AddCreator(b.ncnn_type, b.create);
}

}  // inline namespace b200
}  // namespace feather
