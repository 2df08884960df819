// Built-in layer registrations — the same 13 ncnn type names as the reference
// (/root/reference/src/layer_factory.cpp:53-68).
#include <feather/layer_factory.h>

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

DEFINE_LAYER_CREATOR(Input)
DEFINE_LAYER_CREATOR(Conv)
DEFINE_LAYER_CREATOR(Relu)
DEFINE_LAYER_CREATOR(Pooling)
DEFINE_LAYER_CREATOR(InnerProduct)
DEFINE_LAYER_CREATOR(Dropout)
DEFINE_LAYER_CREATOR(Softmax)
DEFINE_LAYER_CREATOR(BatchNorm)
DEFINE_LAYER_CREATOR(Scale)
DEFINE_LAYER_CREATOR(Split)
DEFINE_LAYER_CREATOR(Eltwise)
DEFINE_LAYER_CREATOR(Concat)

void register_layer_creators() {
    REGISTER_LAYER_CREATOR(Input, Input);
    REGISTER_LAYER_CREATOR(Convolution, Conv);
    REGISTER_LAYER_CREATOR(ConvolutionDepthWise, Conv);
    REGISTER_LAYER_CREATOR(ReLU, Relu);
    REGISTER_LAYER_CREATOR(Pooling, Pooling);
    REGISTER_LAYER_CREATOR(InnerProduct, InnerProduct);
    REGISTER_LAYER_CREATOR(Dropout, Dropout);
    REGISTER_LAYER_CREATOR(Softmax, Softmax);
    REGISTER_LAYER_CREATOR(BatchNorm, BatchNorm);
    REGISTER_LAYER_CREATOR(Scale, Scale);
    REGISTER_LAYER_CREATOR(Split, Split);
    REGISTER_LAYER_CREATOR(Eltwise, Eltwise);
    REGISTER_LAYER_CREATOR(Concat, Concat);
}

}  // inline namespace b200
}  // namespace feather
