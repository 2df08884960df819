// ReluLayer (mirrors /root/reference/src/layers/relu_layer.h:21-43; ncnn's slope parameter is ignored there too).
#pragma once

#include <fcuda.h>
#include <feather/layer.h>

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class ReluLayer : public Layer {
public:
    explicit ReluLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param) {}
    int Forward() { return fcuda_relu_forward(tops[0]->data(), bottoms[0]->data(), bottoms[0]->data_size(), stream()); }
};

}  // inline namespace b200
}  // namespace feather
