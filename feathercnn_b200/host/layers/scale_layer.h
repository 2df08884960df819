// ScaleLayer (mirrors /root/reference/src/layers/scale_layer.h:22-115).
#pragma once

#include <fcuda.h>
#include <feather/layer.h>

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class ScaleLayer : public Layer {
public:
    explicit ScaleLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param), channels(0), bias_term(0), scale_data_size(0) {}

    int LoadParam(const ncnn::ParamDict& pd) {
        scale_data_size = pd.get(0, 0);
        bias_term = pd.get(1, 0);
        if (scale_data_size < 0) {
            LOGE("feather doesn't accept negative scale data size, please use ncnn to run this model.");
            return FEATHER_ERR_WEIGHTS;
        }
        return 0;
    }

    int LoadWeights(const ncnn::ModelBin& mb) {
        if (scale_data_size == -233) return 0;
        ncnn::Mat scale_mat = mb.load(scale_data_size, 1);
        if (scale_mat.empty()) return FEATHER_ERR_WEIGHTS;
        channels = scale_data_size;
        Blob<float>* scale_blob = NewWeightBlob(this->name + "_scale", 1, 1, 1, static_cast<int>(channels));
        scale_blob->CopyDataFromMat(scale_mat);
        weights.push_back(scale_blob);
        if (bias_term) {
            ncnn::Mat bias_mat = mb.load(scale_data_size, 1);
            if (bias_mat.empty()) return FEATHER_ERR_WEIGHTS;
            Blob<float>* bias_blob = NewWeightBlob(this->name + "_bias", 1, 1, 1, static_cast<int>(channels));
            bias_blob->CopyDataFromMat(bias_mat);
            weights.push_back(bias_blob);
        }
        return 0;
    }

    int Forward() {
        if (weights.empty()) return FEATHER_ERR_WEIGHTS;
        if (channels != bottoms[0]->channels()) return FEATHER_ERR_WEIGHTS;
        const size_t stride = bottoms[0]->width() * bottoms[0]->height();
        return fcuda_scale_forward(tops[0]->data(), bottoms[0]->data(), static_cast<int>(channels), stride,
                                   weights[0]->data(), bias_term ? weights[1]->data() : NULL, bottoms[0]->num(), stream());
    }

private:
    size_t channels;
    int bias_term;
    int scale_data_size;
};

}  // inline namespace b200
}  // namespace feather
