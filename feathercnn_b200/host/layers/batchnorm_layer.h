// BatchNormLayer (mirrors /root/reference/src/layers/batchnorm_layer.h:24-192): statistics are folded into
// alpha/beta at load time (:70-75); Forward is y = beta*x + alpha, optionally followed by a fused Scale layer
// and ReLU (the reference's Fuse is dead code that never stored the Scale weights, :114-137 — here it works).
#pragma once

#include <fcuda.h>
#include <feather/layer.h>
#include <math.h>

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class BatchNormLayer : public Layer {
public:
    explicit BatchNormLayer(RuntimeParameter<float>* rt_param)
        : Layer(rt_param), channels(0), eps(0.f), fuse_scale(false), scale_bias_term(false), fuse_relu(false),
          scale_layer(NULL) {
        _fusible = true;
    }

    int LoadParam(const ncnn::ParamDict& pd) {
        this->channels = pd.get(0, 0);
        this->eps = pd.get(1, 0.f);
        return 0;
    }

    int LoadWeights(const ncnn::ModelBin& mb) {
        ncnn::Mat slope_data = mb.load(channels, 1);
        if (slope_data.empty()) return FEATHER_ERR_WEIGHTS;
        ncnn::Mat mean_data = mb.load(channels, 1);
        if (mean_data.empty()) return FEATHER_ERR_WEIGHTS;
        ncnn::Mat var_data = mb.load(channels, 1);
        if (var_data.empty()) return FEATHER_ERR_WEIGHTS;
        ncnn::Mat bias_data = mb.load(channels, 1);
        if (bias_data.empty()) return FEATHER_ERR_WEIGHTS;
        ncnn::Mat alpha(channels), beta(channels);
        for (int i = 0; i < channels; i++) {
            const float sqrt_var = sqrt(var_data[i] + this->eps);
            alpha[i] = bias_data[i] - slope_data[i] * mean_data[i] / sqrt_var;
            beta[i] = slope_data[i] / sqrt_var;
        }
        Blob<float>* alpha_blob = NewWeightBlob(this->name + "_alpha", 1, 1, 1, channels);
        Blob<float>* beta_blob = NewWeightBlob(this->name + "_beta", 1, 1, 1, channels);
        alpha_blob->CopyDataFromMat(alpha);
        beta_blob->CopyDataFromMat(beta);
        this->weights.push_back(alpha_blob);
        this->weights.push_back(beta_blob);
        return 0;
    }

    int Init() {
        const Blob<float>* p_blob = this->bottoms[0];
        if (this->channels != static_cast<int>(p_blob->channels())) {
            LOGE("Mismatch channel in layer %s, expected %d but the bottom %s has %zu channels.", this->name.c_str(),
                 this->channels, p_blob->name.c_str(), p_blob->channels());
            return FEATHER_ERR_WEIGHTS;
        }
        return 0;
    }

    int Forward() {
        const float* scale_data = NULL;
        const float* scale_bias_data = NULL;
        if (fuse_scale && scale_layer) {
            scale_data = scale_layer->weights[0]->data();
            if (scale_bias_term) scale_bias_data = scale_layer->weights[1]->data();
        }
        const size_t stride = bottoms[0]->width() * bottoms[0]->height();
        return fcuda_batchnorm_forward(tops[0]->data(), bottoms[0]->data(), channels, stride, weights[0]->data(),
                                       weights[1]->data(), scale_data, scale_bias_data, fuse_relu ? 1 : 0,
                                       bottoms[0]->num(), stream());
    }

    int Fuse(Layer* next_layer) {
        if (next_layer->type.compare("Scale") == 0 && !fuse_scale && !fuse_relu) {
            fuse_scale = true;
            scale_layer = next_layer;  // its weight blobs stay alive inside the (skipped) Scale layer
            scale_bias_term = next_layer->weights.size() > 1;
            return 1;
        }
        if (next_layer->type.compare("ReLU") == 0 && !fuse_relu) {
            fuse_relu = true;
            return 1;
        }
        return 0;
    }

private:
    int channels;
    float eps;
    bool fuse_scale;
    bool scale_bias_term;
    bool fuse_relu;
    Layer* scale_layer;
};

}  // inline namespace b200
}  // namespace feather
