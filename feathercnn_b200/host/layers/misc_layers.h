// Input, Split, Concat, Softmax, Dropout layers (mirror /root/reference/src/layers/input_layer.h:25-52,
// split_layer.h:22-54, concat_layer.h:22-83, softmax_layer.h:24-57, dropout_layer.h:22-60).
#pragma once

#include <fcuda.h>
#include <feather/layer.h>

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class InputLayer : public Layer {
public:
    explicit InputLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param), w(0), h(0), c(0) {}
    int LoadParam(const ncnn::ParamDict& pd) {
        w = pd.get(0, 0);
        h = pd.get(1, 0);
        c = pd.get(2, 0);
        return 0;
    }
    int Reshape() { return 0; }  // shape comes from FeedInput (input_layer.h:40-44)
    int Init() { return 0; }
    int w, h, c;  // declared input shape, used by the README-era Forward(float*)
};

// Split: the reference memcpy()s the bottom into every top (split_layer.h:43-53).  No layer in this engine
// writes into its bottoms, so the tops are zero-copy views of the bottom instead.
class SplitLayer : public Layer {
public:
    explicit SplitLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param) {}
    int Reshape() {
        const Blob<float>* b = bottoms[0];
        for (size_t i = 0; i < tops.size(); ++i) tops[i]->ViewExternal(b->data(), b->num(), b->channels(), b->height(), b->width());
        return 0;
    }
    int Forward() { return 0; }
};

class ConcatLayer : public Layer {
public:
    explicit ConcatLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param), axis(0) {}
    int LoadParam(const ncnn::ParamDict& pd) {
        this->axis = pd.get(0, 0);
        return 0;
    }
    // ncnn axis of a 3-D blob: 0 = channels, 1 = height, 2 = width.  The reference handles axis 0 only
    // (concat_layer.h:50-54); 1 and 2 are this engine's extension.
    int Reshape() {
        const Blob<float>* first_blob = this->bottoms[0];
        if (axis < 0 || axis > 2) {
            LOGE("Concat layer %s: unsupported axis %d", this->name.c_str(), axis);
            return FEATHER_ERR_WEIGHTS;
        }
        size_t dims[3] = {first_blob->channels(), first_blob->height(), first_blob->width()};
        const size_t num = first_blob->num();
        for (size_t i = 1; i < bottoms.size(); ++i) {
            const Blob<float>* p_blob = bottoms[i];
            const size_t d[3] = {p_blob->channels(), p_blob->height(), p_blob->width()};
            for (int a = 0; a < 3; ++a)
                if (a != axis && d[a] != dims[a]) {
                    LOGE("Images of different shapes cannot be concatenated together");
                    return FEATHER_ERR_WEIGHTS;
                }
            if (num != p_blob->num()) return FEATHER_ERR_WEIGHTS;
            dims[axis] += d[axis];
        }
        tops[0]->ReshapeWithRealloc(num, dims[0], dims[1], dims[2]);
        return 0;
    }
    int Forward() {
        // [outer][mid][inner] copy: outer = everything before the axis, mid = the axis, inner = everything after it
        int offset = 0, rc = 0;
        const Blob<float>* t = tops[0];
        for (size_t i = 0; i < bottoms.size() && rc == 0; ++i) {
            const Blob<float>* b = bottoms[i];
            size_t outer, mid, dst_mid, inner;
            if (axis == 0) { outer = b->num(); mid = b->channels(); dst_mid = t->channels(); inner = t->height() * t->width(); }
            else if (axis == 1) { outer = b->num() * b->channels(); mid = b->height(); dst_mid = t->height(); inner = t->width(); }
            else { outer = b->num() * b->channels() * b->height(); mid = b->width(); dst_mid = t->width(); inner = 1; }
            rc = fcuda_copy_channels(t->data(), static_cast<int>(dst_mid), offset, b->data(), static_cast<int>(mid), inner,
                                     static_cast<int>(outer), stream());
            offset += static_cast<int>(mid);
        }
        return rc;
    }

private:
    int axis;
};

class SoftmaxLayer : public Layer {
public:
    explicit SoftmaxLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param) {}
    int Forward() {
        const Blob<float>* b = bottoms[0];
        return fcuda_softmax_forward(tops[0]->data(), b->data(), b->channels() * b->height() * b->width(), b->num(), stream());
    }
};

class DropoutLayer : public Layer {
public:
    explicit DropoutLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param), scale(1.f) {}
    int LoadParam(const ncnn::ParamDict& pd) {
        scale = pd.get(0, 1.f);
        return 0;
    }
    int Reshape() {
        const Blob<float>* b = bottoms[0];
        if (scale == 1.f) {  // the reference memcpy()s (dropout_layer.h:40-43); a view is equivalent
            tops[0]->ViewExternal(b->data(), b->num(), b->channels(), b->height(), b->width());
            return 0;
        }
        tops[0]->ReshapeWithRealloc(b->num(), b->channels(), b->height(), b->width());
        return 0;
    }
    int Forward() {
        if (scale == 1.f) return 0;
        return fcuda_dropout_forward(tops[0]->data(), bottoms[0]->data(), bottoms[0]->data_size(), scale, stream());
    }

private:
    float scale;
};

}  // inline namespace b200
}  // namespace feather
