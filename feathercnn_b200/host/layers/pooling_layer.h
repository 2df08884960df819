// PoolingLayer (mirrors /root/reference/src/layers/pooling_layer.h:27-155: same ncnn ids, ceil-mode output
// size, and the window start that subtracts both pads of an axis).
#pragma once

#include <fcuda.h>
#include <feather/layer.h>

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class PoolingLayer : public Layer {
public:
    explicit PoolingLayer(RuntimeParameter<float>* rt_param) : Layer(rt_param), stride_h(1), stride_w(1) {}

    int LoadParam(const ncnn::ParamDict& pd) {  // pooling_layer.h:93-110
        pooling_type = pd.get(0, 0);
        kernel_w = pd.get(1, 0);
        kernel_h = pd.get(11, kernel_w);
        stride_w = pd.get(2, 1);
        stride_h = pd.get(12, stride_w);
        pad_left = pd.get(3, 0);
        pad_right = pd.get(14, pad_left);
        pad_top = pd.get(13, pad_left);
        pad_bottom = pd.get(15, pad_top);
        global_pooling = pd.get(4, 0);
        tf_pad_mode = pd.get(5, 0);
        return 0;
    }

    int Reshape() {  // pooling_layer.h:112-134
        const Blob<float>* bottom_blob = bottoms[0];
        input_h = bottom_blob->height();
        input_w = bottom_blob->width();
        input_channels = bottom_blob->channels();
        output_channels = input_channels;
        if (global_pooling) {
            kernel_h = input_h;
            kernel_w = input_w;
            output_h = 1;
            output_w = 1;
        } else {
            if (kernel_h <= 0 || kernel_w <= 0 || stride_h <= 0 || stride_w <= 0) return FEATHER_ERR_WEIGHTS;
            output_h = fcuda_pooling_out_dim(input_h, pad_top, pad_bottom, kernel_h, stride_h);
            output_w = fcuda_pooling_out_dim(input_w, pad_left, pad_right, kernel_w, stride_w);
        }
        this->tops[0]->ReshapeWithRealloc(bottom_blob->num(), output_channels, output_h, output_w);
        return 0;
    }

    int Forward() {
        return fcuda_pooling_forward(tops[0]->data(), bottoms[0]->data(), input_channels, input_h, input_w, pooling_type,
                                     kernel_h, kernel_w, stride_h, stride_w, pad_left, pad_right, pad_top, pad_bottom,
                                     global_pooling ? 1 : 0, bottoms[0]->num(), stream());
    }

    // 2x2 / stride 2 / no padding / max: the pooling a convolution can absorb into its epilogue (Net::ApplyFusion)
    bool IsMax2x2Stride2() const {
        return !global_pooling && pooling_type == 0 && kernel_h == 2 && kernel_w == 2 && stride_h == 2 && stride_w == 2 &&
               pad_left == 0 && pad_right == 0 && pad_top == 0 && pad_bottom == 0;
    }

private:
    int input_h, input_w, input_channels, output_h, output_w, output_channels;
    int pad_left, pad_bottom, pad_right, pad_top;
    int kernel_h, kernel_w, stride_h, stride_w;
    bool global_pooling;
    int pooling_type;
    int tf_pad_mode;
};

}  // inline namespace b200
}  // namespace feather
