// ncnn::ModelBin — weight-blob reader of the ncnn .bin format (/root/reference/src/ncnn/modelbin.h,
// modelbin.cpp:47-293).  type 0 blobs carry a 4-byte tag (0 = raw fp32, 0x01306B47 = fp16, else a 256-entry
// LUT + uint8 indices), type 1 blobs are raw fp32.  int8 (0x000D4B38) is rejected like the reference's
// ConvLayer does (conv_layer.h:49-54).
#pragma once

#include <stdio.h>

#include "mat.h"

namespace ncnn {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class ModelBin {
public:
    virtual ~ModelBin() {}
    virtual Mat load(int w, int type) const = 0;
};

class ModelBinFromStdio : public ModelBin {
public:
    explicit ModelBinFromStdio(FILE* fp) : binfp(fp) {}
    Mat load(int w, int type) const override;

private:
    FILE* binfp;
};

class ModelBinFromMemory : public ModelBin {
public:
    // `end` (one past the last byte) is optional: the reference's reader takes no size (modelbin.cpp:200-293) and a
    // truncated buffer makes it read past the end; with `end` a short buffer yields an empty Mat like a short fread.
    explicit ModelBinFromMemory(const unsigned char*& _mem, const unsigned char* _end = nullptr) : mem(_mem), end(_end) {}
    Mat load(int w, int type) const override;

private:
    const unsigned char*& mem;
    const unsigned char* end;
};

// Returns zero-filled blobs of the requested size without touching any file: lets every rank of a
// multi-GPU job lay out its weight arena identically before rank 0's arena is broadcast over NCCL.
class ModelBinSizesOnly : public ModelBin {
public:
    Mat load(int w, int type) const override;
};

}  // inline namespace b200
}  // namespace ncnn
