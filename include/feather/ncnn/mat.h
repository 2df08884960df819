// Minimal ncnn::Mat — only what feather::Net's public API and the model loader need
// (the reference vendors the full class: /root/reference/src/ncnn/mat.h:28-290).  Host memory, fp32
// (or raw bytes), reference-counted, channel step aligned to 16 bytes like ncnn (mat.h:288,668) so
// user code that fills Mats channel by channel behaves the same.
#pragma once

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <memory>

namespace ncnn {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

inline size_t alignSize(size_t sz, int n) { return (sz + n - 1) & -static_cast<size_t>(n); }

class Mat {
public:
    Mat() : data(nullptr), elemsize(0), dims(0), w(0), h(0), c(0), cstep(0) {}
    explicit Mat(int _w, size_t _elemsize = 4u) : Mat() { create(_w, _elemsize); }
    Mat(int _w, int _h, size_t _elemsize = 4u) : Mat() { create(_w, _h, _elemsize); }
    Mat(int _w, int _h, int _c, size_t _elemsize = 4u) : Mat() { create(_w, _h, _c, _elemsize); }
    // external (non-owning) data
    Mat(int _w, void* _data, size_t _elemsize = 4u)
        : data(_data), elemsize(_elemsize), dims(1), w(_w), h(1), c(1), cstep(static_cast<size_t>(_w)) {}
    Mat(int _w, int _h, int _c, void* _data, size_t _elemsize = 4u)
        : data(_data), elemsize(_elemsize), dims(3), w(_w), h(_h), c(_c) {
        cstep = alignSize(static_cast<size_t>(w) * h * elemsize, 16) / elemsize;
    }

    void create(int _w, size_t _elemsize = 4u) {
        dims = 1; w = _w; h = 1; c = 1; elemsize = _elemsize; cstep = static_cast<size_t>(w);
        alloc(cstep * elemsize);
    }
    void create(int _w, int _h, size_t _elemsize = 4u) {
        dims = 2; w = _w; h = _h; c = 1; elemsize = _elemsize; cstep = static_cast<size_t>(w) * h;
        alloc(cstep * elemsize);
    }
    void create(int _w, int _h, int _c, size_t _elemsize = 4u) {
        dims = 3; w = _w; h = _h; c = _c; elemsize = _elemsize;
        cstep = alignSize(static_cast<size_t>(w) * h * elemsize, 16) / elemsize;
        alloc(cstep * elemsize * c);
    }

    bool empty() const { return data == nullptr || total() == 0; }
    size_t total() const { return cstep * c; }

    Mat channel(int _c) const {
        Mat m;
        m.owner = owner;
        m.data = static_cast<unsigned char*>(data) + cstep * _c * elemsize;
        m.elemsize = elemsize; m.dims = 2; m.w = w; m.h = h; m.c = 1; m.cstep = static_cast<size_t>(w) * h;
        return m;
    }
    float* row(int y) { return static_cast<float*>(data) + static_cast<size_t>(w) * y; }
    const float* row(int y) const { return static_cast<const float*>(data) + static_cast<size_t>(w) * y; }

    Mat reshape(int _w, int _h, int _c) const {  // dense copy into the aligned-cstep layout
        Mat m(_w, _h, _c, elemsize);
        const size_t plane = static_cast<size_t>(_w) * _h;
        for (int i = 0; i < _c; ++i)
            memcpy(static_cast<unsigned char*>(m.data) + m.cstep * i * elemsize,
                   static_cast<const unsigned char*>(data) + plane * i * elemsize, plane * elemsize);
        return m;
    }

    operator float*() { return static_cast<float*>(data); }
    operator const float*() const { return static_cast<const float*>(data); }
    float& operator[](size_t i) { return static_cast<float*>(data)[i]; }
    const float& operator[](size_t i) const { return static_cast<const float*>(data)[i]; }

    void* data;
    size_t elemsize;
    int dims;
    int w, h, c;
    size_t cstep;

private:
    void alloc(size_t bytes) {
        void* p = nullptr;
        if (bytes && posix_memalign(&p, 64, bytes) != 0) p = nullptr;
        owner = std::shared_ptr<void>(p, free);
        data = p;
    }
    std::shared_ptr<void> owner;
};

}  // inline namespace b200
}  // namespace ncnn
