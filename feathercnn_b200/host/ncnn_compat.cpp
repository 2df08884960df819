// ncnn .param / .bin readers (see include/feather/ncnn/*.h for the reference citations).
#include <feather/ncnn/modelbin.h>
#include <feather/ncnn/paramdict.h>

#include <ctype.h>
#include <stdint.h>
#include <string.h>

#include <vector>

namespace ncnn {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

void ParamDict::clear() {
    for (int i = 0; i < NCNN_MAX_PARAM_COUNT; i++) {
        params[i].loaded = 0;
        params[i].i = 0;
        params[i].v = Mat();
    }
}

static bool vstr_is_float(const char* vstr) {
    for (int j = 0; j < 16 && vstr[j] != '\0'; j++)
        if (vstr[j] == '.' || tolower(vstr[j]) == 'e') return true;
    return false;
}

static int store_value(ParamDict* pd, int id, const char* vstr) {
    if (id < 0 || id >= NCNN_MAX_PARAM_COUNT) return -1;
    if (vstr_is_float(vstr)) {
        float f;
        if (sscanf(vstr, "%f", &f) != 1) return -1;
        pd->set(id, f);
    } else {
        int i;
        if (sscanf(vstr, "%d", &i) != 1) return -1;
        pd->set(id, i);
    }
    return 0;
}

int ParamDict::load_param(FILE* fp) {
    clear();
    int id = 0;
    // "0=100 1=1.250000 -23303=5,0.1,0.2,0.4,0.8,1.0"; stops at the first token that is not "<int>="
    while (fscanf(fp, "%d=", &id) == 1) {
        const bool is_array = id <= -23300;
        if (is_array) id = -id - 23300;
        if (id < 0 || id >= NCNN_MAX_PARAM_COUNT) return -1;
        if (is_array) {
            int len = 0;
            if (fscanf(fp, "%d", &len) != 1 || len < 0) return -1;
            Mat v(len);
            for (int j = 0; j < len; j++) {
                char vstr[16];
                if (fscanf(fp, ",%15[^,\n ]", vstr) != 1) return -1;
                if (vstr_is_float(vstr)) {
                    if (sscanf(vstr, "%f", &static_cast<float*>(v)[j]) != 1) return -1;
                } else {
                    if (sscanf(vstr, "%d", &reinterpret_cast<int*>(v.data)[j]) != 1) return -1;
                }
            }
            set(id, v);
        } else {
            char vstr[16];
            if (fscanf(fp, "%15s", vstr) != 1) return -1;
            if (store_value(this, id, vstr)) return -1;
        }
    }
    return 0;
}

int ParamDict::load_param_mem(const char*& mem) {
    clear();
    for (;;) {
        int id = 0, consumed = 0;
        if (sscanf(mem, "%d=%n", &id, &consumed) != 1 || consumed == 0) break;
        mem += consumed;
        const bool is_array = id <= -23300;
        if (is_array) id = -id - 23300;
        if (id < 0 || id >= NCNN_MAX_PARAM_COUNT) return -1;
        if (is_array) {
            int len = 0;
            consumed = 0;
            if (sscanf(mem, "%d%n", &len, &consumed) != 1 || len < 0) return -1;
            mem += consumed;
            Mat v(len);
            for (int j = 0; j < len; j++) {
                char vstr[16];
                consumed = 0;
                if (sscanf(mem, ",%15[^,\n ]%n", vstr, &consumed) != 1) return -1;
                mem += consumed;
                if (vstr_is_float(vstr)) sscanf(vstr, "%f", &static_cast<float*>(v)[j]);
                else sscanf(vstr, "%d", &reinterpret_cast<int*>(v.data)[j]);
            }
            set(id, v);
        } else {
            char vstr[16];
            consumed = 0;
            if (sscanf(mem, "%15s%n", vstr, &consumed) != 1) return -1;
            mem += consumed;
            if (store_value(this, id, vstr)) return -1;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
static float half_to_float(uint16_t h) {
    const uint32_t sign = (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, out;
    if (exp == 0) {
        if (man == 0) out = sign;
        else {  // subnormal
            exp = 127 - 15 + 1;
            while ((man & 0x400u) == 0) { man <<= 1; --exp; }
            man &= 0x3ffu;
            out = sign | (exp << 23) | (man << 13);
        }
    } else if (exp == 31) {
        out = sign | 0x7f800000u | (man << 13);
    } else {
        out = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &out, 4);
    return f;
}

struct Reader {
    virtual ~Reader() {}
    virtual bool read(void* dst, size_t n) = 0;
};
struct FileReader : Reader {
    FILE* fp;
    explicit FileReader(FILE* f) : fp(f) {}
    bool read(void* dst, size_t n) override { return n == 0 || fread(dst, n, 1, fp) == 1; }
};
struct MemReader : Reader {
    const unsigned char*& mem;
    const unsigned char* end;  // one past the buffer, or NULL when the caller gave no size (the reference's contract)
    MemReader(const unsigned char*& m, const unsigned char* e) : mem(m), end(e) {}
    bool read(void* dst, size_t n) override {
        if (end && (mem > end || n > static_cast<size_t>(end - mem))) return false;  // truncated / corrupt model
        memcpy(dst, mem, n);
        mem += n;
        return true;
    }
};

static Mat load_blob(Reader& r, int w, int type) {
    if (w < 0) return Mat();
    if (type == 0) {
        unsigned char flag[4];
        if (!r.read(flag, 4)) { fprintf(stderr, "ModelBin read flag_struct failed\n"); return Mat(); }
        uint32_t tag;
        memcpy(&tag, flag, 4);
        const unsigned sum = flag[0] + flag[1] + flag[2] + flag[3];
        if (tag == 0x01306B47u) {  // fp16
            std::vector<uint16_t> h(alignSize(static_cast<size_t>(w) * 2, 4) / 2);
            if (!r.read(h.data(), h.size() * 2)) return Mat();
            Mat m(w);
            for (int i = 0; i < w; i++) m[i] = half_to_float(h[i]);
            return m;
        }
        if (tag == 0x000D4B38u) {  // int8 weights: not supported by FeatherCNN (conv_layer.h:49-54)
            fprintf(stderr, "ModelBin: int8 weight blobs are not supported\n");
            return Mat();
        }
        Mat m(w);
        if (sum != 0 && tag != 0x0002C056u) {  // 256-entry LUT + uint8 indices
            float table[256];
            if (!r.read(table, sizeof(table))) return Mat();
            std::vector<unsigned char> idx(alignSize(static_cast<size_t>(w), 4));
            if (!r.read(idx.data(), idx.size())) return Mat();
            for (int i = 0; i < w; i++) m[i] = table[idx[i]];
            return m;
        }
        if (!r.read(m.data, static_cast<size_t>(w) * 4)) { fprintf(stderr, "ModelBin read weight_data failed\n"); return Mat(); }
        return m;
    }
    if (type == 1) {
        Mat m(w);
        if (!r.read(m.data, static_cast<size_t>(w) * 4)) { fprintf(stderr, "ModelBin read weight_data failed\n"); return Mat(); }
        return m;
    }
    fprintf(stderr, "ModelBin load type %d not implemented\n", type);
    return Mat();
}

Mat ModelBinFromStdio::load(int w, int type) const {
    if (!binfp) return Mat();
    FileReader r(binfp);
    return load_blob(r, w, type);
}

Mat ModelBinFromMemory::load(int w, int type) const {
    if (!mem) return Mat();
    MemReader r(mem, end);
    return load_blob(r, w, type);
}

Mat ModelBinSizesOnly::load(int w, int type) const {
    (void)type;
    Mat m(w > 0 ? w : 1);
    memset(m.data, 0, static_cast<size_t>(w > 0 ? w : 1) * 4);
    if (w <= 0) return Mat();
    return m;
}

}  // inline namespace b200
}  // namespace ncnn
