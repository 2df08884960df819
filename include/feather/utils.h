// Logging macros and error codes (mirrors /root/reference/src/utils.h:33-35; quiet by default —
// the reference's unconditional hot-path printf()s are not reproduced).
#pragma once

#include <stdio.h>

#define LOGE(fmt, ...) fprintf(stderr, "[feather E] " fmt "\n", ##__VA_ARGS__)
#ifdef FEATHER_VERBOSE
#define LOGI(fmt, ...) fprintf(stderr, "[feather I] " fmt "\n", ##__VA_ARGS__)
#define LOGD(fmt, ...) fprintf(stderr, "[feather D] " fmt "\n", ##__VA_ARGS__)
#else
#define LOGI(fmt, ...) ((void)0)
#define LOGD(fmt, ...) ((void)0)
#endif

// Layer-level return codes used by the reference (SURVEY.md §8b "Errors").
enum {
    FEATHER_OK = 0,
    FEATHER_ERR_GENERIC = -1,
    FEATHER_ERR_WEIGHTS = -100,      // weights / shape problems (conv_layer.h:121, inner_product_layer.h:110)
    FEATHER_ERR_UNSUPPORTED = -200,  // unsupported parameter (conv_layer.h:46)
    FEATHER_ERR_TOPOLOGY = -300,     // missing producer blob (net.cpp:130)
    FEATHER_ERR_BASE_LAYER = -400,   // base-class Reshape misuse (layer.cpp:111)
    FEATHER_ERR_BAD_DIMS = -500,     // Mat/blob shape mismatch (blob.cpp:84)
    FEATHER_ERR_CUDA = -700,
};

// Param-file magic check, /root/reference/src/utils.cpp:27-43.
inline namespace feather_b200 {
int ChkParamHeader(FILE* fp);
}
