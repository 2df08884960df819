// CommonMemPool — one shared DEVICE scratch buffer sized by the largest request
// (mirrors /root/reference/src/mempool.h:32-57, mempool.cpp:32-109: Request() records the maximum,
// GetPtr() allocates lazily).  One Net = one in-flight Forward (SURVEY.md §8b "Threading").
#pragma once

#include <stddef.h>
#include <stdio.h>

#define MEMPOOL_CHECK_RETURN(var)                                               \
    {                                                                           \
        if (!(var)) {                                                           \
            fprintf(stderr, "Err in file %s line %d\n", __FILE__, __LINE__);    \
            return false;                                                       \
        }                                                                       \
    }

inline namespace feather_b200 {  // ABI tag (the reference declares the same global template)

template <typename PTR_TYPE>
class CommonMemPool {
public:
    CommonMemPool() : common_size(0), allocated_size(0), common_memory(nullptr) {}
    ~CommonMemPool() { Free(); }
    bool Request(size_t size_byte);
    bool GetPtr(PTR_TYPE** ptr);
    bool Reset();
    bool Free();
    bool Alloc();
    size_t size_bytes() const { return common_size; }

private:
    size_t common_size;
    size_t allocated_size;
    PTR_TYPE* common_memory;
};

}  // inline namespace feather_b200
