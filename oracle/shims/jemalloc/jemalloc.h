// Stand-in for jemalloc (absent from this image).  TEST INFRASTRUCTURE ONLY.
// The reference's alloc.h needs an aligned allocation and a sized free.
#pragma once
#include <cstdlib>
static inline void* je_aligned_alloc(std::size_t alignment, std::size_t size) {
    return std::aligned_alloc(alignment, size);
}
static inline void je_sdallocx(void* p, std::size_t, int) { std::free(p); }
