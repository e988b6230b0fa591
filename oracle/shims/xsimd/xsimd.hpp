// Stand-in for the xsimd headers (absent from this image; no network).
// TEST INFRASTRUCTURE ONLY: lets the reference's own C++ kernels compile for the
// oracle build (oracle/build_ref.py).  Exposes exactly the names the reference
// templates use: XSIMD_VERSION_MAJOR, xsimd::simd_type<F>::size, broadcast,
// load_aligned, fma, reduce_add, batch::store_aligned.  Implemented with GCC
// vector extensions, so the SIMD width follows the -march flag of the build.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>

#define XSIMD_VERSION_MAJOR 13

namespace xsimd {

#if defined(__AVX512F__)
constexpr std::size_t kVecBytes = 64;
#elif defined(__AVX__)
constexpr std::size_t kVecBytes = 32;
#else
constexpr std::size_t kVecBytes = 16;
#endif

template <typename F>
struct batch {
    static constexpr std::size_t size = kVecBytes / sizeof(F);
    typedef F vec_t __attribute__((vector_size(kVecBytes)));
    vec_t v;
    batch() = default;
    explicit batch(vec_t x) : v(x) {}
    void store_aligned(F* p) const { *reinterpret_cast<vec_t*>(p) = v; }
};

template <typename F>
using simd_type = batch<F>;

template <typename F>
inline batch<F> broadcast(F x) {
    typename batch<F>::vec_t v;
    for (std::size_t i = 0; i < batch<F>::size; ++i) v[i] = x;
    return batch<F>(v);
}

template <typename F>
inline batch<F> load_aligned(const F* p) {
    return batch<F>(*reinterpret_cast<const typename batch<F>::vec_t*>(p));
}

template <typename F>
inline batch<F> fma(const batch<F>& a, const batch<F>& b, const batch<F>& c) {
    return batch<F>(a.v * b.v + c.v);
}

template <typename F>
inline F reduce_add(const batch<F>& a) {
    F s = 0;
    for (std::size_t i = 0; i < batch<F>::size; ++i) s += a.v[i];
    return s;
}

}  // namespace xsimd
