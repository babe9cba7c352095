// Common definitions shared by the sm_100a kernels and by the host-side
// thread-emulation harness under tests/emu (which compiles the very same
// per-CTA phase functions with g++ to check their index arithmetic on a
// machine without a GPU; it is test infrastructure, never a product path).
#pragma once
#include <cstdint>
#include <cmath>
#include "../../include/ta_b200.h"

#if defined(__CUDACC__)
#define TA_HD __host__ __device__ __forceinline__
#define TA_D __device__ __forceinline__
#else
#define TA_HD inline
#define TA_D inline
#endif

namespace ta {

template <typename R>
struct alignas(2 * sizeof(R)) cplx {
    R x, y;
};

template <typename R> TA_HD cplx<R> cmake(R a, R b) { cplx<R> c; c.x = a; c.y = b; return c; }
template <typename R> TA_HD cplx<R> cadd(cplx<R> a, cplx<R> b) { return cmake<R>(a.x + b.x, a.y + b.y); }
template <typename R> TA_HD cplx<R> csub(cplx<R> a, cplx<R> b) { return cmake<R>(a.x - b.x, a.y - b.y); }
template <typename R> TA_HD cplx<R> cmul(cplx<R> a, cplx<R> b) {
    return cmake<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
template <typename R> TA_HD cplx<R> cmulc(cplx<R> a, cplx<R> b) {
    return cmake<R>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename R> TA_HD cplx<R> cconj(cplx<R> a) { return cmake<R>(a.x, -a.y); }
// multiply by -i (forward rotation) / +i
template <typename R> TA_HD cplx<R> cmul_mi(cplx<R> a) { return cmake<R>(a.y, -a.x); }
template <typename R> TA_HD cplx<R> cmul_pi(cplx<R> a) { return cmake<R>(-a.y, a.x); }
template <typename R> TA_HD R cnorm2(cplx<R> a) { return a.x * a.x + a.y * a.y; }

// Error codes (TA_OK, TA_ERR_*) come from the public header.

// Analysis kinds for the windowed kernel.
enum : int { TA_WIN_PRODUCT = 0, TA_WIN_SQDIFF = 1 };

constexpr int TA_MAX_PASSES = 12;

}  // namespace ta
