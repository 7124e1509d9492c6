// Strictly rounded scalar arithmetic for the bit-exact solvers.
//
// Parity with the reference is defined bit for bit against its C++ built with
// -ffp-contract=off: every product, sum and
// quotient is rounded on its own.  nvcc contracts a*b+c into FFMA/DFMA by
// default, so the solvers are written on Strict<T>, whose operators lower to
// the __f*_rn / __d*_rn intrinsics that the compiler never fuses (the build
// also passes -fmad=false, -prec-div=true, -ftz=false as a second fence).
#pragma once
#include <cuda_runtime.h>

namespace sksb {

template <typename T>
struct Strict;

template <>
struct Strict<float> {
    float v;
    __device__ __forceinline__ Strict() {}
    __device__ __forceinline__ Strict(float x) : v(x) {}
    friend __device__ __forceinline__ Strict operator+(Strict a, Strict b) { return __fadd_rn(a.v, b.v); }
    friend __device__ __forceinline__ Strict operator-(Strict a, Strict b) { return __fsub_rn(a.v, b.v); }
    friend __device__ __forceinline__ Strict operator*(Strict a, Strict b) { return __fmul_rn(a.v, b.v); }
    friend __device__ __forceinline__ Strict operator/(Strict a, Strict b) { return __fdiv_rn(a.v, b.v); }
    __device__ __forceinline__ Strict operator-() const { return -v; }
};

template <>
struct Strict<double> {
    double v;
    __device__ __forceinline__ Strict() {}
    __device__ __forceinline__ Strict(double x) : v(x) {}
    friend __device__ __forceinline__ Strict operator+(Strict a, Strict b) { return __dadd_rn(a.v, b.v); }
    friend __device__ __forceinline__ Strict operator-(Strict a, Strict b) { return __dsub_rn(a.v, b.v); }
    friend __device__ __forceinline__ Strict operator*(Strict a, Strict b) { return __dmul_rn(a.v, b.v); }
    friend __device__ __forceinline__ Strict operator/(Strict a, Strict b) { return __ddiv_rn(a.v, b.v); }
    __device__ __forceinline__ Strict operator-() const { return -v; }
};

template <typename T>
__device__ __forceinline__ bool finite_val(T x);
template <>
__device__ __forceinline__ bool finite_val<float>(float x)
{
    return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u;
}
template <>
__device__ __forceinline__ bool finite_val<double>(double x)
{
    return ((unsigned)__double2hiint(x) & 0x7ff00000u) != 0x7ff00000u;
}

}  // namespace sksb
