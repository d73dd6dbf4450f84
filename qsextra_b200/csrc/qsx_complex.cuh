// qsx_complex.cuh -- complex helpers shared by the ahead-of-time and the run-time compiled kernels (double2 = re, im)
#pragma once
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmul_conj(double2 a, double2 b) {   // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
