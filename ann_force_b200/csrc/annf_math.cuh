// fp64 tanh with a short dependency chain (host + device; the host build exists so the CPU test-suite can pin it).
//
// The single-replica kernel is a chain of dependent steps, and libdevice's tanh(double) is ~80 dependent fp64
// instructions (~765 cycles measured on B200, 20 % of a whole 21-40-2 evaluation).  This version keeps full double
// accuracy (<= 4 ulp, checked against long double in tests/test_oracle.py::test_device_tanh_matches_long_double) with a
// dependency depth of ~20: tanh(x) = -e / (e + 2) with e = expm1(-2|x|), expm1 by 2^k-scaled range reduction and a degree-13
// Taylor polynomial evaluated with Estrin's scheme.  Computing through expm1 (not exp) keeps tiny arguments exact.
#pragma once

#if defined(__CUDACC__)
#define ANNF_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#define ANNF_HD inline
#endif

namespace annf {

ANNF_HD double annf_pow2i(int k) {                 // 2^k for -1022 <= k <= 1023
#if defined(__CUDA_ARCH__)
    return __hiloint2double((k + 1023) << 20, 0);
#else
    const uint64_t bits = (uint64_t) (k + 1023) << 52;
    double d;
    memcpy(&d, &bits, sizeof(d));
    return d;
#endif
}

ANNF_HD double annf_tanh(double x) {
    const double ax = fabs(x);
    const double y = -2.0 * (ax < 20.0 ? ax : 20.0);                 // tanh(20) rounds to 1; also keeps k in range
    // k = round(y / ln 2), r = y - k ln 2 (two-part ln 2), |r| <= 0.3466
    const double shifter = 6755399441055744.0;                       // 1.5 * 2^52
    const double t = fma(y, 1.4426950408889634074, shifter);
    const double kd = t - shifter;
#if defined(__CUDA_ARCH__)
    const int k = __double2loint(t);
#else
    const int k = (int) kd;
#endif
    double r = fma(kd, -6.93147180369123816490e-01, y);
    r = fma(kd, -1.90821492927058770002e-10, r);
    // expm1(r) = r + r^2 * s(r),  s(r) = sum_{i=0}^{11} r^i / (i + 2)!
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = fma(1.66666666666666666667e-01, r, 5.0e-01);
    const double a1 = fma(8.33333333333333333333e-03, r, 4.16666666666666666667e-02);
    const double a2 = fma(1.98412698412698412698e-04, r, 1.38888888888888888889e-03);
    const double a3 = fma(2.75573192239858906526e-06, r, 2.48015873015873015873e-05);
    const double a4 = fma(2.50521083854417187751e-08, r, 2.75573192239858906526e-07);
    const double a5 = fma(1.60590438368216145994e-10, r, 2.08767569878680989792e-09);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    double s = fma(b1, r4, b0);
    s = fma(b2, r8, s);
    const double p = fma(r2, s, r);
    const double two_k = annf_pow2i(k);
    const double e = fma(two_k, p, two_k - 1.0);                     // expm1(-2|x|), in (-1, 0]
    const double v = -e / (e + 2.0);
    return x < 0.0 ? -v : v;
}

}  // namespace annf
