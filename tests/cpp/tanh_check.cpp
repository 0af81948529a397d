// Pins ann_force_b200/csrc/annf_math.cuh::annf_tanh (the fp64 tanh of the low-latency single-replica kernel) against
// long double tanhl on the host: dense sweeps over the ranges a pre-activation can take, plus the awkward points.
// Prints the worst error in ulps of the exact value and exits non-zero above 4 ulp.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "annf_math.cuh"

static double ulp_of(double v) { return std::nextafter(std::fabs(v), INFINITY) - std::fabs(v); }

int main() {
    double worst = 0.0, worst_x = 0.0;
    auto check = [&](double x) {
        const double got = annf::annf_tanh(x);
        const long double want = tanhl((long double) x);
        const double err = (double) fabsl((long double) got - want) / ulp_of((double) want);
        if (err > worst) { worst = err; worst_x = x; }
        if (x != 0.0 && (got < 0) != (x < 0)) { worst = 1e9; worst_x = x; }
    };
    unsigned long long s = 12345;
    auto uni = [&]() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double) (s >> 11) * (1.0 / 9007199254740992.0); };
    for (int i = 0; i < 2000000; i++) check((uni() * 2 - 1) * 0.5);       // polynomial range
    for (int i = 0; i < 2000000; i++) check((uni() * 2 - 1) * 4.0);       // typical pre-activations
    for (int i = 0; i < 1000000; i++) check((uni() * 2 - 1) * 45.0);      // saturation
    for (int e = -300; e <= 5; e++) { check(std::ldexp(1.0, e)); check(-std::ldexp(1.2345, e)); }   // tiny arguments stay exact
    const double special[] = {0.0, -0.0, 0.17328679513998632, 0.1732867951399864, 0.34657359027997264, 19.999999, 20.0, 20.000001, 1e300, -1e300};
    for (double x : special) check(x);
    if (annf::annf_tanh(0.0) != 0.0 || annf::annf_tanh(1e300) != 1.0 || annf::annf_tanh(-1e300) != -1.0) worst = 1e9;
    printf("{\"worst_ulp\": %.3f, \"at\": %.17g}\n", worst, worst_x);
    return worst <= 4.0 ? 0 : 1;
}
