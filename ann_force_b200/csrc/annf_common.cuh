// Shared host/device definitions of libannforce_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "annforce_b200.h"
#include "annf_math.cuh"

namespace annf {

constexpr int kMaxLayers = ANNF_MAX_LAYERS;
constexpr double kTwoPiRef = 6.2832;   // the reference's 2*pi (ANN_ReferenceKernels.h:106-107, .cpp:282-283)

// Device view of one network.  All pointers are device pointers owned by the handle.
struct NetView {
    int L;                          // node layers
    int nodes[kMaxLayers];
    int types[kMaxLayers];          // types[l] = activation producing layer l+1
    int node_off[kMaxLayers + 1];   // prefix sums of nodes (activation offsets)
    long long w_off[kMaxLayers];    // offset of connection l in the flat weight arrays
    int b_off[kMaxLayers];          // offset of connection l in the flat bias arrays
    const double* W64;              // [n_out][n_in] per connection   (backward: lanes over n_in)
    const double* Wt64;             // [n_in][n_out] per connection   (forward:  lanes over n_out)
    const float* W32;
    const float* Wt32;
    const double* b64;
    const float* b32;
    const double* center;           // [num_center] device copy (annf_update_parameters rewrites it)
    const double* k_dev;            // force constant on the device (same reason)
    int num_center;
    double scale;                   // scaling_factor
    int input_type;
    int n_sel;                      // atoms this force reads: the Cartesian selection, or the union of the atoms of all
                                    // pairs / dihedrals ("feature atoms")
    const int* sel_atoms;           // 0-based original atom index per feature atom
    const int* sel_slots;           // position-buffer slot per feature atom (OpenMM reordering)
    int n_pairs;
    const int* pairs;               // [2*n_pairs] indices INTO the feature-atom list
    int n_dihedrals;
    const int* dihedrals;           // [4*n_dihedrals] indices INTO the feature-atom list
};

// Low-latency single-replica kernel (annf_single.cu, annf_single_fast_kernel): what the launcher needs to know; the work
// split itself travels inside the per-CTA parameter blobs.
struct FastPlan {
    int cl;                         // CTAs in the cluster (1 = plain launch)
    int blob_doubles;               // size of one CTA's parameter blob, in doubles
    int smem_doubles;               // dynamic shared memory of one CTA, in doubles
};

constexpr int kStreamMaxPass = 2 * (kMaxLayers - 1);

// Host-built description of the weight stream and of the work split of every pass (= one contraction) of the streamed
// batched kernel (annf_stream.cu).  pass p < L-1 is forward connection l = p; pass p >= L-1 is backward connection
// l = 2(L-1)-1-p.  A copy lives in device memory and is bulk-copied into shared memory at kernel start, so that the kernel's
// constant-bank argument block stays small (every first touch of a constant line costs a serialised miss).
struct StreamPlan {
    int L, total, npass, scratch_unit;   // scratch_unit: K-split scratch in elements per replica of the tile
    int nodes[kMaxLayers];
    int types[kMaxLayers];
    int node_off[kMaxLayers + 1];
    int n_lane[kStreamMaxPass];      // outputs of the pass (lane dimension)
    int n_loop[kStreamMaxPass];      // contraction length (rows of the streamed matrix)
    int ld[kStreamMaxPass];          // row stride in elements = groups * nc * 32 (zero padded: lanes never leave the row)
    int rows_per[kStreamMaxPass];    // rows per chunk
    int nchunks[kStreamMaxPass];
    int nc[kStreamMaxPass], groups[kStreamMaxPass], ks[kStreamMaxPass];
    int pad[3];
    long long base[kStreamMaxPass];  // element offset of the pass in the stream
};
static_assert(sizeof(StreamPlan) % 16 == 0, "StreamPlan is bulk-copied: its size must be a multiple of 16 bytes");

template <typename T> struct WeightsOf;
template <> struct WeightsOf<double> {
    __device__ static const double* W(const NetView& n) { return n.W64; }
    __device__ static const double* Wt(const NetView& n) { return n.Wt64; }
    __device__ static const double* b(const NetView& n) { return n.b64; }
};
template <> struct WeightsOf<float> {
    __device__ static const float* W(const NetView& n) { return n.W32; }
    __device__ static const float* Wt(const NetView& n) { return n.Wt32; }
    __device__ static const float* b(const NetView& n) { return n.b32; }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double act_tanh(double x) { return annf_tanh(x); }   // annf_math.cuh: <= 4 ulp, ~20-deep chain
__device__ __forceinline__ float act_tanh(float x) { return tanhf(x); }

// ---------------------------------------------------------------------------------------------
// Output head in fp64, one replica, executed by ONE thread (n_last is a handful of CVs).
//   in : z[n]   pre-activations of the output layer (stride zs)
//   out: dz[n]  dE/dz of the output layer (stride ds); returns the umbrella energy.
// Follows update_and_get_potential_energy (ANN_ReferenceKernels.h:89-119) and the output-layer part
// of back_prop (ANN_ReferenceKernels.cpp:265-296) composed with that layer's own Jacobian
// (:307-318 Circular, :343-348 Tanh, :349-354 Linear, :355-365 Softmax).
// ---------------------------------------------------------------------------------------------
__device__ inline double output_head(int type, int n, const double* z, int zs, const double* center, double k,
                                     double* dz, int ds) {
    double e = 0.0;
    if (type == ANNF_CIRCULAR) {
        for (int j = 0; j < n / 2; j++) {
            const double p = z[(2 * j) * zs], q = z[(2 * j + 1) * zs];
            // one reciprocal instead of five divisions (each ~300 cycles of dependent fp64 latency in the single-replica
            // kernel): x / r as x * (1/r) is within 1 ulp of the reference's quotient
            const double inv_r = rsqrt(p * p + q * q);                // one MUFU + Newton steps instead of sqrt + divide: <= 1 ulp apart
            const double c = p * inv_r, s = q * inv_r;               // :213-216
            const int sign = s > 0 ? 1 : -1;                         // :279, .h:102
            const double angle = acos(c) * sign;
            const double d1 = angle - center[j], d2 = d1 + kTwoPiRef, d3 = d1 - kTwoPiRef;
            double dist = d1;                                         // :285-291
            if (fabs(d2) < fabs(dist)) dist = d2;
            if (fabs(d3) < fabs(dist)) dist = d3;
            e += 0.5 * k * fmin(fmin(d1 * d1, d2 * d2), d3 * d3);     // .h:109-114
            const double d_cos = k * dist * (-1) * rsqrt(1 - c * c) * sign;  // :292-293 ; d_sin = 0 (:294)
            // Circular Jacobian (:309-317) applied to (d_cos, 0): q / r^3 * (q d_cos), p / r^3 * (-q d_cos)
            const double t = q * d_cos * (inv_r * inv_r * inv_r);
            dz[(2 * j) * ds] = q * t;
            dz[(2 * j + 1) * ds] = -p * t;
        }
    } else if (type == ANNF_SOFTMAX) {
        double sum = 0.0;                                             // no max-shift, like :226-234
        for (int j = 0; j < n; j++) sum += exp(z[j * zs]);
        double dot = 0.0;
        for (int j = 0; j < n; j++) {
            const double a = exp(z[j * zs]) / sum, d = (a - center[j]) * k;
            e += 0.5 * k * (a - center[j]) * (a - center[j]);
            dot += a * d;
        }
        for (int j = 0; j < n; j++) {                                 // sum_s (delta_js - a_j) a_s d_s  (:355-365)
            const double a = exp(z[j * zs]) / sum, d = (a - center[j]) * k;
            dz[j * ds] = a * d - a * dot;
        }
    } else {
        for (int j = 0; j < n; j++) {
            const double zz = z[j * zs];
            const double a = (type == ANNF_TANH) ? annf_tanh(zz) : zz;
            const double d = (a - center[j]) * k;                      // :266-269
            e += 0.5 * k * (a - center[j]) * (a - center[j]);          // .h:93-96
            dz[j * ds] = (type == ANNF_TANH) ? d * (1 - a * a) : d;    // :343-354
        }
    }
    return e;
}

// ---------------------------------------------------------------------------------------------
// Dihedral featuriser (data_type_in_input_layer == 0), fp64.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void v3sub(const double* a, const double* b, double* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
__device__ __forceinline__ double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void v3cross(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
// sign convention of the reference: (sum of components of n1 x n2) * (sum of components of the middle bond) > 0
// (ANN_ReferenceKernels.cpp:456,612)
__device__ __forceinline__ int dihedral_sign(const double* n1xn2, const double* d2) {
    return (n1xn2[0] + n1xn2[1] + n1xn2[2]) * (d2[0] + d2[1] + d2[2]) > 0 ? 1 : -1;
}

// cos / sin of the dihedral p0-p1-p2-p3: get_cos_and_sin_for_four_atoms, ANN_ReferenceKernels.cpp:599-636
// (normals normalised by multiplying with the reciprocal norm, like OpenMM's Vec3::operator/=)
__device__ inline void dihedral_cos_sin(const double* p0, const double* p1, const double* p2, const double* p3, double& c, double& s) {
    double d1[3], d2[3], d3[3], n1[3], n2[3], sv[3];
    v3sub(p1, p0, d1); v3sub(p2, p1, d2); v3sub(p3, p2, d3);
    v3cross(d1, d2, n1); v3cross(d2, d3, n2);
    const double r1 = 1.0 / sqrt(v3dot(n1, n1)), r2 = 1.0 / sqrt(v3dot(n2, n2));
    for (int i = 0; i < 3; i++) { n1[i] *= r1; n2[i] *= r2; }
    c = v3dot(n1, n2);
    v3cross(n1, n2, sv);
    s = sqrt(v3dot(sv, sv)) * dihedral_sign(sv, d2);
}

// Force contributions of one dihedral from dE/dcos, dE/dsin: the chain rule of
// get_force_from_derivative_of_first_layer (ANN_ReferenceKernels.cpp:425-577) in vector form:
//   cos = n1.n2/(|n1||n2|), sin = sign |n1 x n2|/(|n1||n2|), n1 = d1 x d2, n2 = d2 x d3;
//   dE/dd1 = d2 x g1, dE/dd2 = g1 x d1 + d3 x g2, dE/dd3 = g2 x d2  with g1 = dE/dn1, g2 = dE/dn2;
//   force[atom k] += dE/dd_k, force[atom k+1] -= dE/dd_k (:546-553).   f[4][3] receives the four contributions.
__device__ inline void dihedral_forces(const double* p0, const double* p1, const double* p2, const double* p3,
                                       double g_cos, double g_sin, double (&f)[4][3]) {
    double d1[3], d2[3], d3[3], n1[3], n2[3], c[3];
    v3sub(p1, p0, d1); v3sub(p2, p1, d2); v3sub(p3, p2, d3);
    v3cross(d1, d2, n1); v3cross(d2, d3, n2); v3cross(n1, n2, c);
    const int sign = dihedral_sign(c, d2);
    const double a = v3dot(n1, n1), b = v3dot(n2, n2), n1n2 = v3dot(n1, n2), cc = v3dot(c, c);
    const double la = sqrt(a), lb = sqrt(b), lc = sqrt(cc);
    double n2xc[3], cxn1[3], g1[3], g2[3];
    v3cross(n2, c, n2xc); v3cross(c, n1, cxn1);
    for (int i = 0; i < 3; i++) {
        const double dcos_dn1 = n2[i] / (la * lb) - n1[i] * n1n2 / (a * la * lb);
        const double dcos_dn2 = n1[i] / (la * lb) - n2[i] * n1n2 / (b * la * lb);
        const double dsin_dn1 = sign * (n2xc[i] / (lc * la * lb) - n1[i] * lc / (a * la * lb));
        const double dsin_dn2 = sign * (cxn1[i] / (lc * la * lb) - n2[i] * lc / (b * la * lb));
        g1[i] = g_cos * dcos_dn1 + g_sin * dsin_dn1;
        g2[i] = g_cos * dcos_dn2 + g_sin * dsin_dn2;
    }
    double e1[3], e2a[3], e2b[3], e3[3];
    v3cross(d2, g1, e1);
    v3cross(g1, d1, e2a); v3cross(d3, g2, e2b);
    v3cross(g2, d2, e3);
    for (int i = 0; i < 3; i++) {
        const double e2 = e2a[i] + e2b[i];
        f[0][i] = e1[i];
        f[1][i] = e2 - e1[i];
        f[2][i] = e3[i] - e2;
        f[3][i] = -e3[i];
    }
}

// The same head, spread over the threads of a CTA for the element-wise output types: thread t < n handles
// output t (Tanh / Linear) or pair t < n/2 (Circular) and returns ITS energy term; Softmax stays on thread 0.
// Callers sum the returned terms (fixed order) to get the energy.
__device__ inline double output_head_parallel(int type, int n, const double* z, int zs, const double* center, double k,
                                              double* dz, int ds, int t) {
    if (type == ANNF_SOFTMAX) return t == 0 ? output_head(type, n, z, zs, center, k, dz, ds) : 0.0;
    if (type == ANNF_CIRCULAR) {
        if (t >= n / 2) return 0.0;
        return output_head(type, 2, z + (size_t) 2 * t * zs, zs, center + t, k, dz + (size_t) 2 * t * ds, ds);
    }
    if (t >= n) return 0.0;
    return output_head(type, 1, z + (size_t) t * zs, zs, center + t, k, dz + (size_t) t * ds, ds);
}

}  // namespace annf
