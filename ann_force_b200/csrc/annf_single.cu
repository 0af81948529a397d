// Single-replica fused kernel on OpenMM's device buffers (sm_100a).
//
// One launch = one ANN_Force evaluation: gather posq -> scale / remove centroid -> MLP forward ->
// umbrella energy -> analytic backward -> centring Jacobian -> atomicAdd into the fixed-point force
// buffer.  It replaces the 60-"bond" generated snippet the legacy plugin handed to
// CudaBondedUtilities (/root/reference/platforms/cuda/src/CudaANN_Kernels.cpp:180-359) and follows
// the arithmetic of ReferenceCalcANN_ForceKernel::candidate_2
// (/root/reference/platforms/reference/ANN_ReferenceKernels.cpp:122-171) in fp64.
//
// Parallel decomposition: ONE thread-block cluster of `CL` CTAs (1 for small nets, up to 16 for wide
// layers).  Every CTA keeps the full activation vectors in its own shared memory; each computes a
// contiguous slice of a layer's rows (forward) or columns (backward) and publishes the slice into
// every peer's shared memory through DSMEM stores, followed by one cluster barrier.  Within a CTA a
// forward row is a warp-wide dot product reduced with shuffles; a backward column is one thread
// walking the (coalesced) weight rows.
// Small networks (one CTA): ALL weights and biases are pulled into shared memory by one cp.async burst
// issued before the position gather, so the whole evaluation pays one global-memory round trip instead
// of one per dependent loop (measured: the L2 round trips were ~70% of the kernel).  Wide networks (cluster):
// weights stream from L2 (each is used once per direction).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "annf_common.cuh"
#include "annf_math.cuh"

namespace cg = cooperative_groups;

namespace annf {

// Diagnostic: clock64() of thread 0 / CTA 0 at the phase boundaries of the last launch (annf_debug_single_phases).
__device__ long long g_single_phase_clock[16];
#ifdef ANNF_PHASE_CLOCKS        // diagnostic build only (make EXTRA=-DANNF_PHASE_CLOCKS): the product kernels carry no instrumentation
#define ANNF_PHASE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_single_phase_clock[i] = clock64(); } while (0)
#else
#define ANNF_PHASE(i) do { } while (0)
#endif

constexpr int kSingleThreads = 256;
constexpr int kSingleWarps = kSingleThreads / 32;

struct SingleArgs {
    NetView net;
    const void* posq;
    const void* posq_corr;
    unsigned long long* force;
    int padded;
    void* energy;
    unsigned flags;
    int red_doubles;            // size of the reduction / scratch region in doubles
};

// Position of one atom slot in nm, fp64, honouring single / mixed (+correction) / double layouts.
__device__ __forceinline__ double3 load_pos(const void* posq, const void* posq_corr, unsigned flags, int slot) {
    if (flags & ANNF_FLAG_POSQ_DOUBLE) {
        const double4 p = reinterpret_cast<const double4*>(posq)[slot];
        return make_double3(p.x, p.y, p.z);
    }
    const float4 p = __ldg(reinterpret_cast<const float4*>(posq) + slot);
    double3 r = make_double3((double) p.x, (double) p.y, (double) p.z);
    if (posq_corr != nullptr) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(posq_corr) + slot);
        r.x += (double) c.x; r.y += (double) c.y; r.z += (double) c.z;
    }
    return r;
}
__device__ __forceinline__ double3 load_pos(const SingleArgs& a, int slot) { return load_pos(a.posq, a.posq_corr, a.flags, slot); }

// Write v into the same shared-memory location of every CTA of the cluster.
__device__ __forceinline__ void publish(cg::cluster_group& cluster, int cl, double* local_ptr, double v, int lane) {
    if (cl == 1) {
        if (lane == 0) *local_ptr = v;
    } else if (lane < cl) {
        *cluster.map_shared_rank(local_ptr, lane) = v;
    }
}

__device__ __forceinline__ void cluster_or_block_sync(cg::cluster_group& cluster, int cl) {
    if (cl == 1) __syncthreads();
    else cluster.sync();
}

// Lanes per output of a small mat-vec: the largest power of two (<= 32) such that all outputs fit in one pass of the CTA
// and every lane still has a few terms to add.  Few lanes per output = few shuffle rounds, which is what bounds these
// latency-critical contractions (a 40 x 21 layer took 3,800 cycles with one warp per row and 5 x 2 shuffles per sum).
__device__ __forceinline__ int lanes_per_output(int n_outputs, int n_terms) {
    int g = 32;
    while (g > 1 && (g * n_outputs > kSingleThreads || 4 * g > n_terms)) g >>= 1;
    return g;
}

// In-place activation of a full layer held in shared memory (every CTA does this redundantly).
//   z -> a ; types per ANN_ReferenceKernels.cpp:200-234
__device__ void activate_layer(int type, int n, const double* z, double* a) {
    const int tid = threadIdx.x;
    if (type == ANNF_LINEAR) {
        for (int j = tid; j < n; j += blockDim.x) a[j] = z[j];
    } else if (type == ANNF_TANH) {
        for (int j = tid; j < n; j += blockDim.x) a[j] = annf_tanh(z[j]);
    } else if (type == ANNF_CIRCULAR) {
        for (int j = tid; j < n / 2; j += blockDim.x) {
            const double p = z[2 * j], q = z[2 * j + 1], r = sqrt(p * p + q * q);
            a[2 * j] = p / r;
            a[2 * j + 1] = q / r;
        }
    } else {
        for (int j = tid; j < n; j += blockDim.x) {
            double sum = 0.0;
            for (int s = 0; s < n; s++) sum += exp(z[s]);
            a[j] = exp(z[j]) / sum;
        }
    }
}

// dE/da -> dE/dz of a hidden layer, full vector in shared memory, in place in d[].
// (ANN_ReferenceKernels.cpp:307-318 Circular, :343-348 Tanh, :349-354 Linear, :355-365 Softmax)
__device__ void activation_backward(int type, int n, const double* z, const double* a, double* d, double* scratch) {
    const int tid = threadIdx.x;
    if (type == ANNF_TANH) {
        for (int j = tid; j < n; j += blockDim.x) d[j] *= (1 - a[j] * a[j]);
    } else if (type == ANNF_CIRCULAR) {
        for (int j = tid; j < n / 2; j += blockDim.x) {
            const double p = z[2 * j], q = z[2 * j + 1], r = sqrt(p * p + q * q);
            const double dp = d[2 * j], dq = d[2 * j + 1];
            d[2 * j] = q / (r * r * r) * (q * dp - p * dq);
            d[2 * j + 1] = p / (r * r * r) * (p * dq - q * dp);
        }
    } else if (type == ANNF_SOFTMAX) {
        for (int j = tid; j < n; j += blockDim.x) {
            double dot = 0.0;
            for (int s = 0; s < n; s++) dot += a[s] * d[s];
            scratch[j] = a[j] * d[j] - a[j] * dot;
        }
        __syncthreads();
        for (int j = tid; j < n; j += blockDim.x) d[j] = scratch[j];
    }
}

__global__ void __launch_bounds__(kSingleThreads, 1) annf_single_kernel(const SingleArgs args) {
    cg::cluster_group cluster = cg::this_cluster();
    const int cl = (int) cluster.num_blocks();
    const int rank = (int) cluster.block_rank();
    const NetView& net = args.net;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = net.L, total = net.node_off[L];

    ANNF_PHASE(0);
    extern __shared__ __align__(16) double smem[];
    double* a = smem;                 // activations  a[node_off[l] + j]
    double* z = a + total;            // pre-activations
    double* d = z + total;            // dE/dz (hidden/output layers), dE/dx (layer 0)
    double* red = d + total;          // [8] small reductions / broadcasts, then scratch
    __shared__ double s_energy;

    // every CTA of the cluster must be resident before anyone stores into a peer's shared memory
    if (cl > 1) cluster.sync();

    // umbrella centre and force constant: fetched now, needed only after the forward pass
    __shared__ double s_k;
    double* cen_s = red + args.red_doubles + ((3 * total + args.red_doubles) & 1);
    for (int i = tid; i < net.num_center; i += kSingleThreads) cen_s[i] = net.center[i];
    if (tid == 0) s_k = *net.k_dev;
    // pair / dihedral inputs: feature-atom positions and a per-atom force accumulator
    const int n_fpos = net.input_type == ANNF_INPUT_CARTESIAN ? 0 : 3 * net.n_sel;
    double* fpos = cen_s + net.num_center + (net.num_center & 1);
    double* facc = fpos + n_fpos;

    // ---- one CTA: stage every weight and bias in shared memory (asynchronously) ---------------------
    const double* Wbase = net.W64;
    const double* bbase = net.b64;
    if (cl == 1) {
        const int nw = (int) net.w_off[L - 2] + net.nodes[L - 2] * net.nodes[L - 1];
        const int nb = net.b_off[L - 2] + net.nodes[L - 1];
        double* wsm = facc + n_fpos;                                                    // 2 * n_fpos is even: stays 16-byte aligned
        double* bsm = wsm + nw + (nw & 1);
        for (int i = 2 * tid; i + 1 < nw; i += 2 * kSingleThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(wsm + i)), "l"(net.W64 + i) : "memory");
        for (int i = 2 * tid; i + 1 < nb; i += 2 * kSingleThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(bsm + i)), "l"(net.b64 + i) : "memory");
        if (tid == 0 && (nw & 1)) wsm[nw - 1] = net.W64[nw - 1];
        if (tid == 0 && (nb & 1)) bsm[nb - 1] = net.b64[nb - 1];
        asm volatile("cp.async.commit_group;" ::: "memory");
        Wbase = wsm;
        bbase = bsm;
    }

    // ---- feature stage: ANN_ReferenceKernels.cpp:128-151 (Cartesian) -------------------------
    // every CTA builds the full, centred input redundantly (it is 3*A values)
    const int n0 = net.nodes[0];
    if (net.input_type == ANNF_INPUT_CARTESIAN) {
        const int A = net.n_sel;
        const double inv_scale = 1.0 / net.scale;           // x / s as x * (1/s): 1 ulp from the reference's division
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int i = tid; i < A; i += kSingleThreads) {
            const double3 p = load_pos(args, net.sel_slots[i]);
            const double x = p.x * inv_scale, y = p.y * inv_scale, w = p.z * inv_scale;   // :139
            a[3 * i] = x; a[3 * i + 1] = y; a[3 * i + 2] = w;
            sx += x; sy += y; sz += w;
        }
        if (A <= 32) {
            // the whole selection sits in warp 0: three interleaved shuffle reductions, no block-level hop
            if (warp == 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sx += __shfl_xor_sync(0xffffffffu, sx, o);
                    sy += __shfl_xor_sync(0xffffffffu, sy, o);
                    sz += __shfl_xor_sync(0xffffffffu, sz, o);
                }
                if (lane == 0) { red[0] = sx / A; red[1] = sy / A; red[2] = sz / A; }   // unweighted centroid :137-141
            }
        } else {
            sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
            if (lane == 0) { red[8 + 3 * warp] = sx; red[8 + 3 * warp + 1] = sy; red[8 + 3 * warp + 2] = sz; }
            __syncthreads();
            if (tid < 3) {
                double t = 0.0;
                for (int w = 0; w < kSingleWarps; w++) t += red[8 + 3 * w + tid];
                red[tid] = t / A;
            }
        }
        __syncthreads();
        for (int i = tid; i < n0; i += kSingleThreads) a[i] -= red[i % 3];                // :143-145
    } else {
        // pair distances (:152-160) or dihedral cos / sin (:125-127, :579-636): stage the feature atoms' positions
        double* pos_s = fpos;
        for (int i = tid; i < net.n_sel; i += kSingleThreads) {
            const double3 p = load_pos(args, net.sel_slots[i]);
            pos_s[3 * i] = p.x; pos_s[3 * i + 1] = p.y; pos_s[3 * i + 2] = p.z;
            facc[3 * i] = 0.0; facc[3 * i + 1] = 0.0; facc[3 * i + 2] = 0.0;
        }
        __syncthreads();
        if (net.input_type == ANNF_INPUT_PAIR_DISTANCES) {
            for (int p = tid; p < net.n_pairs; p += kSingleThreads) {
                const double* pa = pos_s + 3 * net.pairs[2 * p];
                const double* pb = pos_s + 3 * net.pairs[2 * p + 1];
                const double inv = 1.0 / net.scale;                                       // Vec3 / scalar multiplies by the reciprocal
                const double dx = pa[0] * inv - pb[0] * inv, dy = pa[1] * inv - pb[1] * inv, dz = pa[2] * inv - pb[2] * inv;
                a[p] = sqrt(dx * dx + dy * dy + dz * dz);
            }
        } else {
            for (int q = tid; q < net.n_dihedrals; q += kSingleThreads) {
                const int* id = net.dihedrals + 4 * q;
                dihedral_cos_sin(pos_s + 3 * id[0], pos_s + 3 * id[1], pos_s + 3 * id[2], pos_s + 3 * id[3], a[2 * q], a[2 * q + 1]);
            }
        }
    }
    ANNF_PHASE(1);                                         // features done (this thread)
    if (cl == 1) asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    ANNF_PHASE(2);                                         // weights staged, everyone here

    // ---- forward: calculate_output_of_each_layer, ANN_ReferenceKernels.cpp:179-260 -------------
    for (int l = 0; l + 1 < L; l++) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        const double* W = Wbase + net.w_off[l];
        const double* b = bbase + net.b_off[l];
        const double* ain = a + net.node_off[l];
        double* zout = z + net.node_off[l + 1];
        const int per = (n_out + cl - 1) / cl;
        const int j0 = rank * per, j1 = min(n_out, j0 + per);
        // each output row is shared by G lanes (power of two, inside one warp); all rows of the slice go in as few passes as possible
        const int G = lanes_per_output(j1 - j0, n_in);
        const int rows_per_pass = kSingleThreads / G, g = tid & (G - 1), slot = tid / G;
        const int passes = (max(j1 - j0, 0) + rows_per_pass - 1) / rows_per_pass;
        for (int pass = 0; pass < passes; pass++) {
            const int j = j0 + pass * rows_per_pass + slot;
            const bool valid = j < j1;
            double acc0 = 0.0, acc1 = 0.0;
            if (valid) {
                const double* wrow = W + (size_t) j * n_in;
                int i = g;
                for (; i + G < n_in; i += 2 * G) {
                    acc0 = fma(wrow[i], ain[i], acc0);
                    acc1 = fma(wrow[i + G], ain[i + G], acc1);
                }
                if (i < n_in) acc0 = fma(wrow[i], ain[i], acc0);
            }
            double sum = acc0 + acc1;
            for (int o = G >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (valid && g == 0) {                                                         // :193-198
                const double v = sum + b[j];
                if (cl == 1) zout[j] = v;
                else for (int r = 0; r < cl; r++) *cluster.map_shared_rank(zout + j, r) = v;
            }
        }
        cluster_or_block_sync(cluster, cl);
        if (l < 2) ANNF_PHASE(7 + 2 * l);                  // diagnostic: connection l contracted
        if (l + 2 < L) {   // hidden layer: activate the whole vector locally
            activate_layer(net.types[l], n_out, zout, a + net.node_off[l + 1]);
            __syncthreads();
        }
        if (l < 2) ANNF_PHASE(8 + 2 * l);                  // diagnostic: layer l + 1 activated
    }

    ANNF_PHASE(3);                                         // forward done
    // ---- energy + output-layer gradient (fp64; one output or Circular pair per thread; identical in every CTA) -----
    {
        const int n_last = net.nodes[L - 1];
        double* eterm = red + 8;
        for (int t = tid; t < n_last; t += kSingleThreads)
            eterm[t] = output_head_parallel(net.types[L - 2], n_last, z + net.node_off[L - 1], 1, cen_s, s_k,
                                            d + net.node_off[L - 1], 1, t);
        __syncthreads();
        if (tid == 0) {
            double e = 0.0;
            for (int t = 0; t < n_last; t++) e += eterm[t];
            s_energy = e;
        }
    }
    __syncthreads();

    ANNF_PHASE(4);                                         // head done
    // ---- backward: back_prop, ANN_ReferenceKernels.cpp:299-373 ----------------------------------
    for (int l = L - 2; l >= 0; l--) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        const double* W = Wbase + net.w_off[l];
        const double* dz = d + net.node_off[l + 1];
        double* dout = d + net.node_off[l];
        const int per = (n_in + cl - 1) / cl;
        const int m0 = rank * per, m1 = min(n_in, m0 + per);
        // column m of W^T * dz is shared by G lanes, each walking rows k = g, g + G, ... (consecutive columns in neighbouring groups)
        const int G = lanes_per_output(m1 - m0, n_out);
        const int cols_per_pass = kSingleThreads / G, g = tid & (G - 1), slot = tid / G;
        const int passes = (max(m1 - m0, 0) + cols_per_pass - 1) / cols_per_pass;
        for (int pass = 0; pass < passes; pass++) {
            const int m = m0 + pass * cols_per_pass + slot;
            const bool valid = m < m1;
            double acc0 = 0.0, acc1 = 0.0;
            if (valid) {
                int k = g;
                for (; k + G < n_out; k += 2 * G) {
                    acc0 = fma(W[(size_t) k * n_in + m], dz[k], acc0);
                    acc1 = fma(W[(size_t) (k + G) * n_in + m], dz[k + G], acc1);
                }
                if (k < n_out) acc0 = fma(W[(size_t) k * n_in + m], dz[k], acc0);
            }
            double v = acc0 + acc1;
            for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (valid && g == 0) {
                if (cl == 1) dout[m] = v;
                else for (int r = 0; r < cl; r++) *cluster.map_shared_rank(dout + m, r) = v;
            }
        }
        cluster_or_block_sync(cluster, cl);
        if (l >= 1) {
            activation_backward(net.types[l - 1], n_in, z + net.node_off[l], a + net.node_off[l], dout, red + 8);
            __syncthreads();
        }
    }

    ANNF_PHASE(5);                                         // backward done
    // ---- forces: get_all_forces_from_derivative_of_first_layer, ANN_ReferenceKernels.cpp:397-410 --
    if (!(args.flags & ANNF_FLAG_SKIP_FORCES) && net.input_type == ANNF_INPUT_CARTESIAN) {
        const int A = net.n_sel;
        const double* g = d;   // dE/dx, layer 0
        // mean gradient per component = the Jacobian term of centroid removal (:402-406), O(A) not O(A^2)
        if (warp < 3) {
            double t = 0.0;
            for (int i = lane; i < A; i += 32) t += g[3 * i + warp];
            t = warp_sum(t);
            if (lane == 0) red[warp] = t / A;
        }
        __syncthreads();
        const int per = (A + cl - 1) / cl;
        const int i0 = rank * per, i1 = min(A, i0 + per);
        const double scale_to_fixed = -(double) 0x100000000LL / net.scale;             // -(1/s) folded into the fixed-point scale
        for (int t = 3 * i0 + tid; t < 3 * i1; t += kSingleThreads) {
            const int i = t / 3, c = t - 3 * i;
            atomicAdd(&args.force[net.sel_slots[i] + c * args.padded],
                      (unsigned long long) ((long long) ((g[t] - red[c]) * scale_to_fixed)));
        }
    }
    if (!(args.flags & ANNF_FLAG_SKIP_FORCES) && net.input_type != ANNF_INPUT_CARTESIAN && rank == 0) {
        // get_all_forces_from_derivative_of_first_layer: dihedral branch :392-396 (+ :425-577), pair branch :411-421.
        // Contributions are summed per feature atom in shared memory first, then one fixed-point atomicAdd per component.
        const double* g = d;
        const double* pos_s = fpos;
        if (net.input_type == ANNF_INPUT_PAIR_DISTANCES) {
            for (int p = tid; p < net.n_pairs; p += kSingleThreads) {
                const int i0 = net.pairs[2 * p], i1 = net.pairs[2 * p + 1];
                const double dx = pos_s[3 * i0] - pos_s[3 * i1], dy = pos_s[3 * i0 + 1] - pos_s[3 * i1 + 1], dz = pos_s[3 * i0 + 2] - pos_s[3 * i1 + 2];
                const double f = g[p] / (net.scale * sqrt(dx * dx + dy * dy + dz * dz));
                atomicAdd(&facc[3 * i0], -dx * f); atomicAdd(&facc[3 * i0 + 1], -dy * f); atomicAdd(&facc[3 * i0 + 2], -dz * f);
                atomicAdd(&facc[3 * i1], dx * f); atomicAdd(&facc[3 * i1 + 1], dy * f); atomicAdd(&facc[3 * i1 + 2], dz * f);
            }
        } else {
            for (int q = tid; q < net.n_dihedrals; q += kSingleThreads) {
                const int* id = net.dihedrals + 4 * q;
                double f[4][3];
                dihedral_forces(pos_s + 3 * id[0], pos_s + 3 * id[1], pos_s + 3 * id[2], pos_s + 3 * id[3], g[2 * q], g[2 * q + 1], f);
                for (int k = 0; k < 4; k++)
                    for (int c = 0; c < 3; c++) atomicAdd(&facc[3 * id[k] + c], f[k][c]);
            }
        }
        __syncthreads();
        for (int t = tid; t < 3 * net.n_sel; t += kSingleThreads) {
            const int i = t / 3, c = t - 3 * i;
            atomicAdd(&args.force[net.sel_slots[i] + c * args.padded],
                      (unsigned long long) ((long long) (facc[t] * (double) 0x100000000LL)));
        }
    }
    ANNF_PHASE(6);                                         // forces issued
    if (rank == 0 && tid == 0 && args.energy != nullptr && !(args.flags & ANNF_FLAG_SKIP_ENERGY)) {
        // fire-and-forget reduction instead of a load + store round trip at the tail of the kernel
        if (args.flags & ANNF_FLAG_ENERGY_DOUBLE) atomicAdd(reinterpret_cast<double*>(args.energy), s_energy);
        else atomicAdd(reinterpret_cast<float*>(args.energy), (float) s_energy);
    }
}

static int single_red_doubles(const NetView& net) {
    int max_n = 0;
    for (int l = 0; l < net.L; l++) max_n = net.nodes[l] > max_n ? net.nodes[l] : max_n;
    return 8 + (max_n > 3 * kSingleWarps ? max_n : 3 * kSingleWarps);
}

cudaError_t read_single_phase_clocks(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_single_phase_clock, sizeof(long long) * 16);
}

size_t single_smem_bytes(const NetView& net, int cluster_size) {
    const int total = net.node_off[net.L];
    size_t doubles = 3 * (size_t) total + single_red_doubles(net) + 1 + net.num_center + 1;
    if (net.input_type != ANNF_INPUT_CARTESIAN) doubles += 6 * (size_t) net.n_sel;      // feature-atom positions + force accumulator
    if (cluster_size == 1) {   // all weights + biases staged in shared memory
        const int L = net.L;
        doubles += (size_t) net.w_off[L - 2] + (size_t) net.nodes[L - 2] * net.nodes[L - 1] + net.b_off[L - 2] + net.nodes[L - 1] + 4;
    }
    return sizeof(double) * doubles;
}

cudaError_t launch_single(const NetView& net, int cluster_size, const void* posq, const void* posq_corr,
                          unsigned long long* force, int padded, void* energy, unsigned flags, cudaStream_t stream) {
    SingleArgs args;
    args.net = net;
    args.posq = posq;
    args.posq_corr = posq_corr;
    args.force = force;
    args.padded = padded;
    args.energy = energy;
    args.flags = flags;
    args.red_doubles = single_red_doubles(net);
    if (cluster_size == 1) {   // plain launch: no cluster attribute on the latency-critical small-network path
        annf_single_kernel<<<1, kSingleThreads, single_smem_bytes(net, 1), stream>>>(args);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    cfg.blockDim = dim3(kSingleThreads);
    cfg.dynamicSmemBytes = single_smem_bytes(net, cluster_size);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, annf_single_kernel, args);
}

cudaError_t prepare_single(const NetView& net, int cluster_size) {
    // function attributes are per device, not per handle: always ask for the architectural maximum so
    // that handles of different sizes can coexist
    (void) net;
    cudaError_t err = cudaFuncSetAttribute(annf_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (err != cudaSuccess) return err;
    if (cluster_size > 8)
        err = cudaFuncSetAttribute(annf_single_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return err;
}

// =====================================================================================================================
// Low-latency variant for the common case: Cartesian input, Tanh / Linear hidden layers, any output layer.
//
// One evaluation of a small network is a chain of dependent steps (C2: 4 kFLOP), so the kernel is organised around the
// number of serialised round trips, not around throughput:
//   * ONE global-memory round trip for parameters: every CTA pulls exactly the weight rows it will use — its row slice of
//     every W_l for the forward pass, its row slice of every W_l^T for the backward pass, biases, centre, k — into shared
//     memory with cp.async before it touches the positions; the copies land while the positions are gathered;
//   * ONE barrier per layer: the lane that finishes a row's dot product applies bias + activation (forward) or the
//     activation derivative (backward) itself and publishes the result (DSMEM stores when the network spans a cluster);
//     a Tanh / Linear output layer gets its energy term and output gradient in the same epilogue;
//   * the first connection's backward matrix is stored pre-centred and pre-scaled,
//        W0c^T[t][k] = -(1/s) (W_0[k][t] - (1/A) sum_i W_0[k][3i+c(t)]),
//     i.e. the Jacobian of "divide by s, subtract the centroid" (ANN_ReferenceKernels.cpp:137-145, :397-410) folded into
//     the weights, so the last mat-vec yields forces and its epilogue is the fixed-point atomicAdd: no mean, no extra pass;
//   * a dot product is shared by 2^lg lanes chosen on the host per layer (few lanes = few shuffle rounds).
// =====================================================================================================================
// Per-CTA parameter block ("blob"), built on the host at create time and fetched with ONE bulk copy (TMA engine) per
// CTA: a 512-byte integer header (network shape + work split of this CTA) followed by exactly the weight rows the CTA
// uses, already in their shared-memory layout.  Keeping the kernel's argument struct to a few dozen bytes matters as much
// as the single copy: every first touch of a 64-byte line of the constant-bank argument block is a serialised miss
// (measured: ~2,000 cycles of the old prologue went into walking per-layer arrays passed by value).
enum {
    kHdrL = 0, kHdrTotal, kHdrNL, kHdrNodes = 8, kHdrTypes = 16, kHdrNodeOff = 24,
    kHdrFRows = 40, kHdrFLg = 48, kHdrFOff = 56, kHdrFbOff = 64, kHdrFJ0 = 72,
    kHdrBRows = 80, kHdrBLg = 88, kHdrBOff = 96, kHdrBM0 = 104, kHdrInts = 128
};

struct FastArgs {
    const double* blob;         // cluster_size blobs, blob_doubles apart
    int blob_doubles;
    int A, num_center, total, nL;
    const int* sel_slots;
    const double* center;
    const double* k_dev;
    double inv_scale, inv_A;
    const void* posq;
    const void* posq_corr;
    unsigned long long* force;
    int padded;
    unsigned flags;
    void* energy;
    int input_type;             // ANNF_INPUT_*: Cartesian, or pair distances / dihedral cos-sin over the feature atoms
    int n_feat;                 // pairs or dihedrals
    const int* feat;            // [2 * n_feat] or [4 * n_feat] indices INTO the feature-atom list
    double scale;
};

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t) __cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// out[r] = sum_c M[r][c] * v[c] for r < nrows (M rows contiguous in shared memory); 2^lg lanes of one warp share a row and
// the lane with g == 0 receives the sum in epi(r, sum).  Called by every thread of the CTA.
// The terms of a lane are taken 8 at a time with every load issued before the first FMA: the loop is latency-bound, and
// a strided loop with run-time bounds costs ~90 cycles per 4 terms.  (16 at a time measured 4 % slower end to end: this
// code runs a handful of times per launch, so its size counts as much as its schedule.)
#ifndef ANNF_MATVEC_BLOCK
#define ANNF_MATVEC_BLOCK 8
#endif
constexpr int kMatvecBlock = ANNF_MATVEC_BLOCK;           // terms of a lane in flight together

template <typename Epi>
__device__ __forceinline__ void fast_matvec(const double* M, int ncols, int nrows, const double* v, int lg, Epi epi) {
    const int tid = threadIdx.x;
    const int G = 1 << lg, g = tid & (G - 1), slot = tid >> lg, rows_per_pass = kSingleThreads >> lg;
    for (int r0 = 0; r0 < nrows; r0 += rows_per_pass) {        // trip count is uniform over the CTA
        const int r = r0 + slot;
        const bool valid = r < nrows;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        if (valid) {
            const double* row = M + (size_t) r * ncols;
            for (int c0 = g; c0 < ncols; c0 += kMatvecBlock * G) {
                double w[kMatvecBlock], x[kMatvecBlock];
#pragma unroll
                for (int u = 0; u < kMatvecBlock; u++) {
                    const int c = c0 + u * G;
                    const bool in = c < ncols;
                    w[u] = in ? row[c] : 0.0;
                    x[u] = in ? v[c] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kMatvecBlock; u += 4) {
                    acc0 = fma(w[u], x[u], acc0);
                    acc1 = fma(w[u + 1], x[u + 1], acc1);
                    acc2 = fma(w[u + 2], x[u + 2], acc2);
                    acc3 = fma(w[u + 3], x[u + 3], acc3);
                }
            }
        }
        double sum = (acc0 + acc1) + (acc2 + acc3);
        for (int o = G >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (valid && g == 0) epi(r, sum);
    }
}

template <bool kCluster>
__global__ void __launch_bounds__(kSingleThreads, 1) annf_single_fast_kernel(const FastArgs args) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cl = 1, rank = 0;
    if (kCluster) {
        asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cl));
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        // every CTA must be running before anyone stores into a peer's shared memory: arrive now, wait after the prologue
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    }
    // Programmatic dependent launch: the next kernel of the stream may start its own prologue now; this one fetches its
    // (static) parameters first and only then waits for the kernels before it — the integrator that wrote the positions.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    ANNF_PHASE(0);

    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long blob_bar;
    const int total = args.total, nL = args.nL, A = args.A;
    double* blob = smem;                                  // header + this CTA's weight rows
    const int* hdr = reinterpret_cast<const int*>(blob);
    double* a = blob + args.blob_doubles;                 // activations, all layers
    double* d = a + total;                                // dE/dz, all layers
    double* zL = d + total;                               // output pre-activations (non-elementwise heads only)
    double* eterm = zL + nL;                              // energy terms of the outputs
    double* cen = eterm + nL;                             // umbrella centre
    double* kval = cen + args.num_center;                 // force constant
    double* red = kval + 1;                               // [3 * warps] centroid partial sums, [3] centroid
    const bool cartesian = args.input_type == ANNF_INPUT_CARTESIAN;
    double* fpos = red + 3 * kSingleWarps + 4;            // pair / dihedral inputs: positions of the feature atoms [3 * A] ...
    double* facc = fpos + (cartesian ? 0 : 3 * A);        // ... and their force accumulators [3 * A]
    int* slots_s = reinterpret_cast<int*>(facc + (cartesian ? 0 : 3 * A));   // position-buffer slot of every selected atom

    // ---- parameters: one bulk copy + two tiny cp.async, consumed after the positions are in -----------------------
    const uint32_t bar = (uint32_t) __cvta_generic_to_shared(&blob_bar);
    if (tid == kSingleThreads - 32) {                     // first lane of the last warp: warp 0 goes straight to the positions
        const uint32_t bytes = (uint32_t) args.blob_doubles * 8u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t) __cvta_generic_to_shared(blob)),
                       "l"(__cvta_generic_to_global(args.blob + (size_t) rank * args.blob_doubles)), "r"(bytes), "r"(bar) : "memory");
    }
    if (warp == 1) {
        for (int i = lane; i < args.num_center; i += 32) cp_async8(cen + i, args.center + i);
        if (lane == 31) cp_async8(kval, args.k_dev);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // The slot of this thread's first atom is fetched BEFORE the wait: sel_slots is written only by annf_set_atom_order, and the
    // launcher drops the PDL attribute for the first evaluation after a reorder, so under PDL the kernel ahead of this one never
    // is its writer.  It takes one of the two dependent global round trips (slot -> position) off the critical path.
    const int my_slot = tid < A ? __ldg(args.sel_slots + tid) : 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");    // everything below may read what earlier kernels wrote
    ANNF_PHASE(11);

    // ---- feature stage: ANN_ReferenceKernels.cpp:128-151 (every CTA builds the full centred input) ----------------
    if (!cartesian) {
        // pair distances (:152-160) or dihedral cos / sin (:125-127, :579-636) over the staged positions of the feature atoms
        for (int i = tid; i < A; i += kSingleThreads) {
            const int slot = i == tid ? my_slot : args.sel_slots[i];
            slots_s[i] = slot;
            const double3 p = load_pos(args.posq, args.posq_corr, args.flags, slot);
            fpos[3 * i] = p.x; fpos[3 * i + 1] = p.y; fpos[3 * i + 2] = p.z;
            facc[3 * i] = 0.0; facc[3 * i + 1] = 0.0; facc[3 * i + 2] = 0.0;
        }
        __syncthreads();
        if (args.input_type == ANNF_INPUT_PAIR_DISTANCES) {
            for (int p = tid; p < args.n_feat; p += kSingleThreads) {
                const double* pa = fpos + 3 * args.feat[2 * p];
                const double* pb = fpos + 3 * args.feat[2 * p + 1];
                const double inv = args.inv_scale;                                          // Vec3 / scalar multiplies by the reciprocal
                const double dx = pa[0] * inv - pb[0] * inv, dy = pa[1] * inv - pb[1] * inv, dz = pa[2] * inv - pb[2] * inv;
                a[p] = sqrt(dx * dx + dy * dy + dz * dz);
            }
        } else {
            for (int q = tid; q < args.n_feat; q += kSingleThreads) {
                const int* id = args.feat + 4 * q;
                dihedral_cos_sin(fpos + 3 * id[0], fpos + 3 * id[1], fpos + 3 * id[2], fpos + 3 * id[3], a[2 * q], a[2 * q + 1]);
            }
        }
    } else if (A <= 32) {
        // the whole selection sits in warp 0: butterfly sums leave the centroid in every lane, no block-level hop
        if (warp == 0) {
            double x = 0.0, y = 0.0, w = 0.0;
            if (lane < A) {
                const int slot = my_slot;                 // warp 0: tid == lane
                slots_s[lane] = slot;
                const double3 p = load_pos(args.posq, args.posq_corr, args.flags, slot);
                x = p.x * args.inv_scale; y = p.y * args.inv_scale; w = p.z * args.inv_scale;      // :139
            }
            double sx = x, sy = y, sz = w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
            }
            if (lane < A) {                                                                         // :137-145
                a[3 * lane] = x - sx * args.inv_A;
                a[3 * lane + 1] = y - sy * args.inv_A;
                a[3 * lane + 2] = w - sz * args.inv_A;
            }
        }
    } else {
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int i = tid; i < A; i += kSingleThreads) {
            const int slot = i == tid ? my_slot : args.sel_slots[i];
            slots_s[i] = slot;
            const double3 p = load_pos(args.posq, args.posq_corr, args.flags, slot);
            const double x = p.x * args.inv_scale, y = p.y * args.inv_scale, w = p.z * args.inv_scale;
            a[3 * i] = x; a[3 * i + 1] = y; a[3 * i + 2] = w;
            sx += x; sy += y; sz += w;
        }
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
        if (lane == 0) { red[3 * warp] = sx; red[3 * warp + 1] = sy; red[3 * warp + 2] = sz; }
        __syncthreads();
        if (tid < 3) {
            double t = 0.0;
            for (int w = 0; w < kSingleWarps; w++) t += red[3 * w + tid];
            red[3 * kSingleWarps + tid] = t * args.inv_A;
        }
        __syncthreads();
        for (int i = tid; i < 3 * A; i += kSingleThreads) a[i] -= red[3 * kSingleWarps + i % 3];
    }
    ANNF_PHASE(1);
    if (warp == 1) asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                      // also orders the mbarrier init before everyone's wait
    {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0, 0x989680;\n\t"
            "@P1 bra DONE_%=;\n\t"
            "bra WAIT_%=;\n\t"
            "DONE_%=:\n\t"
            "}" ::"r"(bar) : "memory");
    }
    if (kCluster) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    ANNF_PHASE(2);

    auto publish = [&](double* local, double v) {
        if (!kCluster) {
            *local = v;
        } else {
            const uint32_t addr = (uint32_t) __cvta_generic_to_shared(local);
            for (int r = 0; r < cl; r++) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(r));
                asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
            }
        }
    };
    auto sync_all = [&]() {
        if (!kCluster) {
            __syncthreads();
        } else {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    };

    // ---- forward: calculate_output_of_each_layer, ANN_ReferenceKernels.cpp:179-260 --------------------------------
    const int L = hdr[kHdrL];
    const double kk = *kval;
    for (int l = 0; l + 1 < L; l++) {
        const int n_in = hdr[kHdrNodes + l];
        const int j0 = hdr[kHdrFJ0 + l], nf = hdr[kHdrFRows + l];
        const double* bias = blob + hdr[kHdrFbOff + l];
        const int type = hdr[kHdrTypes + l];
        const bool last = l + 2 == L;
        double* out = a + hdr[kHdrNodeOff + l + 1];
        double* dzL = d + hdr[kHdrNodeOff + L - 1];
        fast_matvec(blob + hdr[kHdrFOff + l], n_in, nf, a + hdr[kHdrNodeOff + l], hdr[kHdrFLg + l], [&](int r, double sum) {
            const int j = j0 + r;
            const double zz = sum + bias[r];                                                        // :193-198
            if (!last) {
                publish(out + j, type == ANNF_TANH ? annf_tanh(zz) : zz);                           // :200-209
            } else if (type == ANNF_TANH || type == ANNF_LINEAR) {
                // output head of an element-wise layer, in the lane that owns the output (same expressions as output_head)
                const double av = (type == ANNF_TANH) ? annf_tanh(zz) : zz;
                const double dd = (av - cen[j]) * kk;                                               // :266-269
                publish(eterm + j, 0.5 * kk * (av - cen[j]) * (av - cen[j]));                       // .h:93-96
                publish(dzL + j, (type == ANNF_TANH) ? dd * (1 - av * av) : dd);                    // :343-354
            } else {
                publish(zL + j, zz);
            }
        });
        sync_all();
        if (l < 2) ANNF_PHASE(7 + 2 * l);
        if (last && !(type == ANNF_TANH || type == ANNF_LINEAR)) {
            // Circular / Softmax output: needs the whole output vector (identical in every CTA)
            for (int t = tid; t < nL; t += kSingleThreads)
                eterm[t] = output_head_parallel(type, nL, zL, 1, cen, kk, dzL, 1, t);
            __syncthreads();
        }
    }
    ANNF_PHASE(3);
    if (rank == 0 && tid == kSingleThreads - 1 && args.energy != nullptr && !(args.flags & ANNF_FLAG_SKIP_ENERGY)) {
        double e = 0.0;
        for (int t = 0; t < nL; t++) e += eterm[t];
        if (args.flags & ANNF_FLAG_ENERGY_DOUBLE) atomicAdd(reinterpret_cast<double*>(args.energy), e);
        else atomicAdd(reinterpret_cast<float*>(args.energy), (float) e);
    }
    ANNF_PHASE(4);

    // ---- backward: back_prop, ANN_ReferenceKernels.cpp:299-373, ending in the forces (:397-410) --------------------
    if (!(args.flags & ANNF_FLAG_SKIP_FORCES)) {
        for (int l = L - 2; l >= 0; l--) {
            const int n_out = hdr[kHdrNodes + l + 1];
            const int m0 = hdr[kHdrBM0 + l], nb = hdr[kHdrBRows + l];
            const int type = l >= 1 ? hdr[kHdrTypes + l - 1] : ANNF_LINEAR;
            const double* al = a + hdr[kHdrNodeOff + l];
            double* dl = d + hdr[kHdrNodeOff + l];
            fast_matvec(blob + hdr[kHdrBOff + l], n_out, nb, d + hdr[kHdrNodeOff + l + 1], hdr[kHdrBLg + l], [&](int r, double sum) {
                const int m = m0 + r;
                if (l >= 1) {
                    const double av = al[m];
                    publish(dl + m, type == ANNF_TANH ? sum * (1 - av * av) : sum);                 // :343-354
                } else if (cartesian) {
                    const int i = m / 3, c = m - 3 * i;
                    atomicAdd(&args.force[slots_s[i] + c * args.padded],
                              (unsigned long long) ((long long) (sum * (double) 0x100000000LL)));
                } else {
                    publish(dl + m, sum);                                                           // dE/d(feature): the chain rule follows
                }
            });
            if (l >= 1 || !cartesian) sync_all();
        }
        if (!cartesian && rank == 0) {
            // get_all_forces_from_derivative_of_first_layer: pair branch :411-421, dihedral branch :392-396 (+ :425-577).
            // Contributions are summed per feature atom in shared memory first, then one fixed-point atomicAdd per component.
            const double* g = d;
            if (args.input_type == ANNF_INPUT_PAIR_DISTANCES) {
                for (int p = tid; p < args.n_feat; p += kSingleThreads) {
                    const int i0 = args.feat[2 * p], i1 = args.feat[2 * p + 1];
                    const double dx = fpos[3 * i0] - fpos[3 * i1], dy = fpos[3 * i0 + 1] - fpos[3 * i1 + 1], dz = fpos[3 * i0 + 2] - fpos[3 * i1 + 2];
                    const double f = g[p] / (args.scale * sqrt(dx * dx + dy * dy + dz * dz));
                    atomicAdd(&facc[3 * i0], -dx * f); atomicAdd(&facc[3 * i0 + 1], -dy * f); atomicAdd(&facc[3 * i0 + 2], -dz * f);
                    atomicAdd(&facc[3 * i1], dx * f); atomicAdd(&facc[3 * i1 + 1], dy * f); atomicAdd(&facc[3 * i1 + 2], dz * f);
                }
            } else {
                for (int q = tid; q < args.n_feat; q += kSingleThreads) {
                    const int* id = args.feat + 4 * q;
                    double f[4][3];
                    dihedral_forces(fpos + 3 * id[0], fpos + 3 * id[1], fpos + 3 * id[2], fpos + 3 * id[3], g[2 * q], g[2 * q + 1], f);
                    for (int k = 0; k < 4; k++)
                        for (int c = 0; c < 3; c++) atomicAdd(&facc[3 * id[k] + c], f[k][c]);
                }
            }
            __syncthreads();
            for (int t = tid; t < 3 * A; t += kSingleThreads) {
                const int i = t / 3, c = t - 3 * i;
                atomicAdd(&args.force[slots_s[i] + c * args.padded], (unsigned long long) ((long long) (facc[t] * (double) 0x100000000LL)));
            }
        }
    }
    ANNF_PHASE(5);
    ANNF_PHASE(6);
    // a CTA must not exit while peers may still store into its shared memory: the last publish is followed by sync_all above
}

// Latency model of one mat-vec (cycles), calibrated on B200 phase clocks: blocks of <= 16 terms per lane whose loads are
// all in flight together, a shuffle round per halving, per pass.
static int pick_lanes_log2(int nrows, int ncols) {
    int forced = -1;
    if (const char* env = getenv("ANNF_SINGLE_LG")) forced = atoi(env);
    int best = 0;
    double best_cost = 1e30;
    for (int lg = 0; lg <= 5; lg++) {
        const int G = 1 << lg;
        if (lg > 0 && G > ncols) break;
        const int terms = (ncols + G - 1) / G;
        const int blocks = (terms + kMatvecBlock - 1) / kMatvecBlock;
        const int passes = (nrows * G + kSingleThreads - 1) / kSingleThreads;
        const double cost = (passes > 0 ? passes : 1) * (blocks * (40.0 + 6.5 * (terms < kMatvecBlock ? terms : kMatvecBlock)) + 65.0 * lg);
        if (cost < best_cost) { best_cost = cost; best = lg; }
    }
    if (forced >= 0 && forced <= 5 && (1 << forced) <= (ncols > 1 ? ncols : 1)) best = forced;
    return best;
}

// Eligibility, work split and the per-CTA parameter blobs (host vector; the caller uploads it).
// w[l] are the fp64 weights of connection l, row-major [n_out][n_in]; b64 the flat biases (net.b_off).
bool fast_single_plan(const NetView& net, int cluster_size, const std::vector<std::vector<double>>& w, const std::vector<double>& b64,
                      FastPlan* plan, std::vector<double>* blobs) {
    const bool cartesian = net.input_type == ANNF_INPUT_CARTESIAN;
    const int L = net.L;
    for (int l = 0; l + 2 < L; l++)
        if (net.types[l] != ANNF_TANH && net.types[l] != ANNF_LINEAR) return false;
    if (net.nodes[L - 1] > 256) return false;
    if (getenv("ANNF_SINGLE_GENERIC")) return false;
    const int cl = cluster_size, A = net.n_sel;
    // pre-centred, pre-scaled transpose of the first connection: F_t = -(1/s) (g_t - mean_c), g = W_0^T dz_1
    const int n0 = net.nodes[0], n1 = net.nodes[1];
    std::vector<double> wsum(3 * (size_t) n1, 0.0), wt0c((size_t) n0 * n1);
    for (int k = 0; k < n1; k++)
        for (int t = 0; t < n0; t++) wsum[(size_t) (t % 3) * n1 + k] += w[0][(size_t) k * n0 + t];
    for (int t = 0; t < n0; t++)
        for (int k = 0; k < n1; k++)
            wt0c[(size_t) t * n1 + k] = cartesian ? -(w[0][(size_t) k * n0 + t] - wsum[(size_t) (t % 3) * n1 + k] / A) / net.scale
                                                  : w[0][(size_t) k * n0 + t];       // pair / dihedral inputs: plain W_0^T, the chain rule follows in the kernel
    // layout of one blob (identical offsets in every CTA)
    int f_per[kMaxLayers], b_per[kMaxLayers], f_off[kMaxLayers], fb_off[kMaxLayers], b_off[kMaxLayers];
    size_t off = kHdrInts / 2;
    for (int l = 0; l + 1 < L; l++) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        f_per[l] = (n_out + cl - 1) / cl;
        b_per[l] = (n_in + cl - 1) / cl;
        f_off[l] = (int) off;  off += (size_t) f_per[l] * n_in;
        fb_off[l] = (int) off; off += f_per[l];
        b_off[l] = (int) off;  off += (size_t) b_per[l] * n_out;
    }
    off += off & 1;                                           // bulk copies move multiples of 16 bytes
    const size_t blob_doubles = off;
    blobs->assign(blob_doubles * cl, 0.0);
    for (int rank = 0; rank < cl; rank++) {
        double* blob = blobs->data() + blob_doubles * rank;
        int* hdr = reinterpret_cast<int*>(blob);
        hdr[kHdrL] = L; hdr[kHdrTotal] = net.node_off[L]; hdr[kHdrNL] = net.nodes[L - 1];
        for (int l = 0; l < L; l++) { hdr[kHdrNodes + l] = net.nodes[l]; hdr[kHdrNodeOff + l] = net.node_off[l]; }
        hdr[kHdrNodeOff + L] = net.node_off[L];
        for (int l = 0; l + 1 < L; l++) {
            const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
            const int j0 = rank * f_per[l], nf = std::max(0, std::min(f_per[l], n_out - j0));
            const int m0 = rank * b_per[l], nb = std::max(0, std::min(b_per[l], n_in - m0));
            hdr[kHdrTypes + l] = net.types[l];
            hdr[kHdrFRows + l] = nf; hdr[kHdrFLg + l] = pick_lanes_log2(f_per[l], n_in); hdr[kHdrFOff + l] = f_off[l];
            hdr[kHdrFbOff + l] = fb_off[l]; hdr[kHdrFJ0 + l] = j0;
            hdr[kHdrBRows + l] = nb; hdr[kHdrBLg + l] = pick_lanes_log2(b_per[l], n_out); hdr[kHdrBOff + l] = b_off[l];
            hdr[kHdrBM0 + l] = m0;
            for (int r = 0; r < nf; r++) {
                for (int c = 0; c < n_in; c++) blob[f_off[l] + (size_t) r * n_in + c] = w[l][(size_t) (j0 + r) * n_in + c];
                blob[fb_off[l] + r] = b64[net.b_off[l] + j0 + r];
            }
            for (int r = 0; r < nb; r++)
                for (int c = 0; c < n_out; c++)
                    blob[b_off[l] + (size_t) r * n_out + c] = l == 0 ? wt0c[(size_t) (m0 + r) * n_out + c] : w[l][(size_t) c * n_in + m0 + r];
        }
    }
    plan->cl = cl;
    plan->blob_doubles = (int) blob_doubles;
    const int total = net.node_off[L], nL = net.nodes[L - 1];
    plan->smem_doubles = (int) (blob_doubles + 2 * (size_t) total + 2 * (size_t) nL + net.num_center + 1 + 3 * kSingleWarps + 4 +
                                (cartesian ? 0 : 6 * (size_t) A) + ((size_t) A + 1) / 2 + 2);
    return (size_t) plan->smem_doubles * sizeof(double) <= 200 * 1024;
}

cudaError_t prepare_single_fast() {
    cudaError_t err = cudaFuncSetAttribute(annf_single_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (err != cudaSuccess) return err;
    return cudaFuncSetAttribute(annf_single_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
}

cudaError_t launch_single_fast(const NetView& net, const FastPlan& plan, const double* blobs, const void* posq, const void* posq_corr,
                               unsigned long long* force, int padded, void* energy, unsigned flags, cudaStream_t stream, bool allow_pdl) {
    FastArgs args;
    args.blob = blobs;
    args.blob_doubles = plan.blob_doubles;
    args.A = net.n_sel;
    args.num_center = net.num_center;
    args.total = net.node_off[net.L];
    args.nL = net.nodes[net.L - 1];
    args.sel_slots = net.sel_slots;
    args.center = net.center;
    args.k_dev = net.k_dev;
    args.inv_scale = 1.0 / net.scale;
    args.inv_A = 1.0 / net.n_sel;
    args.posq = posq;
    args.posq_corr = posq_corr;
    args.force = force;
    args.padded = padded;
    args.energy = energy;
    args.flags = flags;
    args.input_type = net.input_type;
    args.n_feat = net.input_type == ANNF_INPUT_PAIR_DISTANCES ? net.n_pairs : net.n_dihedrals;
    args.feat = net.input_type == ANNF_INPUT_PAIR_DISTANCES ? net.pairs : net.dihedrals;
    args.scale = net.scale;
    const size_t smem = sizeof(double) * (size_t) plan.smem_doubles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.cl);
    cfg.blockDim = dim3(kSingleThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see griddepcontrol in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = plan.cl;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    // allow_pdl == false (the first evaluation after annf_set_atom_order): a plain, fully serialised launch — the kernel reads
    // sel_slots before griddepcontrol.wait
    cfg.attrs = allow_pdl ? attr : attr + 1;
    cfg.numAttrs = (allow_pdl ? 1 : 0) + (plan.cl == 1 ? 0 : 1);
    if (plan.cl == 1) return cudaLaunchKernelEx(&cfg, annf_single_fast_kernel<false>, args);
    return cudaLaunchKernelEx(&cfg, annf_single_fast_kernel<true>, args);
}

}  // namespace annf
