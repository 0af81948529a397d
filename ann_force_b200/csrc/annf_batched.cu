// Batched-replica fused kernel (sm_100a): many umbrella windows / walkers per launch.
//
// One CTA evaluates a tile of TR replicas end to end — gather + centre, every layer forward, the
// fp64 output head (energy + output gradient), every layer backward, the centring Jacobian and the
// force write — with all activations of the tile resident in shared memory, laid out
// [node][replica] so that one 128-bit broadcast load feeds 4 fp32 (2 fp64) replicas.
// A layer is a [nodes x TR] register-tiled contraction: a warp owns 32 consecutive output nodes
// (forward, weights read from the transposed copy) or 32 consecutive input nodes (backward,
// weights read from the original copy), i.e. weight loads are always unit-stride across the warp
// and each weight is re-used for all TR replicas from registers.  When a layer has fewer
// 32-node chunks than the CTA has warps, the contraction dimension is split across warps and the
// partial sums are combined through shared memory.
//
// Arithmetic follows ReferenceCalcANN_ForceKernel::candidate_2
// (/root/reference/platforms/reference/ANN_ReferenceKernels.cpp:122-171); T = double gives the
// reference's own precision, T = float runs the hidden contractions in fp32 FMA while centring,
// the output layer, the energy and the output gradient stay fp64.
#include "annf_common.cuh"

namespace annf {

constexpr int kBatchThreads = 256;
constexpr int kBatchWarps = kBatchThreads / 32;

struct BatchedArgs {
    NetView net;
    int R;
    const float* coords;
    int atoms_per_replica;
    const float* centers;      // [R][num_center] or nullptr
    float* forces;
    double* energies;
    int x_off[kMaxLayers];     // per hidden layer: offset (in T elements / TR) of its 2*n extra slots, or -1
    int extra_nodes;           // total extra slot rows
    int ws_off[kMaxLayers];    // shared-memory weight staging: element offset of connection l ...
    int ws_ld[kMaxLayers];     // ... and its (odd) row stride; rows are [n_out][ws_ld]
    int ws_total;              // staged elements in total, 0 = weights stay in global memory
};

template <typename T, int TR>
__device__ __forceinline__ void load_row(const T* p, T (&x)[TR]) {
    if constexpr ((sizeof(T) * TR) % 16 == 0) {
        constexpr int V = (int) (sizeof(T) * TR / 16);
        const int4* p4 = reinterpret_cast<const int4*>(p);
        int4* x4 = reinterpret_cast<int4*>(x);
#pragma unroll
        for (int v = 0; v < V; v++) x4[v] = p4[v];
    } else {
#pragma unroll
        for (int r = 0; r < TR; r++) x[r] = p[r];
    }
}

// acc[r] += sum_{i in [i0,i1)} w(i, j) * in[i][r]   with w(i, j) = Wp[i * s_loop + j * s_lane].
// Weight loads are issued eight at a time ahead of their use: the loop is latency-bound on them otherwise.
template <typename T, typename ACC, int TR>
__device__ __forceinline__ void accumulate_range(const T* __restrict__ Wp, size_t s_loop, size_t s_lane, int j, bool active,
                                                 int i0, int i1, const T* in, ACC (&acc)[TR]) {
    constexpr int U = 8;
    const T* wj = Wp + (size_t) j * s_lane;
    int i = i0;
    for (; i + U <= i1; i += U) {
        T w[U];
#pragma unroll
        for (int u = 0; u < U; u++) w[u] = active ? wj[(size_t) (i + u) * s_loop] : T(0);
#pragma unroll
        for (int u = 0; u < U; u++) {
            T x[TR];
            load_row<T, TR>(in + (size_t) (i + u) * TR, x);
#pragma unroll
            for (int r = 0; r < TR; r++) acc[r] = fma((ACC) w[u], (ACC) x[r], acc[r]);
        }
    }
    for (; i < i1; i++) {
        const T w = active ? wj[(size_t) i * s_loop] : T(0);
        T x[TR];
        load_row<T, TR>(in + (size_t) i * TR, x);
#pragma unroll
        for (int r = 0; r < TR; r++) acc[r] = fma((ACC) w, (ACC) x[r], acc[r]);
    }
}

// out[j][r] = sum_i w(i, j) * in[i][r]   for j < n_lane, r < TR; epilogue(j, acc) consumes it.
// Must be called by all threads of the CTA (contains __syncthreads when the loop dimension is split).
template <typename T, typename ACC, int TR, typename Epilogue>
__device__ __forceinline__ void dense_tile(const T* __restrict__ Wp, size_t s_loop, size_t s_lane, int n_lane, int n_loop,
                                           const T* in, ACC* scratch, Epilogue epilogue) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (n_lane + 31) >> 5;
    if (nchunks >= kBatchWarps) {
        for (int chunk = warp; chunk < nchunks; chunk += kBatchWarps) {
            const int j = chunk * 32 + lane;
            const bool active = j < n_lane;
            ACC acc[TR];
#pragma unroll
            for (int r = 0; r < TR; r++) acc[r] = ACC(0);
            accumulate_range<T, ACC, TR>(Wp, s_loop, s_lane, active ? j : 0, active, 0, n_loop, in, acc);
            if (active) epilogue(j, acc);
        }
        return;
    }
    const int ks = kBatchWarps / nchunks;            // warps per chunk along the loop dimension
    const bool busy = warp < nchunks * ks;
    const int chunk = warp / ks, part = warp - chunk * ks;
    const int j = chunk * 32 + lane;
    const bool active = busy && j < n_lane;
    ACC acc[TR];
#pragma unroll
    for (int r = 0; r < TR; r++) acc[r] = ACC(0);
    if (busy) {
        const int per = (n_loop + ks - 1) / ks;
        const int i0 = part * per, i1 = min(n_loop, i0 + per);
        accumulate_range<T, ACC, TR>(Wp, s_loop, s_lane, active ? j : 0, active, i0, i1, in, acc);
    }
    if (ks > 1) {
        if (busy && part > 0) {
#pragma unroll
            for (int r = 0; r < TR; r++) scratch[((size_t) warp * TR + r) * 32 + lane] = acc[r];
        }
        __syncthreads();
        if (busy && part == 0) {
            for (int p = 1; p < ks; p++) {
#pragma unroll
                for (int r = 0; r < TR; r++) acc[r] += scratch[((size_t) (warp + p) * TR + r) * 32 + lane];
            }
        }
    }
    if (active && part == 0) epilogue(j, acc);
}

template <typename T, int TR>
__global__ void __launch_bounds__(kBatchThreads) annf_batched_kernel(const BatchedArgs args) {
    const NetView& net = args.net;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = net.L, total = net.node_off[L];
    const int n0 = net.nodes[0], nL = net.nodes[L - 1];
    const int r0 = blockIdx.x * TR;
    const int N = args.atoms_per_replica;
    const double scale = net.scale;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* act = reinterpret_cast<T*>(smem_raw);                         // [total][TR]
    T* extra = act + (size_t) total * TR;                            // [extra_nodes][TR]
    const size_t t_bytes = (sizeof(T) * (size_t) (total + args.extra_nodes) * TR + 15) & ~(size_t) 15;
    double* zout = reinterpret_cast<double*>(smem_raw + t_bytes);   // [nL][TR]
    double* dzout = zout + (size_t) nL * TR;                         // [nL][TR]
    double* cen = dzout + (size_t) nL * TR;                          // [TR][num_center]
    double* com = cen + (size_t) TR * net.num_center;                // [TR*3]
    double* scratch = com + TR * 3 + 1;                              // [kBatchWarps*TR*32] (as ACC)
    const int n_fpos = net.input_type == ANNF_INPUT_CARTESIAN ? 0 : 3 * net.n_sel;
    double* fpos = scratch + (size_t) kBatchWarps * TR * 32;         // [3*n_feat][TR] positions of the feature atoms (pair / dihedral inputs)
    double* facc = fpos + (size_t) n_fpos * TR;                      // [3*n_feat][TR] per-atom force accumulators
    T* wstage = reinterpret_cast<T*>(facc + (size_t) n_fpos * TR);   // [ws_total] weights, rows padded to an odd stride

    // ---- weights that fit are staged in shared memory once per CTA: [n_out][odd stride] serves the forward
    //      (lane = output row, stride odd -> conflict-free) and the backward (lane = input column) ----------------
    const bool staged = args.ws_total > 0;
    if (staged) {
        for (int l = 0; l + 1 < L; l++) {
            const int n_in = net.nodes[l], n_out = net.nodes[l + 1], ld = args.ws_ld[l];
            const T* src = WeightsOf<T>::W(net) + net.w_off[l];
            T* dst = wstage + args.ws_off[l];
            for (int idx = tid; idx < n_in * n_out; idx += kBatchThreads) {
                const int row = idx / n_in, col = idx - row * n_in;
                dst[(size_t) row * ld + col] = __ldg(src + idx);
            }
        }
    }

    // ---- umbrella centres of this tile -----------------------------------------------------------
    for (int idx = tid; idx < TR * net.num_center; idx += kBatchThreads) {
        const int r = idx / net.num_center, c = idx - r * net.num_center;
        const int rr = min(r0 + r, args.R - 1);
        cen[idx] = args.centers ? (double) args.centers[(size_t) rr * net.num_center + c] : net.center[c];
    }

    const int A = net.n_sel;
    if (net.input_type == ANNF_INPUT_CARTESIAN) {
        // ---- feature stage (Cartesian): ANN_ReferenceKernels.cpp:128-151 --------------------------------
        for (int p = warp; p < TR * 3; p += kBatchWarps) {
            const int r = p / 3, c = p - 3 * r, rr = r0 + r;
            double s = 0.0;
            if (rr < args.R) {
                const float* base = args.coords + (size_t) rr * N * 3;
                for (int i = lane; i < A; i += 32) s += (double) __ldg(base + (size_t) net.sel_atoms[i] * 3 + c) / scale;
            }
            s = warp_sum(s);
            if (lane == 0) com[p] = s / A;                                // unweighted centroid (:137-141)
        }
        __syncthreads();
        for (int idx = tid; idx < n0 * TR; idx += kBatchThreads) {
            const int r = idx / n0, i = idx - r * n0, rr = r0 + r;
            const int atom = i / 3, c = i - 3 * atom;
            double v = 0.0;
            if (rr < args.R)
                v = (double) __ldg(args.coords + ((size_t) rr * N + net.sel_atoms[atom]) * 3 + c) / scale - com[r * 3 + c];
            act[(size_t) i * TR + r] = (T) v;
        }
    } else {
        // ---- pair distances (:152-160) / dihedral cos, sin (:125-127, :579-636): stage the feature atoms -------
        for (int idx = tid; idx < n_fpos * TR; idx += kBatchThreads) {
            const int r = idx / n_fpos, i = idx - r * n_fpos, rr = min(r0 + r, args.R - 1);
            const int atom = i / 3, c = i - 3 * atom;
            fpos[(size_t) i * TR + r] = (double) __ldg(args.coords + ((size_t) rr * N + net.sel_atoms[atom]) * 3 + c);
            facc[(size_t) i * TR + r] = 0.0;
        }
        __syncthreads();
        if (net.input_type == ANNF_INPUT_PAIR_DISTANCES) {
            const double inv = 1.0 / scale;
            for (int idx = tid; idx < net.n_pairs * TR; idx += kBatchThreads) {
                const int p = idx / TR, r = idx - p * TR;
                const int i0 = net.pairs[2 * p], i1 = net.pairs[2 * p + 1];
                double d2 = 0.0;
                for (int c = 0; c < 3; c++) {
                    const double d = fpos[(size_t) (3 * i0 + c) * TR + r] * inv - fpos[(size_t) (3 * i1 + c) * TR + r] * inv;
                    d2 += d * d;
                }
                act[(size_t) p * TR + r] = (T) sqrt(d2);
            }
        } else {
            for (int idx = tid; idx < net.n_dihedrals * TR; idx += kBatchThreads) {
                const int q = idx / TR, r = idx - q * TR;
                double pt[4][3];
                for (int k = 0; k < 4; k++)
                    for (int c = 0; c < 3; c++) pt[k][c] = fpos[(size_t) (3 * net.dihedrals[4 * q + k] + c) * TR + r];
                double cs, sn;
                dihedral_cos_sin(pt[0], pt[1], pt[2], pt[3], cs, sn);
                act[(size_t) (2 * q) * TR + r] = (T) cs;
                act[(size_t) (2 * q + 1) * TR + r] = (T) sn;
            }
        }
    }
    __syncthreads();

    // ---- forward: calculate_output_of_each_layer, ANN_ReferenceKernels.cpp:179-260 --------------------
    for (int l = 0; l + 1 < L; l++) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        // forward weight w(i = input, j = output): transposed global copy [n_in][n_out], or staged [n_out][ld]
        const T* Wt = staged ? wstage + args.ws_off[l] : WeightsOf<T>::Wt(net) + net.w_off[l];
        const size_t s_loop = staged ? 1 : (size_t) n_out, s_lane = staged ? (size_t) args.ws_ld[l] : 1;
        const T* b = WeightsOf<T>::b(net) + net.b_off[l];
        const T* in = act + (size_t) net.node_off[l] * TR;
        T* out = act + (size_t) net.node_off[l + 1] * TR;
        const int type = net.types[l];
        if (l + 2 == L) {
            // output layer: fp64 accumulation, pre-activations kept in fp64 for the head
            dense_tile<T, double, TR>(Wt, s_loop, s_lane, n_out, n_in, in, scratch, [&](int j, const double (&acc)[TR]) {
                const double bj = (double) b[j];
#pragma unroll
                for (int r = 0; r < TR; r++) zout[(size_t) j * TR + r] = acc[r] + bj;
            });
        } else if (type == ANNF_TANH || type == ANNF_LINEAR) {
            dense_tile<T, T, TR>(Wt, s_loop, s_lane, n_out, n_in, in, reinterpret_cast<T*>(scratch), [&](int j, const T (&acc)[TR]) {
                const T bj = b[j];
#pragma unroll
                for (int r = 0; r < TR; r++) {
                    const T zz = acc[r] + bj;
                    out[(size_t) j * TR + r] = (type == ANNF_TANH) ? act_tanh(zz) : zz;
                }
            });
        } else {
            // Circular / Softmax hidden layer: keep z, then transform the whole vector
            T* zbuf = extra + (size_t) args.x_off[l + 1] * TR;
            dense_tile<T, T, TR>(Wt, s_loop, s_lane, n_out, n_in, in, reinterpret_cast<T*>(scratch), [&](int j, const T (&acc)[TR]) {
                const T bj = b[j];
#pragma unroll
                for (int r = 0; r < TR; r++) zbuf[(size_t) j * TR + r] = acc[r] + bj;
            });
            __syncthreads();
            if (type == ANNF_CIRCULAR) {
                for (int idx = tid; idx < (n_out / 2) * TR; idx += kBatchThreads) {
                    const int j = idx / TR, r = idx - j * TR;
                    const T p = zbuf[(size_t) (2 * j) * TR + r], q = zbuf[(size_t) (2 * j + 1) * TR + r];
                    const T rad = sqrt(p * p + q * q);
                    out[(size_t) (2 * j) * TR + r] = p / rad;
                    out[(size_t) (2 * j + 1) * TR + r] = q / rad;
                }
            } else {
                for (int idx = tid; idx < n_out * TR; idx += kBatchThreads) {
                    const int j = idx / TR, r = idx - j * TR;
                    T sum = T(0);
                    for (int s = 0; s < n_out; s++) sum += exp(zbuf[(size_t) s * TR + r]);
                    out[(size_t) j * TR + r] = exp(zbuf[(size_t) j * TR + r]) / sum;
                }
            }
        }
        __syncthreads();
    }

    // ---- fp64 head: energy + dE/dz of the output layer; one thread per (replica, output or Circular pair) --------
    {
        double* eterm = scratch;                                     // [TR][nL] energy terms
        for (int idx = tid; idx < TR * nL; idx += kBatchThreads) {
            const int r = idx / nL, t = idx - r * nL;
            eterm[idx] = output_head_parallel(net.types[L - 2], nL, zout + r, TR, cen + (size_t) r * net.num_center, *net.k_dev,
                                              dzout + r, TR, t);
        }
        __syncthreads();
        if (tid < TR && r0 + tid < args.R) {
            double e = 0.0;
            for (int t = 0; t < nL; t++) e += eterm[tid * nL + t];
            args.energies[r0 + tid] = e;
        }
    }
    __syncthreads();
    {
        T* dL = act + (size_t) net.node_off[L - 1] * TR;
        for (int idx = tid; idx < nL * TR; idx += kBatchThreads) dL[idx] = (T) dzout[idx];
    }
    __syncthreads();

    // ---- backward: back_prop, ANN_ReferenceKernels.cpp:299-373 --------------------------------------------
    for (int l = L - 2; l >= 0; l--) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        // backward weight w(k = output, m = input): original copy [n_out][n_in], or staged [n_out][ld]
        const T* W = staged ? wstage + args.ws_off[l] : WeightsOf<T>::W(net) + net.w_off[l];
        const size_t s_loop = staged ? (size_t) args.ws_ld[l] : (size_t) n_in;
        const T* dz = act + (size_t) net.node_off[l + 1] * TR;     // dE/dz of layer l+1 (overwrote its activations)
        T* slot = act + (size_t) net.node_off[l] * TR;            // holds a_l, receives dE/dz_l (or dE/dx for l = 0)
        const int type = l >= 1 ? net.types[l - 1] : ANNF_LINEAR;
        if (type == ANNF_TANH || type == ANNF_LINEAR) {
            dense_tile<T, T, TR>(W, s_loop, 1, n_in, n_out, dz, reinterpret_cast<T*>(scratch), [&](int m, const T (&acc)[TR]) {
#pragma unroll
                for (int r = 0; r < TR; r++) {
                    if (type == ANNF_TANH) {
                        const T a = slot[(size_t) m * TR + r];
                        slot[(size_t) m * TR + r] = acc[r] * (T(1) - a * a);      // :343-348
                    } else {
                        slot[(size_t) m * TR + r] = acc[r];                        // :349-354
                    }
                }
            });
        } else {
            T* zbuf = extra + (size_t) args.x_off[l] * TR;
            T* dbuf = zbuf + (size_t) n_in * TR;
            dense_tile<T, T, TR>(W, s_loop, 1, n_in, n_out, dz, reinterpret_cast<T*>(scratch), [&](int m, const T (&acc)[TR]) {
#pragma unroll
                for (int r = 0; r < TR; r++) dbuf[(size_t) m * TR + r] = acc[r];
            });
            __syncthreads();
            if (type == ANNF_CIRCULAR) {                                            // :307-318
                for (int idx = tid; idx < (n_in / 2) * TR; idx += kBatchThreads) {
                    const int j = idx / TR, r = idx - j * TR;
                    const T p = zbuf[(size_t) (2 * j) * TR + r], q = zbuf[(size_t) (2 * j + 1) * TR + r];
                    const T dp = dbuf[(size_t) (2 * j) * TR + r], dq = dbuf[(size_t) (2 * j + 1) * TR + r];
                    const T rad = sqrt(p * p + q * q), r3 = rad * rad * rad;
                    slot[(size_t) (2 * j) * TR + r] = q / r3 * (q * dp - p * dq);
                    slot[(size_t) (2 * j + 1) * TR + r] = p / r3 * (p * dq - q * dp);
                }
            } else {                                                                // Softmax :355-365
                for (int idx = tid; idx < n_in * TR; idx += kBatchThreads) {
                    const int j = idx / TR, r = idx - j * TR;
                    T dot = T(0);
                    for (int s = 0; s < n_in; s++) dot += slot[(size_t) s * TR + r] * dbuf[(size_t) s * TR + r];
                    const T a = slot[(size_t) j * TR + r];
                    zbuf[(size_t) j * TR + r] = a * dbuf[(size_t) j * TR + r] - a * dot;
                }
                __syncthreads();
                for (int idx = tid; idx < n_in * TR; idx += kBatchThreads) slot[idx] = zbuf[idx];
            }
        }
        __syncthreads();
    }

    if (net.input_type == ANNF_INPUT_CARTESIAN) {
        // ---- forces: ANN_ReferenceKernels.cpp:397-410 (centring Jacobian as one mean, O(A)) -----------------
        for (int p = warp; p < TR * 3; p += kBatchWarps) {
            const int r = p / 3, c = p - 3 * r;
            double s = 0.0;
            for (int i = lane; i < A; i += 32) s += (double) act[(size_t) (3 * i + c) * TR + r];
            s = warp_sum(s);
            if (lane == 0) com[p] = s / A;
        }
        __syncthreads();
        for (int idx = tid; idx < n0 * TR; idx += kBatchThreads) {
            const int r = idx / n0, i = idx - r * n0, rr = r0 + r;
            if (rr >= args.R) continue;
            const int atom = i / 3, c = i - 3 * atom;
            const double f = -((double) act[(size_t) i * TR + r] - com[r * 3 + c]) / scale;
            args.forces[((size_t) rr * N + net.sel_atoms[atom]) * 3 + c] = (float) f;
        }
    } else {
        // ---- forces for pair (:411-421) / dihedral (:392-396, :425-577) inputs: accumulate per feature atom, then write -----
        if (net.input_type == ANNF_INPUT_PAIR_DISTANCES) {
            for (int idx = tid; idx < net.n_pairs * TR; idx += kBatchThreads) {
                const int p = idx / TR, r = idx - p * TR;
                const int i0 = net.pairs[2 * p], i1 = net.pairs[2 * p + 1];
                double d[3], d2 = 0.0;
                for (int c = 0; c < 3; c++) {
                    d[c] = fpos[(size_t) (3 * i0 + c) * TR + r] - fpos[(size_t) (3 * i1 + c) * TR + r];
                    d2 += d[c] * d[c];
                }
                const double f = (double) act[(size_t) p * TR + r] / (scale * sqrt(d2));
                for (int c = 0; c < 3; c++) {
                    atomicAdd(&facc[(size_t) (3 * i0 + c) * TR + r], -d[c] * f);
                    atomicAdd(&facc[(size_t) (3 * i1 + c) * TR + r], d[c] * f);
                }
            }
        } else {
            for (int idx = tid; idx < net.n_dihedrals * TR; idx += kBatchThreads) {
                const int q = idx / TR, r = idx - q * TR;
                double pt[4][3], f[4][3];
                for (int k = 0; k < 4; k++)
                    for (int c = 0; c < 3; c++) pt[k][c] = fpos[(size_t) (3 * net.dihedrals[4 * q + k] + c) * TR + r];
                dihedral_forces(pt[0], pt[1], pt[2], pt[3], (double) act[(size_t) (2 * q) * TR + r],
                                (double) act[(size_t) (2 * q + 1) * TR + r], f);
                for (int k = 0; k < 4; k++)
                    for (int c = 0; c < 3; c++) atomicAdd(&facc[(size_t) (3 * net.dihedrals[4 * q + k] + c) * TR + r], f[k][c]);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < n_fpos * TR; idx += kBatchThreads) {
            const int r = idx / n_fpos, i = idx - r * n_fpos, rr = r0 + r;
            if (rr >= args.R) continue;
            const int atom = i / 3, c = i - 3 * atom;
            args.forces[((size_t) rr * N + net.sel_atoms[atom]) * 3 + c] = (float) facc[(size_t) i * TR + r];
        }
    }
}

template <typename T>
static size_t batched_smem_bytes(const NetView& net, int extra_nodes, int tr) {
    const size_t total = net.node_off[net.L];
    const size_t nL = net.nodes[net.L - 1];
    const size_t t_bytes = (sizeof(T) * (total + extra_nodes) * tr + 15) & ~(size_t) 15;
    const size_t n_fpos = net.input_type == ANNF_INPUT_CARTESIAN ? 0 : 3 * (size_t) net.n_sel;
    return t_bytes + sizeof(double) * (2 * nL * tr + (size_t) tr * net.num_center + tr * 3 + 1)
           + sizeof(double) * kBatchWarps * tr * 32 + sizeof(double) * 2 * n_fpos * tr;
}

size_t batched_smem_bytes(const NetView& net, int extra_nodes, int tr, bool fp64) {
    return fp64 ? batched_smem_bytes<double>(net, extra_nodes, tr) : batched_smem_bytes<float>(net, extra_nodes, tr);
}

// Elements needed to stage every connection as [n_out][odd stride]; fills the per-connection offsets.
static size_t staging_elements(const NetView& net, int* off, int* ld) {
    size_t total = 0;
    for (int l = 0; l + 1 < net.L; l++) {
        ld[l] = net.nodes[l] | 1;
        off[l] = (int) total;
        total += (size_t) net.nodes[l + 1] * ld[l];
    }
    return total;
}

template <typename T, int TR>
static cudaError_t launch_t(const BatchedArgs& args, size_t smem, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(annf_batched_kernel<T, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (err != cudaSuccess) return err;
        configured[dev] = true;
    }
    const int grid = (args.R + TR - 1) / TR;
    annf_batched_kernel<T, TR><<<grid, kBatchThreads, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_batched(const NetView& net, const int* x_off, int extra_nodes, int tr, bool fp64, int R,
                           const float* coords, int atoms_per_replica, const float* centers, float* forces,
                           double* energies, cudaStream_t stream) {
    BatchedArgs args;
    args.net = net;
    args.R = R;
    args.coords = coords;
    args.atoms_per_replica = atoms_per_replica;
    args.centers = centers;
    args.forces = forces;
    args.energies = energies;
    for (int l = 0; l < kMaxLayers; l++) args.x_off[l] = x_off[l];
    args.extra_nodes = extra_nodes;
    size_t smem = batched_smem_bytes(net, extra_nodes, tr, fp64);
    const size_t stage = staging_elements(net, args.ws_off, args.ws_ld) * (fp64 ? sizeof(double) : sizeof(float));
    if (smem + stage <= 200 * 1024) {
        args.ws_total = (int) (stage / (fp64 ? sizeof(double) : sizeof(float)));
        smem += stage;
    } else {
        args.ws_total = 0;
    }
    if (fp64) {
        switch (tr) {
        case 8: return launch_t<double, 8>(args, smem, stream);
        case 4: return launch_t<double, 4>(args, smem, stream);
        case 2: return launch_t<double, 2>(args, smem, stream);
        default: return launch_t<double, 1>(args, smem, stream);
        }
    }
    switch (tr) {
    case 8: return launch_t<float, 8>(args, smem, stream);
    case 4: return launch_t<float, 4>(args, smem, stream);
    case 2: return launch_t<float, 2>(args, smem, stream);
    default: return launch_t<float, 1>(args, smem, stream);
    }
}

}  // namespace annf
