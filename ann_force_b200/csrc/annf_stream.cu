// Batched-replica kernel with TMA-streamed weights (sm_100a): the path for BASELINE configs 3 and 4
// (umbrella windows / walkers of a small or medium network, Cartesian input, Tanh / Linear hidden layers).
//
// One CTA evaluates a tile of TR replicas end to end, as annf_batched.cu does, but the weights never go through the
// register file on their way in: the handle holds ONE linear "weight stream" — every matrix in exactly the order and
// orientation the evaluation consumes it (W_0^T, W_1^T, ... for the forward pass, then W_{L-2}, ..., W_0 for the backward
// pass), rows padded to 16 bytes — and a dedicated producer warp walks it with cp.async.bulk (the TMA engine, 1-D bulk
// copies of <= 16 KB) into a 4-stage shared-memory ring guarded by full / empty mbarriers.  The eight compute warps
// never wait on L2 latency: they read weights and activations from shared memory only, and the next layer's weights
// are already landing while the current layer's epilogue (bias, tanh, K-split reduction) runs.
//   * coordinates of the replica tile arrive the same way: one bulk copy of TR x N x 3 floats (when 16-byte aligned);
//   * the first connection's backward matrix in the stream is pre-centred and pre-scaled,
//     W0c[k][t] = -(1/s) (W_0[k][t] - (1/A) sum_i W_0[k][3i+c(t)]), the Jacobian of "divide by s, subtract the centroid"
//     (ANN_ReferenceKernels.cpp:137-145, :397-410), so the last contraction yields forces;
//   * a contraction maps lane -> node (32 consecutive nodes per warp pass, `nc` such chunks per warp sharing each
//     broadcast activation load), TR accumulators per node in registers; when a layer has fewer node chunks than warps
//     the rows of every weight chunk are dealt round-robin to `ks` warps and the partial sums meet in shared memory,
//     where ALL threads run the epilogue (one tanh per thread instead of TR per lane).
// Arithmetic follows ReferenceCalcANN_ForceKernel::candidate_2 (ANN_ReferenceKernels.cpp:122-171); T = double is the
// reference's own precision, T = float keeps centring, the output layer, the energy and the output gradient in fp64.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "annf_common.cuh"
#include "annf_ptx.cuh"

namespace annf {

constexpr int kStreamComputeWarps = 16;
constexpr int kStreamComputeThreads = 32 * kStreamComputeWarps;
constexpr int kStreamThreads = kStreamComputeThreads + 32;          // + one producer warp
constexpr int kStreamStages = 4;                                     // power of two
constexpr int kStreamStageBytes = 16 * 1024;
constexpr int kStreamMaxNC = 2;                                      // 32-node chunks per warp

// Kept small on purpose: everything layer-dependent is in the StreamPlan that arrives in shared memory.
struct StreamArgs {
    const void* statics;             // device blob: StreamPlan | T bias[total nodes] (16-byte padded) | int sel_atoms[A] (padded)
    int static_bytes;
    const void* stream;              // T[]: the weight stream
    const double* center;
    const double* k_dev;
    double inv_scale, inv_A;
    const float* coords;
    const float* centers;
    float* forces;
    double* energies;
    int R, atoms_per_replica, A, num_center;
    int L, total, n0, nL;
    int scratch_elems;               // K-split scratch of this launch, in doubles
    int coord_stage_bytes;           // > 0: the tile's coordinates are bulk-copied into shared memory (fits the scratch region)
    int out_type;                    // activation of the output layer
};

// Diagnostic: clock64() of thread 0 / CTA 0 at the phase boundaries of the last launch (annf_debug_stream_phases):
// [0] start, [1] barriers + prologue loads, [2] features, [3 + p] pass p done, [3 + npass] head, [4 + npass] end.
__device__ long long g_stream_phase_clock[64];
#ifdef ANNF_PHASE_CLOCKS        // diagnostic build only (make EXTRA=-DANNF_PHASE_CLOCKS): the product kernels carry no instrumentation
#define ANNF_SPHASE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_stream_phase_clock[i] = clock64(); } while (0)
#else
#define ANNF_SPHASE(i) do { } while (0)
#endif

namespace {

using namespace ptx;

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kStreamComputeThreads) : "memory"); }

template <typename T, int TR>
__device__ __forceinline__ void load_tile_row(const T* p, T (&x)[TR]) {
    if constexpr ((sizeof(T) * TR) % 16 == 0) {
        constexpr int V = (int) (sizeof(T) * TR / 16);
        const int4* p4 = reinterpret_cast<const int4*>(p);
        int4* x4 = reinterpret_cast<int4*>(x);
#pragma unroll
        for (int v = 0; v < V; v++) x4[v] = p4[v];
    } else if constexpr ((sizeof(T) * TR) % 8 == 0) {
        constexpr int V = (int) (sizeof(T) * TR / 8);
        const int2* p2 = reinterpret_cast<const int2*>(p);
        int2* x2 = reinterpret_cast<int2*>(x);
#pragma unroll
        for (int v = 0; v < V; v++) x2[v] = p2[v];
    } else {
#pragma unroll
        for (int r = 0; r < TR; r++) x[r] = p[r];
    }
}

// Rows part, part + ks, ... of one weight chunk: acc[q][r] += w[row][32 q + lane] * in[row][r].  Rows are zero padded to
// a multiple of 32 * NC columns, so no lane needs a bounds check and the loads sit ahead of the FMAs.
template <typename T, typename ACC, int TR, int NC>
__device__ __forceinline__ void stream_rows(const T* wc, const T* in, int part, int nrows, int ks, int ld, ACC (&acc)[kStreamMaxNC][TR]) {
#pragma unroll 2
    for (int row = part; row < nrows; row += ks) {
        T x[TR], w[NC];
        load_tile_row<T, TR>(in + (size_t) row * TR, x);
#pragma unroll
        for (int q = 0; q < NC; q++) w[q] = wc[(size_t) row * ld + 32 * q];
#pragma unroll
        for (int q = 0; q < NC; q++)
#pragma unroll
            for (int r = 0; r < TR; r++) acc[q][r] = fma((ACC) w[q], (ACC) x[r], acc[q][r]);
    }
}

// One contraction out[j][r] = sum_i M[i][j] * in[i][r], M streamed through the ring in chunks of rows.
// epi(j, r, value) is called exactly once per (j < n_lane, r < TR).  Executed by the compute warps only.
template <typename T, typename ACC, int TR, typename Epi>
__device__ __forceinline__ void stream_pass(const StreamPlan& plan, int p, int& chunk, const unsigned char* ring, uint32_t full0,
                                            uint32_t empty0, const T* in, double* scratch_raw, Epi epi) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nc = plan.nc[p], groups = plan.groups[p], ks = plan.ks[p];
    const int n_lane = plan.n_lane[p], n_loop = plan.n_loop[p], ld = plan.ld[p], rows_per = plan.rows_per[p], nch = plan.nchunks[p];
    const bool busy = warp < groups * ks;
    const int grp = warp / ks, part = warp - grp * ks;
    const int jbase = grp * nc * 32 + lane;
    ACC acc[kStreamMaxNC][TR];
#pragma unroll
    for (int q = 0; q < kStreamMaxNC; q++)
#pragma unroll
        for (int r = 0; r < TR; r++) acc[q][r] = ACC(0);

    // every CTA walks the chunks of a pass from its own starting point (see the producer)
    int tt = blockIdx.x % nch;
    for (int t = 0; t < nch; t++, chunk++) {
        const int s = chunk & (kStreamStages - 1);
        bar_wait(full0 + 8 * s, (chunk / kStreamStages) & 1);
        if (busy) {
            const T* wc = reinterpret_cast<const T*>(ring + (size_t) s * kStreamStageBytes) + jbase;
            const int row0 = tt * rows_per, nrows = min(rows_per, n_loop - row0);
            const T* xin = in + (size_t) row0 * TR;
            if (nc == 2) stream_rows<T, ACC, TR, 2>(wc, xin, part, nrows, ks, ld, acc);
            else stream_rows<T, ACC, TR, 1>(wc, xin, part, nrows, ks, ld, acc);
        }
        __syncwarp();
        if (lane == 0) bar_arrive(empty0 + 8 * s);
        if (++tt == nch) tt = 0;
    }

    if (ks > 1) {
        // partial sums meet in shared memory: scratch[part][r][node], node stride padded so that the epilogue's
        // (r fastest) reads are conflict-free; then every thread finishes whole (node, replica) elements
        ACC* scratch = reinterpret_cast<ACC*>(scratch_raw);
        const int sstride = ld + 32 / TR;
        if (busy) {
#pragma unroll
            for (int q = 0; q < kStreamMaxNC; q++) {
                if (q < nc) {
#pragma unroll
                    for (int r = 0; r < TR; r++) scratch[((size_t) part * TR + r) * sstride + jbase + 32 * q] = acc[q][r];
                }
            }
        }
        compute_sync();
        for (int idx = tid; idx < n_lane * TR; idx += kStreamComputeThreads) {
            const int j = idx / TR, r = idx - j * TR;
            ACC v = scratch[(size_t) r * sstride + j];
            for (int pp = 1; pp < ks; pp++) v += scratch[((size_t) pp * TR + r) * sstride + j];
            epi(j, r, v);
        }
    } else if (busy) {
#pragma unroll
        for (int q = 0; q < kStreamMaxNC; q++) {
            if (q < nc) {
                const int j = jbase + 32 * q;
                if (j < n_lane) {
#pragma unroll
                    for (int r = 0; r < TR; r++) epi(j, r, acc[q][r]);
                }
            }
        }
    }
    compute_sync();
}

}  // namespace

template <typename T, int TR>
__global__ void __launch_bounds__(kStreamThreads, 1) annf_stream_kernel(const StreamArgs args) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = args.L, total = args.total, n0 = args.n0, nL = args.nL, A = args.A;
    const int r0 = blockIdx.x * TR;
    // programmatic dependent launch: the next kernel of the stream may begin its prologue; this one waits for its
    // predecessors (whoever wrote the coordinates) only after its own barriers are set up
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    ANNF_SPHASE(0);
    const int N = args.atoms_per_replica;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                                   // [stages][16 KB]
    unsigned char* ptr = ring + (size_t) kStreamStages * kStreamStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ptr);                                // full[S], empty[S], coords, plan
    ptr += 16 * sizeof(uint64_t);
    auto carve = [&](size_t bytes) { unsigned char* q = ptr; ptr += (bytes + 15) & ~(size_t) 15; return q; };   // 16-byte regions
    // statics: the plan, the biases (node of layer l >= 1 at node_off[l] + j) and the selection, one bulk copy
    const StreamPlan& plan = *reinterpret_cast<const StreamPlan*>(carve(sizeof(StreamPlan)));
    const T* bias_s = reinterpret_cast<const T*>(carve(sizeof(T) * (size_t) total));
    const int* sel_s = reinterpret_cast<const int*>(carve(sizeof(int) * (size_t) A));
    double* zout = reinterpret_cast<double*>(carve(sizeof(double) * nL * TR));     // [nL][TR] output pre-activations
    double* dzout = reinterpret_cast<double*>(carve(sizeof(double) * nL * TR));    // [nL][TR] dE/dz of the output layer
    double* eterm = reinterpret_cast<double*>(carve(sizeof(double) * nL * TR));    // [TR][nL] energy terms
    double* cen = reinterpret_cast<double*>(carve(sizeof(double) * TR * args.num_center));
    double* com = reinterpret_cast<double*>(carve(sizeof(double) * TR * 3));
    double* scratch = reinterpret_cast<double*>(carve(sizeof(double) * args.scratch_elems));
    double* csum = reinterpret_cast<double*>(carve(sizeof(double) * kStreamComputeWarps * 3));   // centroid partial sums
    T* act = reinterpret_cast<T*>(carve(sizeof(T) * (size_t) total * TR));         // [total][TR]
    const uint32_t full0 = s_u32(bars), empty0 = s_u32(bars + kStreamStages), cbar = s_u32(bars + 2 * kStreamStages),
                   pbar = s_u32(bars + 2 * kStreamStages + 1);

    // the tile's coordinates can take the bulk path when the tile is complete and 16-byte aligned in global memory
    const float* tile_coords = args.coords + (size_t) r0 * N * 3;
    const bool coords_staged = args.coord_stage_bytes > 0 && r0 + TR <= args.R && (reinterpret_cast<uintptr_t>(tile_coords) & 15) == 0;

    if (tid == kStreamComputeThreads) {
        for (int s = 0; s < kStreamStages; s++) {
            bar_init(full0 + 8 * s, 1);
            bar_init(empty0 + 8 * s, kStreamComputeWarps);
        }
        bar_init(cbar, 1);
        bar_init(pbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the statics never change after annf_create: fetch them before waiting for the kernels ahead of this one
        bar_expect_tx(pbar, (uint32_t) args.static_bytes);
        bulk_load(s_u32(&plan), args.statics, (uint32_t) args.static_bytes, pbar);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (coords_staged) {
            bar_expect_tx(cbar, (uint32_t) args.coord_stage_bytes);
            bulk_load(s_u32(scratch), tile_coords, (uint32_t) args.coord_stage_bytes, cbar);
        }
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    __syncthreads();

    if (warp == kStreamComputeWarps) {
        // ---- producer warp: one lane walks the weight stream ------------------------------------------------------
        if (lane == 0) {
            bar_wait(pbar, 0);
            const T* stream = reinterpret_cast<const T*>(args.stream);
            int chunk = 0;
            for (int p = 0; p < plan.npass; p++) {
                const T* src = stream + plan.base[p];
                const int rows_per = plan.rows_per[p], n_loop = plan.n_loop[p], ld = plan.ld[p], nch = plan.nchunks[p];
                // All CTAs run in near lock-step and read the SAME weights: if they also read them in the same order, a
                // chunk's few L2 slices serve 100+ SMs at once (measured: ~15 B/clk/SM, a third of the L2 cap).  The sum
                // over rows is order-free, so CTA b starts every pass at chunk b mod nch and wraps around.
                int tt = blockIdx.x % nch;
                for (int t = 0; t < nch; t++, chunk++) {
                    const int s = chunk & (kStreamStages - 1);
                    if (chunk >= kStreamStages) bar_wait(empty0 + 8 * s, ((chunk / kStreamStages) - 1) & 1);
                    const int row0 = tt * rows_per, nrows = min(rows_per, n_loop - row0);
                    if (++tt == nch) tt = 0;
                    const uint32_t bytes = (uint32_t) (sizeof(T) * (size_t) nrows * ld);
                    bar_expect_tx(full0 + 8 * s, bytes);
                    bulk_load(s_u32(ring + (size_t) s * kStreamStageBytes), src + (size_t) row0 * ld, bytes, full0 + 8 * s);
                }
            }
        }
        return;
    }

    // ---- compute warps -------------------------------------------------------------------------------------------
    for (int idx = tid; idx < TR * args.num_center; idx += kStreamComputeThreads) {
        const int r = idx / args.num_center, c = idx - r * args.num_center;
        const int rr = min(r0 + r, args.R - 1);
        cen[idx] = args.centers ? (double) args.centers[(size_t) rr * args.num_center + c] : args.center[c];
    }
    const double kk = *args.k_dev;
    bar_wait(pbar, 0);                      // plan, biases, selection
    ANNF_SPHASE(1);

    // ---- feature stage: ANN_ReferenceKernels.cpp:128-151 ------------------------------------------------------------
    {
        const float* src = tile_coords;
        if (coords_staged) {
            bar_wait(cbar, 0);
            src = reinterpret_cast<const float*>(scratch);
        }
        // centroid: the warps of the CTA are dealt evenly to the replicas of the tile; a lane sums x, y, z of its atoms
        constexpr int kWarpsPerReplica = kStreamComputeWarps / TR;
        {
            const int r = warp / kWarpsPerReplica, part = warp - r * kWarpsPerReplica;
            double sx = 0.0, sy = 0.0, sz = 0.0;
            if (r0 + r < args.R) {
                const float* base = src + (size_t) r * N * 3;
                for (int i = part * 32 + lane; i < A; i += 32 * kWarpsPerReplica) {
                    const float* q = base + (size_t) sel_s[i] * 3;
                    sx += (double) q[0] * args.inv_scale;                              // :139
                    sy += (double) q[1] * args.inv_scale;
                    sz += (double) q[2] * args.inv_scale;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
            }
            if (lane == 0) { csum[warp * 3] = sx; csum[warp * 3 + 1] = sy; csum[warp * 3 + 2] = sz; }
        }
        compute_sync();
        if (tid < TR * 3) {
            const int r = tid / 3, c = tid - 3 * r;
            double t = 0.0;
            for (int w = 0; w < kWarpsPerReplica; w++) t += csum[(r * kWarpsPerReplica + w) * 3 + c];
            com[tid] = t * args.inv_A;                                                 // unweighted centroid :137-141
        }
        compute_sync();
        for (int r = 0; r < TR; r++) {
            const bool live = r0 + r < args.R;
            const float* base = src + (size_t) r * N * 3;
            for (int a = tid; a < A; a += kStreamComputeThreads) {
                const float* q = base + (size_t) sel_s[a] * 3;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double v = live ? (double) q[c] * args.inv_scale - com[r * 3 + c] : 0.0;   // :143-145
                    act[(size_t) (3 * a + c) * TR + r] = (T) v;
                }
            }
        }
        compute_sync();
    }
    ANNF_SPHASE(2);

    // ---- forward: calculate_output_of_each_layer, ANN_ReferenceKernels.cpp:179-260 ---------------------------------
    int chunk = 0;
    for (int l = 0; l + 1 < L; l++) {
        const T* in = act + (size_t) plan.node_off[l] * TR;
        T* out = act + (size_t) plan.node_off[l + 1] * TR;
        const T* b = bias_s + plan.node_off[l + 1];
        const int type = plan.types[l];
        if (l + 2 == L) {
            // output layer: fp64 accumulation, pre-activations kept in fp64 for the head
            stream_pass<T, double, TR>(plan, l, chunk, ring, full0, empty0, in, scratch, [&](int j, int r, double v) {
                zout[(size_t) j * TR + r] = v + (double) b[j];                         // :193-198
            });
        } else {
            stream_pass<T, T, TR>(plan, l, chunk, ring, full0, empty0, in, scratch, [&](int j, int r, T v) {
                const T zz = v + b[j];
                out[(size_t) j * TR + r] = (type == ANNF_TANH) ? act_tanh(zz) : zz;   // :200-209
            });
        }
        ANNF_SPHASE(3 + l);
    }

    // ---- fp64 head: energy + dE/dz of the output layer; one thread per (replica, output or Circular pair) -----------
    for (int idx = tid; idx < TR * nL; idx += kStreamComputeThreads) {
        const int r = idx / nL, t = idx - r * nL;
        eterm[idx] = output_head_parallel(args.out_type, nL, zout + r, TR, cen + (size_t) r * args.num_center, kk, dzout + r, TR, t);
    }
    compute_sync();
    if (tid < TR && r0 + tid < args.R) {
        double e = 0.0;
        for (int t = 0; t < nL; t++) e += eterm[tid * nL + t];
        args.energies[r0 + tid] = e;
    }
    {
        T* dL = act + (size_t) plan.node_off[L - 1] * TR;
        for (int idx = tid; idx < nL * TR; idx += kStreamComputeThreads) dL[idx] = (T) dzout[idx];
    }
    compute_sync();
    ANNF_SPHASE(3 + 2 * (L - 1));

    // ---- backward: back_prop, ANN_ReferenceKernels.cpp:299-373; the last contraction yields forces (:397-410) -------
    for (int l = L - 2; l >= 0; l--) {
        const int p = 2 * (L - 1) - 1 - l;
        const T* dz = act + (size_t) plan.node_off[l + 1] * TR;     // dE/dz of layer l+1 (overwrote its activations)
        T* slot = act + (size_t) plan.node_off[l] * TR;            // holds a_l, receives dE/dz_l (l >= 1) or the forces (l = 0)
        const int type = l >= 1 ? plan.types[l - 1] : ANNF_LINEAR;
        stream_pass<T, T, TR>(plan, p, chunk, ring, full0, empty0, dz, scratch, [&](int m, int r, T v) {
            if (type == ANNF_TANH) {
                const T av = slot[(size_t) m * TR + r];
                slot[(size_t) m * TR + r] = v * (T(1) - av * av);                     // :343-348
            } else {
                slot[(size_t) m * TR + r] = v;                                         // :349-354
            }
        });
        ANNF_SPHASE(3 + p);
    }
    for (int r = 0; r < TR; r++) {
        if (r0 + r >= args.R) break;
        float* out = args.forces + (size_t) (r0 + r) * N * 3;
        for (int i = tid; i < n0; i += kStreamComputeThreads) {
            const int atom = i / 3, c = i - 3 * atom;
            out[(size_t) sel_s[atom] * 3 + c] = (float) act[(size_t) i * TR + r];
        }
    }
    ANNF_SPHASE(4 + 2 * (L - 1));
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------------
cudaError_t read_stream_phase_clocks(long long* out32) {
    return cudaMemcpyFromSymbol(out32, g_stream_phase_clock, sizeof(long long) * 64);
}

template <typename T>
static size_t stream_smem_bytes_t(const NetView& net, const StreamPlan& plan, int tr) {
    const size_t total = net.node_off[net.L], nL = net.nodes[net.L - 1];
    auto up = [](size_t b) { return (b + 15) & ~(size_t) 15; };
    size_t bytes = (size_t) kStreamStages * kStreamStageBytes + 16 * sizeof(uint64_t);
    bytes += up(sizeof(StreamPlan)) + up(sizeof(T) * total) + up(sizeof(int) * (size_t) net.n_sel);        // statics
    bytes += 3 * up(sizeof(double) * nL * tr) + up(sizeof(double) * tr * net.num_center) + up(sizeof(double) * tr * 3) +
             up(sizeof(double) * (size_t) plan.scratch_unit * tr) + up(sizeof(double) * kStreamComputeWarps * 3);
    bytes += up(sizeof(T) * total * tr);
    return bytes + 128;
}

size_t stream_smem_bytes(const NetView& net, const StreamPlan& plan, int tr, bool fp64) {
    return fp64 ? stream_smem_bytes_t<double>(net, plan, tr) : stream_smem_bytes_t<float>(net, plan, tr);
}

// Eligibility + plan + the stream itself (host vectors; the caller uploads them).  `w` are the fp64 weights per
// connection, row-major [n_out][n_in].
template <typename T>
static bool build_stream_t(const NetView& net, const std::vector<std::vector<double>>& w, StreamPlan* plan, std::vector<T>* stream,
                           std::vector<unsigned char>* statics, const std::vector<double>& b64, const std::vector<int>& sel) {
    const int L = net.L;
    if (net.input_type != ANNF_INPUT_CARTESIAN) return false;
    for (int l = 0; l + 2 < L; l++)
        if (net.types[l] != ANNF_TANH && net.types[l] != ANNF_LINEAR) return false;
    if (net.nodes[L - 1] > 256) return false;
    *plan = StreamPlan{};
    plan->L = L;
    plan->total = net.node_off[L];
    plan->npass = 2 * (L - 1);
    for (int l = 0; l < L; l++) { plan->nodes[l] = net.nodes[l]; plan->node_off[l] = net.node_off[l]; }
    plan->node_off[L] = net.node_off[L];
    for (int l = 0; l + 1 < L; l++) plan->types[l] = net.types[l];
    long long off = 0;
    int scratch = 0;
    stream->clear();
    for (int p = 0; p < plan->npass; p++) {
        const bool fwd = p < L - 1;
        const int l = fwd ? p : 2 * (L - 1) - 1 - p;
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        const int n_lane = fwd ? n_out : n_in, n_loop = fwd ? n_in : n_out;
        const int nchunks32 = (n_lane + 31) / 32;
        int nc = (nchunks32 + kStreamComputeWarps - 1) / kStreamComputeWarps;
        if (nc < 2 && nchunks32 >= 2) nc = 2;
        if (nc > kStreamMaxNC) return false;
        const int groups = (nchunks32 + nc - 1) / nc;
        const int ks = std::max(1, kStreamComputeWarps / groups);
        const int ld = groups * nc * 32;
        const int rows_per = kStreamStageBytes / (ld * (int) sizeof(T));
        if (rows_per < 1) return false;
        plan->n_lane[p] = n_lane; plan->n_loop[p] = n_loop; plan->ld[p] = ld; plan->rows_per[p] = rows_per;
        plan->nchunks[p] = (n_loop + rows_per - 1) / rows_per;
        plan->base[p] = off;
        plan->nc[p] = nc; plan->groups[p] = groups; plan->ks[p] = ks;
        if (ks > 1) scratch = std::max(scratch, ks * (ld + 32));
        // matrix of the pass: rows = loop index, columns = lane index, zero padded to ld
        stream->resize((size_t) off + (size_t) n_loop * ld, T(0));
        T* dst = stream->data() + off;
        const std::vector<double>& wl = w[l];
        if (fwd) {
            for (int i = 0; i < n_in; i++)
                for (int j = 0; j < n_out; j++) dst[(size_t) i * ld + j] = (T) wl[(size_t) j * n_in + i];
        } else if (l > 0) {
            for (int k = 0; k < n_out; k++)
                for (int m = 0; m < n_in; m++) dst[(size_t) k * ld + m] = (T) wl[(size_t) k * n_in + m];
        } else {
            // first connection, backward: centring + scaling Jacobian folded in (fp64 on the host)
            const int A = net.n_sel;
            for (int k = 0; k < n_out; k++) {
                double wsum[3] = {0.0, 0.0, 0.0};
                for (int t = 0; t < n_in; t++) wsum[t % 3] += wl[(size_t) k * n_in + t];
                for (int t = 0; t < n_in; t++) dst[(size_t) k * ld + t] = (T) (-(wl[(size_t) k * n_in + t] - wsum[t % 3] / A) / net.scale);
            }
        }
        off += (long long) n_loop * ld;
    }
    plan->scratch_unit = scratch;
    // statics blob, in the kernel's shared-memory layout: plan | biases | selection, each padded to 16 bytes
    auto up = [](size_t b) { return (b + 15) & ~(size_t) 15; };
    const size_t total = net.node_off[L];
    statics->assign(up(sizeof(StreamPlan)) + up(sizeof(T) * total) + up(sizeof(int) * sel.size()), 0);
    memcpy(statics->data(), plan, sizeof(StreamPlan));
    T* bias = reinterpret_cast<T*>(statics->data() + up(sizeof(StreamPlan)));
    for (int l = 0; l + 1 < L; l++)
        for (int j = 0; j < net.nodes[l + 1]; j++) bias[net.node_off[l + 1] + j] = (T) b64[net.b_off[l] + j];
    memcpy(statics->data() + up(sizeof(StreamPlan)) + up(sizeof(T) * total), sel.data(), sizeof(int) * sel.size());
    return true;
}

bool build_stream(const NetView& net, const std::vector<std::vector<double>>& w, const std::vector<double>& b64,
                  const std::vector<int>& sel_atoms, bool fp64, StreamPlan* plan, std::vector<double>* stream64,
                  std::vector<float>* stream32, std::vector<unsigned char>* statics) {
    if (getenv("ANNF_BATCHED_GENERIC")) return false;
    return fp64 ? build_stream_t<double>(net, w, plan, stream64, statics, b64, sel_atoms)
                : build_stream_t<float>(net, w, plan, stream32, statics, b64, sel_atoms);
}

template <typename T, int TR>
static cudaError_t launch_stream_t(const StreamArgs& args, size_t smem, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(annf_stream_kernel<T, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (err != cudaSuccess) return err;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((args.R + TR - 1) / TR);
    cfg.blockDim = dim3(kStreamThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see griddepcontrol in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, annf_stream_kernel<T, TR>, args);
}

// Replicas per CTA: the largest tile that still gives (nearly) every SM a CTA — weights are streamed once per CTA, so
// bigger tiles amortise them, but an idle SM amortises nothing.
int stream_pick_tile(const NetView& net, const StreamPlan& plan, bool fp64, int R, int sm_count) {
    if (const char* env = getenv("ANNF_STREAM_TR")) {
        const int tr = atoi(env);
        if ((tr == 1 || tr == 2 || tr == 4 || tr == 8) && stream_smem_bytes(net, plan, tr, fp64) <= 226 * 1024) return tr;
    }
    int best = 0;
    for (int tr : {8, 4, 2, 1}) {
        if (stream_smem_bytes(net, plan, tr, fp64) > 226 * 1024) continue;
        if (best == 0) best = tr;
        if ((R + tr - 1) / tr >= (3 * sm_count) / 4) return tr;
        best = tr;
    }
    return best;       // fewer replicas than SMs: the smallest tile that fits
}

cudaError_t launch_stream(const NetView& net, const StreamPlan& plan, const void* statics_dev, int static_bytes, const void* stream_dev,
                          bool fp64, int tr, int R, const float* coords, int atoms_per_replica,
                          const float* centers, float* forces, double* energies, cudaStream_t stream) {
    StreamArgs args;
    args.statics = statics_dev;
    args.static_bytes = static_bytes;
    args.stream = stream_dev;
    args.center = net.center;
    args.k_dev = net.k_dev;
    args.inv_scale = 1.0 / net.scale;
    args.inv_A = 1.0 / net.n_sel;
    args.coords = coords;
    args.centers = centers;
    args.forces = forces;
    args.energies = energies;
    args.R = R;
    args.atoms_per_replica = atoms_per_replica;
    args.A = net.n_sel;
    args.num_center = net.num_center;
    args.L = net.L;
    args.total = net.node_off[net.L];
    args.n0 = net.nodes[0];
    args.nL = net.nodes[net.L - 1];
    args.scratch_elems = plan.scratch_unit * tr;
    args.out_type = net.types[net.L - 2];
    const size_t smem = stream_smem_bytes(net, plan, tr, fp64);
    // coordinates share the K-split scratch region (free until the first contraction ends)
    const size_t cbytes = sizeof(float) * 3 * (size_t) atoms_per_replica * tr;
    args.coord_stage_bytes = (cbytes % 16 == 0 && cbytes <= sizeof(double) * (size_t) args.scratch_elems) ? (int) cbytes : 0;
    if (fp64) {
        switch (tr) {
        case 8: return launch_stream_t<double, 8>(args, smem, stream);
        case 4: return launch_stream_t<double, 4>(args, smem, stream);
        case 2: return launch_stream_t<double, 2>(args, smem, stream);
        default: return launch_stream_t<double, 1>(args, smem, stream);
        }
    }
    switch (tr) {
    case 8: return launch_stream_t<float, 8>(args, smem, stream);
    case 4: return launch_stream_t<float, 4>(args, smem, stream);
    case 2: return launch_stream_t<float, 2>(args, smem, stream);
    default: return launch_stream_t<float, 1>(args, smem, stream);
    }
}

}  // namespace annf
