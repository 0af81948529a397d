// fp64 batched-replica kernel on the FP64 tensor path (sm_100a): BASELINE configs 3 and 4 in the reference's own precision.
//
// Why this shape.  A tile of 8 replicas through a layer is D[8 x n] = X^T[8 x K] * M[K x n] in fp64.  Measured on B200, the
// register-tiled FMA version of this kernel (annf_stream.cu, still used for fp32) is bound by shared-memory wavefronts, not
// by the FP64 pipe: every DFMA needs an activation that all 32 lanes read from the same address, and a broadcast load
// still costs a wavefront per 4-8 bytes (fwd0 of 180-100-100-3: 7.2 k LSU cycles against 2.3 k FP64-pipe cycles).
// mma.sync.m8n8k4.f64 does 256 FMAs per warp instruction from ONE 8-byte operand per lane per matrix, i.e. 4 wavefronts per
// 256 FMAs instead of 10: the contraction becomes FP64-pipe bound.  tcgen05 has no fp64 kind; DMMA is the FP64 tensor
// instruction of sm_100.
//
// Structure (shared with annf_stream.cu): one CTA = one tile of 8 replicas end to end; a producer warp walks the handle's
// linear weight stream with cp.async.bulk (TMA engine) into a shared-memory ring guarded by full / empty mbarriers; 16
// compute warps consume.  The stream holds every matrix in consumption order, already in DMMA B-fragment order
// ([k-block of 4][node tile of 8][lane]), so a lane's operand is one conflict-free LDS.64 at base + lane.  Activations
// live in shared memory as [node][8 replicas], which IS the A-fragment order of a k-block (lane -> node 4kb + lane%4,
// replica lane/4).  Warps split a pass by node tiles first and by k-blocks only if tiles run out (no K-split and no
// reduction for the BASELINE shapes' hidden layers).  The first connection's backward matrix is pre-centred and
// pre-scaled (the Jacobian of "divide by s, subtract the centroid", ANN_ReferenceKernels.cpp:137-145, :397-410), so
// the last contraction yields forces.  Statics (plan, biases, selection) arrive in one bulk copy; the launch is
// PDL-enabled.  Arithmetic follows ReferenceCalcANN_ForceKernel::candidate_2 (ANN_ReferenceKernels.cpp:122-171) in fp64.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "annf_common.cuh"
#include "annf_ptx.cuh"

namespace annf {

constexpr int kDWarps = 16;
constexpr int kDComputeThreads = 32 * kDWarps;
constexpr int kDThreads = kDComputeThreads + 32;      // + one producer warp
// the weight ring: ANNF_DMMA_STAGES x ANNF_DMMA_STAGE_KB (compile-time; the product build's values were picked by measurement,
// profiles/r02_dmma_ring.md)
#ifndef ANNF_DMMA_STAGES
#define ANNF_DMMA_STAGES 2
#endif
#ifndef ANNF_DMMA_STAGE_KB
#define ANNF_DMMA_STAGE_KB 64
#endif
constexpr int kDStages = ANNF_DMMA_STAGES;
constexpr int kDStageBytes = ANNF_DMMA_STAGE_KB * 1024;
constexpr int kDBarSlots = (2 * kDStages + 2 + 15) / 16 * 16;   // full[S], empty[S], coords, statics (16-byte multiples)
constexpr int kDTile = 8;                             // replicas per CTA = rows of the DMMA A operand
constexpr int kDMaxNT = 4;                            // node tiles (of 8) per warp

struct DmmaArgs {
    const void* statics;             // device blob: StreamPlan | double bias[padded nodes] | int sel_atoms[A], 16-byte padded each
    int static_bytes;
    const double* stream;            // the weight stream, B-fragment order
    const double* center;
    const double* k_dev;
    double inv_scale, inv_A;
    const float* coords;
    const float* centers;
    float* forces;
    double* energies;
    int R, atoms_per_replica, A, num_center;
    int L, total, n0, nL;            // total: node slots including the padding of every layer to a multiple of 4
    int scratch_elems;               // K-split scratch, in doubles
    int coord_stage_bytes;           // > 0: the tile's coordinates are bulk-copied into shared memory (into the ring's last stage)
    int out_type;
};

// Diagnostic: clock64() of thread 0 / CTA 0 at the phase boundaries of the last launch (annf_debug_stream_phases).
__device__ long long g_dmma_phase_clock[64];
#ifdef ANNF_PHASE_CLOCKS        // diagnostic build only (make EXTRA=-DANNF_PHASE_CLOCKS): the product kernels carry no instrumentation
#define ANNF_DPHASE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_dmma_phase_clock[i] = clock64(); } while (0)
#else
#define ANNF_DPHASE(i) do { } while (0)
#endif

namespace {

using namespace ptx;

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kDComputeThreads) : "memory"); }

// D[8x8] += A[8x4] * B[4x8], fp64: lane holds A[lane/4][lane%4], B[lane%4][lane/4], D[lane/4][2(lane%4) + {0,1}]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// k-blocks part, part + ks, ... of one weight chunk for NT node tiles; two accumulator sets alternate so that consecutive
// DMMAs of a tile never depend on each other.  wb: this warp's first tile of the chunk + lane; ain: A operand of the
// chunk's first k-block for this lane; tile_stride = doubles per k-block in the stream.
template <int NT>
__device__ __forceinline__ void dmma_chunk(const double* wb, const double* ain, int part, int nkb, int ks, int tile_stride,
                                           double (&c)[2][kDMaxNT][2]) {
    for (int kb = part; kb < nkb; kb += 2 * ks) {
        const bool two = kb + ks < nkb;
        const double a0 = ain[kb * 32];
        const double a1 = two ? ain[(kb + ks) * 32] : 0.0;
        const double* w0 = wb + kb * tile_stride;
        const double* w1 = w0 + ks * tile_stride;
        double b0[NT], b1[NT];
#pragma unroll
        for (int q = 0; q < NT; q++) { b0[q] = w0[q * 32]; b1[q] = two ? w1[q * 32] : 0.0; }
#pragma unroll
        for (int q = 0; q < NT; q++) { dmma(c[0][q], a0, b0[q]); dmma(c[1][q], a1, b1[q]); }
    }
}

// What a contraction does with its results.
enum PassMode { kFwdHidden = 0, kFwdOutput = 1, kBwd = 2 };

struct PassCtx {
    const StreamPlan* plan;
    const unsigned char* ring;
    uint32_t full0, empty0;
    int stage;
    uint32_t phase;
    double* scratch;
    double* act;                    // [padded nodes][8]
    const double* bias;             // by padded node offset
    double* zout;                   // [nL][8]
};

// One contraction out[j][r] = sum_i M[i][j] * in[i][r] on the FP64 tensor path, M streamed through the ring.
// The epilogue is selected at run time so that the kernel holds two copies of this code (forward loop, backward loop)
// instead of one per pass: six copies with their unrolled tile loops overflowed the instruction cache (stall_no_inst was
// the second largest stall reason).  Executed by the compute warps only.
//   kFwdHidden: act[out_off + j] = f(sum + bias)         (ANN_ReferenceKernels.cpp:193-209)
//   kFwdOutput: zout[j] = sum + bias                      (kept in fp64 for the head)
//   kBwd:       act[out_off + m] = sum * f'(a_m)          (:343-354; for connection 0 the sum already IS the force)
__device__ __forceinline__ void dmma_pass(PassCtx& ctx, int p, int mode, int type, int in_off, int out_off) {
    const StreamPlan& plan = *ctx.plan;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = plan.nc[p], groups = plan.groups[p], ks = plan.ks[p];
    const int n_lane = plan.n_lane[p], nkb_total = (plan.n_loop[p] + 3) >> 2, kb_per = plan.rows_per[p], nch = plan.nchunks[p];
    const int tile_stride = plan.ld[p];
    const bool busy = warp < groups * ks;
    const int grp = warp / ks, part = warp - grp * ks;
    const int tile0 = grp * NT;
    const double* in = ctx.act + (size_t) in_off * kDTile;
    double c[2][kDMaxNT][2];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int q = 0; q < kDMaxNT; q++) c[h][q][0] = c[h][q][1] = 0.0;

    // every CTA walks the chunks of a pass from its own starting point (the sum over k is order-free)
    int tt = blockIdx.x % nch;
    int stage = ctx.stage;
    uint32_t phase = ctx.phase;
    for (int t = 0; t < nch; t++) {
        bar_wait(ctx.full0 + 8 * stage, phase);
        if (busy) {
            const int kb0 = tt * kb_per, nkb = min(kb_per, nkb_total - kb0);
            const double* wb = reinterpret_cast<const double*>(ctx.ring + (size_t) stage * kDStageBytes) + tile0 * 32 + lane;
            const double* ain = in + (4 * kb0 + (lane & 3)) * kDTile + (lane >> 2);
            if (NT == 1) dmma_chunk<1>(wb, ain, part, nkb, ks, tile_stride, c);
            else if (NT == 2) dmma_chunk<2>(wb, ain, part, nkb, ks, tile_stride, c);
            else dmma_chunk<4>(wb, ain, part, nkb, ks, tile_stride, c);
        }
        __syncwarp();
        if (lane == 0) bar_arrive(ctx.empty0 + 8 * stage);
        if (++stage == kDStages) { stage = 0; phase ^= 1; }
        if (++tt == nch) tt = 0;
    }
    ctx.stage = stage;
    ctx.phase = phase;

    double* out = ctx.act + (size_t) out_off * kDTile;
    const double* b = ctx.bias + out_off;
    auto epi = [&](int j, int r, double v) {
        if (mode == kFwdHidden) {
            const double zz = v + b[j];
            out[j * kDTile + r] = (type == ANNF_TANH) ? annf_tanh(zz) : zz;
        } else if (mode == kFwdOutput) {
            ctx.zout[j * kDTile + r] = v + b[j];
        } else if (type == ANNF_TANH) {
            const double av = out[j * kDTile + r];
            out[j * kDTile + r] = v * (1.0 - av * av);
        } else {
            out[j * kDTile + r] = v;
        }
    };
    const int r = lane >> 2, jl = 2 * (lane & 3);
    if (ks > 1) {
        // partial sums meet in shared memory, scratch[part][r][node]; then every thread finishes whole elements
        double* scratch = ctx.scratch;
        const int sstride = tile_stride / 4 + 4;
        if (busy) {
#pragma unroll
            for (int q = 0; q < kDMaxNT; q++) {
                if (q < NT) {
                    double* dst = scratch + (part * kDTile + r) * sstride + (tile0 + q) * 8 + jl;
                    dst[0] = c[0][q][0] + c[1][q][0];
                    dst[1] = c[0][q][1] + c[1][q][1];
                }
            }
        }
        compute_sync();
        for (int idx = tid; idx < n_lane * kDTile; idx += kDComputeThreads) {
            const int j = idx >> 3, rr = idx & 7;
            double v = scratch[rr * sstride + j];
            for (int pp = 1; pp < ks; pp++) v += scratch[(pp * kDTile + rr) * sstride + j];
            epi(j, rr, v);
        }
    } else if (busy) {
#pragma unroll
        for (int q = 0; q < kDMaxNT; q++) {
            if (q < NT) {
                const int j = (tile0 + q) * 8 + jl;
                const double v0 = c[0][q][0] + c[1][q][0], v1 = c[0][q][1] + c[1][q][1];
                if (mode == kFwdHidden && type == ANNF_TANH) {
                    // the lane's two tanh chains (~25 dependent fp64 instructions each) side by side, stores predicated;
                    // bias rows are padded, so b[j + 1] is readable
                    const double t0 = annf_tanh(v0 + b[j]), t1 = annf_tanh(v1 + b[j + 1]);
                    if (j < n_lane) out[j * kDTile + r] = t0;
                    if (j + 1 < n_lane) out[(j + 1) * kDTile + r] = t1;
                } else {
                    if (j < n_lane) epi(j, r, v0);
                    if (j + 1 < n_lane) epi(j + 1, r, v1);
                }
            }
        }
    }
    compute_sync();
}

}  // namespace

__global__ void __launch_bounds__(kDThreads, 1) annf_dmma_kernel(const DmmaArgs args) {
    constexpr int TR = kDTile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = args.L, total = args.total, n0 = args.n0, nL = args.nL, A = args.A;
    const int r0 = blockIdx.x * TR;
    const int N = args.atoms_per_replica;
    // programmatic dependent launch: the next kernel of the stream may begin its prologue; this one waits for its
    // predecessors (whoever wrote the coordinates) only after its barriers are set up and its statics are on their way
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    ANNF_DPHASE(0);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                                   // [stages][32 KB]
    unsigned char* ptr = ring + (size_t) kDStages * kDStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ptr);                                // full[S], empty[S], coords, statics
    ptr += kDBarSlots * sizeof(uint64_t);
    auto carve = [&](size_t bytes) { unsigned char* q = ptr; ptr += (bytes + 15) & ~(size_t) 15; return q; };   // 16-byte regions
    const StreamPlan& plan = *reinterpret_cast<const StreamPlan*>(carve(sizeof(StreamPlan)));
    const double* bias_s = reinterpret_cast<const double*>(carve(sizeof(double) * (size_t) total));
    const int* sel_s = reinterpret_cast<const int*>(carve(sizeof(int) * (size_t) A));
    double* zout = reinterpret_cast<double*>(carve(sizeof(double) * nL * TR));     // [nL][TR] output pre-activations
    double* dzout = reinterpret_cast<double*>(carve(sizeof(double) * nL * TR));    // [nL][TR] dE/dz of the output layer
    double* eterm = reinterpret_cast<double*>(carve(sizeof(double) * nL * TR));    // [TR][nL] energy terms
    double* cen = reinterpret_cast<double*>(carve(sizeof(double) * TR * args.num_center));
    double* com = reinterpret_cast<double*>(carve(sizeof(double) * TR * 3));
    double* csum = reinterpret_cast<double*>(carve(sizeof(double) * kDWarps * 3));
    double* scratch = reinterpret_cast<double*>(carve(sizeof(double) * args.scratch_elems));
    double* act = reinterpret_cast<double*>(carve(sizeof(double) * (size_t) total * TR));   // [padded nodes][TR]
    const uint32_t full0 = s_u32(bars), empty0 = s_u32(bars + kDStages), cbar = s_u32(bars + 2 * kDStages), pbar = s_u32(bars + 2 * kDStages + 1);

    // the tile's coordinates can take the bulk path when the tile is complete and 16-byte aligned in global memory;
    // they land in the ring's LAST stage, which the weight stream does not reach before the feature stage is over
    const float* tile_coords = args.coords + (size_t) r0 * N * 3;
    const bool coords_staged = args.coord_stage_bytes > 0 && r0 + TR <= args.R && (reinterpret_cast<uintptr_t>(tile_coords) & 15) == 0;
    unsigned char* coord_stage = ring + (size_t) (kDStages - 1) * kDStageBytes;

    if (tid == kDComputeThreads) {
        for (int s = 0; s < kDStages; s++) {
            bar_init(full0 + 8 * s, 1);
            bar_init(empty0 + 8 * s, kDWarps);
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
            bulk_load(s_u32(coord_stage), tile_coords, (uint32_t) args.coord_stage_bytes, cbar);
        }
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    __syncthreads();

    if (warp == kDWarps) {
        // ---- producer warp: one lane walks the weight stream ------------------------------------------------------
        if (lane == 0) {
            bar_wait(pbar, 0);
            int stage = 0, chunk = 0;
            uint32_t phase = 0;
            for (int p = 0; p < plan.npass; p++) {
                const double* src = args.stream + plan.base[p];
                const int kb_per = plan.rows_per[p], nkb_total = (plan.n_loop[p] + 3) >> 2, tile_stride = plan.ld[p], nch = plan.nchunks[p];
                int tt = blockIdx.x % nch;
                for (int t = 0; t < nch; t++, chunk++) {
                    // the last stage first carries the coordinates: one extra phase of its empty barrier (arrived after
                    // the feature stage), so its parity runs one phase ahead of the other stages'
                    const bool last = stage == kDStages - 1;
                    if (chunk >= kDStages) bar_wait(empty0 + 8 * stage, (phase ^ 1u) ^ ((last && coords_staged) ? 1u : 0u));
                    else if (last && coords_staged) bar_wait(empty0 + 8 * stage, 0);
                    const int kb0 = tt * kb_per, nkb = min(kb_per, nkb_total - kb0);
                    const uint32_t bytes = (uint32_t) (sizeof(double) * (size_t) nkb * tile_stride);
                    bar_expect_tx(full0 + 8 * stage, bytes);
                    bulk_load(s_u32(ring + (size_t) stage * kDStageBytes), src + (size_t) kb0 * tile_stride, bytes, full0 + 8 * stage);
                    if (++stage == kDStages) { stage = 0; phase ^= 1; }
                    if (++tt == nch) tt = 0;
                }
            }
        }
        return;
    }

    // ---- compute warps -------------------------------------------------------------------------------------------
    for (int i = tid; i < total * TR; i += kDComputeThreads) act[i] = 0.0;    // layer padding rows must read as zero
    for (int idx = tid; idx < TR * args.num_center; idx += kDComputeThreads) {
        const int r = idx / args.num_center, c = idx - r * args.num_center;
        const int rr = min(r0 + r, args.R - 1);
        cen[idx] = args.centers ? (double) args.centers[(size_t) rr * args.num_center + c] : args.center[c];
    }
    const double kk = *args.k_dev;
    bar_wait(pbar, 0);                      // plan, biases, selection
    compute_sync();                         // act is zeroed before anyone writes features
    ANNF_DPHASE(1);

    // ---- feature stage: ANN_ReferenceKernels.cpp:128-151 ------------------------------------------------------------
    {
        const float* src = tile_coords;
        if (coords_staged) {
            bar_wait(cbar, 0);
            src = reinterpret_cast<const float*>(coord_stage);
        }
        // centroid: two warps per replica; a lane sums x, y, z of its atoms
        constexpr int kWarpsPerReplica = kDWarps / TR;
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
        for (int idx = tid; idx < TR * A; idx += kDComputeThreads) {
            const int r = idx / A, a = idx - r * A;
            if (r0 + r >= args.R) continue;                                            // dead replicas stay zero
            const float* q = src + ((size_t) r * N + sel_s[a]) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++)
                act[(size_t) (3 * a + c) * TR + r] = (double) q[c] * args.inv_scale - com[r * 3 + c];   // :143-145
        }
        compute_sync();
        if (coords_staged && lane == 0) bar_arrive(empty0 + 8 * (kDStages - 1));      // the coordinate stage joins the ring
    }
    ANNF_DPHASE(2);

    // ---- forward: calculate_output_of_each_layer, ANN_ReferenceKernels.cpp:179-260 ---------------------------------
    PassCtx ctx;
    ctx.plan = &plan; ctx.ring = ring; ctx.full0 = full0; ctx.empty0 = empty0; ctx.stage = 0; ctx.phase = 0;
    ctx.scratch = scratch; ctx.act = act; ctx.bias = bias_s; ctx.zout = zout;
    for (int l = 0; l + 1 < L; l++) {
        dmma_pass(ctx, l, l + 2 == L ? kFwdOutput : kFwdHidden, plan.types[l], plan.node_off[l], plan.node_off[l + 1]);
        ANNF_DPHASE(3 + l);
    }

    // ---- fp64 head: energy + dE/dz of the output layer; one thread per (replica, output or Circular pair) -----------
    for (int idx = tid; idx < TR * nL; idx += kDComputeThreads) {
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
        double* dL = act + (size_t) plan.node_off[L - 1] * TR;
        for (int idx = tid; idx < nL * TR; idx += kDComputeThreads) dL[idx] = (r0 + (idx & (TR - 1)) < args.R) ? dzout[idx] : 0.0;
    }
    compute_sync();
    ANNF_DPHASE(3 + 2 * (L - 1));

    // ---- backward: back_prop, ANN_ReferenceKernels.cpp:299-373; the last contraction yields forces (:397-410) -------
    for (int l = L - 2; l >= 0; l--) {
        const int p = 2 * (L - 1) - 1 - l;
        // input: dE/dz of layer l+1 (overwrote its activations); output slot holds a_l, receives dE/dz_l or the forces
        dmma_pass(ctx, p, kBwd, l >= 1 ? plan.types[l - 1] : ANNF_LINEAR, plan.node_off[l + 1], plan.node_off[l]);
        ANNF_DPHASE(3 + p);
    }
    {
        const int r = warp & (TR - 1);                    // two warps per replica
        if (r0 + r < args.R) {
            float* out = args.forces + (size_t) (r0 + r) * N * 3;
            for (int i = (warp >> 3) * 32 + lane; i < n0; i += 32 * (kDWarps / TR)) {
                const int atom = i / 3, c = i - 3 * atom;
                out[(size_t) sel_s[atom] * 3 + c] = (float) act[(size_t) i * TR + r];
            }
        }
    }
    ANNF_DPHASE(4 + 2 * (L - 1));
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------------
cudaError_t read_dmma_phase_clocks(long long* out64) {
    return cudaMemcpyFromSymbol(out64, g_dmma_phase_clock, sizeof(long long) * 64);
}

size_t dmma_smem_bytes(const NetView& net, const StreamPlan& plan) {
    auto up = [](size_t b) { return (b + 15) & ~(size_t) 15; };
    const size_t total = plan.total, nL = net.nodes[net.L - 1], tr = kDTile;
    size_t bytes = (size_t) kDStages * kDStageBytes + kDBarSlots * sizeof(uint64_t);
    bytes += up(sizeof(StreamPlan)) + up(sizeof(double) * total) + up(sizeof(int) * (size_t) net.n_sel);        // statics
    bytes += 3 * up(sizeof(double) * nL * tr) + up(sizeof(double) * tr * net.num_center) + up(sizeof(double) * tr * 3) +
             up(sizeof(double) * kDWarps * 3) + up(sizeof(double) * (size_t) plan.scratch_unit);
    bytes += up(sizeof(double) * total * tr);
    return bytes + 128;
}

// Eligibility + plan + the weight stream in B-fragment order + the statics blob (host vectors; the caller uploads them).
// `w` are the fp64 weights per connection, row-major [n_out][n_in].
bool build_dmma_stream(const NetView& net, const std::vector<std::vector<double>>& w, const std::vector<double>& b64,
                       const std::vector<int>& sel, StreamPlan* plan, std::vector<double>* stream, std::vector<unsigned char>* statics) {
    const int L = net.L;
    if (getenv("ANNF_BATCHED_GENERIC") || getenv("ANNF_BATCHED_SIMT") || getenv("ANNF_STREAM_TR")) return false;   // experiment switches
    if (net.input_type != ANNF_INPUT_CARTESIAN) return false;
    for (int l = 0; l + 2 < L; l++)
        if (net.types[l] != ANNF_TANH && net.types[l] != ANNF_LINEAR) return false;
    if (net.nodes[L - 1] > 256) return false;
    *plan = StreamPlan{};
    plan->L = L;
    plan->npass = 2 * (L - 1);
    plan->node_off[0] = 0;
    for (int l = 0; l < L; l++) {
        plan->nodes[l] = net.nodes[l];
        plan->node_off[l + 1] = plan->node_off[l] + (net.nodes[l] + 3) / 4 * 4;     // every layer padded to whole k-blocks
    }
    plan->total = plan->node_off[L];
    for (int l = 0; l + 1 < L; l++) plan->types[l] = net.types[l];
    long long off = 0;
    int scratch = 0;
    stream->clear();
    for (int p = 0; p < plan->npass; p++) {
        const bool fwd = p < L - 1;
        const int l = fwd ? p : 2 * (L - 1) - 1 - p;
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        const int n_lane = fwd ? n_out : n_in, n_loop = fwd ? n_in : n_out;
        const int ntiles = (n_lane + 7) / 8, nkb = (n_loop + 3) / 4;
        int NT = 1;
        while (NT * kDWarps < ntiles) NT *= 2;                                     // 1, 2 or 4 tiles per warp
        if (NT > kDMaxNT) return false;
        const int groups = (ntiles + NT - 1) / NT;
        const int ks = std::max(1, std::min(kDWarps / groups, nkb));
        const int tile_stride = groups * NT * 32;                                   // doubles per k-block (zero tiles pad the last group)
        const int kb_per = kDStageBytes / (tile_stride * (int) sizeof(double));
        if (kb_per < 1) return false;
        plan->n_lane[p] = n_lane; plan->n_loop[p] = n_loop; plan->ld[p] = tile_stride; plan->rows_per[p] = kb_per;
        plan->nchunks[p] = (nkb + kb_per - 1) / kb_per;
        plan->base[p] = off;
        plan->nc[p] = NT; plan->groups[p] = groups; plan->ks[p] = ks;
        if (ks > 1) scratch = std::max(scratch, ks * kDTile * (groups * NT * 8 + 4));
        // M[k][n], k = contraction index, n = output of the pass
        const std::vector<double>& wl = w[l];
        std::vector<double> w0c;
        if (!fwd && l == 0) {
            // first connection, backward: centring + scaling Jacobian folded in (fp64 on the host)
            const int A = net.n_sel;
            w0c.resize((size_t) n_out * n_in);
            for (int k = 0; k < n_out; k++) {
                double wsum[3] = {0.0, 0.0, 0.0};
                for (int t = 0; t < n_in; t++) wsum[t % 3] += wl[(size_t) k * n_in + t];
                for (int t = 0; t < n_in; t++) w0c[(size_t) k * n_in + t] = -(wl[(size_t) k * n_in + t] - wsum[t % 3] / A) / net.scale;
            }
        }
        auto M = [&](int k, int n) -> double {
            if (k >= n_loop || n >= n_lane) return 0.0;
            if (fwd) return wl[(size_t) n * n_in + k];                               // W_l^T
            return l == 0 ? w0c[(size_t) k * n_in + n] : wl[(size_t) k * n_in + n];   // W_l
        };
        stream->resize((size_t) off + (size_t) nkb * tile_stride, 0.0);
        double* dst = stream->data() + off;
        for (int kb = 0; kb < nkb; kb++)
            for (int tile = 0; tile < groups * NT; tile++)
                for (int lane = 0; lane < 32; lane++)
                    dst[(size_t) kb * tile_stride + tile * 32 + lane] = M(4 * kb + (lane & 3), 8 * tile + (lane >> 2));
        off += (long long) nkb * tile_stride;
    }
    plan->scratch_unit = scratch;
    auto up = [](size_t b) { return (b + 15) & ~(size_t) 15; };
    const size_t total = plan->total;
    statics->assign(up(sizeof(StreamPlan)) + up(sizeof(double) * total) + up(sizeof(int) * sel.size()), 0);
    memcpy(statics->data(), plan, sizeof(StreamPlan));
    double* bias = reinterpret_cast<double*>(statics->data() + up(sizeof(StreamPlan)));
    for (int l = 0; l + 1 < L; l++)
        for (int j = 0; j < net.nodes[l + 1]; j++) bias[plan->node_off[l + 1] + j] = b64[net.b_off[l] + j];
    memcpy(statics->data() + up(sizeof(StreamPlan)) + up(sizeof(double) * total), sel.data(), sizeof(int) * sel.size());
    return dmma_smem_bytes(net, *plan) <= 226 * 1024;
}

cudaError_t launch_dmma(const NetView& net, const StreamPlan& plan, const void* statics_dev, int static_bytes, const double* stream_dev,
                        int R, const float* coords, int atoms_per_replica, const float* centers, float* forces, double* energies,
                        cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(annf_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (err != cudaSuccess) return err;
        configured[dev] = true;
    }
    DmmaArgs args;
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
    args.total = plan.total;
    args.n0 = net.nodes[0];
    args.nL = net.nodes[net.L - 1];
    args.scratch_elems = plan.scratch_unit;
    args.out_type = net.types[net.L - 2];
    const size_t cbytes = sizeof(float) * 3 * (size_t) atoms_per_replica * kDTile;
    args.coord_stage_bytes = (cbytes % 16 == 0 && cbytes <= (size_t) kDStageBytes) ? (int) cbytes : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((R + kDTile - 1) / kDTile);
    cfg.blockDim = dim3(kDThreads);
    cfg.dynamicSmemBytes = dmma_smem_bytes(net, plan);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see griddepcontrol in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, annf_dmma_kernel, args);
}

}  // namespace annf
