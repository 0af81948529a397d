// Layer-wise tensor-core path for wide networks (sm_100a): tcgen05.mma on operand PAIRS (kind::f16 on row-scaled fp16 pairs by
// default, kind::tf32 on TF32 pairs selectable), TMA-fed shared-memory pipeline, fp32 accumulators in TMEM, fused epilogues.
//
// For a batch of R replicas the hidden layers are real dense contractions
//     forward   X_{l+1} = act(X_l  * W_l^T + b_l)        M = R, N = n_{l+1}, K = n_l
//     backward  D_l     = (D_{l+1} * W_l) (.) act'(X_l)   M = R, N = n_l,     K = n_{l+1}
// (ANN_ReferenceKernels.cpp:193-209 and :340-354 with the replica index as the GEMM M dimension).
// Every fp32 operand v is stored as a pair  hi = rn(v), lo = rn(v - hi)  (TF32, or fp16 of the value times a per-row power
// of two: KindOf, ExpChain) so that
//     v*w ~= hi_v*hi_w + hi_v*lo_w + lo_v*hi_w      (three tcgen05.mma per K-step, error ~2^-22)
// Weights are split once at plan creation (both W and W^T, K-major for the tensor core); activations
// are split by the epilogue that produces them.  The tiny output layer, the umbrella energy and its
// gradient stay fp64 in a SIMT "head" kernel; gather/centre ("prep": prep_bulk_kernel for a contiguous selection — one bulk
// copy per replica row — prep_kernel otherwise) is a bandwidth-bound SIMT kernel;
// the Jacobian of "divide by s, subtract the centroid" is folded into the backward copy of the first connection at plan
// creation, so the last backward contraction's accumulator is the force and its epilogue writes it straight out.
//
// GEMM kernel anatomy (one 128 x BN output tile per CTA, BN = 128 or 256):
//   epilogue warps (BN/32 of them, first): tcgen05.ld TMEM -> registers -> bias / tanh / tanh' -> hi/lo or force stores
//   producer warp : TMA - cp.async.bulk.tensor.2d of A_hi, A_lo, B_hi, B_lo (128B swizzle) per stage
//   MMA warp      : TMEM allocator + single-thread tcgen05.mma issuer, tcgen05.commit -> mbarriers
// (BN = 256: setmaxnreg moves registers from the producer/MMA warpgroup to the eight epilogue warps.)
// Every kernel of the step is launched with programmatic dependent launch: its prologue (barrier init,
// TMEM allocation, descriptor prefetch) overlaps the tail of the kernel before it.
// Accumulation: the tensor core's fp32 accumulator does not round to nearest, so a K = 4500 chain in
// TMEM costs ~30x the error of the products themselves (measured).  The K loop is therefore cut into
// chunks of kChunk K-blocks: a chunk accumulates in one of two TMEM buffers while the epilogue warps
// drain the other one into fp32 REGISTER accumulators with round-to-nearest adds ("promotion").
// Split-K: S CTAs of one cluster (cluster dims 1x1xS) each take a K-range, park their partial tile in
// their own shared memory, and after one cluster barrier every CTA reduces a 128/S-row slice over
// all peers through DSMEM in a fixed order (deterministic), then runs the epilogue on it.
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "annf_common.cuh"
#include "annf_profiler.h"
#include "annf_ptx.cuh"

namespace cg = cooperative_groups;

namespace annf {

constexpr int BM = 128;               // replicas per tile = UMMA M
constexpr int kBlockBytes = 128;      // bytes of one operand row in a K-block = one swizzle span
constexpr int kStepBytes = 32;        // bytes along K consumed by one MMA instruction (8 tf32 or 16 f16 values)
constexpr int kStepsPerBlock = kBlockBytes / kStepBytes;

// Operand kind of the contraction.  Both split an fp32 value v into a pair hi + lo with 11 significant bits each, and a
// product costs three MMAs (lo*hi, hi*lo, hi*hi); they differ in the container:
//   KIND_TF32  hi, lo are fp32 words rounded to TF32 (cvt.rna); tcgen05 kind::tf32, 8 values per 32-byte K-step
//   KIND_F16   hi, lo are IEEE halves of v * 2^e, e a per-row power of two that puts the row's largest magnitude just below
//              2^15 (fp16 has TF32's 10-bit mantissa but a 5-bit exponent: the scale buys the range back); tcgen05 kind::f16,
//              16 values per K-step: HALF the operand bytes through L2 / shared memory and TWICE the MMA rate of kind::tf32.
//              The epilogue multiplies the fp32 accumulator by 2^-(e_row + e_col), an exact operation.
enum { KIND_TF32 = 0, KIND_F16 = 1 };
template <int KIND> struct KindOf;
template <> struct KindOf<KIND_TF32> { typedef float elem; static constexpr int kBlockElems = 32; };
template <> struct KindOf<KIND_F16> { typedef __half elem; static constexpr int kBlockElems = 64; };
constexpr int kExpTarget = 14;        // KIND_F16: a row is scaled so that its magnitude bound lies in [2^13, 2^14]
constexpr int kExpClamp = 60;         // |e| <= 60, so 2^-(e_row + e_col) is always a normal fp32 number

constexpr int kChunkDefault = 2;      // K-blocks accumulated in TMEM between promotions (24 chained MMAs); ANNF_TENSOR_CHUNK overrides

enum { EPI_STORE = 0, EPI_BIAS_TANH = 1, EPI_BIAS_LINEAR = 2, EPI_TANH_GRAD = 3, EPI_LINEAR_GRAD = 4, EPI_FORCE = 5 };

struct GemmParams {
    int M, N, K;
    int epi;
    const float* bias;                // [N]               (EPI_BIAS_*)
    const void* aux_hi;               // [M][ld_aux] activations of the layer being differentiated (EPI_TANH_GRAD); elem of the kind
    const void* aux_lo;
    int ld_aux;
    void* out_hi;                     // [M][ld_out] elem of the kind (EPI_STORE: the plain fp32 result; EPI_FORCE: fp32 forces [M][atoms][3])
    void* out_lo;
    int ld_out;                       //              (EPI_FORCE: 3 * atoms_per_replica)
    // KIND_F16 only: binary exponents of the power-of-two scales.  Stored operand = value * 2^exp.
    const int* a_exp;                 // [M] rows of A (activations / gradients going in)
    const int* b_exp;                 // [N] rows of B (weights)
    const int* out_exp;               // [M] rows of the result (activations / gradients coming out); unused by EPI_FORCE / EPI_STORE
    const int* aux_exp;               // [M] rows of aux
    const int* sel_atoms;             // selected atom -> atom index, or nullptr when the selection is 0..A-1 (EPI_FORCE)
    // Split-K THROUGH GLOBAL MEMORY (S = 1 kernels only): gridDim.z CTAs (or CTA pairs) share an output tile, each over its own
    // K range; partial tiles meet in `gs_scratch`, CTAs meet on the counters, every one finishes 1/gridDim.z of the rows.
    int gsplit;                       // number of K ranges (0 / 1: off)
    float* gs_scratch;                // [tiles][gsplit][BM][BN] fp32 partial tiles
    int* gs_arrive;                   // [tiles] partial tiles published so far (self-resetting)
    int* gs_done;                     // [tiles] CTAs that have finished reading (the last one resets both counters)
    int no_b_prefetch;                // experiments (ANNF_TENSOR_B_PREFETCH=0): request the weights only after the dependency wait
    int direct_store;                 // EPI_FORCE, unsplit, identity selection, 16-byte aligned rows: the tile leaves through tm_out (TMA store)
    int dbg;                          // diagnostic builds: slot of g_tensor_clock this launch records into (-1: none)
    int chunk;                        // K-blocks accumulated in TMEM between two promotions into the fp32 register accumulators
    int dry;                          // experiments (ANNF_TENSOR_DRY=1): return immediately, to measure the launch floor
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// L2 eviction priorities (compile-time: -DANNF_L2_HINTS=mask, default 15).  A step streams 18 MB of coordinates in and 18 MB of
// forces out through an L2 that also has to keep 18 MB of weights and the 18 MB operand pair prep rewrites in place every step;
// with default priorities the streams push those out (steady state, ncu --cache-control none: 13 MB of weights re-read from
// DRAM and 15 MB of dead operand rows written back per step; profiles/r02_l2_hints.md).
//   bit 0  coordinates in: evict_first          bit 2  operand pairs (plan buffers, rewritten every step): evict_last
//   bit 1  forces out: evict_first              bit 3  weights: evict_last
// The unhinted instruction is used where a bit is off: a created evict_normal policy is NOT the default (hinting every access
// "normal" cost the C5 step 4 us), and a run-time choice between the two forms cost 1.6 us (branches in prep's store loop).
#ifndef ANNF_L2_HINTS
#define ANNF_L2_HINTS 15
#endif
constexpr int kL2Hints = ANNF_L2_HINTS;
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <bool HINT>
__device__ __forceinline__ void bulk_load_h(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    if (HINT)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
template <bool HINT>
__device__ __forceinline__ void st_v2_h(uint2* ptr, uint2 v, uint64_t pol) {
    if (HINT) asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(ptr), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
    else *ptr = v;
}
template <bool HINT>
__device__ __forceinline__ void st_v4_h(float4* ptr, float4 v, uint64_t pol) {
    if (HINT) asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
    else *ptr = v;
}
template <bool HINT>
__device__ __forceinline__ void tma_load_2d_h(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t pol) {
    if (HINT)
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(pol) : "memory");
    else tma_load_2d(dst, map, bar, c0, c1);
}
// shared -> global tile store by the TMA engine (bulk async-group completion)
template <bool HINT>
__device__ __forceinline__ void tma_store_2d_h(const CUtensorMap* map, uint32_t src, int c0, int c1, uint64_t pol) {
    if (HINT)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                     ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, N = BN, K = 32 bytes (8 tf32 or 16 f16 values), issued by ONE thread
template <int KIND>
__device__ __forceinline__ void umma_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO<<16 | SBO<<32 | version=1<<46 | SWIZZLE_128B=2<<61)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    return (uint64_t) ((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a / b format at bits 7 / 10, both K-major, N>>3 at bit 17, M>>4 at bit 24

__device__ __forceinline__ float tf32_round(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = tf32_round(v);
    lo = tf32_round(v - hi);
}

// CG = CTAs per MMA (tcgen05 cta_group): 1, or 2 = a CTA pair computes a 256 x BN tile, each CTA holding 128 rows of A
// and BN/2 rows of B; the leader CTA issues every MMA and both CTAs' tensor cores execute it.
template <int BN, int CG> struct GemmCfg {
    static constexpr int kEpiWarps = BN / 32;                                   // each owns 32 TMEM lanes x 128 columns
    static constexpr int kThreads = (kEpiWarps + 4) * 32;                       // + producer warp, MMA warp, 2 idle (warpgroup padding)
    static constexpr int kBRows = BN / CG;                                      // rows of B this CTA stages
    static constexpr int kStageBytes = (2 * BM + 2 * kBRows) * kBlockBytes;
    static constexpr int kStages = kStageBytes > 80 * 1024 ? 2 : (kStageBytes > 48 * 1024 ? 3 : 4);
    static constexpr int kStagingLd = BN + 4;                                   // fp32 elements, padded against bank conflicts
    static constexpr int kStagingBytes = BM * kStagingLd * 4;
    static constexpr int kDataBytes = kStageBytes * kStages > kStagingBytes ? kStageBytes * kStages : kStagingBytes;
    static constexpr int kAccBufs = 512 / BN;                                   // TMEM accumulator buffers (all 512 columns are used)
    static constexpr int kVecBytes = (2 * BN + 3 * BM) * 4;                     // per-tile epilogue vectors: b_exp, bias | a_exp, out_exp, aux_exp
    static constexpr int kSmemBytes = kDataBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + kVecBytes;
};

// Per-tile epilogue operands staged in shared memory by an otherwise idle warp while the main loop runs, so that the
// epilogue itself issues no global loads (each one was a full L2 round trip on the critical path of a short kernel).
struct EpiVec {
    const int* b_exp;                 // [BN] scale exponents of the tile's B rows = output columns   (KIND_F16)
    const float* bias;                // [BN]                                                          (EPI_BIAS_*)
    const int* a_exp;                 // [BM] scale exponents of the tile's A rows                      (KIND_F16)
    const int* out_exp;               // [BM] scale exponents of the rows coming out                    (KIND_F16)
    const int* aux_exp;               // [BM] scale exponents of the aux rows                           (KIND_F16, EPI_TANH_GRAD)
};

// Diagnostic build only (make EXTRA=-DANNF_PHASE_CLOCKS): %globaltimer (ns) at the phase boundaries of CTA (0,0,0) of every
// kernel of the step, plus first / last CTA start and last CTA end over the whole grid.  annf_debug_tensor_phases reads it.
//   [slot][0..15] phases of CTA 0, [16] min start, [17] max start, [18] max end, over all CTAs
__device__ unsigned long long g_tensor_clock[8][24];
#ifdef ANNF_PHASE_CLOCKS
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ANNF_TMARK(slot, i) do { if ((slot) >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tensor_clock[slot][i] = gtimer(); } while (0)
#define ANNF_TGRID_START(slot) do { if ((slot) >= 0 && threadIdx.x == 0) { const unsigned long long t_ = gtimer(); \
        atomicMin(&g_tensor_clock[slot][16], t_); atomicMax(&g_tensor_clock[slot][17], t_); } } while (0)
#define ANNF_TGRID_END(slot) do { if ((slot) >= 0 && threadIdx.x == 0) atomicMax(&g_tensor_clock[slot][18], gtimer()); } while (0)
// SM cycles CTA 0's single-thread roles spent waiting: [19] MMA issuer on accumulator buffers, [20] MMA issuer on operands,
// [21] TMA producer on free stages, [22] epilogue warp 0 on finished chunks
// per-CTA marks of the GEMM launched last (diagnostic of CTA-to-CTA skew): [i][linear block id], i = 0 main loop done, 1 drained,
// 2 stores read, 3 exit
__device__ __forceinline__ unsigned smid_() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ unsigned long long g_cta_clock[10][1024];  // [4] first operands landed, [5] SM id, [6] entry, [7] prologue done, [8] dependency resolved
__device__ int g_cta_slot = 5;                        // which kernel of the step records per-CTA marks (annf_debug_cta_clocks sets it)
__device__ unsigned long long g_sm_exit[8][256];      // [kernel slot][SM id]: when the (last) CTA of that kernel left that SM
#define ANNF_TCTA(i) do { const unsigned b_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); if (b_ < 1024 && p.dbg == g_cta_slot) { g_cta_clock[i][b_] = gtimer(); g_cta_clock[5][b_] = smid_(); } } while (0)
#define ANNF_TSM_EXIT(slot) do { if ((slot) >= 0 && (slot) < 8) g_sm_exit[slot][smid_() & 255] = gtimer(); } while (0)
// start-up of one LATE CTA of the force GEMM (block (20, 0): launched when the GEMM before it leaves the SM), into slot 7:
// 0 entry, 1 barriers initialised, 2 TMEM allocated, 3 static vectors loaded, 4 CTA / pair synchronised, 5 B requested,
// 6 dependency resolved, 7 A requested, 8 first operands landed
#define ANNF_TLATE(i) do { if (p.dbg == 5 && blockIdx.x == 20 && blockIdx.y == 0) g_tensor_clock[7][i] = gtimer(); } while (0)
#define ANNF_TWAIT_BEGIN() const long long tw_ = clock64()
#define ANNF_TWAIT_END(slot, i) do { if ((slot) >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tensor_clock[slot][i] += (unsigned long long) (clock64() - tw_); } while (0)
#else
#define ANNF_TWAIT_BEGIN() do { } while (0)
#define ANNF_TWAIT_END(slot, i) do { } while (0)
#define ANNF_TMARK(slot, i) do { } while (0)
#define ANNF_TCTA(i) do { } while (0)
#define ANNF_TSM_EXIT(slot) do { } while (0)
#define ANNF_TLATE(i) do { } while (0)
#define ANNF_TGRID_START(slot) do { } while (0)
#define ANNF_TGRID_END(slot) do { } while (0)
#endif

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- cta_group::2 flavours of the PTX wrappers -------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completes its bytes on the mbarrier at `bar_cluster` (the leader's)
template <bool HINT>
__device__ __forceinline__ void tma_load_2d_pair_h(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, uint64_t pol) {
    if (HINT)
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(pol) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_mma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// arrive (once the MMAs issued so far retire) on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
// 1-CTA MMAs, but the completion is signalled in every CTA of `cta_mask` (CTAs that share multicast operand tiles)
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}
// TMA load whose box lands at the same shared-memory offset, and completes its bytes on the mbarrier at the same offset, in
// every CTA of `cta_mask`
template <bool HINT>
__device__ __forceinline__ void tma_load_2d_mc_h(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t cta_mask, uint64_t pol) {
    if (HINT)
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5, %6;"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask), "l"(pol) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask) : "memory");
}
// a / b format field: kind::tf32 -> 2 (TF32); kind::f16 -> 0 (F16)
__host__ __device__ constexpr uint32_t umma_idesc_mn(int kind, int m, int n) {
    return (1u << 4) | ((kind == KIND_TF32 ? 2u : 0u) << 7) | ((kind == KIND_TF32 ? 2u : 0u) << 10) | ((uint32_t) (n >> 3) << 17) |
           ((uint32_t) (m >> 4) << 24);
}

// One epilogue element group = 4 consecutive columns of one output row, in two steps so that callers can issue the global
// loads of several groups before any of the dependent math (the epilogue loop is otherwise bound by one L2 round trip per group).
struct EpiAux { float4 a, b, unscale; float out_scale; };

__device__ __forceinline__ float exp2i(int k) { return __int_as_float((k + 127) << 23); }   // 2^k exactly, -126 <= k <= 127

__device__ __forceinline__ float4 half4_to_float4(uint2 raw) {
    const __half2 p0 = *reinterpret_cast<const __half2*>(&raw.x), p1 = *reinterpret_cast<const __half2*>(&raw.y);
    const float2 f0 = __half22float2(p0), f1 = __half22float2(p1);
    return make_float4(f0.x, f0.y, f1.x, f1.y);
}

// v (already multiplied by the row's power-of-two scale) -> hi = rn_f16(v), lo = rn_f16(v - hi), four values packed
__device__ __forceinline__ void split_f16x4(float4 v, uint2& hi, uint2& lo) {
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

// (m, n) global row / first column of the group; (lr, lc) the same inside the tile (indices into the staged vectors)
template <int KIND>
__device__ __forceinline__ EpiAux epilogue_load(const GemmParams& p, const EpiVec& vec, int m, int n, int lr, int lc) {
    EpiAux x;
    x.a = make_float4(0.f, 0.f, 0.f, 0.f);
    x.b = x.a;
    x.unscale = make_float4(1.f, 1.f, 1.f, 1.f);
    x.out_scale = 1.f;
    if (KIND == KIND_F16) {
        const int ea = vec.a_exp[lr];
        const int4 eb = *reinterpret_cast<const int4*>(vec.b_exp + lc);
        x.unscale = make_float4(exp2i(-(ea + eb.x)), exp2i(-(ea + eb.y)), exp2i(-(ea + eb.z)), exp2i(-(ea + eb.w)));
        if (p.epi != EPI_STORE && p.epi != EPI_FORCE) x.out_scale = exp2i(vec.out_exp[lr]);
    }
    if (p.epi == EPI_BIAS_TANH || p.epi == EPI_BIAS_LINEAR) {
        x.a = *reinterpret_cast<const float4*>(vec.bias + lc);
    } else if (p.epi == EPI_TANH_GRAD) {
        if (KIND == KIND_F16) {
            // a = (hi + lo) * 2^-e: the sum of the pair is exact in fp32 and so is the power-of-two unscale
            const float un = exp2i(-vec.aux_exp[lr]);
            const float4 h = half4_to_float4(__ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(p.aux_hi) + (size_t) m * p.ld_aux + n)));
            const float4 l = half4_to_float4(__ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(p.aux_lo) + (size_t) m * p.ld_aux + n)));
            x.a = make_float4((h.x + l.x) * un, (h.y + l.y) * un, (h.z + l.z) * un, (h.w + l.w) * un);
        } else {
            x.a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.aux_hi) + (size_t) m * p.ld_aux + n));
            x.b = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.aux_lo) + (size_t) m * p.ld_aux + n));
        }
    }
    return x;
}

template <int KIND>
__device__ __forceinline__ void epilogue_finish(const GemmParams& p, int m, int n, float4 v, const EpiAux& x) {
    if (KIND == KIND_F16) { v.x *= x.unscale.x; v.y *= x.unscale.y; v.z *= x.unscale.z; v.w *= x.unscale.w; }
    if (p.epi == EPI_STORE) {
        *reinterpret_cast<float4*>(static_cast<float*>(p.out_hi) + (size_t) m * p.ld_out + n) = v;
        return;
    }
    if (p.epi == EPI_FORCE) {
        // column n = 3 * atom + component.  The B operand of this GEMM is W0c^T = -(1/s) (W_0 - (1/A) sum_i W_0[.][3i+c])^T, the
        // first connection with the Jacobian of "divide by s, subtract the centroid" folded in at plan creation
        // (ANN_ReferenceKernels.cpp:137-145, :397-410), so the accumulator already IS the force.
        const float4 f = v;
        float* row = static_cast<float*>(p.out_hi) + (size_t) m * p.ld_out;
        if (p.sel_atoms == nullptr) {
            *reinterpret_cast<float4*>(row + n) = f;
        } else {
            const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int e = n + j, atom = e / 3;
                row[(size_t) __ldg(p.sel_atoms + atom) * 3 + (e - 3 * atom)] = fv[j];
            }
        }
        return;
    }
    if (p.epi == EPI_BIAS_TANH || p.epi == EPI_BIAS_LINEAR) {
        v.x += x.a.x; v.y += x.a.y; v.z += x.a.z; v.w += x.a.w;
        if (p.epi == EPI_BIAS_TANH) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
    } else if (p.epi == EPI_TANH_GRAD) {
        const float a0 = x.a.x + x.b.x, a1 = x.a.y + x.b.y, a2 = x.a.z + x.b.z, a3 = x.a.w + x.b.w;
        v.x *= (1.f - a0 * a0); v.y *= (1.f - a1 * a1); v.z *= (1.f - a2 * a2); v.w *= (1.f - a3 * a3);
    }
    if (KIND == KIND_F16) {
        uint2 hi, lo;
        split_f16x4(make_float4(v.x * x.out_scale, v.y * x.out_scale, v.z * x.out_scale, v.w * x.out_scale), hi, lo);
        *reinterpret_cast<uint2*>(static_cast<__half*>(p.out_hi) + (size_t) m * p.ld_out + n) = hi;
        *reinterpret_cast<uint2*>(static_cast<__half*>(p.out_lo) + (size_t) m * p.ld_out + n) = lo;
        return;
    }
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(static_cast<float*>(p.out_hi) + (size_t) m * p.ld_out + n) = hi;
    *reinterpret_cast<float4*>(static_cast<float*>(p.out_lo) + (size_t) m * p.ld_out + n) = lo;
}

// Cooperative epilogue over the parked tile(s): thread t of kGroupThreads walks (row, 4-column) groups with consecutive lanes
// along a row — 512-byte coalesced global accesses — summing the S split-K partials (own + peers', through DSMEM) in a fixed
// order.  The loop is ROLLED and software-pipelined by hand (the loads of group g + 1 are issued before the math of group g):
// this code runs a handful of times per CTA, so its size is paid in instruction-cache misses — ~170 cycles per 128-byte
// line of cold code — and an unrolled body with 16 inlined tanhf was most of the 4-6 us this epilogue used to take.
template <int KIND, int BN, int S, int CG, int kGroupThreads, int CX = CG>
__device__ __forceinline__ void cooperative_epilogue(const GemmParams& p, const EpiVec& vec, unsigned char* data, int rank, int pair_rank, int m0, int n0, int t) {
    using Cfg = GemmCfg<BN, CG>;
    constexpr int kRows = (BM + S - 1) / S;                                        // rows this CTA finishes (S = 3: 43, 43, 42)
    constexpr int kVecPerRow = BN / 4;
    float* staging = reinterpret_cast<float*>(data);
    const float* peer[S];
    if (S > 1) {
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int r = 0; r < S; r++) peer[r] = cluster.map_shared_rank(staging, pair_rank + CX * r);   // same x position, split r
    } else {
        peer[0] = staging;
    }
    constexpr int kGroups = (kRows * kVecPerRow + kGroupThreads - 1) / kGroupThreads;   // groups per thread
    struct Group { EpiAux aux; float4 sum; int m, n; bool ok; };
    auto fetch = [&](int g) {
        Group q;
        const int v = t + g * kGroupThreads;
        const int row = rank * kRows + v / kVecPerRow, col = (v % kVecPerRow) * 4;
        q.m = m0 + row; q.n = n0 + col;
        q.ok = g < kGroups && v < kRows * kVecPerRow && row < BM && q.m < p.M && q.n < p.N;
        q.sum = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q.ok) {
            q.aux = epilogue_load<KIND>(p, vec, q.m, q.n, row, col);
#pragma unroll
            for (int r = 0; r < S; r++) {                                          // fixed order: deterministic
                const float4 x = *reinterpret_cast<const float4*>(peer[r] + (size_t) row * Cfg::kStagingLd + col);
                q.sum.x += x.x; q.sum.y += x.y; q.sum.z += x.z; q.sum.w += x.w;
            }
        }
        return q;
    };
    // one group ahead.  (Two ahead measured slower at BASELINE C5: epilogues 2.6 / 2.3 / 3.0 -> 2.7 / 2.7 / 4.3 us on the three
    // split-K GEMMs — more DSMEM requests in flight per thread did not shorten the round trips.)
    Group next = fetch(0);
#pragma unroll 1
    for (int g = 0; g < kGroups; g++) {
        const Group cur = next;
        next = fetch(g + 1);
        if (cur.ok) epilogue_finish<KIND>(p, cur.m, cur.n, cur.sum, cur.aux);
    }
}

// Epilogue of a tile whose K ranges were computed by `gs` different CTAs (split-K through global memory): this CTA finishes rows
// [grank * rows, ...) of the tile, summing the gs partial tiles in a fixed order (deterministic).  Same rolled, software-
// pipelined structure as cooperative_epilogue; the partials are read with ld.global.cg (they were written by other SMs).
template <int KIND, int BN, int kGroupThreads>
__device__ __forceinline__ void global_split_epilogue(const GemmParams& p, const EpiVec& vec, const float* tile_partials, int gs, int grank,
                                                      int m0, int n0, int t) {
    constexpr int kVecPerRow = BN / 4;
    const int rows = (BM + gs - 1) / gs;
    const int groups = (rows * kVecPerRow + kGroupThreads - 1) / kGroupThreads;
    struct Group { EpiAux aux; float4 sum; int m, n; bool ok; };
    auto fetch = [&](int g) {
        Group q;
        const int v = t + g * kGroupThreads;
        const int row = grank * rows + v / kVecPerRow, col = (v % kVecPerRow) * 4;
        q.m = m0 + row; q.n = n0 + col;
        q.ok = g < groups && v < rows * kVecPerRow && row < BM && q.m < p.M && q.n < p.N;
        q.sum = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q.ok) {
            q.aux = epilogue_load<KIND>(p, vec, q.m, q.n, row, col);
            const float* src = tile_partials + (size_t) row * BN + col;
            for (int r = 0; r < gs; r++) {                                         // fixed order: deterministic
                const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (size_t) r * BM * BN));
                q.sum.x += x.x; q.sum.y += x.y; q.sum.z += x.z; q.sum.w += x.w;
            }
        }
        return q;
    };
    Group next = fetch(0);
#pragma unroll 1
    for (int g = 0; g < groups; g++) {
        const Group cur = next;
        next = fetch(g + 1);
        if (cur.ok) epilogue_finish<KIND>(p, cur.m, cur.n, cur.sum, cur.aux);
    }
}

// Block- or cluster-wide rendezvous used by every warp role at the same two points of the kernel.
// kOrderMemory = false: execution ordering only ("nobody leaves while peers may still read its shared memory").  The release
// form of the cluster barrier compiles to MEMBAR.ALL.GPU, which at the end of a kernel waits for every global store of the
// epilogue to be acknowledged by L2 — visibility the NEXT kernel gets from the grid dependency anyway.
template <bool kCluster, bool kOrderMemory = true> __device__ __forceinline__ void role_sync() {
    if (kCluster) {
        if (kOrderMemory) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        }
    } else {
        asm volatile("bar.sync 0;" ::: "memory");
    }
}

// NX = 2 (1-CTA tiles only): the two CTAs at x = 2i, 2i + 1 of a cluster compute neighbouring N tiles of the SAME rows, so
// they need the same A tile: each loads half of its rows and TMA-multicasts them into both CTAs (48 instead of 64 KB of
// L2 -> SM traffic per K-block and CTA).  A stage is free again when BOTH CTAs' MMAs have retired (multicast commit).
template <int KIND, int BN, int S, int CG, int NX = 1>
__global__ void __launch_bounds__(GemmCfg<BN, CG>::kThreads, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
              const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
              const __grid_constant__ CUtensorMap tm_out, const GemmParams p) {
    using Cfg = GemmCfg<BN, CG>;
    constexpr int BK = KindOf<KIND>::kBlockElems;           // operand values per K-block (one 128-byte swizzle span)
    constexpr int kStages = Cfg::kStages;
    constexpr int kEpiWarps = Cfg::kEpiWarps;
    constexpr int kColsPerWarp = 128;                       // accumulator columns held in registers by one epilogue thread
    constexpr int kEpiThreads = BN >= 256 ? 32 * kEpiWarps : Cfg::kThreads;   // threads sharing the cooperative epilogue
    constexpr int kBRows = Cfg::kBRows;
    constexpr int kBBox = kBRows < 128 ? kBRows : 128;      // rows of one TMA box of B (the 256 x 128 pair tile stages 64 per CTA)
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment: required by the 128B swizzle atoms (TMA and UMMA must agree on address bits 7..9)
    unsigned char* data = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    // barriers: full[kStages], empty[kStages], acc_full[NB], acc_empty[NB]   (CG = 2: full and acc_empty are used in the leader only)
    constexpr int NB = Cfg::kAccBufs;                       // TMEM accumulator buffers: 4 (BN = 128) or 2 (BN = 256)
    uint64_t* bars = reinterpret_cast<uint64_t*>(data + Cfg::kDataBytes);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + kStages;
    uint64_t* bar_acc_full = bars + 2 * kStages;
    uint64_t* bar_acc_empty = bars + 2 * kStages + NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * NB);
    static_assert((2 * 4 + 2 * 4) * 8 + 4 <= 256, "barrier region");
    // per-tile epilogue vectors (EpiVec), staged by warp kEpiWarps + 2
    int* s_b_exp = reinterpret_cast<int*>(data + Cfg::kDataBytes + 256);
    float* s_bias = reinterpret_cast<float*>(s_b_exp + BN);
    int* s_a_exp = reinterpret_cast<int*>(s_bias + BN);
    int* s_out_exp = s_a_exp + BM;
    int* s_aux_exp = s_out_exp + BM;
    const EpiVec vec = {s_b_exp, s_bias, s_a_exp, s_out_exp, s_aux_exp};

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(NX == 1 || (CG == 1 && NX == 2), "operand multicast is implemented for pairs of 1-CTA tiles");
    constexpr int CX = CG * NX;                                                   // cluster extent along x
    const int x_rank = CX == 2 ? (int) (blockIdx.x & 1) : 0;                      // position in the cluster along x
    const int pair_rank = CG == 2 ? x_rank : 0;                                   // 0 = leader of the CTA pair
    const int n0 = (CG == 2 ? (int) (blockIdx.x >> 1) : (int) blockIdx.x) * BN;   // first output column of the tile
    const int m0 = (CG == 2 ? (int) blockIdx.y * 2 + pair_rank : (int) blockIdx.y) * BM;
    const int b_row0 = n0 + pair_rank * kBRows;                                   // first row of B this CTA stages
    const int rank = S > 1 ? (int) blockIdx.z : 0;                                // split-K index
    const uint32_t my_cta = (CX == 2 || S > 1) ? cluster_ctarank() : 0u;
    const uint16_t mc_mask = (uint16_t) (3u << (my_cta & ~1u));                   // NX = 2: this CTA and its x neighbour
    const uint32_t leader_cta = my_cta & ~1u;
    const int kb_total = (p.K + BK - 1) / BK;
    const int gs = (S == 1 && p.gsplit > 1) ? p.gsplit : 1;                       // split-K through global memory (gridDim.z ranges)
    const int grank = gs > 1 ? (int) blockIdx.z : 0;
    const int ksplits = S > 1 ? S : gs, krank = S > 1 ? rank : grank;
    const int kb0 = (int) ((long long) kb_total * krank / ksplits), kb1 = (int) ((long long) kb_total * (krank + 1) / ksplits);
    const int num_kb = kb1 - kb0;
    const int kChunk = p.chunk;
    const int num_chunks = (num_kb + kChunk - 1) / kChunk;
    const bool direct = p.direct_store != 0;                // unsplit force GEMM: registers -> swizzled tile -> TMA store

    ANNF_TGRID_START(p.dbg);
    if (threadIdx.x == 0) { ANNF_TMARK(p.dbg, 0); ANNF_TCTA(6); ANNF_TLATE(0); }
    pdl_launch_dependents();                                // the next kernel may start its own prologue now
    if (p.dry & 1) return;
    if (warp == kEpiWarps && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
        for (int s = 0; s < kStages; s++) {
            mbar_init(smem_u32(&bar_full[s]), CG);                  // leader's arrive.expect_tx (+ the peer's remote arrive)
            mbar_init(smem_u32(&bar_empty[s]), NX);                 // NX = 2: both CTAs' MMA commits
        }
        for (int b = 0; b < NB; b++) {
            mbar_init(smem_u32(&bar_acc_full[b]), 1);
            mbar_init(smem_u32(&bar_acc_empty[b]), CG * kEpiWarps);   // one arrival per epilogue warp (of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ANNF_TLATE(1);
    }
    if (warp == kEpiWarps + 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t) (NB * BN)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t) (NB * BN)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        if (lane == 0) ANNF_TLATE(2);
    }
    tcgen05_fence_before();
    if (CX == 2) cg::this_cluster().sync();                 // the peer arrives on / multicasts into this CTA: its barriers must exist first
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) { ANNF_TMARK(p.dbg, 1); ANNF_TCTA(7); ANNF_TLATE(4); }
    // One pipeline stage of K-block i: arm the stage's barrier, load the A rows (activations: written by the kernel before this
    // one) and / or the B rows (weights: static).  Called by the producer thread only.
    constexpr bool kHintA = (kL2Hints & 4) != 0, kHintB = (kL2Hints & 8) != 0;
    const uint64_t pol_last = (kHintA || kHintB) ? l2_evict_last() : 0;
    auto issue_stage = [&](int i, bool do_a, bool do_b, bool arm) {
        const int s = i % kStages;
        const uint32_t full = smem_u32(&bar_full[s]);
        const uint32_t base = smem_u32(data + (size_t) s * Cfg::kStageBytes);
        const int kc = (kb0 + i) * BK;
        if (CG == 2) {
            const uint32_t full_leader = map_to_cta(full, leader_cta);
            if (arm) {
                if (pair_rank == 0) mbar_expect_tx(full, (uint32_t) (2 * Cfg::kStageBytes));   // both CTAs' bytes land here
                else mbar_arrive_cluster(full_leader);
            }
            if (do_a) {
                tma_load_2d_pair_h<kHintA>(base, &tm_a_hi, full_leader, kc, m0, pol_last);
                tma_load_2d_pair_h<kHintA>(base + BM * kBlockBytes, &tm_a_lo, full_leader, kc, m0, pol_last);
            }
            if (do_b) {
#pragma unroll
                for (int h = 0; h < kBRows / kBBox; h++) {
                    tma_load_2d_pair_h<kHintB>(base + 2 * BM * kBlockBytes + h * kBBox * kBlockBytes, &tm_b_hi, full_leader, kc, b_row0 + h * kBBox, pol_last);
                    tma_load_2d_pair_h<kHintB>(base + 2 * BM * kBlockBytes + kBRows * kBlockBytes + h * kBBox * kBlockBytes, &tm_b_lo, full_leader, kc, b_row0 + h * kBBox, pol_last);
                }
            }
        } else {
            if (arm) mbar_expect_tx(full, (uint32_t) Cfg::kStageBytes);
            if (do_a) {
                if (NX == 2) {
                    // own half of the A rows to both CTAs (the maps have 64-row boxes); the other half arrives from the peer
                    const uint32_t half = (uint32_t) x_rank * (BM / 2) * kBlockBytes;
                    tma_load_2d_mc_h<kHintA>(base + half, &tm_a_hi, full, kc, m0 + x_rank * (BM / 2), mc_mask, pol_last);
                    tma_load_2d_mc_h<kHintA>(base + BM * kBlockBytes + half, &tm_a_lo, full, kc, m0 + x_rank * (BM / 2), mc_mask, pol_last);
                } else {
                    tma_load_2d_h<kHintA>(base, &tm_a_hi, full, kc, m0, pol_last);
                    tma_load_2d_h<kHintA>(base + BM * kBlockBytes, &tm_a_lo, full, kc, m0, pol_last);
                }
            }
            if (do_b) {
#pragma unroll
                for (int h = 0; h < kBRows / kBBox; h++) {                        // TMA boxes are at most 256 rows; use 128-row boxes
                    tma_load_2d_h<kHintB>(base + 2 * BM * kBlockBytes + h * kBBox * kBlockBytes, &tm_b_hi, full, kc, b_row0 + h * kBBox, pol_last);
                    tma_load_2d_h<kHintB>(base + 2 * BM * kBlockBytes + kBRows * kBlockBytes + h * kBBox * kBlockBytes, &tm_b_lo, full, kc, b_row0 + h * kBBox, pol_last);
                }
            }
        }
    };
    // The weights do not depend on the kernel before this one: the first stages are armed and their B halves requested BEFORE the
    // dependency wait, so that after it only the A halves are still on their way (the wait-to-first-MMA gap was 1.6 - 1.9 us).
    const int npre = (p.dry || p.no_b_prefetch) ? 0 : min(kStages, num_kb);
    if (warp == kEpiWarps && lane == 0) {
        for (int i = 0; i < npre; i++) issue_stage(i, false, true, true);
        ANNF_TLATE(5);
    }
    if (warp == kEpiWarps + 2) {
        // static per-tile vectors (weights' scale exponents, biases): not produced by an earlier kernel of the step.  AFTER the
        // prologue rendezvous, and all loads in flight at once: as a rolled loop before it, these were BN / 32 dependent L2 round
        // trips (2.9 us at BN = 256) that every warp of a CTA launched late — on an SM the previous GEMM had just left — sat out
        // before its first TMA request (profiles/r02_cta_startup.md).
        int be[BN / 32];
        float bi[BN / 32];
        const bool has_bias = p.epi == EPI_BIAS_TANH || p.epi == EPI_BIAS_LINEAR;
#pragma unroll
        for (int j = 0; j < BN / 32; j++) {
            const int n = n0 + lane + 32 * j;
            be[j] = (KIND == KIND_F16 && n < p.N) ? __ldg(p.b_exp + n) : 0;
            bi[j] = (has_bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < BN / 32; j++) { s_b_exp[lane + 32 * j] = be[j]; s_bias[lane + 32 * j] = bi[j]; }
        if (lane == 0) ANNF_TLATE(3);
    }
    pdl_wait();                                             // everything below reads what earlier kernels wrote
    if (threadIdx.x == 0) { ANNF_TMARK(p.dbg, 2); ANNF_TCTA(8); ANNF_TLATE(6); }
    if (warp == kEpiWarps + 2) {
        // per-row scale exponents written by earlier kernels of the step; published to the epilogue by named barrier 1
        if (KIND == KIND_F16) {
            const bool has_out = p.epi != EPI_STORE && p.epi != EPI_FORCE;
            int ea[BM / 32], eo[BM / 32], ex[BM / 32];      // all loads in flight at once (a short main loop ends 2.5 us after the wait)
#pragma unroll
            for (int j = 0; j < BM / 32; j++) {
                const int m = m0 + lane + 32 * j;
                ea[j] = m < p.M ? p.a_exp[m] : 0;
                eo[j] = (has_out && m < p.M) ? p.out_exp[m] : 0;
                ex[j] = (p.epi == EPI_TANH_GRAD && m < p.M) ? p.aux_exp[m] : 0;
            }
#pragma unroll
            for (int j = 0; j < BM / 32; j++) { s_a_exp[lane + 32 * j] = ea[j]; s_out_exp[lane + 32 * j] = eo[j]; s_aux_exp[lane + 32 * j] = ex[j]; }
        }
        __syncwarp();
        asm volatile("bar.arrive 1, %0;" ::"n"(32 * (kEpiWarps + 1)) : "memory");
        if (p.epi == EPI_TANH_GRAD && gs == 1 && !direct) {
            // act'() needs the activations of the rows this CTA will finish ((BM / S) x BN of the pair, written two kernels ago):
            // pull those lines into L1 while the main loop runs, so that the epilogue's loads are L1 hits instead of one L2
            // round trip per (software-pipelined) group — the epilogue of the backward GEMM was 3.3 us against 2.1 us for the
            // forward one, which has no such loads.
            typedef typename KindOf<KIND>::elem elem;
            constexpr int kRows = (BM + S - 1) / S;
            constexpr int kLinesPerRow = BN * (int) sizeof(elem) / 128;
            for (int i = lane; i < kRows * kLinesPerRow; i += 32) {
                const int row = rank * kRows + i / kLinesPerRow, m = m0 + row;
                const int col = n0 + (i % kLinesPerRow) * (128 / (int) sizeof(elem));
                if (row < BM && m < p.M && col < p.N) {
                    const size_t off = (size_t) m * p.ld_aux + col;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(static_cast<const elem*>(p.aux_hi) + off));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(static_cast<const elem*>(p.aux_lo) + off));
                }
            }
        }
    }

    // Rendezvous + epilogue, the same for every warp role.  BN = 128: ONE instance after the roles re-join (the epilogue's code is
    // large: see cooperative_epilogue), and the producer / MMA warps, idle by then and still holding their registers, take a
    // share of it.  BN = 256: one instance per role, INSIDE the region governed by that role's setmaxnreg (the producer
    // warpgroup gave its registers away and only keeps the rendezvous; code after the roles re-join could not assume either
    // register budget).
    auto finish_tile = [&](bool participates) {
        role_sync<(S > 1 || NX == 2)>();                    // every partial tile (of the split) is parked; NX = 2: the peer no longer multicasts into this CTA's stages
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 7);
        if (S == 1 && gs > 1) {
            if (participates && !(p.dry & 8)) {
                // publish this CTA's partial tile, meet the other K ranges of the tile, finish 1 / gs of its rows
                const int n_tiles = (int) gridDim.x / CG;
                const int tile = ((int) blockIdx.y * n_tiles + n0 / BN) * CG + pair_rank;
                float* mine = p.gs_scratch + ((size_t) tile * gs + grank) * BM * BN;
                const float* staging = reinterpret_cast<const float*>(data);
                for (int v = (int) threadIdx.x; v < BM * (BN / 4); v += kEpiThreads) {
                    const int row = v / (BN / 4), col = (v % (BN / 4)) * 4;
                    __stcg(reinterpret_cast<float4*>(mine + (size_t) row * BN + col),
                           *reinterpret_cast<const float4*>(staging + (size_t) row * Cfg::kStagingLd + col));
                }
                __threadfence();
                asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
                if (threadIdx.x == 0) {
                    atomicAdd(p.gs_arrive + tile, 1);
                    // every CTA of the grid is resident (the launcher sizes the grid to the SM count), so this wait ends; the
                    // bound turns a violated assumption into a reported error instead of a hang
                    const long long t_start = clock64();
                    while (*reinterpret_cast<volatile int*>(p.gs_arrive + tile) < gs)
                        if (clock64() - t_start > (1ll << 31)) __trap();
                    __threadfence();
                }
                asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
                if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 10);
                global_split_epilogue<KIND, BN, kEpiThreads>(p, vec, p.gs_scratch + (size_t) tile * gs * BM * BN, gs, grank, m0, n0, (int) threadIdx.x);
                asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
                if (threadIdx.x == 0 && atomicAdd(p.gs_done + tile, 1) == gs - 1) {     // the last reader re-arms the tile for the next launch
                    p.gs_done[tile] = 0;
                    p.gs_arrive[tile] = 0;
                    __threadfence();
                }
            }
        } else if (participates && !(p.dry & 8) && !(S == 1 && direct))
            cooperative_epilogue<KIND, BN, S, CG, kEpiThreads, CX>(p, vec, data, rank, x_rank, m0, n0, (int) threadIdx.x);
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 8);
        role_sync<(S > 1 || CX == 2), false>();             // nobody leaves while peers still read / share TMEM (execution order only)
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 9);
    };
    if (warp >= kEpiWarps) {
        // ===== producer / MMA warpgroup: needs few registers =====
        if (BN >= 256) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
        if (warp == kEpiWarps) {
            // ----- TMA producer (every CTA stages its own operand rows) -----
            if (lane == 0) {
                for (int i = 0; i < num_kb; i++) {
                    if (i < npre) {                                               // armed, and its B half issued, before the dependency wait
                        if (i == 0) ANNF_TMARK(p.dbg, 3);
                        issue_stage(i, true, false, false);
                        if (i == 0) ANNF_TLATE(7);
                        continue;
                    }
                    const int s = i % kStages;
                    const uint32_t ph = (uint32_t) (i / kStages) & 1u;
                    { ANNF_TWAIT_BEGIN();
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);                  // slot free (in both CTAs of a pair)
                    ANNF_TWAIT_END(p.dbg, 21); }
                    if (p.dry & 4) {                                              // experiments: no TMA traffic
                        const uint32_t full = smem_u32(&bar_full[s]);
                        if (CG == 2 && pair_rank == 1) mbar_arrive_cluster(map_to_cta(full, leader_cta));
                        else if (CG == 2) mbar_expect_tx(full, 0);
                        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");
                        continue;
                    }
                    if (i == 0) ANNF_TMARK(p.dbg, 3);
                    issue_stage(i, true, true, true);
                }
            }
        } else if (warp == kEpiWarps + 1 && pair_rank == 0) {
            // ----- MMA issuer (one thread of the leader CTA).  Chunk c accumulates into TMEM buffer c & 1. -----
            if (lane == 0) {
                constexpr uint32_t idesc = umma_idesc_mn(KIND, BM * CG, BN);
                const uint16_t pair_mask = (uint16_t) (3u << leader_cta);          // both CTAs of this pair
                int i = 0;
                for (int c = 0; c < num_chunks; c++) {
                    const int buf = c % NB;
                    const uint32_t use = (uint32_t) (c / NB);
                    { ANNF_TWAIT_BEGIN();
                    mbar_wait(smem_u32(&bar_acc_empty[buf]), (use & 1u) ^ 1u);    // epilogue warps (of both CTAs) drained this buffer
                    ANNF_TWAIT_END(p.dbg, 19); }
                    tcgen05_fence_after();
                    const uint32_t tmem_acc = tmem_base + (uint32_t) (buf * BN);
                    const int i_end = min(num_kb, (c + 1) * kChunk);
                    bool first = true;
                    for (; i < i_end; i++) {
                        const int s = i % kStages;
                        const uint32_t ph = (uint32_t) (i / kStages) & 1u;
                        { ANNF_TWAIT_BEGIN();
                        mbar_wait(smem_u32(&bar_full[s]), ph);                     // TMA bytes landed (both CTAs)
                        ANNF_TWAIT_END(p.dbg, 20); }
                        tcgen05_fence_after();
                        if (i == 0) { ANNF_TMARK(p.dbg, 4); ANNF_TCTA(4); ANNF_TLATE(8); }
                        const uint32_t base = smem_u32(data + (size_t) s * Cfg::kStageBytes);
                        const uint32_t a_hi = base, a_lo = base + BM * kBlockBytes;
                        const uint32_t b_hi = base + 2 * BM * kBlockBytes, b_lo = b_hi + kBRows * kBlockBytes;
#pragma unroll
                        for (int kk = 0; kk < kStepsPerBlock; kk++) {
                            if (p.dry & 2) break;                                  // experiments: no tensor-core work
                            const uint32_t off = kk * kStepBytes;                   // 32 bytes along K inside the swizzle span
                            const uint64_t dah = umma_smem_desc(a_hi + off), dal = umma_smem_desc(a_lo + off);
                            const uint64_t dbh = umma_smem_desc(b_hi + off), dbl = umma_smem_desc(b_lo + off);
                            if (CG == 2) {
                                umma_mma_pair<KIND>(tmem_acc, dal, dbh, idesc, first ? 0u : 1u);   // small terms first
                                umma_mma_pair<KIND>(tmem_acc, dah, dbl, idesc, 1u);
                                umma_mma_pair<KIND>(tmem_acc, dah, dbh, idesc, 1u);
                            } else {
                                umma_mma<KIND>(tmem_acc, dal, dbh, idesc, first ? 0u : 1u);
                                umma_mma<KIND>(tmem_acc, dah, dbl, idesc, 1u);
                                umma_mma<KIND>(tmem_acc, dah, dbh, idesc, 1u);
                            }
                            first = false;
                        }
                        if (CG == 2) umma_commit_pair(smem_u32(&bar_empty[s]), pair_mask);   // frees the slot in both CTAs
                        else if (NX == 2) umma_commit_mc(smem_u32(&bar_empty[s]), mc_mask);
                        else umma_commit(smem_u32(&bar_empty[s]));
                    }
                    if (CG == 2) umma_commit_pair(smem_u32(&bar_acc_full[buf]), pair_mask);   // chunk complete, both CTAs' epilogues
                    else umma_commit(smem_u32(&bar_acc_full[buf]));
                }
                ANNF_TMARK(p.dbg, 5);
                ANNF_TCTA(0);
            }
        }
        if (BN >= 256) finish_tile(false);
    } else {
        // ===== epilogue warps: promote TMEM chunks into fp32 registers (round-to-nearest adds) =====
        if (BN >= 256) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
        const int q = warp & 3;                                                    // TMEM lane quarter this warp may read
        const int col0 = (warp >> 2) * kColsPerWarp;                               // column range of this warp
        const int row = q * 32 + lane;
        float acc[kColsPerWarp];
#pragma unroll
        for (int j = 0; j < kColsPerWarp; j++) acc[j] = 0.f;
        for (int c = 0; c < num_chunks; c++) {
            const int buf = c % NB;
            const uint32_t use = (uint32_t) (c / NB);
            { ANNF_TWAIT_BEGIN();
            mbar_wait(smem_u32(&bar_acc_full[buf]), use & 1u);
            if (threadIdx.x == 0) ANNF_TWAIT_END(p.dbg, 22); }
            tcgen05_fence_after();
#pragma unroll
            for (int cc = 0; cc < kColsPerWarp / 32; cc++) {
                if (p.dry & 16) break;                                             // experiments: no TMEM drain
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (buf * BN + col0 + cc * 32), r);
#pragma unroll
                for (int j = 0; j < 32; j++) acc[cc * 32 + j] += __uint_as_float(r[j]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                const uint32_t bar = smem_u32(&bar_acc_empty[buf]);
                if (CG == 2) mbar_arrive_cluster(map_to_cta(bar, leader_cta));     // the MMA issuer lives in the leader
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            }
        }
        if (threadIdx.x == 0) { ANNF_TMARK(p.dbg, 6); ANNF_TCTA(1); }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * (kEpiWarps + 1)) : "memory");    // the staged per-tile vectors are in place
        if (S == 1 && direct) {
            // Unsplit force GEMM writing whole rows: every thread turns its own accumulators into forces (x 2^-(e_row + e_col)),
            // parks them in the layout of a 128-byte-swizzled TMA box (32 columns x 128 rows per box: lane = row, and the
            // 16-byte unit index is XORed with row % 8, so the 32 lanes of a store hit every bank exactly four times), and ONE
            // thread hands the boxes to the TMA engine.  The thread-per-group store loop this replaces was latency-bound:
            // 10.8 us for 18 MB at BASELINE C5 (profiles/r02_c5_phases.md).  Rows >= M and columns >= N are clipped by the TMA unit.
            if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 11);
            const int ea = KIND == KIND_F16 ? s_a_exp[row] : 0;
            unsigned char* dst = data + (size_t) row * 128;
#pragma unroll
            for (int j4 = 0; j4 < kColsPerWarp / 4; j4++) {
                const int c = col0 + 4 * j4;
                float4 v = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
                if (KIND == KIND_F16) {
                    const int4 eb = *reinterpret_cast<const int4*>(s_b_exp + c);
                    v.x *= exp2i(-(ea + eb.x)); v.y *= exp2i(-(ea + eb.y)); v.z *= exp2i(-(ea + eb.z)); v.w *= exp2i(-(ea + eb.w));
                }
                *reinterpret_cast<float4*>(dst + (size_t) (c >> 5) * (BM * 128) + (((j4 & 7) ^ (row & 7)) << 4)) = v;
            }
            fence_proxy_async_smem();                                              // generic-proxy writes -> visible to the TMA engine
            asm volatile("bar.sync 2, %0;" ::"n"(32 * kEpiWarps) : "memory");
            if (threadIdx.x == 0) {
                ANNF_TMARK(p.dbg, 12);
                const uint64_t pol_out = (kL2Hints & 2) ? l2_evict_first() : 0;
                for (int cb = 0; cb < BN / 32; cb++)
                    if (n0 + cb * 32 < p.N) tma_store_2d_h<(kL2Hints & 2) != 0>(&tm_out, smem_u32(data + (size_t) cb * (BM * 128)), n0 + cb * 32, m0, pol_out);
                ANNF_TMARK(p.dbg, 13);
                tma_store_commit_and_wait_read();                                  // shared memory has been read: the CTA may leave
                ANNF_TMARK(p.dbg, 14);
                ANNF_TCTA(2);
            }
        } else {
            // park the tile in shared memory (all TMA loads have landed and been consumed: the stage buffers are free).
            // Unsplit GEMMs do this too: the epilogue below then walks ROWS with consecutive lanes, i.e. 512-byte coalesced
            // global stores, instead of every lane writing 16 bytes into a different row (measured: 47 -> 27 us on the
            // 1024 x 4500 x 512 force GEMM).
            float* dst = reinterpret_cast<float*>(data) + (size_t) row * Cfg::kStagingLd + col0;
#pragma unroll
            for (int j = 0; j < kColsPerWarp; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
        if (BN >= 256) finish_tile(true);
    }
    if (BN < 256) finish_tile(true);

    if (warp == kEpiWarps + 1) {
        tcgen05_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) (NB * BN)) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) (NB * BN)) : "memory");
    }
    if (threadIdx.x == 0) { ANNF_TCTA(3); ANNF_TSM_EXIT(p.dbg); }
    ANNF_TGRID_END(p.dbg);
}

// ------------------------------------------------------------------------------------------------
// SIMT companions
// ------------------------------------------------------------------------------------------------
// KIND_F16 bookkeeping: how the per-row scale exponents are derived without a second pass over the data.  The exponent of
// the input row comes from its exact largest magnitude; every later tensor gets an exponent from a BOUND on its magnitudes,
// propagated through the layers with the weights' L1 norms (fp16 keeps 2^-24 .. 2^15: a bound that is 2^10 too generous
// still leaves every significant value with its full 11 + 11 bits).
struct ExpChain {
    float fwd_gain[kMaxLayers];     // max_n sum_k |W_l[n][k]|       (connection l, forward)
    float fwd_bias[kMaxLayers];     // max_n |b_l[n]|
    float bwd_gain[kMaxLayers];     // max_i sum_j |W_l[j][i]|       (connection l, backward)
    int* x_exp[kMaxLayers];         // [R] per forward tensor X_l, l = 0..L-2   (written by prep)
    int* d_exp[kMaxLayers];         // [R] per gradient tensor D_l, l = 1..L-2  (written by head)
    double inv_scale;               // 1 / scaling_factor (a double division per thread of prep otherwise)
    float inv_hi, inv_lo;           // the same as a float pair: hi = fl(1/s), lo = fl(1/s - hi); scaling by 2^e keeps the pair exact
};

__device__ __forceinline__ int exp_for_bound(float bound) {
    if (!(bound > 0.f) || !(bound < 3.0e38f)) return 0;
    // bound = f * 2^q with 0.5 <= f < 1: q = biased exponent - 126.  (A subnormal bound reads q = -126 instead of its true,
    // smaller q: both end at the upper clamp.)
    const int q = (int) ((__float_as_uint(bound) >> 23) & 0xffu) - 126;
    return max(-kExpClamp, min(kExpClamp, kExpTarget - q));
}
// warp-wide min / max of floats by one integer redux.sync: key(x) is monotonic in x (sign-magnitude -> two's-complement order)
__device__ __forceinline__ float warp_min_float(float v) {
    const int b = __float_as_int(v);
    const int key = b >= 0 ? b : (int) (0x80000000u - (unsigned) b);
    const int r = __reduce_min_sync(0xffffffffu, key);
    return __int_as_float(r >= 0 ? r : (int) (0x80000000u - (unsigned) r));
}
__device__ __forceinline__ float warp_max_float(float v) {
    const int b = __float_as_int(v);
    const int key = b >= 0 ? b : (int) (0x80000000u - (unsigned) b);
    const int r = __reduce_max_sync(0xffffffffu, key);
    return __int_as_float(r >= 0 ? r : (int) (0x80000000u - (unsigned) r));
}

// prep: gather + scale + remove centroid (ANN_ReferenceKernels.cpp:128-151), write the operand pair.
// One CTA per replica; the replica's selected coordinates are staged once in shared memory.  Two passes and ONE block barrier:
//   pass 1  global -> shared, per-component sums and min / max of d = p - p_ref (the largest |p - centroid| of a component is
//           attained at its min or max, so the row's scale exponent needs no pass of its own);
//   pass 2  shared -> (d - mean(d)) / s -> x 2^e -> operand pair -> global.
// Arithmetic is fp32 RELATIVE TO A REFERENCE ATOM of the replica (p_ref = its first selected atom): d = p - p_ref is exact or
// rounded relative to the molecule's extent, not to its distance from the origin, so p - centroid = d - mean(d) keeps every bit
// a far-away molecule has (the fp64 centring this replaces gave the same guarantee, but its conversions — three per element —
// ran on the FP64 pipe and made the kernel conversion-bound: 12.9 us for 18 MB in, 18 MB out; profiles/r02_prep.md).
// 1 / s is carried as a float pair.  Work is walked in groups of 12 floats = 4 atoms, so that the component of every element
// is static (n_0 = 3A is a multiple of 4 on this path, hence of 12).
struct PrepRow {                                            // what pass 2 needs, identical in every thread of the CTA
    float ref[3], mean[3];                                  // reference atom; mean of d per component
    float inv_hi, inv_lo;                                   // 2^e / s as a float pair
    float bound;                                            // largest |x| of the row (before the 2^e scale)
    int e0;
};

// finishes the block reduction of pass 1 (every thread redundantly: broadcast reads instead of a second barrier)
template <int KIND>
__device__ __forceinline__ PrepRow prep_finish_reduction(const float (*red_sum)[3], const float (*red_min)[3], const float (*red_max)[3], int nwarps,
                                                         const float* ref, int A, double scale) {
    PrepRow row;
    float bound = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float t = 0.f, lo = 3.0e38f, hi = -3.0e38f;
        for (int w = 0; w < nwarps; w++) { t += red_sum[w][c]; lo = fminf(lo, red_min[w][c]); hi = fmaxf(hi, red_max[w][c]); }
        row.ref[c] = ref[c];
        row.mean[c] = t / (float) A;                        // centroid - p_ref
        bound = fmaxf(bound, fmaxf(fabsf(hi - row.mean[c]), fabsf(lo - row.mean[c])));
    }
    const double inv = 1.0 / scale;
    row.bound = bound * (float) fabs(inv) * 1.000002f;      // head-room for the roundings on the way
    row.e0 = KIND == KIND_F16 ? exp_for_bound(row.bound) : 0;
    const double inv_scaled = ldexp(inv, row.e0);
    row.inv_hi = (float) inv_scaled;
    row.inv_lo = (float) (inv_scaled - (double) row.inv_hi);
    return row;
}

__device__ __forceinline__ float prep_feature(const PrepRow& row, float e, int c) {
    const float t = (e - row.ref[c]) - row.mean[c];
    return fmaf(t, row.inv_hi, t * row.inv_lo);
}

template <int KIND>
__global__ void __launch_bounds__(256, 7) prep_kernel(NetView net, int R, const float* __restrict__ coords, int N,
                                                  void* __restrict__ x_hi_v, void* __restrict__ x_lo_v, int ld, int contiguous,
                                                  ExpChain chain, int dbg) {
    extern __shared__ float xs[];                           // [3 * A]
    ANNF_TGRID_START(dbg);
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 0);
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 2);
    if (contiguous & 2) return;                             // experiments: launch-floor dry run
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;         // 128 or 256 (ANNF_PREP_THREADS)
    const int A = net.n_sel, quads = A / 4;
    const float* base = coords + (size_t) r * N * 3;
    __shared__ float red_sum[8][3], red_min[8][3], red_max[8][3];
    const float* ref_ptr = base + (size_t) ((contiguous & 1) ? 0 : net.sel_atoms[0]) * 3;
    const float ref[3] = {__ldg(ref_ptr), __ldg(ref_ptr + 1), __ldg(ref_ptr + 2)};
    float s[3] = {0.f, 0.f, 0.f};
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    // Coalesced walk (contiguous selection, 256 threads): a trip covers 768 consecutive 16-byte vectors, thread t owning vectors
    // t, t + 256 and t + 512 of it, so a warp's loads and stores are whole 512- / 256-byte runs (whole 32-byte sectors; the
    // 4-atom groups below touch every sector two or three times: 1.74 M store sectors for 0.58 M of data in
    // profiles/r02_prep.md).  768 is a multiple of 3, so element i of a thread's m-th vector always has component
    // (t + m + i) mod 3: statistics and constants are kept ROTATED by t mod 3, which keeps every index static.
    const bool coalesced = (contiguous & 4) != 0;
    const int rot = tid % 3, nvec = 3 * quads;
    if (coalesced) {
        const float4* src = reinterpret_cast<const float4*>(base);
        float4* dst = reinterpret_cast<float4*>(xs);
        const float refl[3] = {ref[rot], ref[(rot + 1) % 3], ref[(rot + 2) % 3]};
        float sl[3] = {0.f, 0.f, 0.f}, mnl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mxl[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (int q0 = tid; q0 < nvec; q0 += 768) {
            float4 p[3];
#pragma unroll
            for (int m = 0; m < 3; m++)
                if (q0 + 256 * m < nvec) p[m] = __ldg(src + q0 + 256 * m);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                if (q0 + 256 * m < nvec) {
                    dst[q0 + 256 * m] = p[m];
                    const float e[4] = {p[m].x, p[m].y, p[m].z, p[m].w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int j = (i + m) % 3;
                        const float d = e[i] - refl[j];
                        sl[j] += d;
                        mnl[j] = fminf(mnl[j], d);
                        mxl[j] = fmaxf(mxl[j], d);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {                       // rotate back: component c sits at local index (c - rot) mod 3
            const int j = (c + 3 - rot) % 3;
            s[c] = j == 0 ? sl[0] : (j == 1 ? sl[1] : sl[2]);
            mn[c] = j == 0 ? mnl[0] : (j == 1 ? mnl[1] : mnl[2]);
            mx[c] = j == 0 ? mxl[0] : (j == 1 ? mxl[1] : mxl[2]);
        }
    }
    for (int g = tid; g < quads && !coalesced; g += nthreads) {
        float e[12];
        if (contiguous & 1) {
            // selection is atoms 0..A-1 and rows are 16-byte aligned: three vector loads per 4 atoms
            const float4* src = reinterpret_cast<const float4*>(base) + 3 * g;
            const float4 p0 = __ldg(src), p1 = __ldg(src + 1), p2 = __ldg(src + 2);
            e[0] = p0.x; e[1] = p0.y; e[2] = p0.z; e[3] = p0.w; e[4] = p1.x; e[5] = p1.y; e[6] = p1.z; e[7] = p1.w;
            e[8] = p2.x; e[9] = p2.y; e[10] = p2.z; e[11] = p2.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float* p = base + (size_t) net.sel_atoms[4 * g + k] * 3;
                e[3 * k] = __ldg(p); e[3 * k + 1] = __ldg(p + 1); e[3 * k + 2] = __ldg(p + 2);
            }
        }
        float4* dst = reinterpret_cast<float4*>(xs) + 3 * g;
        dst[0] = make_float4(e[0], e[1], e[2], e[3]);
        dst[1] = make_float4(e[4], e[5], e[6], e[7]);
        dst[2] = make_float4(e[8], e[9], e[10], e[11]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d0 = e[c] - ref[c], d1 = e[3 + c] - ref[c], d2 = e[6 + c] - ref[c], d3 = e[9 + c] - ref[c];
            s[c] += (d0 + d1) + (d2 + d3);
            mn[c] = fminf(fminf(mn[c], fminf(d0, d1)), fminf(d2, d3));
            mx[c] = fmaxf(fmaxf(mx[c], fmaxf(d0, d1)), fmaxf(d2, d3));
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        s[c] = warp_sum(s[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) { red_sum[warp][c] = s[c]; red_min[warp][c] = mn[c]; red_max[warp][c] = mx[c]; }
    }
    __syncthreads();
    // warp 0 finishes the reduction and publishes the row's constants; a second barrier is cheaper than the 72 broadcast
    // reads + arithmetic of every thread finishing it for itself (the kernel is instruction-issue bound: profiles/r02_prep.md)
    __shared__ PrepRow row_s;
    if (warp == 0) {
        const PrepRow r0 = prep_finish_reduction<KIND>(red_sum, red_min, red_max, nwarps, ref, A, net.scale);
        if (lane == 0) row_s = r0;
    }
    __syncthreads();
    const PrepRow row = row_s;
    if (KIND == KIND_F16 && tid == 0) {
        chain.x_exp[0][r] = row.e0;
        float b = row.bound;
        for (int l = 0; l + 2 < net.L; l++) {                // X_{l+1} = act(W_l X_l + b_l)
            b = net.types[l] == ANNF_TANH ? 1.f : chain.fwd_bias[l] + b * chain.fwd_gain[l];
            chain.x_exp[l + 1][r] = exp_for_bound(b);
        }
    }
    if (coalesced) {
        PrepRow rowl = row;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int c = (rot + j) % 3;
            rowl.ref[j] = c == 0 ? row.ref[0] : (c == 1 ? row.ref[1] : row.ref[2]);
            rowl.mean[j] = c == 0 ? row.mean[0] : (c == 1 ? row.mean[1] : row.mean[2]);
        }
        const float4* src = reinterpret_cast<const float4*>(xs);
        for (int q0 = tid; q0 < nvec; q0 += 768) {
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const int q = q0 + 256 * m;
                if (q < nvec) {
                    const float4 pv = src[q];
                    const float4 x = make_float4(prep_feature(rowl, pv.x, m % 3), prep_feature(rowl, pv.y, (m + 1) % 3),
                                                 prep_feature(rowl, pv.z, (m + 2) % 3), prep_feature(rowl, pv.w, m % 3));
                    if (KIND == KIND_F16) {
                        uint2 hi, lo;
                        split_f16x4(x, hi, lo);
                        reinterpret_cast<uint2*>(static_cast<__half*>(x_hi_v) + (size_t) r * ld)[q] = hi;
                        reinterpret_cast<uint2*>(static_cast<__half*>(x_lo_v) + (size_t) r * ld)[q] = lo;
                    } else {
                        float hi[4], lo[4];
                        split_tf32(x.x, hi[0], lo[0]); split_tf32(x.y, hi[1], lo[1]);
                        split_tf32(x.z, hi[2], lo[2]); split_tf32(x.w, hi[3], lo[3]);
                        reinterpret_cast<float4*>(static_cast<float*>(x_hi_v) + (size_t) r * ld)[q] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        reinterpret_cast<float4*>(static_cast<float*>(x_lo_v) + (size_t) r * ld)[q] = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
        }
    }
    for (int g = tid; g < quads && !coalesced; g += nthreads) {
        const float4* src = reinterpret_cast<const float4*>(xs) + 3 * g;
        const float4 p0 = src[0], p1 = src[1], p2 = src[2];
        const float e[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
        float x[12];
#pragma unroll
        for (int j = 0; j < 12; j++) x[j] = prep_feature(row, e[j], j % 3);
        if (KIND == KIND_F16) {
            uint2* oh = reinterpret_cast<uint2*>(static_cast<__half*>(x_hi_v) + (size_t) r * ld) + 3 * g;
            uint2* ol = reinterpret_cast<uint2*>(static_cast<__half*>(x_lo_v) + (size_t) r * ld) + 3 * g;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                uint2 hi, lo;
                split_f16x4(make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]), hi, lo);
                oh[q] = hi;
                ol[q] = lo;
            }
        } else {
            float4* oh = reinterpret_cast<float4*>(static_cast<float*>(x_hi_v) + (size_t) r * ld) + 3 * g;
            float4* ol = reinterpret_cast<float4*>(static_cast<float*>(x_lo_v) + (size_t) r * ld) + 3 * g;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; j++) split_tf32(x[4 * q + j], hi[j], lo[j]);
                oh[q] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                ol[q] = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 9);
    ANNF_TGRID_END(dbg);
}

// prep for a contiguous selection (atoms 0..A-1, 16-byte aligned rows) — the default on that layout.  One CTA of W warps per
// replica; the whole row of coordinates arrives in shared memory by ONE bulk copy (TMA engine, mbarrier), so a row exposes
// a single memory latency instead of one per trip of loads, and pass 1 costs no load / store-to-shared instructions at all.
// W = 1 needs no block barrier (shuffles only) and pays the per-thread fixed work (prologue, reductions, constants: about as
// many instructions as the 17 elements a thread of prep_kernel owns) once per row instead of eight times; more warps per row
// shorten the serial chain of one row.  Pass 2 is the coalesced walk of prep_kernel: trips of 3 * T vectors, thread t owning
// vectors t, t + T, t + 2T of a trip, T = 32 W; element i of the m-th vector has component (t + (T mod 3) m + i) mod 3.
// Measured at BASELINE C5 (18 MB in, 18 MB out; profiles/r02_prep.md).
template <int KIND, int W>
__global__ void __launch_bounds__(32 * W) prep_bulk_kernel(NetView net, int R, const float* __restrict__ coords, int N,
                                                        void* __restrict__ x_hi_v, void* __restrict__ x_lo_v, int ld, ExpChain chain, int dbg) {
    extern __shared__ __align__(16) float xs[];             // [3 * A]
    __shared__ uint64_t bar;
    __shared__ float red[W][9];
    constexpr int T = 32 * W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x, A = net.n_sel, quads = A / 4, nvec = 3 * quads;
    ANNF_TGRID_START(dbg);
    if (tid == 0) ANNF_TMARK(dbg, 0);
    pdl_launch_dependents();
    if (tid == 0) {
        ptx::bar_init(ptx::s_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (W > 1) __syncthreads(); else __syncwarp();
    pdl_wait();
    if (tid == 0) {
        ANNF_TMARK(dbg, 2);
        ptx::bar_expect_tx(ptx::s_u32(&bar), (uint32_t) (12 * A));
        // the coordinates are read once (evict_first); the operand pair is rewritten in place every step (evict_last)
        bulk_load_h<(kL2Hints & 1) != 0>(ptx::s_u32(xs), coords + (size_t) r * N * 3, (uint32_t) (12 * A), ptx::s_u32(&bar), (kL2Hints & 1) ? l2_evict_first() : 0);
    }
    ptx::bar_wait(ptx::s_u32(&bar), 0);
    if (tid == 0) ANNF_TMARK(dbg, 3);
    const float ref[3] = {xs[0], xs[1], xs[2]};
    float s[3] = {0.f, 0.f, 0.f};
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int g = tid; g < quads; g += T) {                  // 4 atoms = 3 vectors, 48-byte stride: conflict-free shared loads
        const float4* src = reinterpret_cast<const float4*>(xs) + 3 * g;
        const float4 p0 = src[0], p1 = src[1], p2 = src[2];
        const float e[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d0 = e[c] - ref[c], d1 = e[3 + c] - ref[c], d2 = e[6 + c] - ref[c], d3 = e[9 + c] - ref[c];
            s[c] += (d0 + d1) + (d2 + d3);
            mn[c] = fminf(fminf(mn[c], fminf(d0, d1)), fminf(d2, d3));
            mx[c] = fmaxf(fmaxf(mx[c], fmaxf(d0, d1)), fmaxf(d2, d3));
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {                           // every lane ends up with the warp's sums, minima, maxima
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
        mn[c] = warp_min_float(mn[c]);                      // one redux.sync each on order-preserving integer keys
        mx[c] = warp_max_float(mx[c]);
    }
    if (W > 1) {
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) { red[warp][c] = s[c]; red[warp][3 + c] = mn[c]; red[warp][6 + c] = mx[c]; }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 3; c++) {
            s[c] = red[0][c]; mn[c] = red[0][3 + c]; mx[c] = red[0][6 + c];
#pragma unroll
            for (int w = 1; w < W; w++) { s[c] += red[w][c]; mn[c] = fminf(mn[c], red[w][3 + c]); mx[c] = fmaxf(mx[c], red[w][6 + c]); }
        }
    }
    PrepRow row;
    float bound = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        row.ref[c] = ref[c];
        row.mean[c] = s[c] / (float) A;                     // centroid - p_ref
        bound = fmaxf(bound, fmaxf(fabsf(mx[c] - row.mean[c]), fabsf(mn[c] - row.mean[c])));
    }
    row.bound = bound * fabsf(chain.inv_hi) * 1.000002f;    // head-room for the roundings on the way
    row.e0 = KIND == KIND_F16 ? exp_for_bound(row.bound) : 0;
    row.inv_hi = chain.inv_hi * exp2i(row.e0);              // |e0| <= 60: exact
    row.inv_lo = chain.inv_lo * exp2i(row.e0);
    if (KIND == KIND_F16 && tid == 0) {
        chain.x_exp[0][r] = row.e0;
        float b = row.bound;
        for (int l = 0; l + 2 < net.L; l++) {                // X_{l+1} = act(W_l X_l + b_l)
            b = net.types[l] == ANNF_TANH ? 1.f : chain.fwd_bias[l] + b * chain.fwd_gain[l];
            chain.x_exp[l + 1][r] = exp_for_bound(b);
        }
    }
    if (tid == 0) ANNF_TMARK(dbg, 4);
    const int rot = tid % 3;
    PrepRow rowl = row;                                     // constants rotated so that local index j means component (rot + j) mod 3
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int c = (rot + j) % 3;
        rowl.ref[j] = c == 0 ? row.ref[0] : (c == 1 ? row.ref[1] : row.ref[2]);
        rowl.mean[j] = c == 0 ? row.mean[0] : (c == 1 ? row.mean[1] : row.mean[2]);
    }
    const float4* src = reinterpret_cast<const float4*>(xs);
    constexpr bool kHintX = (kL2Hints & 4) != 0;
    const uint64_t pol_x = kHintX ? l2_evict_last() : 0;
    for (int q0 = tid; q0 < nvec; q0 += 3 * T) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const int q = q0 + T * m;
            constexpr int step = T % 3;
            if (q < nvec) {
                const float4 pv = src[q];
                const float4 x = make_float4(prep_feature(rowl, pv.x, (step * m) % 3), prep_feature(rowl, pv.y, (step * m + 1) % 3),
                                             prep_feature(rowl, pv.z, (step * m + 2) % 3), prep_feature(rowl, pv.w, (step * m) % 3));
                if (KIND == KIND_F16) {
                    uint2 hi, lo;
                    split_f16x4(x, hi, lo);
                    st_v2_h<kHintX>(reinterpret_cast<uint2*>(static_cast<__half*>(x_hi_v) + (size_t) r * ld) + q, hi, pol_x);
                    st_v2_h<kHintX>(reinterpret_cast<uint2*>(static_cast<__half*>(x_lo_v) + (size_t) r * ld) + q, lo, pol_x);
                } else {
                    float hi[4], lo[4];
                    split_tf32(x.x, hi[0], lo[0]); split_tf32(x.y, hi[1], lo[1]);
                    split_tf32(x.z, hi[2], lo[2]); split_tf32(x.w, hi[3], lo[3]);
                    st_v4_h<kHintX>(reinterpret_cast<float4*>(static_cast<float*>(x_hi_v) + (size_t) r * ld) + q, make_float4(hi[0], hi[1], hi[2], hi[3]), pol_x);
                    st_v4_h<kHintX>(reinterpret_cast<float4*>(static_cast<float*>(x_lo_v) + (size_t) r * ld) + q, make_float4(lo[0], lo[1], lo[2], lo[3]), pol_x);
                }
            }
        }
    }
    if (tid == 0) ANNF_TMARK(dbg, 9);
    ANNF_TGRID_END(dbg);
}

template <int KIND>
static bool prep_bulk_attributes() {
    const int bytes = 200 * 1024;
    return cudaFuncSetAttribute(prep_bulk_kernel<KIND, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess &&
           cudaFuncSetAttribute(prep_bulk_kernel<KIND, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess &&
           cudaFuncSetAttribute(prep_bulk_kernel<KIND, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess &&
           cudaFuncSetAttribute(prep_bulk_kernel<KIND, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
}

// prep, persistent flavour (KIND_F16, selection = atoms 0..A-1, 16-byte aligned rows): 2 CTAs per SM walk the replicas; the
// NEXT row of coordinates arrives by one bulk copy (TMA engine, mbarrier) while the current one is reduced, centred, scaled
// and split.  Opt-in (ANNF_PREP_PIPELINED=1): measured SLOWER than one CTA per replica at BASELINE C5 (step 130.9 against
// 112.9 us, same session) — a row is a chain of dependent phases, and 16 warps per SM hide less of it than the 40 - 56 warps
// of seven independent one-row CTAs (profiles/r02_prep.md).
constexpr int kPrepPipeThreads = 256;
__global__ void __launch_bounds__(kPrepPipeThreads) prep_pipelined_kernel(NetView net, int R, const float* __restrict__ coords, int N,
                                                                       __half* __restrict__ x_hi, __half* __restrict__ x_lo, int ld,
                                                                       ExpChain chain, int dbg) {
    extern __shared__ __align__(128) unsigned char prep_smem[];
    const int A = net.n_sel, quads = A / 4, row_bytes = 12 * A;      // 3A floats; A % 4 == 0, so a row is a multiple of 48 bytes
    const int buf_stride = (row_bytes + 127) & ~127;
    __shared__ uint64_t bar[2];
    __shared__ float red_sum[8][3], red_min[8][3], red_max[8][3];
    ANNF_TGRID_START(dbg);
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 0);
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        ptx::bar_init(ptx::s_u32(&bar[0]), 1);
        ptx::bar_init(ptx::s_u32(&bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 2);
    int r = blockIdx.x;
    if (r < R && tid == 0) {
        ptx::bar_expect_tx(ptx::s_u32(&bar[0]), (uint32_t) row_bytes);
        ptx::bulk_load(ptx::s_u32(prep_smem), coords + (size_t) r * N * 3, (uint32_t) row_bytes, ptx::s_u32(&bar[0]));
    }
    for (int i = 0; r < R; i++, r += gridDim.x) {
        const int cur = i & 1, rn = r + gridDim.x;
        if (rn < R && tid == 0) {                           // buffer cur ^ 1 was released by the barrier that ended iteration i - 1
            ptx::bar_expect_tx(ptx::s_u32(&bar[cur ^ 1]), (uint32_t) row_bytes);
            ptx::bulk_load(ptx::s_u32(prep_smem + (size_t) (cur ^ 1) * buf_stride), coords + (size_t) rn * N * 3, (uint32_t) row_bytes, ptx::s_u32(&bar[cur ^ 1]));
        }
        ptx::bar_wait(ptx::s_u32(&bar[cur]), (uint32_t) (i >> 1) & 1u);
        const float4* src = reinterpret_cast<const float4*>(prep_smem + (size_t) cur * buf_stride);
        const float* first = reinterpret_cast<const float*>(src);
        const float ref[3] = {first[0], first[1], first[2]};
        float s[3] = {0.f, 0.f, 0.f};
        float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (int g = tid; g < quads; g += kPrepPipeThreads) {
            const float4 p0 = src[3 * g], p1 = src[3 * g + 1], p2 = src[3 * g + 2];
            const float e[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float d0 = e[c] - ref[c], d1 = e[3 + c] - ref[c], d2 = e[6 + c] - ref[c], d3 = e[9 + c] - ref[c];
                s[c] += (d0 + d1) + (d2 + d3);
                mn[c] = fminf(fminf(mn[c], fminf(d0, d1)), fminf(d2, d3));
                mx[c] = fmaxf(fmaxf(mx[c], fmaxf(d0, d1)), fmaxf(d2, d3));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            s[c] = warp_sum(s[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) { red_sum[warp][c] = s[c]; red_min[warp][c] = mn[c]; red_max[warp][c] = mx[c]; }
        }
        __syncthreads();
        const PrepRow row = prep_finish_reduction<KIND_F16>(red_sum, red_min, red_max, kPrepPipeThreads / 32, ref, A, net.scale);
        if (tid == 0) {
            chain.x_exp[0][r] = row.e0;
            float b = row.bound;
            for (int l = 0; l + 2 < net.L; l++) {
                b = net.types[l] == ANNF_TANH ? 1.f : chain.fwd_bias[l] + b * chain.fwd_gain[l];
                chain.x_exp[l + 1][r] = exp_for_bound(b);
            }
        }
        uint2* oh = reinterpret_cast<uint2*>(x_hi + (size_t) r * ld);
        uint2* ol = reinterpret_cast<uint2*>(x_lo + (size_t) r * ld);
        for (int g = tid; g < quads; g += kPrepPipeThreads) {
            const float4 p0 = src[3 * g], p1 = src[3 * g + 1], p2 = src[3 * g + 2];
            const float e[12] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
            float x[12];
#pragma unroll
            for (int j = 0; j < 12; j++) x[j] = prep_feature(row, e[j], j % 3);
#pragma unroll
            for (int q = 0; q < 3; q++) {
                uint2 hi, lo;
                split_f16x4(make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]), hi, lo);
                oh[3 * g + q] = hi;
                ol[3 * g + q] = lo;
            }
        }
        __syncthreads();                                    // buf[cur] and the reduction scratch are free again
    }
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 9);
    ANNF_TGRID_END(dbg);
}

// activation m of a row stored as an operand pair: hi + lo is exact in fp32, and so is the power-of-two unscale
template <int KIND>
__device__ __forceinline__ float load_pair(const void* hi, const void* lo, int m, float unscale) {
    if (KIND == KIND_F16)
        return (__half2float(__ldg(static_cast<const __half*>(hi) + m)) + __half2float(__ldg(static_cast<const __half*>(lo) + m))) * unscale;
    return __ldg(static_cast<const float*>(hi) + m) + __ldg(static_cast<const float*>(lo) + m);
}

// head: output layer forward (fp64 accumulate), umbrella energy + gradient (output_head), and the
// backward contraction through the output connection, fused with act'() of the last hidden layer.
// One 128-thread CTA per replica: every thread owns hidden units tid, tid+128, ... so the dependent
// chains are short; the nL output sums are reduced with shuffles + one shared-memory hop.
// (Tried and rejected on measurement: a warp-per-replica variant, 8 replicas per CTA with the output connection staged in
// shared memory.  Fully unrolled it was instruction-fetch bound — 58 % stall_no_inst, ~170 cycles per 128-byte line of code
// that runs once per CTA; rolled, its 128 fat CTAs delayed the cluster GEMM that follows under PDL: step 110 -> 121 us.)
constexpr int kHeadThreads = 128;
constexpr int kHeadMaxOut = 64;
template <int KIND>
__global__ void __launch_bounds__(kHeadThreads, 7) head_kernel(NetView net, int R, const float* __restrict__ centers,
                                                           const void* __restrict__ a_hi, const void* __restrict__ a_lo, int ld,
                                                           void* __restrict__ d_hi, void* __restrict__ d_lo, double* __restrict__ energies,
                                                           ExpChain chain, int dbg) {
    typedef typename KindOf<KIND>::elem elem;
    ANNF_TGRID_START(dbg);
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 0);
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 2);
    if (R < 0) return;                                      // experiments: launch-floor dry run
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r_index = r;
    const int L = net.L, nH = net.nodes[L - 2], nL = net.nodes[L - 1];
    __shared__ double part[4][kHeadMaxOut];
    __shared__ double zs[kHeadMaxOut], dzs[kHeadMaxOut], cen[kHeadMaxOut];
    __shared__ float dzf[kHeadMaxOut];
    const double* W = net.W64 + net.w_off[L - 2];           // [nL][nH]
    const float* W32 = net.W32 + net.w_off[L - 2];
    const double* b = net.b64 + net.b_off[L - 2];
    const elem* ah = static_cast<const elem*>(a_hi) + (size_t) r * ld;
    const elem* al = static_cast<const elem*>(a_lo) + (size_t) r * ld;
    const float a_unscale = KIND == KIND_F16 ? exp2i(-chain.x_exp[L - 2][r]) : 1.f;
    // Every thread owns FOUR CONSECUTIVE hidden units per pass (nH is a multiple of 4 on this path): the activation pair arrives
    // as two 8-byte loads, an output's weights as two 16-byte loads — a quarter of the load instructions of a strided mapping,
    // and whole 256-byte lines per warp.  (The kernel is instruction-issue bound: 5.0 M warp-instructions for 8.4 MFLOP before.)
    auto load_act4 = [&](int m4, float (&av)[4]) {
        if (KIND == KIND_F16) {
            const float4 h = half4_to_float4(__ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(ah) + m4)));
            const float4 l = half4_to_float4(__ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(al) + m4)));
            av[0] = (h.x + l.x) * a_unscale; av[1] = (h.y + l.y) * a_unscale; av[2] = (h.z + l.z) * a_unscale; av[3] = (h.w + l.w) * a_unscale;
        } else {
            const float4 h = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ah) + m4));
            const float4 l = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(al) + m4));
            av[0] = h.x + l.x; av[1] = h.y + l.y; av[2] = h.z + l.z; av[3] = h.w + l.w;      // hi + lo is exact in fp32
        }
    };
    for (int j0 = 0; j0 < nL; j0 += 8) {                    // 8 outputs at a time: 8 independent accumulators per thread
        double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int m4 = 4 * tid; m4 < nH; m4 += 4 * kHeadThreads) {
            float av[4];
            load_act4(m4, av);
            double2 w[8][2];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const bool in = j0 + u < nL;
                const double2* wp = reinterpret_cast<const double2*>(W + (size_t) (j0 + u) * nH + m4);
                w[u][0] = in ? __ldg(wp) : make_double2(0.0, 0.0);
                w[u][1] = in ? __ldg(wp + 1) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                acc[u] = fma(w[u][0].x, (double) av[0], acc[u]);
                acc[u] = fma(w[u][0].y, (double) av[1], acc[u]);
                acc[u] = fma(w[u][1].x, (double) av[2], acc[u]);
                acc[u] = fma(w[u][1].y, (double) av[3], acc[u]);
            }
        }
        // Eight warp-wide sums as ONE transposed butterfly: every round halves the values a lane carries (it keeps one half and
        // sends the other to its partner), so 18 64-bit shuffles + 9 adds replace the 80 + 40 of eight separate butterflies
        // (the shuffle unit was the busiest pipe of this kernel).  Lane l ends with the sum of output u = bits 4, 3, 2 of l.
        {
            double b4[4], b2[2];
            const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0, up4 = (lane & 4) != 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double send = up16 ? acc[i] : acc[i + 4], keep = up16 ? acc[i + 4] : acc[i];
                b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double send = up8 ? b4[i] : b4[i + 2], keep = up8 ? b4[i + 2] : b4[i];
                b2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            const double send = up4 ? b2[0] : b2[1], keep = up4 ? b2[1] : b2[0];
            double d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            const int u = (lane >> 2) & 7;                  // bit 4 -> u bit 2, bit 3 -> u bit 1, bit 2 -> u bit 0
            if ((lane & 3) == 0 && j0 + u < nL) part[warp][j0 + u] = d;
        }
    }
    for (int c = tid; c < net.num_center; c += kHeadThreads)
        cen[c] = centers ? (double) centers[(size_t) r * net.num_center + c] : net.center[c];
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 3);
    __syncthreads();
    if (tid < nL) zs[tid] = ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid])) + b[tid];
    __syncthreads();
    // energy + output gradient: one output (or Circular pair) per thread, terms summed in a fixed order
    if (tid < kHeadMaxOut) part[0][tid] = output_head_parallel(net.types[L - 2], nL, zs, 1, cen, *net.k_dev, dzs, 1, tid);
    __syncthreads();
    if (tid == 0) {
        double e = 0.0;
        for (int j = 0; j < nL; j++) e += part[0][j];
        energies[r] = e;
    }
    if (tid < nL) dzf[tid] = (float) dzs[tid];
    __syncthreads();
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 4);
    const int hidden_type = net.types[L - 3];
    // KIND_F16: scale exponent of this replica's gradient rows, from a bound: |D[m]| <= max_j |dz_j| * max_m sum_j |W[j][m]|
    float d_scale = 1.f;
    if (KIND == KIND_F16) {
        float mx = 0.f;
        for (int j = 0; j < nL; j++) mx = fmaxf(mx, fabsf(dzf[j]));
        float bound = mx * chain.bwd_gain[L - 2];
        const int e_d = exp_for_bound(bound);
        d_scale = exp2i(e_d);
        if (tid == 0) {
            chain.d_exp[L - 2][r] = e_d;
            for (int l = L - 3; l >= 1; l--) {               // D_l = (D_{l+1} W_l) (.) act'(X_l), |act'| <= 1
                bound *= chain.bwd_gain[l];
                chain.d_exp[l][r] = exp_for_bound(bound);
            }
        }
    }
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 5);
    // the result is re-split to operand pairs: fp32 FMA is enough here.  Four consecutive hidden units per thread, vector loads.
    for (int m4 = 4 * tid; m4 < nH; m4 += 4 * kHeadThreads) {
        float av[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
        load_act4(m4, av);
        for (int j = 0; j < nL; j++) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(W32 + (size_t) j * nH + m4));
            const float dj = dzf[j];
            acc[0] = fmaf(w.x, dj, acc[0]); acc[1] = fmaf(w.y, dj, acc[1]); acc[2] = fmaf(w.z, dj, acc[2]); acc[3] = fmaf(w.w, dj, acc[3]);
        }
        if (hidden_type == ANNF_TANH) {
#pragma unroll
            for (int v = 0; v < 4; v++) acc[v] *= (1.f - av[v] * av[v]);
        }
        if (KIND == KIND_F16) {
            uint2 hi, lo;
            split_f16x4(make_float4(acc[0] * d_scale, acc[1] * d_scale, acc[2] * d_scale, acc[3] * d_scale), hi, lo);
            *reinterpret_cast<uint2*>(static_cast<__half*>(d_hi) + (size_t) r_index * ld + m4) = hi;
            *reinterpret_cast<uint2*>(static_cast<__half*>(d_lo) + (size_t) r_index * ld + m4) = lo;
        } else {
            float hi[4], lo[4];
#pragma unroll
            for (int v = 0; v < 4; v++) split_tf32(acc[v], hi[v], lo[v]);
            *reinterpret_cast<float4*>(static_cast<float*>(d_hi) + (size_t) r_index * ld + m4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(static_cast<float*>(d_lo) + (size_t) r_index * ld + m4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (threadIdx.x == 0) ANNF_TMARK(dbg, 9);
    ANNF_TGRID_END(dbg);
}

// ------------------------------------------------------------------------------------------------
// Fused middle of the step (KIND_F16): forward through the last hidden connection, output head, backward through the same
// connection.  As three kernels (GEMM, head, GEMM) this was 37 of the 90 us of a BASELINE C5 step for 10 % of its flops —
// three prologues, three first-operand latencies, two launch gaps, a 1024-CTA SIMT kernel in two waves
// (profiles/r02_c5_phases_session3.txt).  Here a cluster of up to 4 CTAs owns one tile of 128 replicas end to end:
//   phase 1   CTA j:  A2[:, 128j : 128j+128] = act(X1 W^T + b)       tcgen05 main loop, K = n1
//   head      every thread-pair of a row: partial output-layer sums over the CTA's 128 columns (fp64), pushed to every CTA of
//             the cluster through DSMEM, summed in a fixed order; umbrella energy + dE/dz (output_head_parallel);
//             D2[:, 128j : ...] = (W_L^T dz) (.) act'(A2), scaled and split to fp16 pairs, written to global
//   phase 2   CTA j:  D1[:, tiles j, j+CL, ..] = (D2 W) (.) act'(X1)   tcgen05 main loop, K = n2 (A through TMA from what the
//             cluster just wrote: fence.proxy.async + cluster barrier)
// Warp roles: 8 epilogue warps (two per TMEM lane quarter, 64 accumulator columns each), TMA producer, MMA issuer, a loader
// for the per-tile vectors, one spare.  One K-block per TMEM chunk, four accumulator buffers.
// ------------------------------------------------------------------------------------------------
constexpr int kMidCluster = 4;
constexpr int kMidEpiWarps = 8;
constexpr int kMidThreads = (kMidEpiWarps + 4) * 32;
constexpr int kMidStages = 3;
constexpr int kMidStageBytes = 4 * BM * kBlockBytes;            // A hi, A lo, B hi, B lo: 128 rows x 128 bytes each
constexpr int kMidDataBytes = kMidStages * kMidStageBytes;      // 192 KB; doubles as the head's scratch and the epilogues' staging
constexpr int kMidMaxOut = 16;                                  // output nodes the fused head supports
constexpr int kMidVecBytes = (2 * 128 + 3 * BM) * 4;
constexpr int kMidSmemBytes = kMidDataBytes + 1024 + 256 + kMidVecBytes;
// head scratch inside `data` (bytes); zpart is dead once zfull exists and is re-used for dz / energy terms
constexpr int kMidStagLd = 132;
constexpr int kMidOffZpart = 68 * 1024;                                        // [4][16][128] doubles = 64 KB
constexpr int kMidOffZh = kMidOffZpart + kMidCluster * kMidMaxOut * BM * 8;   // [16][128] doubles
constexpr int kMidOffZfull = kMidOffZh + kMidMaxOut * BM * 8;                  // [16][128] doubles
constexpr int kMidOffCen = kMidOffZfull + kMidMaxOut * BM * 8;                 // [128][17] doubles
constexpr int kMidCenLd = kMidMaxOut + 1;
static_assert(BM * kMidStagLd * 4 <= kMidOffZpart, "staging tile overlaps the head scratch");
static_assert(kMidOffCen + BM * kMidCenLd * 8 <= kMidDataBytes, "head scratch exceeds the stage buffers");
constexpr int kMidOffDz = kMidOffZpart;                                        // [16][128] doubles   (aliases zpart)
constexpr int kMidOffEpart = kMidOffDz + kMidMaxOut * BM * 8;                  // [16][128] doubles
constexpr int kMidOffDzf = kMidOffEpart + kMidMaxOut * BM * 8;                 // [16][128] floats

struct MidParams {
    int M, L;
    int n1, n2, nL;                   // widths of X_{L-3}, X_{L-2} and of the output layer
    int hidden_type, prev_type, out_type, num_center;
    const float* bias;                // [n2]   connection L-3
    const int* w_exp;                 // [n2]   row exponents of W_{L-3}        (B of phase 1)
    const int* wt_exp;                // [n1]   row exponents of W_{L-3}^T      (B of phase 2)
    const int* x1_exp;                // [M]    row exponents of X_{L-3}        (A of phase 1, aux of phase 2)
    const __half *x1_hi, *x1_lo;      // [M][ld1]
    int ld1;
    __half *d2_hi, *d2_lo;            // [M][ld2]   D_{L-2}, written here and read back through TMA
    int ld2;
    __half *d1_hi, *d1_lo;            // [M][ld1]   D_{L-3}
    const double* WL;                 // [nL][n2]   output connection, fp64
    const float* WL32;
    const double* bL;                 // [nL]
    const float* centers;             // [M][num_center] or nullptr
    const double* center_dev;         // [num_center]
    const double* k_dev;
    double* energies;                 // [M]
    int dbg;
};

__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64(uint32_t cluster_addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(cluster_addr), "d"(v) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(32 * kMidEpiWarps) : "memory"); }

// NLP = output nodes rounded up to 4, 8 or 16: the register accumulators of the partial output-layer sums
template <int NLP>
__global__ void __launch_bounds__(kMidThreads, 1)
mid_kernel(const __grid_constant__ CUtensorMap tm_x1_hi, const __grid_constant__ CUtensorMap tm_x1_lo,
           const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
           const __grid_constant__ CUtensorMap tm_d2_hi, const __grid_constant__ CUtensorMap tm_d2_lo,
           const __grid_constant__ CUtensorMap tm_wt_hi, const __grid_constant__ CUtensorMap tm_wt_lo,
           const MidParams p, const ExpChain chain) {
    constexpr int BK = KindOf<KIND_F16>::kBlockElems;
    constexpr int NB = 4;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* data = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(data + kMidDataBytes);
    uint64_t* bar_full = bars;                              // [kMidStages]
    uint64_t* bar_empty = bars + kMidStages;                // [kMidStages]
    uint64_t* bar_acc_full = bars + 2 * kMidStages;         // [NB]
    uint64_t* bar_acc_empty = bars + 2 * kMidStages + NB;   // [NB]
    uint64_t* bar_epi_free = bars + 2 * kMidStages + 2 * NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMidStages + 2 * NB + 1);
    int* s_b_exp = reinterpret_cast<int*>(data + kMidDataBytes + 256);
    float* s_bias = reinterpret_cast<float*>(s_b_exp + 128);
    int* s_a_exp = reinterpret_cast<int*>(s_bias + 128);
    int* s_out_exp = s_a_exp + BM;
    int* s_aux_exp = s_out_exp + BM;
    const EpiVec vec = {s_b_exp, s_bias, s_a_exp, s_out_exp, s_aux_exp};

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CL = (int) cluster_nctarank();
    const int rank = (int) cluster_ctarank();               // = blockIdx.x
    const int m0 = (int) blockIdx.y * BM;
    const int n0 = rank * 128;                              // this CTA's columns of X_{L-2} / D_{L-2}
    const int nkb1 = (p.n1 + BK - 1) / BK, nkb2 = (p.n2 + BK - 1) / BK;
    const int tiles1 = (p.n1 + 127) / 128;                  // output tiles of phase 2; this CTA takes rank, rank + CL, ...

    ANNF_TGRID_START(p.dbg);
    if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 0);
    pdl_launch_dependents();
    if (warp == kMidEpiWarps && lane == 0) {
        tma_prefetch_desc(&tm_x1_hi); tma_prefetch_desc(&tm_x1_lo); tma_prefetch_desc(&tm_w_hi); tma_prefetch_desc(&tm_w_lo);
        tma_prefetch_desc(&tm_d2_hi); tma_prefetch_desc(&tm_d2_lo); tma_prefetch_desc(&tm_wt_hi); tma_prefetch_desc(&tm_wt_lo);
        for (int s = 0; s < kMidStages; s++) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int b = 0; b < NB; b++) { mbar_init(smem_u32(&bar_acc_full[b]), 1); mbar_init(smem_u32(&bar_acc_empty[b]), kMidEpiWarps); }
        mbar_init(smem_u32(bar_epi_free), kMidEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMidEpiWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t) (NB * 128)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == kMidEpiWarps + 2) {
        for (int c = lane; c < 128; c += 32) {              // static vectors of the phase-1 tile
            const int n = n0 + c;
            s_b_exp[c] = n < p.n2 ? __ldg(p.w_exp + n) : 0;
            s_bias[c] = n < p.n2 ? __ldg(p.bias + n) : 0.f;
        }
    }
    tcgen05_fence_before();
    cluster_barrier();                                      // peers push into this CTA's shared memory later: it must be live
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 1);
    pdl_wait();
    if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 2);
    if (warp == kMidEpiWarps + 2) {
        for (int r = lane; r < BM; r += 32) s_aux_exp[r] = (m0 + r) < p.M ? p.x1_exp[m0 + r] : 0;
        __syncwarp();
        asm volatile("bar.arrive 1, %0;" ::"n"(32 * (kMidEpiWarps + 1)) : "memory");
    }

    // ---- pipeline state of the three roles (each role's counters live in its own threads) ----
    int it = 0;                                             // K-blocks issued / consumed so far (stage ring position)
    int ch = 0;                                             // TMEM chunks issued / drained so far (accumulator ring position)
    auto produce = [&](const CUtensorMap* a_hi, const CUtensorMap* a_lo, const CUtensorMap* b_hi, const CUtensorMap* b_lo, int nkb, int b_row0) {
        for (int kb = 0; kb < nkb; kb++, it++) {
            const int s = it % kMidStages;
            mbar_wait(smem_u32(&bar_empty[s]), (((uint32_t) (it / kMidStages)) & 1u) ^ 1u);
            const uint32_t full = smem_u32(&bar_full[s]);
            const uint32_t base = smem_u32(data + (size_t) s * kMidStageBytes);
            mbar_expect_tx(full, (uint32_t) kMidStageBytes);
            tma_load_2d(base, a_hi, full, kb * BK, m0);
            tma_load_2d(base + BM * kBlockBytes, a_lo, full, kb * BK, m0);
            tma_load_2d(base + 2 * BM * kBlockBytes, b_hi, full, kb * BK, b_row0);
            tma_load_2d(base + 3 * BM * kBlockBytes, b_lo, full, kb * BK, b_row0);
        }
    };
    auto issue = [&](int nkb) {
        constexpr uint32_t idesc = umma_idesc_mn(KIND_F16, BM, 128);
        for (int kb = 0; kb < nkb; kb++, it++, ch++) {
            const int buf = ch % NB, s = it % kMidStages;
            mbar_wait(smem_u32(&bar_acc_empty[buf]), (((uint32_t) (ch / NB)) & 1u) ^ 1u);
            mbar_wait(smem_u32(&bar_full[s]), ((uint32_t) (it / kMidStages)) & 1u);
            tcgen05_fence_after();
            const uint32_t tmem_acc = tmem_base + (uint32_t) (buf * 128);
            const uint32_t base = smem_u32(data + (size_t) s * kMidStageBytes);
            const uint32_t a_hi = base, a_lo = base + BM * kBlockBytes, b_hi = base + 2 * BM * kBlockBytes, b_lo = base + 3 * BM * kBlockBytes;
#pragma unroll
            for (int kk = 0; kk < kStepsPerBlock; kk++) {
                const uint32_t off = kk * kStepBytes;
                const uint64_t dah = umma_smem_desc(a_hi + off), dal = umma_smem_desc(a_lo + off);
                const uint64_t dbh = umma_smem_desc(b_hi + off), dbl = umma_smem_desc(b_lo + off);
                umma_mma<KIND_F16>(tmem_acc, dal, dbh, idesc, kk == 0 ? 0u : 1u);     // small terms first
                umma_mma<KIND_F16>(tmem_acc, dah, dbl, idesc, 1u);
                umma_mma<KIND_F16>(tmem_acc, dah, dbh, idesc, 1u);
            }
            umma_commit(smem_u32(&bar_empty[s]));
            umma_commit(smem_u32(&bar_acc_full[buf]));
        }
    };
    // epilogue threads: e = 0..255; TMEM lane quarter q, column half h of the CTA's 128 columns
    const int e = (int) threadIdx.x, q = warp & 3, h = warp >> 2, row = q * 32 + lane;
    float acc[64];
    auto drain = [&](int nkb) {
#pragma unroll
        for (int j = 0; j < 64; j++) acc[j] = 0.f;
        for (int kb = 0; kb < nkb; kb++, ch++) {
            const int buf = ch % NB;
            mbar_wait(smem_u32(&bar_acc_full[buf]), ((uint32_t) (ch / NB)) & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (buf * 128 + h * 64 + cc * 32), r);
#pragma unroll
                for (int j = 0; j < 32; j++) acc[cc * 32 + j] += __uint_as_float(r[j]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_acc_empty[buf])) : "memory");
        }
        // park this thread's 64 columns of its row (the stage buffers are free: every MMA of the tile has completed)
        float* dst = reinterpret_cast<float*>(data) + (size_t) row * kMidStagLd + h * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    };

    float* stag = reinterpret_cast<float*>(data);
    double* zpart = reinterpret_cast<double*>(data + kMidOffZpart);
    double* zh = reinterpret_cast<double*>(data + kMidOffZh);
    double* zfull = reinterpret_cast<double*>(data + kMidOffZfull);
    double* cen = reinterpret_cast<double*>(data + kMidOffCen);
    double* dzs = reinterpret_cast<double*>(data + kMidOffDz);
    double* epart = reinterpret_cast<double*>(data + kMidOffEpart);
    float* dzf = reinterpret_cast<float*>(data + kMidOffDzf);

    // =========================== phase 1: A2 slice, partial output-layer sums ===========================
    if (warp == kMidEpiWarps) {
        if (lane == 0) produce(&tm_x1_hi, &tm_x1_lo, &tm_w_hi, &tm_w_lo, nkb1, n0);
        __syncwarp();
    } else if (warp == kMidEpiWarps + 1) {
        if (lane == 0) issue(nkb1);
        __syncwarp();
    } else if (warp < kMidEpiWarps) {
        drain(nkb1);
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 3);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * (kMidEpiWarps + 1)) : "memory");   // staged vectors are in place
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 10);
        const int ea = s_aux_exp[row];
        float* mine = stag + (size_t) row * kMidStagLd + h * 64;
        double zacc[NLP];
#pragma unroll
        for (int j = 0; j < NLP; j++) zacc[j] = 0.0;
        // rolled on purpose (instruction-cache footprint): 16 groups of 4 columns of this thread's row half
#pragma unroll 1
        for (int i = 0; i < 16; i++) {
            const int lc = h * 64 + 4 * i, n = n0 + lc;
            float4 v = *reinterpret_cast<const float4*>(mine + 4 * i);
            const int4 eb = *reinterpret_cast<const int4*>(s_b_exp + lc);
            const float4 bs = *reinterpret_cast<const float4*>(s_bias + lc);
            v.x = v.x * exp2i(-(ea + eb.x)) + bs.x; v.y = v.y * exp2i(-(ea + eb.y)) + bs.y;
            v.z = v.z * exp2i(-(ea + eb.z)) + bs.z; v.w = v.w * exp2i(-(ea + eb.w)) + bs.w;
            if (p.hidden_type == ANNF_TANH) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
            *reinterpret_cast<float4*>(mine + 4 * i) = v;                          // A2 stays here for act'() below
            if (n < p.n2) {                                                        // n2 is a multiple of 4: a group is all in or all out
#pragma unroll
                for (int j = 0; j < NLP; j++) {
                    if (j < p.nL) {
                        const double2 w0 = __ldg(reinterpret_cast<const double2*>(p.WL + (size_t) j * p.n2 + n));
                        const double2 w1 = __ldg(reinterpret_cast<const double2*>(p.WL + (size_t) j * p.n2 + n + 2));
                        zacc[j] = fma(w0.x, (double) v.x, zacc[j]); zacc[j] = fma(w0.y, (double) v.y, zacc[j]);
                        zacc[j] = fma(w1.x, (double) v.z, zacc[j]); zacc[j] = fma(w1.y, (double) v.w, zacc[j]);
                    }
                }
            }
        }
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 11);
        // the two column halves of a row, then push this CTA's partial sums to every CTA of the cluster
        if (h == 1) {
#pragma unroll
            for (int j = 0; j < NLP; j++) if (j < p.nL) zh[j * BM + row] = zacc[j];
        }
        epi_barrier();
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 12);
        if (h == 0) {
#pragma unroll
            for (int j = 0; j < NLP; j++) {
                if (j < p.nL) {
                    const double z = zacc[j] + zh[j * BM + row];
                    const uint32_t local = smem_u32(zpart + ((size_t) rank * kMidMaxOut + j) * BM + row);
                    for (int c = 0; c < CL; c++) st_cluster_f64(map_to_cta(local, (uint32_t) c), z);
                }
            }
        }
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 4);
    }
    cluster_barrier();                                      // A: every CTA holds every partial sum
    if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 13);
    // =========================== head: energy, dE/dz, D2 slice ===========================
    if (warp < kMidEpiWarps) {
        for (int v = e; v < p.nL * BM; v += 32 * kMidEpiWarps) {
            const int j = v / BM, r = v % BM;
            double z = p.bL[j];
            for (int c = 0; c < CL; c++) z += zpart[((size_t) c * kMidMaxOut + j) * BM + r];   // fixed order: deterministic
            zfull[j * BM + r] = z;
        }
        for (int v = e; v < p.num_center * BM; v += 32 * kMidEpiWarps) {
            const int c = v / BM, r = v % BM, m = m0 + r;
            cen[r * kMidCenLd + c] = (p.centers != nullptr && m < p.M) ? (double) p.centers[(size_t) m * p.num_center + c] : p.center_dev[c];
        }
        epi_barrier();
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 14);
        const double kforce = *p.k_dev;
        for (int v = e; v < kMidMaxOut * BM; v += 32 * kMidEpiWarps) {            // item (t, r): output t (or Circular pair t) of row r
            const int t = v / BM, r = v % BM;
            if (t < p.nL) epart[t * BM + r] = output_head_parallel(p.out_type, p.nL, zfull + r, BM, cen + r * kMidCenLd, kforce, dzs + r, BM, t);
        }
        epi_barrier();
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 15);
        if (e < BM) {                                       // one thread per row: energy, scale exponents of the gradient rows
            const int r = e, m = m0 + r;
            double en = 0.0;
            float mx = 0.f;
            for (int j = 0; j < p.nL; j++) {
                en += epart[j * BM + r];
                const float d = (float) dzs[j * BM + r];
                dzf[j * BM + r] = d;
                mx = fmaxf(mx, fabsf(d));
            }
            float bound = mx * chain.bwd_gain[p.L - 2];     // |D2[r][.]| <= max_j |dz_j| * max column L1 norm of W_L
            const int e2 = exp_for_bound(bound);
            bound *= chain.bwd_gain[p.L - 3];               // |D1[r][.]| <= the same through W_{L-3}, |act'| <= 1
            const int e1 = exp_for_bound(bound);
            s_a_exp[r] = e2;                                // A rows of phase 2
            s_out_exp[r] = e1;                              // rows coming out of phase 2
            if (rank == 0 && m < p.M) {
                p.energies[m] = en;
                chain.d_exp[p.L - 2][m] = e2;
                chain.d_exp[p.L - 3][m] = e1;
                for (int l = p.L - 4; l >= 1; l--) {        // deeper networks: the rest of the chain, like head_kernel
                    bound *= chain.bwd_gain[l];
                    chain.d_exp[l][m] = exp_for_bound(bound);
                }
            }
        }
        epi_barrier();
        // D2 slice: lanes along a row (coalesced stores), 32 four-column groups per row
#pragma unroll 1
        for (int v = e; v < BM * 32; v += 32 * kMidEpiWarps) {
            const int r = v >> 5, lc = (v & 31) * 4, n = n0 + lc, m = m0 + r;
            if (m >= p.M || n >= p.n2) continue;
            const float4 a = *reinterpret_cast<const float4*>(stag + (size_t) r * kMidStagLd + lc);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < p.nL; j++) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(p.WL32 + (size_t) j * p.n2 + n));
                const float d = dzf[j * BM + r];
                g.x = fmaf(w.x, d, g.x); g.y = fmaf(w.y, d, g.y); g.z = fmaf(w.z, d, g.z); g.w = fmaf(w.w, d, g.w);
            }
            if (p.hidden_type == ANNF_TANH) { g.x *= (1.f - a.x * a.x); g.y *= (1.f - a.y * a.y); g.z *= (1.f - a.z * a.z); g.w *= (1.f - a.w * a.w); }
            const float sc = exp2i(s_a_exp[r]);
            uint2 hi, lo;
            split_f16x4(make_float4(g.x * sc, g.y * sc, g.z * sc, g.w * sc), hi, lo);
            *reinterpret_cast<uint2*>(p.d2_hi + (size_t) m * p.ld2 + n) = hi;
            *reinterpret_cast<uint2*>(p.d2_lo + (size_t) m * p.ld2 + n) = lo;
        }
        asm volatile("fence.proxy.async;" ::: "memory");    // generic-proxy global writes -> visible to the TMA loads of phase 2
        if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 5);
    }
    cluster_barrier();                                      // B: D2 is complete for this tile of replicas
    if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 6);
    // =========================== phase 2: D1 tiles ===========================
    if (warp == kMidEpiWarps) {
        if (lane == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            int done = 0;
            for (int t = rank; t < tiles1; t += CL, done++) {
                if (done > 0) mbar_wait(smem_u32(bar_epi_free), ((uint32_t) (done - 1)) & 1u);   // the epilogue's staging overlaps the stages
                produce(&tm_d2_hi, &tm_d2_lo, &tm_wt_hi, &tm_wt_lo, nkb2, t * 128);
            }
        }
    } else if (warp == kMidEpiWarps + 1) {
        if (lane == 0)
            for (int t = rank; t < tiles1; t += CL) issue(nkb2);
    } else if (warp < kMidEpiWarps) {
        GemmParams gp;
        gp.M = p.M; gp.N = p.n1; gp.K = p.n2;
        gp.epi = p.prev_type == ANNF_TANH ? EPI_TANH_GRAD : EPI_LINEAR_GRAD;
        gp.bias = nullptr;
        gp.aux_hi = p.x1_hi; gp.aux_lo = p.x1_lo; gp.ld_aux = p.ld1;
        gp.out_hi = p.d1_hi; gp.out_lo = p.d1_lo; gp.ld_out = p.ld1;
        gp.a_exp = nullptr; gp.b_exp = nullptr; gp.out_exp = nullptr; gp.aux_exp = nullptr;   // staged in `vec`
        gp.sel_atoms = nullptr; gp.direct_store = 0; gp.dbg = -1; gp.chunk = 1; gp.dry = 0;
        gp.no_b_prefetch = 0;
        gp.gsplit = 0; gp.gs_scratch = nullptr; gp.gs_arrive = nullptr; gp.gs_done = nullptr;
        for (int t = rank; t < tiles1; t += CL) {
            drain(nkb2);
            if (e < 128) s_b_exp[e] = (t * 128 + e) < p.n1 ? __ldg(p.wt_exp + t * 128 + e) : 0;
            epi_barrier();                                  // tile parked, column exponents staged
            if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 7);
            cooperative_epilogue<KIND_F16, 128, 1, 1, 32 * kMidEpiWarps, 1>(gp, vec, data, 0, 0, m0, t * 128, e);
            if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 8);
            epi_barrier();                                  // staging and s_b_exp may be overwritten by the next tile
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar_epi_free)) : "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) ANNF_TMARK(p.dbg, 9);
    if (warp == kMidEpiWarps + 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) (NB * 128)) : "memory");
    }
    ANNF_TGRID_END(p.dbg);
}

// ------------------------------------------------------------------------------------------------
// Host side: plan
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &res) == cudaSuccess &&
            res == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// [rows][cols] fp32 or fp16, row stride ld elements; box = one 128-byte K-block x box_rows rows, 128B swizzle, zero OOB fill
static bool make_map(CUtensorMap* map, int kind, const void* ptr, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const int esz = kind == KIND_F16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t) cols, (cuuint64_t) rows};
    cuuint64_t strides[1] = {(cuuint64_t) ld * esz};
    cuuint32_t box[2] = {(cuuint32_t) (kBlockBytes / esz), (cuuint32_t) box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, kind == KIND_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims,
              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }


struct GemmOp {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    CUtensorMap b_hi64, b_lo64;          // the same B with 64-row boxes (256 x 128 CTA-pair tiles)
    CUtensorMap a_hi64, a_lo64;          // the same A with 64-row boxes (operand multicast between two N tiles)
    CUtensorMap out;                     // EPI_FORCE with direct_store: the caller's force array, 32-column x 128-row boxes (per call)
    GemmParams p;
    char name[48];
};

constexpr int kGsMaxTiles = 64;                    // output tiles (x CTA of a pair) a globally split GEMM may have
constexpr size_t kGsScratchBytes = (size_t) 160 * BM * 256 * sizeof(float);   // up to 160 partial tiles of 128 x 256 fp32 (21 MB)
constexpr int kPlanSlots = 2;         // independent activation-buffer sets: two chunks of a call may be in flight at once

struct PlanSlot {
    int cap = 0;                                    // replica capacity of the activation buffers
    // activations X_l (l = 0..L-2) and pre-activation gradients D_l (l = 1..L-2) as operand pairs (fp32 words or halves)
    void *x_hi[kMaxLayers] = {}, *x_lo[kMaxLayers] = {}, *d_hi[kMaxLayers] = {}, *d_lo[kMaxLayers] = {};
    int *x_exp[kMaxLayers] = {}, *d_exp[kMaxLayers] = {};   // KIND_F16: per-row scale exponents of the same tensors
    float* gs_scratch = nullptr;                    // split-K through global memory: partial tiles (kGsScratchBytes)
    int* gs_counters = nullptr;                     // 2 x kGsMaxTiles ints: arrive[], done[]
    std::vector<GemmOp> fwd, bwd;
};

struct TensorPlan {
    int L = 0;
    int kind = KIND_TF32;                           // operand container of every contraction of this plan
    int nodes[kMaxLayers];
    int types[kMaxLayers];
    int ld[kMaxLayers];
    int sm_count = 148;
    // weights (operand pairs), forward [n_{l+1}][ld(n_l)] and transposed [n_l][ld(n_{l+1})], l = 0..L-3
    void *w_hi[kMaxLayers] = {}, *w_lo[kMaxLayers] = {}, *wt_hi[kMaxLayers] = {}, *wt_lo[kMaxLayers] = {};
    int *w_exp[kMaxLayers] = {}, *wt_exp[kMaxLayers] = {};   // KIND_F16: per-row scale exponents of the same matrices
    ExpChain chain = {};                            // KIND_F16: norms that propagate magnitude bounds (pointers filled per slot)
    PlanSlot slots[kPlanSlots];
    bool sel_identity = false;                      // selection is atoms 0..A-1 in order
    int force_bn = 0;                               // ANNF_GEMM_BN=128|256 (1-CTA tiles) or 512 (CTA pairs) overrides the tile heuristic
    int force_split = 0;                            // ANNF_GEMM_SPLIT_SHORTK=1|2|4|8 overrides split-K for GEMMs with K <= 1024
    int dry = 0;                                    // ANNF_TENSOR_DRY=1: every kernel returns at once (launch-floor measurement)
    int dbg_next = 0;                               // diagnostic builds: g_tensor_clock slot of the next launch of the step
    int chunk = kChunkDefault;                      // ANNF_TENSOR_CHUNK: K-blocks per TMEM accumulation chunk (accuracy experiments)
    bool no_cluster16 = false;                      // a 16-CTA cluster launch failed on this device
    int prep_mode = 0;                              // ANNF_PREP_MODE: 0 = bulk-copy kernel when the selection is contiguous (default); 1 = prep_kernel, coalesced walk; 2 = prep_kernel, 4-atom groups
    int prep_warps = 8;                             // ANNF_PREP_WARPS: warps per replica of the bulk-copy kernel (1, 2, 4, 8)
    bool prep_pipelined = false;                    // ANNF_PREP_PIPELINED=0: the one-CTA-per-replica prep kernel (experiments)
    int prep_threads = 256;                         // CTA size of prep_kernel (ANNF_PREP_THREADS=128|256)
    int gsplit_mode = 0;                            // ANNF_GEMM_GSPLIT: 0 off (default), -1 auto, n > 1 forces n K ranges (split-K through global memory)
    bool use_mid = false;                           // fused forward / head / backward of the last hidden connection (mid_kernel)
    bool no_b_prefetch = false;                     // ANNF_TENSOR_B_PREFETCH=0
    bool no_direct_store = false;                   // ANNF_TENSOR_DIRECT_STORE=0: force tiles leave through the thread epilogue (experiments)
};

static void free_activations(PlanSlot* t) {
    for (int l = 0; l < kMaxLayers; l++) {
        cudaFree(t->x_hi[l]); cudaFree(t->x_lo[l]); cudaFree(t->d_hi[l]); cudaFree(t->d_lo[l]);
        cudaFree(t->x_exp[l]); cudaFree(t->d_exp[l]);
        t->x_hi[l] = t->x_lo[l] = t->d_hi[l] = t->d_lo[l] = nullptr;
        t->x_exp[l] = t->d_exp[l] = nullptr;
    }
    cudaFree(t->gs_scratch); cudaFree(t->gs_counters);
    t->gs_scratch = nullptr; t->gs_counters = nullptr;
    t->cap = 0;
}

void tensor_plan_destroy(TensorPlan* t) {
    if (!t) return;
    for (PlanSlot& slot : t->slots) free_activations(&slot);
    for (int l = 0; l < kMaxLayers; l++) {
        cudaFree(t->w_hi[l]); cudaFree(t->w_lo[l]); cudaFree(t->wt_hi[l]); cudaFree(t->wt_lo[l]);
        cudaFree(t->w_exp[l]); cudaFree(t->wt_exp[l]);
    }
    delete t;
}

static float host_tf32(float v) {          // round-to-nearest, ties away from zero on the 13 dropped bits (cvt.rna)
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return v;
    u += 0x1000u;
    u &= 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

// [rows][ld] fp32 matrix -> operand pair on the device.  KIND_F16: every row is first multiplied by 2^e, e chosen so that the
// row's largest magnitude lies in [2^13, 2^14); the exponents are uploaded next to the pair.
static cudaError_t upload_pair(int kind, const std::vector<float>& src, int rows, int ld, void** hi, void** lo, int** exps) {
    cudaError_t err;
    if (kind == KIND_TF32) {
        std::vector<float> h(src.size()), l(src.size());
        for (size_t i = 0; i < src.size(); i++) {
            h[i] = host_tf32(src[i]);
            l[i] = host_tf32(src[i] - h[i]);
        }
        if ((err = cudaMalloc(hi, sizeof(float) * src.size())) != cudaSuccess) return err;
        if ((err = cudaMalloc(lo, sizeof(float) * src.size())) != cudaSuccess) return err;
        if ((err = cudaMemcpy(*hi, h.data(), sizeof(float) * src.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return err;
        return cudaMemcpy(*lo, l.data(), sizeof(float) * src.size(), cudaMemcpyHostToDevice);
    }
    std::vector<__half> h(src.size()), l(src.size());
    std::vector<int> e(rows, 0);
    for (int r = 0; r < rows; r++) {
        float mx = 0.f;
        for (int c = 0; c < ld; c++) mx = std::max(mx, std::fabs(src[(size_t) r * ld + c]));
        if (mx > 0.f && std::isfinite(mx)) {
            int q;
            std::frexp(mx, &q);
            e[r] = std::max(-kExpClamp, std::min(kExpClamp, kExpTarget - q));
        }
        const float sc = std::ldexp(1.f, e[r]);
        for (int c = 0; c < ld; c++) {
            const float v = src[(size_t) r * ld + c] * sc;
            const __half vh = __float2half_rn(v);
            h[(size_t) r * ld + c] = vh;
            l[(size_t) r * ld + c] = __float2half_rn(v - __half2float(vh));
        }
    }
    if ((err = cudaMalloc(hi, sizeof(__half) * src.size())) != cudaSuccess) return err;
    if ((err = cudaMalloc(lo, sizeof(__half) * src.size())) != cudaSuccess) return err;
    if ((err = cudaMalloc((void**) exps, sizeof(int) * round_up(rows, 4))) != cudaSuccess) return err;
    if ((err = cudaMemset(*exps, 0, sizeof(int) * round_up(rows, 4))) != cudaSuccess) return err;
    if ((err = cudaMemcpy(*hi, h.data(), sizeof(__half) * src.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return err;
    if ((err = cudaMemcpy(*lo, l.data(), sizeof(__half) * src.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return err;
    return cudaMemcpy(*exps, e.data(), sizeof(int) * rows, cudaMemcpyHostToDevice);
}

TensorPlan* tensor_plan_create(const NetView& net, const std::vector<std::vector<double>>& coeff, const std::vector<double>& bias64,
                               int kind, std::string* why_not) {
    auto no = [&](const char* msg) { if (why_not) *why_not = msg; return (TensorPlan*) nullptr; };
    const int L = net.L;
    if (net.input_type != ANNF_INPUT_CARTESIAN) return no("only Cartesian input is on the tensor-core path");
    if (L < 3) return no("no hidden layer");
    if (net.nodes[L - 1] > 64 || net.num_center > 64) return no("output layer wider than 64 nodes");
    for (int l = 0; l + 2 < L; l++)
        if (net.types[l] != ANNF_TANH && net.types[l] != ANNF_LINEAR) return no("hidden layers must be Tanh or Linear");
    for (int l = 0; l + 1 < L; l++)
        if (net.nodes[l] % 4 != 0) return no("layer widths (except the output layer) must be multiples of 4");
    for (int l = 1; l + 1 < L; l++)
        if (net.nodes[l] < 64) return no("hidden layers narrower than 64 nodes are not a dense contraction worth the tensor pipe");
    if (!encode_fn()) return no("cuTensorMapEncodeTiled is not available from this driver");
    if ((size_t) net.nodes[0] * sizeof(float) > 200 * 1024) return no("input layer too wide for the prep kernel's shared memory");
    if (cudaFuncSetAttribute(prep_kernel<KIND_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(prep_kernel<KIND_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(prep_pipelined_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        !prep_bulk_attributes<KIND_TF32>() || !prep_bulk_attributes<KIND_F16>())
        return no("cannot configure the prep kernel");
    TensorPlan* t = new TensorPlan();
    t->L = L;
    t->kind = kind;
    if (const char* env = getenv("ANNF_GEMM_BN")) t->force_bn = atoi(env);
    if (const char* env = getenv("ANNF_GEMM_SPLIT_SHORTK")) t->force_split = atoi(env);
    if (const char* env = getenv("ANNF_TENSOR_DRY")) t->dry = atoi(env);
    if (const char* env = getenv("ANNF_TENSOR_B_PREFETCH")) t->no_b_prefetch = atoi(env) == 0;
    if (const char* env = getenv("ANNF_TENSOR_DIRECT_STORE")) t->no_direct_store = atoi(env) == 0;
    // fp16 pairs: one K-block (64 values, 12 chained MMAs) per TMEM chunk; TF32 pairs: two (64 values, 24 chained MMAs).
    // Measured at BASELINE C5 (64 replicas against the reference, profiles/r02_chunk_sweep.md): the energy error grows with
    // the chunk length — the tensor core's accumulator truncates — 6.4e-7 / 1.0e-6 / 1.9e-6 / 3.0e-6 worst case at 1 / 2 / 4 / 8.
    t->chunk = kind == KIND_F16 ? 1 : kChunkDefault;
    if (const char* env = getenv("ANNF_TENSOR_CHUNK")) t->chunk = std::max(1, std::min(64, atoi(env)));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev);
    const int ld_unit = kind == KIND_F16 ? 8 : 4;               // rows are 16-byte multiples (TMA global stride)
    for (int l = 0; l < L; l++) { t->nodes[l] = net.nodes[l]; t->ld[l] = round_up(net.nodes[l], ld_unit); }
    for (int l = 0; l + 1 < L; l++) t->types[l] = net.types[l];
    t->chain.inv_scale = 1.0 / net.scale;
    t->chain.inv_hi = (float) t->chain.inv_scale;
    t->chain.inv_lo = (float) (t->chain.inv_scale - (double) t->chain.inv_hi);
    // norms for the magnitude bounds of KIND_F16 (ExpChain): forward max row L1 norm and max |bias|, backward max column L1 norm
    for (int l = 0; l + 1 < L; l++) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        std::vector<double> col(n_in, 0.0);
        double row_max = 0.0, b_max = 0.0;
        for (int j = 0; j < n_out; j++) {
            double row = 0.0;
            for (int i = 0; i < n_in; i++) {
                const double a = std::fabs(coeff[l][(size_t) j * n_in + i]);
                row += a;
                col[i] += a;
            }
            row_max = std::max(row_max, row);
            b_max = std::max(b_max, std::fabs(bias64[net.b_off[l] + j]));
        }
        // 1 % head-room covers the fp32 arithmetic of the bound itself
        t->chain.fwd_gain[l] = (float) (1.01 * row_max);
        t->chain.fwd_bias[l] = (float) (1.01 * b_max);
        t->chain.bwd_gain[l] = (float) (1.01 * *std::max_element(col.begin(), col.end()));
    }
    for (int l = 0; l + 2 < L; l++) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        std::vector<float> w((size_t) n_out * t->ld[l], 0.f), wt((size_t) n_in * t->ld[l + 1], 0.f);
        for (int j = 0; j < n_out; j++) {
            // l == 0: the backward copy carries the centring / scaling Jacobian (see EPI_FORCE), folded in fp64
            double wsum[3] = {0.0, 0.0, 0.0};
            if (l == 0)
                for (int i = 0; i < n_in; i++) wsum[i % 3] += coeff[0][(size_t) j * n_in + i];
            for (int i = 0; i < n_in; i++) {
                const double v = coeff[l][(size_t) j * n_in + i];
                w[(size_t) j * t->ld[l] + i] = (float) v;
                wt[(size_t) i * t->ld[l + 1] + j] = l == 0 ? (float) (-(v - wsum[i % 3] / net.n_sel) / net.scale) : (float) v;
            }
        }
        if (upload_pair(kind, w, n_out, t->ld[l], &t->w_hi[l], &t->w_lo[l], &t->w_exp[l]) != cudaSuccess ||
            upload_pair(kind, wt, n_in, t->ld[l + 1], &t->wt_hi[l], &t->wt_lo[l], &t->wt_exp[l]) != cudaSuccess) {
            tensor_plan_destroy(t);
            return no("out of device memory for the tensor-core weight copies");
        }
    }
    // the fused middle kernel: fp16 pairs, two or more hidden layers, last hidden layer within one 4-CTA cluster, few outputs
    // Opt-in (ANNF_TENSOR_MID=1), not the default: with one cluster of 4 CTAs per 128 replicas only 32 SMs work at BASELINE C5
    // and the element-wise parts (tanh, fp64 output layer, act') are instruction-issue bound there — 60 us against 52 us for
    // GEMM + head + GEMM spread over 64 - 1024 CTAs (profiles/r02_mid_kernel.md).
    const bool mid_ok = kind == KIND_F16 && L >= 4 && net.nodes[L - 2] <= 128 * kMidCluster && net.nodes[L - 1] <= kMidMaxOut;
    t->use_mid = false;
    if (const char* env = getenv("ANNF_GEMM_GSPLIT")) t->gsplit_mode = atoi(env);
    if (const char* env = getenv("ANNF_PREP_MODE")) t->prep_mode = atoi(env);
    if (const char* env = getenv("ANNF_PREP_WARPS")) {
        const int w = atoi(env);
        if (w == 1 || w == 2 || w == 4 || w == 8) t->prep_warps = w;
    }
    if (const char* env = getenv("ANNF_PREP_PIPELINED")) t->prep_pipelined = atoi(env) != 0;
    if (const char* env = getenv("ANNF_PREP_THREADS")) t->prep_threads = atoi(env) == 128 ? 128 : 256;
    if (const char* env = getenv("ANNF_TENSOR_MID")) t->use_mid = mid_ok && atoi(env) != 0;
    // fp32 biases live in NetView::b32 already
    {
        std::vector<int> sel(net.n_sel);
        cudaMemcpy(sel.data(), net.sel_atoms, sizeof(int) * net.n_sel, cudaMemcpyDeviceToHost);
        t->sel_identity = true;
        for (int i = 0; i < net.n_sel; i++) t->sel_identity = t->sel_identity && sel[i] == i;
    }
    return t;
}

int tensor_plan_kind(const TensorPlan* t) { return t->kind; }

// Sizes one activation-buffer set for R replicas (rows padded to 128).  Called from tensor_plan_reserve (annf_create /
// annf_reserve) and, when a batch is larger than anything reserved, from tensor_plan_run: the buffers are zeroed ON `stream`,
// so the kernels enqueued right after on the same stream are ordered behind the zeroing.
static cudaError_t ensure_capacity(TensorPlan* plan, PlanSlot* t, const NetView& net, int R, cudaStream_t stream) {
    if (R <= t->cap) return cudaSuccess;
    // a growing batch: make sure nothing still reads the old buffers (a rare event, never on the per-step path once warm)
    cudaError_t err = t->cap > 0 ? cudaDeviceSynchronize() : cudaSuccess;
    if (err != cudaSuccess) return err;
    free_activations(t);
    const int cap = round_up(R, BM);
    const int L = plan->L, kind = plan->kind;
    const size_t esz = kind == KIND_F16 ? 2 : 4;
    auto alloc = [&](void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, stream);
        return e;
    };
    for (int l = 0; l + 1 < L && err == cudaSuccess; l++) {
        const size_t bytes = esz * (size_t) cap * plan->ld[l];
        err = alloc(&t->x_hi[l], bytes);
        if (err == cudaSuccess) err = alloc(&t->x_lo[l], bytes);
        if (err == cudaSuccess) err = alloc((void**) &t->x_exp[l], sizeof(int) * cap);
        if (l >= 1 && err == cudaSuccess) err = alloc(&t->d_hi[l], bytes);
        if (l >= 1 && err == cudaSuccess) err = alloc(&t->d_lo[l], bytes);
        if (l >= 1 && err == cudaSuccess) err = alloc((void**) &t->d_exp[l], sizeof(int) * cap);
    }
    if (err == cudaSuccess) err = alloc((void**) &t->gs_scratch, kGsScratchBytes);
    if (err == cudaSuccess) err = alloc((void**) &t->gs_counters, sizeof(int) * 2 * kGsMaxTiles);
    if (err != cudaSuccess) return err;
    t->cap = cap;
    t->fwd.clear();
    t->bwd.clear();
    for (int l = 0; l + 2 < L; l++) {                 // forward: X_{l+1} = act(X_l W_l^T + b)
        GemmOp op;
        memset(&op, 0, sizeof(op));
        const int K = plan->nodes[l], N = plan->nodes[l + 1];
        bool ok = make_map(&op.a_hi, kind, t->x_hi[l], cap, K, plan->ld[l], BM) && make_map(&op.a_lo, kind, t->x_lo[l], cap, K, plan->ld[l], BM) &&
                  make_map(&op.a_hi64, kind, t->x_hi[l], cap, K, plan->ld[l], BM / 2) && make_map(&op.a_lo64, kind, t->x_lo[l], cap, K, plan->ld[l], BM / 2) &&
                  make_map(&op.b_hi, kind, plan->w_hi[l], N, K, plan->ld[l], 128) && make_map(&op.b_lo, kind, plan->w_lo[l], N, K, plan->ld[l], 128) &&
                  make_map(&op.b_hi64, kind, plan->w_hi[l], N, K, plan->ld[l], 64) && make_map(&op.b_lo64, kind, plan->w_lo[l], N, K, plan->ld[l], 64);
        if (!ok) return cudaErrorInvalidValue;
        op.p.N = N; op.p.K = K;
        op.p.epi = plan->types[l] == ANNF_TANH ? EPI_BIAS_TANH : EPI_BIAS_LINEAR;
        op.p.bias = net.b32 + net.b_off[l];
        op.p.out_hi = t->x_hi[l + 1]; op.p.out_lo = t->x_lo[l + 1]; op.p.ld_out = plan->ld[l + 1];
        op.p.a_exp = t->x_exp[l]; op.p.b_exp = plan->w_exp[l]; op.p.out_exp = t->x_exp[l + 1];
        snprintf(op.name, sizeof(op.name), "gemm[Rx%dx%d] fwd%d %s", N, K, l, kind == KIND_F16 ? "f16x3" : "tf32x3");
        t->fwd.push_back(op);
    }
    for (int l = L - 3; l >= 0; l--) {                // backward: D_l = (D_{l+1} W_l) (.) act'(X_l);  l = 0 -> G0
        GemmOp op;
        memset(&op, 0, sizeof(op));
        const int K = plan->nodes[l + 1], N = plan->nodes[l];
        bool ok = make_map(&op.a_hi, kind, t->d_hi[l + 1], cap, K, plan->ld[l + 1], BM) && make_map(&op.a_lo, kind, t->d_lo[l + 1], cap, K, plan->ld[l + 1], BM) &&
                  make_map(&op.a_hi64, kind, t->d_hi[l + 1], cap, K, plan->ld[l + 1], BM / 2) && make_map(&op.a_lo64, kind, t->d_lo[l + 1], cap, K, plan->ld[l + 1], BM / 2) &&
                  make_map(&op.b_hi, kind, plan->wt_hi[l], N, K, plan->ld[l + 1], 128) && make_map(&op.b_lo, kind, plan->wt_lo[l], N, K, plan->ld[l + 1], 128) &&
                  make_map(&op.b_hi64, kind, plan->wt_hi[l], N, K, plan->ld[l + 1], 64) && make_map(&op.b_lo64, kind, plan->wt_lo[l], N, K, plan->ld[l + 1], 64);
        if (!ok) return cudaErrorInvalidValue;
        op.p.N = N; op.p.K = K;
        op.p.a_exp = t->d_exp[l + 1]; op.p.b_exp = plan->wt_exp[l];
        if (l == 0) {
            op.p.epi = EPI_FORCE;                     // out_hi / ld_out / sel_atoms are per call (launch_gemm)
        } else {
            op.p.epi = plan->types[l - 1] == ANNF_TANH ? EPI_TANH_GRAD : EPI_LINEAR_GRAD;
            op.p.aux_hi = t->x_hi[l]; op.p.aux_lo = t->x_lo[l]; op.p.ld_aux = plan->ld[l]; op.p.aux_exp = t->x_exp[l];
            op.p.out_hi = t->d_hi[l]; op.p.out_lo = t->d_lo[l]; op.p.ld_out = plan->ld[l]; op.p.out_exp = t->d_exp[l];
        }
        snprintf(op.name, sizeof(op.name), "gemm[Rx%dx%d] bwd%d %s", N, K, l, kind == KIND_F16 ? "f16x3" : "tf32x3");
        t->bwd.push_back(op);
    }
    return cudaSuccess;
}

// Pre-sizes both activation-buffer sets for batches of up to R replicas, so that no later call allocates.
cudaError_t tensor_plan_reserve(TensorPlan* plan, const NetView& net, int R, cudaStream_t stream) {
    for (PlanSlot& slot : plan->slots) {
        const cudaError_t err = ensure_capacity(plan, &slot, net, R, stream);
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

// Launch with programmatic dependent launch allowed (and, for the GEMM, its split-K cluster shape).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              int cluster_z, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
    if (cluster_z > 0) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = cluster_z;
        n++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    if (cluster_z > 0 && getenv("ANNF_CLUSTER_DEBUG")) {     // diagnostics: how many clusters of this shape the device co-schedules
        int max_clusters = -1;
        cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg);
        fprintf(stderr, "[annf] grid (%u,%u,%u) cluster (%d,1,%d) smem %zu: %d clusters of %u needed, %d co-resident\n", grid.x, grid.y, grid.z,
                cluster_x, cluster_z, smem, (int) (grid.x * grid.y * grid.z) / (cluster_x * cluster_z), (unsigned) (cluster_x * cluster_z), max_clusters);
    }
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <int KIND, int BN, int S, int CG, int NX = 1>
static cudaError_t launch_gemm_t(const GemmOp& op, const GemmParams& p_in, cudaStream_t stream) {
    GemmParams p = p_in;
    if (S != 1) p.direct_store = 0;
    using Cfg = GemmCfg<BN, CG>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(gemm3x_kernel<KIND, BN, S, CG, NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (err != cudaSuccess) return err;
        if (CG * NX * S > 8) {
            err = cudaFuncSetAttribute(gemm3x_kernel<KIND, BN, S, CG, NX>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (err != cudaSuccess) return err;
        }
        configured[dev] = true;
    }
    const int n_tiles = (p.N + BN - 1) / BN, m_tiles = (p.M + BM * CG - 1) / (BM * CG);
    constexpr bool kBox64 = Cfg::kBRows < 128;
    if (BN >= 256) p.chunk = std::max(p.chunk, 2);          // two TMEM buffers only: one-block chunks put the drain on the MMA issuer's critical path
    const int grid_z = S > 1 ? S : std::max(1, p.gsplit);   // S = 1: gridDim.z = K ranges of a split through global memory (no cluster along z)
    return launch_pdl(gemm3x_kernel<KIND, BN, S, CG, NX>, dim3(CG * n_tiles, m_tiles, grid_z), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, CG * NX, S,
                      NX == 2 ? op.a_hi64 : op.a_hi, NX == 2 ? op.a_lo64 : op.a_lo, kBox64 ? op.b_hi64 : op.b_hi, kBox64 ? op.b_lo64 : op.b_lo,
                      p.direct_store ? op.out : op.a_hi, p);
}

static int pick_split(int tiles, int k_blocks, int sm_count) {
    int s = 1;
    // at least two K-blocks per CTA.  (Four, until round 2: the 1024 x 512 x 512 GEMMs of BASELINE C5 then ran as 64 CTAs with
    // split 2; split 4 = 128 CTAs halves both the main loop and the rows each CTA finishes: 16.8 -> 14.6 and 18.7 -> 16.2 us.)
    while (s < 8 && tiles * (s * 2) <= sm_count + sm_count / 8 && k_blocks / (s * 2) >= 2) s *= 2;
    return s;
}

template <int KIND>
static cudaError_t launch_gemm(TensorPlan* t, PlanSlot* sl, const GemmOp& op_in, int R, cudaStream_t stream, const NetView& net,
                               float* forces, int atoms_per_replica) {
    constexpr int BK = KindOf<KIND>::kBlockElems;
    GemmOp op = op_in;
    GemmParams p = op.p;
    p.M = R;
    p.dry = t->dry;
    p.no_b_prefetch = t->no_b_prefetch ? 1 : 0;
    p.chunk = t->chunk;
    p.dbg = t->dbg_next < 8 ? t->dbg_next++ : -1;
    if (p.epi == EPI_FORCE) {
        p.out_hi = forces;
        p.ld_out = 3 * atoms_per_replica;
        const bool vec_ok = t->sel_identity && (3 * atoms_per_replica) % 4 == 0 && (reinterpret_cast<uintptr_t>(forces) & 15) == 0;
        p.sel_atoms = vec_ok ? nullptr : net.sel_atoms;
        // whole rows, 16-byte aligned: the tile can leave through the TMA engine (fp32 map, 32-column boxes, 128B swizzle)
        p.direct_store = (vec_ok && !t->no_direct_store && make_map(&op.out, KIND_TF32, forces, R, p.N, p.ld_out, BM)) ? 1 : 0;
    }
    // Tile shape.  CTA pairs (cta_group::2, 256 x 256 tiles): each SM stages 128 rows of A and HALF of the B tile and the MMA
    // reads 25% fewer shared-memory bytes per flop — the two things the 1-CTA kernel is bound by.  They need
    // enough work to fill the machine with 256 x 256 tiles (possibly split along K); small GEMMs stay on 128 x 128 tiles.
    const int kb = (p.K + BK - 1) / BK;
    int mode = t->force_bn;                                     // 0 auto, 128 / 256: 1-CTA tiles, 512: CTA pairs
    // Opt-in experiment (ANNF_GEMM_GSPLIT=-1 auto, n forces n ranges): long-K GEMMs that cannot fill the machine with unsplit
    // tiles (BASELINE C5's 1024 x 512 x 4500) on 256 x 256 CTA-pair tiles, K split over gridDim.z through global memory.
    // Motivation: the 1-CTA 128 x 128 kernel reads 8 KB of shared memory per 64-cycle MMA and is bound by shared-memory
    // bandwidth (160 KB of reads + TMA writes per K-block against 128 B/clk: 1250 cycles per block for 768 of MMA); a pair
    // reads its own A and HALF of B per MMA (1250 against 1536 cycles: MMA-bound), but splitting K inside the cluster would
    // need 16-CTA clusters.  Measured at C5 (9 ranges, 144 CTAs; profiles/r02_c5_phases_gsplit.txt): main loop 14.5 -> 9.9 us,
    // but publishing 18.9 MB of partial tiles, the rendezvous and the 9-way sum from L2 cost 13 us against the 6.7 us tail of
    // the in-cluster DSMEM reduction: 36.6 against 29.1 us per launch.  Hence off by default.
    const bool gs_forced = t->gsplit_mode > 0;                  // experiments / tests: any shape the machine holds
    if (mode == 0 && sl != nullptr && p.epi != EPI_FORCE && t->gsplit_mode != 0 && (gs_forced || (p.N >= 256 && kb >= 32))) {
        const int pair_ctas = 2 * ((p.N + 255) / 256) * ((R + 255) / 256);
        const int gsp = gs_forced ? std::min(t->gsplit_mode, kb) : std::min(t->sm_count / pair_ctas, kb / 4);
        const int tiles = pair_ctas;                              // one counter per CTA of a pair
        if (gsp >= 2 && (gs_forced || pair_ctas * gsp >= (3 * t->sm_count) / 4) && pair_ctas * gsp <= t->sm_count && tiles <= kGsMaxTiles &&
            (size_t) tiles * gsp * BM * 256 * sizeof(float) <= kGsScratchBytes) {
            p.gsplit = gsp;
            p.gs_scratch = sl->gs_scratch;
            p.gs_arrive = sl->gs_counters;
            p.gs_done = sl->gs_counters + kGsMaxTiles;
            return launch_gemm_t<KIND, 256, 1, 2>(op, p, stream);
        }
    }
    if (mode == 0) {
        const int pair_ctas = 2 * ((p.N + 255) / 256) * ((R + 255) / 256);
        const int s2 = pick_split(pair_ctas, kb, t->sm_count);
        const int tiles128 = ((p.N + 127) / 128) * ((R + BM - 1) / BM);
        // measured at BASELINE C5 (1024 replicas): the unsplit 1024x4500x512 force GEMM runs 45.0 / 38.5 / 37.3 us with
        // 128-wide / 256-wide / CTA-pair tiles; the split-K 1024x512x4500 GEMM 41.5 / 72 / 78 us (2x8 clusters and an
        // 8-way DSMEM reduce of 128 KB partials lose).  So: CTA pairs when they fill the machine WITHOUT splitting K.
        if (p.N >= 256 && s2 == 1 && pair_ctas >= t->sm_count / 2) mode = 512;
        else mode = (tiles128 > t->sm_count && pick_split(tiles128, kb, t->sm_count) == 1) ? 256 : 128;
    }
    if (mode == 132 && ((p.N + 127) / 128) % 2 == 0) {
        // 128 x 128 tiles, two N neighbours per cluster sharing the A tile by TMA multicast; split-K <= 4 keeps 2 x S portable
        const int tiles = ((p.N + 127) / 128) * ((R + BM - 1) / BM);
        // B200 co-schedules 33 clusters of 4, 24 of 6 but only 15 of 8 CTAs at 1 CTA / SM (cudaOccupancyMaxActiveClusters,
        // ANNF_CLUSTER_DEBUG=1): a grid of 16 eight-CTA clusters runs in two waves.  So: the largest split whose 2 x S
        // clusters are all co-resident.
        int s = std::min(4, pick_split(tiles, kb, t->sm_count));
        if (t->force_split > 0 && kb <= 32 && p.epi != EPI_FORCE) s = std::min(4, t->force_split);
        if (s == 4 && tiles / 2 > 15) s = 3;
        if (const char* env = getenv("ANNF_GEMM_NX_SPLIT")) s = atoi(env);
        switch (s) {
        case 4: return launch_gemm_t<KIND, 128, 4, 1, 2>(op, p, stream);
        case 3: return launch_gemm_t<KIND, 128, 3, 1, 2>(op, p, stream);
        case 2: return launch_gemm_t<KIND, 128, 2, 1, 2>(op, p, stream);
        default: return launch_gemm_t<KIND, 128, 1, 1, 2>(op, p, stream);
        }
    }
    if (mode == 132) mode = 128;
    if (mode == 130) {
        // 256 x 128 tiles on CTA pairs: each CTA stages its own 128 rows of A and 64 of the 128 rows of B (48 KB per K-block
        // instead of 64), split along K inside the same cluster (2 x S CTAs, S <= 4 keeps the cluster portable)
        const int pair_ctas = 2 * ((p.N + 127) / 128) * ((R + 255) / 256);
        int s = std::min(4, pick_split(pair_ctas, kb, t->sm_count));
        if (t->force_split > 0 && kb <= 32 && p.epi != EPI_FORCE) s = std::min(4, t->force_split);
        // 2 x 4 clusters: a B200 co-schedules 15 of them, so 16 run in two waves; 2 x 3 clusters (96 CTAs at C5) fit in one
        if (s == 4 && pair_ctas / 2 > 15) s = 3;
        if (const char* env = getenv("ANNF_GEMM_PAIR_SPLIT")) s = atoi(env);
        switch (s) {
        case 4: return launch_gemm_t<KIND, 128, 4, 2>(op, p, stream);
        case 3: return launch_gemm_t<KIND, 128, 3, 2>(op, p, stream);
        case 2: return launch_gemm_t<KIND, 128, 2, 2>(op, p, stream);
        default: return launch_gemm_t<KIND, 128, 1, 2>(op, p, stream);
        }
    }
    if (mode == 512) {
        const int pair_ctas = 2 * ((p.N + 255) / 256) * ((R + 255) / 256);
        int s = pick_split(pair_ctas, kb, t->sm_count);
        if (t->force_split > 0 && kb <= 32 && p.epi != EPI_FORCE) s = t->force_split;
        if (s == 8 && !t->no_cluster16) {
            // 2 x 8 = 16 CTAs is a non-portable cluster size: if this device cannot co-schedule it, remember and use 2 x 4
            const cudaError_t e = launch_gemm_t<KIND, 256, 8, 2>(op, p, stream);
            if (e == cudaSuccess) return e;
            cudaGetLastError();
            t->no_cluster16 = true;
        }
        if (s == 8) s = 4;
        switch (s) {
        case 4: return launch_gemm_t<KIND, 256, 4, 2>(op, p, stream);
        case 2: return launch_gemm_t<KIND, 256, 2, 2>(op, p, stream);
        default: return launch_gemm_t<KIND, 256, 1, 2>(op, p, stream);
        }
    }
    const int bn = mode == 256 ? 256 : 128;
    const int tiles = ((p.N + bn - 1) / bn) * ((R + BM - 1) / BM);
    int s = pick_split(tiles, kb, t->sm_count);
    if (t->force_split > 0 && kb <= 32 && p.epi != EPI_FORCE) s = t->force_split;      // experiments: ANNF_GEMM_SPLIT_SHORTK
    if (bn == 256) {
        switch (s) {
        case 8: return launch_gemm_t<KIND, 256, 8, 1>(op, p, stream);
        case 4: return launch_gemm_t<KIND, 256, 4, 1>(op, p, stream);
        case 2: return launch_gemm_t<KIND, 256, 2, 1>(op, p, stream);
        default: return launch_gemm_t<KIND, 256, 1, 1>(op, p, stream);
        }
    }
    switch (s) {
    case 8: return launch_gemm_t<KIND, 128, 8, 1>(op, p, stream);
    case 4: return launch_gemm_t<KIND, 128, 4, 1>(op, p, stream);
    case 2: return launch_gemm_t<KIND, 128, 2, 1>(op, p, stream);
    default: return launch_gemm_t<KIND, 128, 1, 1>(op, p, stream);
    }
}

template <int NLP>
static cudaError_t launch_mid_t(const GemmOp& fwd, const GemmOp& bwd, const MidParams& mp, const ExpChain& chain, int cl, int m_tiles, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !configured[dev]) {
        cudaError_t err = cudaFuncSetAttribute(mid_kernel<NLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMidSmemBytes);
        if (err != cudaSuccess) return err;
        configured[dev] = true;
    }
    return launch_pdl(mid_kernel<NLP>, dim3(cl, m_tiles, 1), dim3(kMidThreads), kMidSmemBytes, stream, cl, 1,
                      fwd.a_hi, fwd.a_lo, fwd.b_hi, fwd.b_lo, bwd.a_hi, bwd.a_lo, bwd.b_hi, bwd.b_lo, mp, chain);
}

static cudaError_t launch_mid(TensorPlan* t, PlanSlot* sl, const NetView& net, int R, const float* centers, double* energies,
                              const ExpChain& chain, cudaStream_t stream) {
    const int L = t->L;
    MidParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.M = R; mp.L = L;
    mp.n1 = t->nodes[L - 3]; mp.n2 = t->nodes[L - 2]; mp.nL = t->nodes[L - 1];
    mp.hidden_type = t->types[L - 3]; mp.prev_type = t->types[L - 4]; mp.out_type = t->types[L - 2]; mp.num_center = net.num_center;
    mp.bias = net.b32 + net.b_off[L - 3];
    mp.w_exp = t->w_exp[L - 3]; mp.wt_exp = t->wt_exp[L - 3]; mp.x1_exp = sl->x_exp[L - 3];
    mp.x1_hi = static_cast<const __half*>(sl->x_hi[L - 3]); mp.x1_lo = static_cast<const __half*>(sl->x_lo[L - 3]); mp.ld1 = t->ld[L - 3];
    mp.d2_hi = static_cast<__half*>(sl->d_hi[L - 2]); mp.d2_lo = static_cast<__half*>(sl->d_lo[L - 2]); mp.ld2 = t->ld[L - 2];
    mp.d1_hi = static_cast<__half*>(sl->d_hi[L - 3]); mp.d1_lo = static_cast<__half*>(sl->d_lo[L - 3]);
    mp.WL = net.W64 + net.w_off[L - 2]; mp.WL32 = net.W32 + net.w_off[L - 2]; mp.bL = net.b64 + net.b_off[L - 2];
    mp.centers = centers; mp.center_dev = net.center; mp.k_dev = net.k_dev; mp.energies = energies;
    mp.dbg = t->dbg_next < 8 ? t->dbg_next++ : -1;
    const GemmOp& fwd = sl->fwd[L - 3];               // forward connection L-3: A = X_{L-3}, B = W_{L-3}
    const GemmOp& bwd = sl->bwd[0];                   // backward connection L-3: A = D_{L-2}, B = W_{L-3}^T
    const int cl = mp.n2 <= 128 ? 1 : (mp.n2 <= 256 ? 2 : 4);
    const int m_tiles = (R + BM - 1) / BM;
    if (mp.nL <= 4) return launch_mid_t<4>(fwd, bwd, mp, chain, cl, m_tiles, stream);
    if (mp.nL <= 8) return launch_mid_t<8>(fwd, bwd, mp, chain, cl, m_tiles, stream);
    return launch_mid_t<16>(fwd, bwd, mp, chain, cl, m_tiles, stream);
}

template <int KIND>
static cudaError_t plan_run_t(TensorPlan* t, Profiler* prof, const NetView& net, int R, const float* coords,
                              int atoms_per_replica, const float* centers, float* forces, double* energies,
                              cudaStream_t stream, long long* launches, PlanSlot* sl) {
    const int L = t->L;
    char name[64];
    auto timed = [&](const char* nm, auto&& fn) {
        prof->begin(nm, stream);
        cudaError_t e = fn();
        prof->end(stream);
        if (launches) (*launches)++;
        return e;
    };
    ExpChain chain = t->chain;
    for (int l = 0; l < kMaxLayers; l++) { chain.x_exp[l] = sl->x_exp[l]; chain.d_exp[l] = sl->d_exp[l]; }
    t->dbg_next = 0;
    cudaError_t err = timed("prep_kernel", [&] {
        const bool whole_rows = t->sel_identity && (3 * atoms_per_replica) % 4 == 0 && (reinterpret_cast<uintptr_t>(coords) & 15) == 0;
        const size_t pipe_smem = 2 * (((size_t) 4 * t->nodes[0] + 127) & ~(size_t) 127);
        if (KIND == KIND_F16 && whole_rows && !t->dry && t->prep_pipelined && pipe_smem <= 200 * 1024) {
            const int grid = std::min(R, 2 * t->sm_count);
            return launch_pdl(prep_pipelined_kernel, dim3(grid), dim3(kPrepPipeThreads), pipe_smem, stream, 1, 0, net, R, coords, atoms_per_replica,
                              static_cast<__half*>(sl->x_hi[0]), static_cast<__half*>(sl->x_lo[0]), t->ld[0], chain, t->dbg_next++);
        }
        const int contiguous = (t->sel_identity && (3 * atoms_per_replica) % 4 == 0 && (reinterpret_cast<uintptr_t>(coords) & 15) == 0 ? 1 : 0) |
                               (t->dry ? 2 : 0);
        const int coalesced = (contiguous & 1) && t->prep_threads == 256 && t->prep_mode != 2 ? 4 : 0;
        const size_t row_smem = sizeof(float) * (size_t) t->nodes[0];
        if (t->prep_mode == 0 && (contiguous & 1) && !t->dry && row_smem <= 200 * 1024) {
            auto bulk = [&](auto kernel, int warps) {
                return launch_pdl(kernel, dim3(R), dim3(32 * warps), row_smem, stream, 1, 0, net, R, coords, atoms_per_replica, sl->x_hi[0], sl->x_lo[0],
                                  t->ld[0], chain, t->dbg_next++);
            };
            if (t->prep_warps == 1) return bulk(prep_bulk_kernel<KIND, 1>, 1);
            if (t->prep_warps == 2) return bulk(prep_bulk_kernel<KIND, 2>, 2);
            if (t->prep_warps == 4) return bulk(prep_bulk_kernel<KIND, 4>, 4);
            return bulk(prep_bulk_kernel<KIND, 8>, 8);
        }
        return launch_pdl(prep_kernel<KIND>, dim3(R), dim3(t->prep_threads), sizeof(float) * t->nodes[0], stream, 1, 0, net, R, coords, atoms_per_replica,
                          sl->x_hi[0], sl->x_lo[0], t->ld[0], contiguous | coalesced, chain, t->dbg_next++);
    });
    if (err != cudaSuccess) return err;
    const bool mid = KIND == KIND_F16 && t->use_mid;
    for (size_t i = 0; i < sl->fwd.size(); i++) {
        if (mid && i + 1 == sl->fwd.size()) break;        // the last hidden connection belongs to the fused middle kernel
        const GemmOp& op = sl->fwd[i];
        snprintf(name, sizeof(name), "gemm[%dx%dx%d] %s", R, op.p.N, op.p.K, strstr(op.name, "] ") + 2);
        err = timed(name, [&] { return launch_gemm<KIND>(t, sl, op, R, stream, net, nullptr, 0); });
        if (err != cudaSuccess) return err;
    }
    if (mid) {
        snprintf(name, sizeof(name), "mid[%dx%dx%d] fwd%d+head+bwd%d", R, t->nodes[L - 2], t->nodes[L - 3], L - 3, L - 3);
        err = timed(name, [&] { return launch_mid(t, sl, net, R, centers, energies, chain, stream); });
        if (err != cudaSuccess) return err;
    } else
    err = timed("head_kernel", [&] {
        return launch_pdl(head_kernel<KIND>, dim3(R), dim3(kHeadThreads), 0, stream, 1, 0, net, t->dry ? -R : R, centers, (const void*) sl->x_hi[L - 2],
                          (const void*) sl->x_lo[L - 2], t->ld[L - 2], sl->d_hi[L - 2], sl->d_lo[L - 2], energies, chain, t->dbg_next < 8 ? t->dbg_next++ : -1);
    });
    if (err != cudaSuccess) return err;
    for (size_t i = mid ? 1 : 0; i < sl->bwd.size(); i++) {
        const GemmOp& op = sl->bwd[i];
        snprintf(name, sizeof(name), "gemm[%dx%dx%d] %s", R, op.p.N, op.p.K, strstr(op.name, "] ") + 2);
        err = timed(name, [&] { return launch_gemm<KIND>(t, sl, op, R, stream, net, forces, atoms_per_replica); });
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

cudaError_t tensor_plan_run(TensorPlan* t, Profiler* prof, const NetView& net, int R, const float* coords,
                            int atoms_per_replica, const float* centers, float* forces, double* energies,
                            cudaStream_t stream, long long* launches, int slot_index) {
    PlanSlot* sl = &t->slots[slot_index % kPlanSlots];
    cudaError_t err = ensure_capacity(t, sl, net, R, stream);
    if (err != cudaSuccess) return err;
    if (t->kind == KIND_F16)
        return plan_run_t<KIND_F16>(t, prof, net, R, coords, atoms_per_replica, centers, forces, energies, stream, launches, sl);
    return plan_run_t<KIND_TF32>(t, prof, net, R, coords, atoms_per_replica, centers, forces, energies, stream, launches, sl);
}

// Diagnostic (ANNF_PHASE_CLOCKS builds): resets, or reads, the 8 x 24 %globaltimer marks of the kernels of the last step.
cudaError_t tensor_phase_clocks(unsigned long long* out192, bool reset) {
    if (reset) {
        std::vector<unsigned long long> init(8 * 24, 0ull);
        for (int s = 0; s < 8; s++) init[s * 24 + 16] = ~0ull;
        return cudaMemcpyToSymbol(g_tensor_clock, init.data(), sizeof(unsigned long long) * 8 * 24);
    }
    return cudaMemcpyFromSymbol(out192, g_tensor_clock, sizeof(unsigned long long) * 8 * 24);
}

}  // namespace annf

#ifdef ANNF_PHASE_CLOCKS
// diagnostic build only, not part of the C-ABI: per-CTA marks of the last GEMM launched (tools/cta_skew.py)
extern "C" __attribute__((visibility("default"))) int annf_debug_cta_clocks(long long* out6144, int next_slot) {
    cudaDeviceSynchronize();
    if (!out6144) {
        static unsigned long long zeros[10 * 1024];
        cudaMemcpyToSymbol(annf::g_cta_clock, zeros, sizeof(zeros));
        return (int) cudaMemcpyToSymbol(annf::g_cta_slot, &next_slot, sizeof(int));
    }
    cudaError_t e = cudaMemcpyFromSymbol(out6144, annf::g_cta_clock, sizeof(unsigned long long) * 10 * 1024);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out6144 + 10 * 1024, annf::g_sm_exit, sizeof(unsigned long long) * 8 * 256);
    return (int) e;
}
#endif
