/* annforce_b200.h — C ABI of the B200-native ANN bias-force library (libannforce_b200.so).
 *
 * This is the drop-in boundary: plain C symbols, plain pointers and sizes, no C++ or torch types,
 * no exceptions (every call returns an annf_status; annf_last_error() explains a failure).  It sits
 * underneath the reference's CUDA platform kernel
 *     CudaCalcANN_ForceKernel          /root/reference/platforms/cuda/include/CudaANN_Kernels.h:45-125
 * and replaces what that class did through OpenMM's CudaBondedUtilities
 *     initialize (upload + code generation) /root/reference/platforms/cuda/src/CudaANN_Kernels.cpp:81-365
 *     execute                               /root/reference/platforms/cuda/src/CudaANN_Kernels.cpp:367-369
 *     copyParametersToContext (empty stub)  /root/reference/platforms/cuda/src/CudaANN_Kernels.cpp:371-395
 * Arithmetic semantics are those of the reference CPU kernel
 *     ReferenceCalcANN_ForceKernel::candidate_2   /root/reference/platforms/reference/ANN_ReferenceKernels.cpp:122-171
 *
 * All device work is enqueued on the caller's stream.  Which calls wait or allocate:
 *   annf_create, annf_destroy            allocate / free, synchronise the device
 *   annf_eval_openmm                     never waits, never allocates
 *   annf_eval_batched                    never waits; on the tensor-core path (batched_path 1) the handle owns activation
 *                                        scratch sized for the largest batch seen so far: a LARGER batch than any before
 *                                        re-allocates it (device-synchronising once).  annf_reserve sizes it up front, after
 *                                        which no evaluation allocates.  Because that scratch belongs to the handle, two
 *                                        annf_eval_batched calls on DIFFERENT streams must not be in flight at once on one
 *                                        handle (use one handle per stream, as one OpenMM Context has one stream).
 *   annf_update_parameters               stream-ordered, returns without waiting (waits only if 4 earlier updates are still queued)
 *   annf_eval_batched_host               the host-buffer convenience entry: owns its streams, waits for the results
 *   annf_debug_*                         synchronise (diagnostic builds only)
 * A handle may be used from one host thread at a time; different handles are independent.
 */
#ifndef ANNFORCE_B200_H_
#define ANNFORCE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ANNF_API __attribute__((visibility("default")))
#else
#define ANNF_API
#endif

#define ANNF_ABI_VERSION 1
#define ANNF_MAX_LAYERS 8            /* node layers including the input layer */

typedef struct annf_context* annf_handle;

/* Layer types: the reference compares strings "Linear"/"Tanh"/"Circular"/"Softmax"
 * (ANN_ReferenceKernels.cpp:200-238); the legacy CUDA path mapped them to ints (CudaANN_Kernels.cpp:115-116). */
enum { ANNF_LINEAR = 0, ANNF_TANH = 1, ANNF_CIRCULAR = 2, ANNF_SOFTMAX = 3 };

/* data_type_in_input_layer (ANN_Force.h:137; ANN_ReferenceKernels.cpp:125,128,152) */
enum { ANNF_INPUT_DIHEDRAL_COS_SIN = 0, ANNF_INPUT_CARTESIAN = 1, ANNF_INPUT_PAIR_DISTANCES = 2 };

/* Arithmetic of the MLP contractions.  Feature centring, the output layer, the energy and the
 * output-layer gradient are always fp64.
 *   FP64    everything fp64 (B200 runs fp64 FMA at half the fp32 rate): parity to ~1e-12
 *   FP32    hidden-layer contractions in fp32 FMA
 *   TF32X3  hidden-layer contractions on tcgen05 tensor cores, 3xTF32 split, fp32 accumulate in TMEM
 *   F16X3   the same 11 + 11-bit operand split carried in fp16 pairs with per-row power-of-two scales
 *           (tcgen05 kind::f16): identical significand budget, half the operand bytes, twice the MMA rate
 *   AUTO    FP64 for the single-replica OpenMM entry and for small networks; F16X3 for layers wide
 *           enough to be a real dense contraction; FP32 otherwise */
enum { ANNF_PREC_AUTO = 0, ANNF_PREC_FP64 = 1, ANNF_PREC_FP32 = 2, ANNF_PREC_TF32X3 = 3, ANNF_PREC_F16X3 = 4 };

typedef enum {
    ANNF_OK = 0,
    ANNF_ERR_INVALID = 1,       /* bad argument / inconsistent description */
    ANNF_ERR_UNSUPPORTED = 2,   /* valid in the reference, not available on the requested path (message says which) */
    ANNF_ERR_CUDA = 3,          /* a CUDA runtime call failed; message carries cudaGetErrorString */
    ANNF_ERR_NO_DEVICE = 4      /* no CUDA device: there is NO CPU fallback */
} annf_status;

/* Everything ANN_Force carries (openmmapi/include/openmm/ANN_Force.h:126-137).  Host pointers, read
 * during annf_create only. */
typedef struct {
    uint32_t struct_size;                 /* sizeof(annf_desc), for ABI checking */
    int32_t device;                       /* CUDA device ordinal, or -1 for the current device */
    int32_t num_layers;                   /* node layers incl. input (the reference's NUM_OF_LAYERS), 2..ANNF_MAX_LAYERS */
    const int32_t* num_of_nodes;          /* [num_layers] */
    const int32_t* layer_types;           /* [num_layers-1] ANNF_LINEAR.. */
    const double* const* coeff;           /* [num_layers-1] row-major [n_out][n_in] (ANN_ReferenceKernels.cpp:81-83) */
    const double* const* bias;            /* [num_layers-1] [n_out] */
    const double* potential_center;       /* [num_center]: n_last, or n_last/2 for a Circular output layer */
    int32_t num_center;
    double force_constant;
    double scaling_factor;
    int32_t data_type_in_input_layer;
    const int32_t* index_of_backbone_atoms;   /* 1-based, as ANN_Force::set_index_of_backbone_atoms takes them */
    int32_t num_backbone_atoms;
    const int32_t* dihedrals;             /* 0-based quadruples (what ANN_Force stores, ANN_Force.cpp:111) */
    int32_t num_dihedrals;
    const int32_t* pairs;                 /* 0-based pairs (ANN_Force.cpp:164) */
    int32_t num_pairs;
    int32_t num_system_atoms;             /* atoms in the System; every index must be below it */
    int32_t precision;                    /* ANNF_PREC_* */
} annf_desc;

typedef struct {
    int32_t precision_single;             /* resolved ANNF_PREC_* of annf_eval_openmm */
    int32_t precision_batched;            /* resolved ANNF_PREC_* of annf_eval_batched */
    int32_t batched_path;                 /* 0 generic fused per-CTA kernel, 1 layer-wise tensor-core path, 2 fused kernel with TMA-streamed weights,
                                           * 3 the same on the FP64 tensor path (DMMA) */
    int32_t cluster_size;                 /* CTAs per cluster of the single-replica kernel */
    int64_t macs_per_eval;                /* M = sum n_l * n_{l+1} */
    int64_t kernel_launches;              /* kernels launched through this handle since creation */
    int32_t sm_count;
    int32_t single_path;                  /* 0 generic cluster kernel, 1 low-latency kernel (Cartesian input, Tanh / Linear hidden layers) */
} annf_info;

/* One named kernel's accumulated device time while profiling is on (annf_set_profiling). */
typedef struct {
    char name[48];
    int64_t launches;
    double total_ms;
} annf_kernel_time;

ANNF_API const char* annf_version(void);
ANNF_API const char* annf_last_error(void);        /* thread-local, valid until the next failing call on this thread */
ANNF_API int annf_device_count(void);              /* 0 when no usable CUDA device / driver */

/* initialize(): validate, upload (fp64 + fp32 + tensor-core operand copies), pick kernels. */
ANNF_API annf_status annf_create(const annf_desc* desc, annf_handle* out);
ANNF_API annf_status annf_destroy(annf_handle h);
ANNF_API annf_status annf_get_info(annf_handle h, annf_info* info);

/* copyParametersToContext(): move the umbrella centre and force constant on the device.
 * Ordered on `stream` (the host arrays are copied before the call returns; the call does not wait for the device); takes
 * effect for evaluations enqueued on `stream` after it.  (Both reference platforms
 * leave this an empty stub: ANN_ReferenceKernels.cpp:96-111, CudaANN_Kernels.cpp:371-395.) */
ANNF_API annf_status annf_update_parameters(annf_handle h, const double* potential_center, int32_t num_center,
                                   double force_constant, void* stream);

/* Pre-size the per-handle scratch of annf_eval_batched for batches of up to max_replicas, so that no later evaluation
 * allocates (see the preamble).  A no-op on paths that keep no per-batch scratch. */
ANNF_API annf_status annf_reserve(annf_handle h, int32_t max_replicas, void* stream);

/* OpenMM's CUDA platform re-orders atoms; atom_index_dev[slot] = original atom index (device
 * pointer, CudaContext::getAtomIndexArray()).  Call once after creation and from a reorder
 * listener.  Without it, slot == original index. */
ANNF_API annf_status annf_set_atom_order(annf_handle h, const int32_t* atom_index_dev, int32_t num_atoms, void* stream);

/* execute() for one replica on OpenMM's device buffers.  One fused launch: gather -> scale/centre ->
 * forward -> energy -> backward -> centring Jacobian -> atomic scatter.
 *   posq             float4[padded] (single/mixed) or double4[padded] (double)       CudaContext::getPosq()
 *   posq_correction  float4[padded] low-order bits in mixed mode, or NULL            CudaContext::getPosqCorrection()
 *   force_buffer     64-bit fixed point (scale 2^32), [x: padded][y: padded][z: padded], ACCUMULATED with atomicAdd
 *   energy_buffer    float* or double* accumulation buffer; this force's energy is ADDED to element 0; may be NULL
 *   flags            ANNF_FLAG_* */
#define ANNF_FLAG_POSQ_DOUBLE    1u
#define ANNF_FLAG_ENERGY_DOUBLE  2u
#define ANNF_FLAG_SKIP_FORCES    4u   /* includeForces == false */
#define ANNF_FLAG_SKIP_ENERGY    8u   /* includeEnergy == false */
ANNF_API annf_status annf_eval_openmm(annf_handle h, const void* posq, const void* posq_correction,
                             unsigned long long* force_buffer, int32_t padded_num_atoms,
                             void* energy_buffer, uint32_t flags, void* stream);

/* Many independent replicas (umbrella windows / walkers) per call, device buffers:
 *   coords   float[R][atoms_per_replica][3]  (nm)            atoms_per_replica >= num_system_atoms
 *   centers  float[R][num_center] per-replica umbrella centres, or NULL to use the handle's centre
 *   forces   float[R][atoms_per_replica][3]  this force's contribution, WRITTEN for the selected atoms only
 *   energies double[R] */
ANNF_API annf_status annf_eval_batched(annf_handle h, int32_t num_replicas, const float* coords, int32_t atoms_per_replica,
                              const float* centers, float* forces, double* energies, void* stream);

/* The same with HOST buffers: copies host->device in chunks, evaluates each chunk as it lands, copies forces and energies
 * device->host and waits.  The copies are cudaMemcpyAsync straight from / to the caller's buffers: with PINNED (page-locked)
 * buffers they overlap the evaluation of neighbouring chunks; with pageable buffers the results are the same but every copy
 * is staged by the driver and the overlap is lost (measured cost in DESIGN.md 5).  forces_host is overwritten completely (zeros for atoms
 * outside the selection).  This is the end-to-end entry bench.py times. */
ANNF_API annf_status annf_eval_batched_host(annf_handle h, int32_t num_replicas, const float* coords_host,
                                   int32_t atoms_per_replica, const float* centers_host, float* forces_host,
                                   double* energies_host);

/* Diagnostics, available in builds made with -DANNF_PHASE_CLOCKS only (otherwise ANNF_ERR_UNSUPPORTED; the product kernels carry
 * no instrumentation).  SM clock (clock64) of thread 0 at the phase boundaries of the most recent annf_eval_openmm launch:
 * [0] start, [1] features, [2] weights staged, [3] forward, [4] energy head, [5] backward, [6] forces issued.  Synchronises. */
ANNF_API annf_status annf_debug_single_phases(annf_handle h, int64_t* clocks16);

/* Diagnostic: the same for CTA 0 of the most recent streamed batched launch (annf_eval_batched, batched_path 2 or 3):
 * [0] start, [1] prologue loads, [2] features, [3 + p] contraction p done (forward then backward), then head and end.
 * 64 values.  Synchronises. */
ANNF_API annf_status annf_debug_stream_phases(annf_handle h, int64_t* clocks64);

/* Diagnostic: the tensor-core path (batched_path 1).  8 kernels x 24 values of %globaltimer (ns) from the most recent
 * annf_eval_batched: [k][0..9] phase boundaries of CTA 0 of kernel k of the step (entry, prologue done, dependencies
 * resolved, first TMA, first operands landed, last MMA issued, accumulators drained, partials parked, epilogue done, exit),
 * [k][16] / [k][17] first / last CTA start, [k][18] last CTA end.  reset != 0 clears the marks instead.  Synchronises. */
ANNF_API annf_status annf_debug_tensor_phases(annf_handle h, int64_t* clocks192, int32_t reset);

/* Per-kernel device timing with CUDA events on the launching stream (for the roofline line). */
ANNF_API annf_status annf_set_profiling(annf_handle h, int32_t enabled);
ANNF_API annf_status annf_get_kernel_times(annf_handle h, annf_kernel_time* out, int32_t capacity, int32_t* count, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* ANNFORCE_B200_H_ */
