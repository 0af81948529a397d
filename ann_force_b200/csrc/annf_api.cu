// C-ABI of libannforce_b200 (declared in include/annforce_b200.h): handle life-cycle, parameter
// upload, kernel selection and the launchers.  No exceptions cross this boundary and there is no
// CPU fallback: without a CUDA device every entry point fails with ANNF_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "annf_common.cuh"
#include "annf_profiler.h"

namespace annf {
// annf_single.cu
size_t single_smem_bytes(const NetView& net, int cluster_size);
cudaError_t prepare_single(const NetView& net, int cluster_size);
cudaError_t read_single_phase_clocks(long long* out16);
cudaError_t launch_single(const NetView& net, int cluster_size, const void* posq, const void* posq_corr,
                          unsigned long long* force, int padded, void* energy, unsigned flags, cudaStream_t stream);
bool fast_single_plan(const NetView& net, int cluster_size, const std::vector<std::vector<double>>& w, const std::vector<double>& b64,
                      FastPlan* plan, std::vector<double>* blobs);
cudaError_t prepare_single_fast();
cudaError_t launch_single_fast(const NetView& net, const FastPlan& plan, const double* blobs, const void* posq, const void* posq_corr,
                               unsigned long long* force, int padded, void* energy, unsigned flags, cudaStream_t stream, bool allow_pdl);
// annf_batched.cu
size_t batched_smem_bytes(const NetView& net, int extra_nodes, int tr, bool fp64);
cudaError_t launch_batched(const NetView& net, const int* x_off, int extra_nodes, int tr, bool fp64, int R,
                           const float* coords, int atoms_per_replica, const float* centers, float* forces,
                           double* energies, cudaStream_t stream);
// annf_stream.cu
bool build_stream(const NetView& net, const std::vector<std::vector<double>>& w, const std::vector<double>& b64,
                  const std::vector<int>& sel_atoms, bool fp64, StreamPlan* plan, std::vector<double>* stream64,
                  std::vector<float>* stream32, std::vector<unsigned char>* statics);
cudaError_t read_stream_phase_clocks(long long* out32);
int stream_pick_tile(const NetView& net, const StreamPlan& plan, bool fp64, int R, int sm_count);
cudaError_t launch_stream(const NetView& net, const StreamPlan& plan, const void* statics_dev, int static_bytes, const void* stream_dev,
                          bool fp64, int tr, int R, const float* coords, int atoms_per_replica,
                          const float* centers, float* forces, double* energies, cudaStream_t stream);
// annf_dmma.cu
bool build_dmma_stream(const NetView& net, const std::vector<std::vector<double>>& w, const std::vector<double>& b64,
                       const std::vector<int>& sel, StreamPlan* plan, std::vector<double>* stream, std::vector<unsigned char>* statics);
cudaError_t launch_dmma(const NetView& net, const StreamPlan& plan, const void* statics_dev, int static_bytes, const double* stream_dev,
                        int R, const float* coords, int atoms_per_replica, const float* centers, float* forces, double* energies,
                        cudaStream_t stream);
cudaError_t read_dmma_phase_clocks(long long* out64);
// annf_tensor.cu
struct TensorPlan;
TensorPlan* tensor_plan_create(const NetView& net, const std::vector<std::vector<double>>& coeff, const std::vector<double>& bias64,
                               int kind, std::string* why_not);     // kind: 0 = TF32 pairs, 1 = scaled fp16 pairs
void tensor_plan_destroy(TensorPlan* plan);
cudaError_t tensor_plan_reserve(TensorPlan* plan, const NetView& net, int R, cudaStream_t stream);
cudaError_t tensor_phase_clocks(unsigned long long* out192, bool reset);
cudaError_t tensor_plan_run(TensorPlan* plan, Profiler* prof, const NetView& net, int R,
                            const float* coords, int atoms_per_replica, const float* centers, float* forces,
                            double* energies, cudaStream_t stream, long long* launches, int slot_index);
}  // namespace annf

using namespace annf;

namespace {

thread_local std::string g_last_error;

annf_status fail(annf_status code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define ANNF_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess)                                                                    \
            return fail(ANNF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    const size_t bytes = sizeof(T) * std::max<size_t>(src.size(), 1);
    cudaError_t err = cudaMalloc((void**) dst, bytes);
    if (err != cudaSuccess) return err;
    if (!src.empty()) err = cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice);
    return err;
}

__global__ void invert_order_kernel(const int* __restrict__ atom_index, int num_atoms, int* __restrict__ inverse) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < num_atoms) inverse[atom_index[slot]] = slot;
}

__global__ void remap_kernel(const int* __restrict__ atoms, int n, const int* __restrict__ inverse, int* __restrict__ slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slots[i] = inverse[atoms[i]];
}

__global__ void set_parameters_kernel(double* center, const double* staged, int n, double* k_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) center[i] = staged[i];
    if (i == 0) *k_dev = staged[n];
}

}  // namespace

struct annf_context {
    int device = 0;
    int sm_count = 0;
    NetView net{};
    std::vector<int> nodes, types;
    int num_system_atoms = 0;
    long long macs = 0;
    int prec_single = ANNF_PREC_FP64, prec_batched = ANNF_PREC_FP64;
    int batched_path = 0;
    int cluster_size = 1;
    bool order_dirty = true;         // sel_slots was (re)written since the last single-replica launch: that launch must not use PDL
    bool fast_single = false;        // low-latency single-replica kernel applies (Cartesian input, Tanh / Linear hidden layers)
    FastPlan fast_plan{};
    double* fast_blobs = nullptr;    // per-CTA parameter blobs of the low-latency single-replica kernel
    bool dmma = false;               // batched path 3: the streamed kernel on the FP64 tensor path (fp64 only)
    bool streamed = false;           // batched path 2: fused kernel with TMA-streamed weights
    StreamPlan stream_plan{};
    void* stream_dev = nullptr;
    unsigned char* stream_statics = nullptr;   // plan | biases | selection: bulk-copied into shared memory by every CTA
    int stream_static_bytes = 0;          // [n_0][n_1] first connection, transposed, with the centring / scaling Jacobian folded in
    int tile_fp64 = 1, tile_fp32 = 1;
    int x_off[kMaxLayers];
    int extra_nodes = 0;
    long long launches = 0;
    // device allocations
    double *W64 = nullptr, *Wt64 = nullptr, *b64 = nullptr, *center = nullptr, *k_dev = nullptr, *staged = nullptr;
    float *W32 = nullptr, *Wt32 = nullptr, *b32 = nullptr;
    int *sel_atoms = nullptr, *sel_slots = nullptr, *inverse = nullptr, *pairs_local = nullptr, *dihedrals_local = nullptr;
    // annf_update_parameters stages through a ring of pinned slots + device slots, each guarded by an event, so the call
    // is stream-ordered and returns without waiting (it only waits if kParamSlots updates are still in flight)
    static constexpr int kParamSlots = 4;
    double* staged_host = nullptr;   // pinned, [kParamSlots][num_center + 1]
    cudaEvent_t param_ev[kParamSlots] = {};
    int param_next = 0;
    TensorPlan* tensor = nullptr;
    int tensor_slot = 0;             // which activation-buffer set of the tensor plan the next batched call uses
    Profiler prof;
    // host-buffer entry
    cudaStream_t host_stream = nullptr, host_stream2 = nullptr, in_stream = nullptr, out_stream = nullptr;
    cudaEvent_t ev_in[8] = {}, ev_comp[8] = {};
    float *h_coords = nullptr, *h_centers = nullptr, *h_forces = nullptr;
    double* h_energies = nullptr;
    size_t h_cap_coords = 0, h_cap_centers = 0, h_cap_energies = 0;
};

namespace {

struct DeviceGuard {
    int previous = -1;
    explicit DeviceGuard(int device) {
        cudaGetDevice(&previous);
        if (previous != device) cudaSetDevice(device);
        else previous = -1;
    }
    ~DeviceGuard() { if (previous >= 0) cudaSetDevice(previous); }
};

void release(annf_context* h) {
    DeviceGuard guard(h->device);
    h->prof.drain();
    if (h->tensor) tensor_plan_destroy(h->tensor);
    cudaFree(h->W64); cudaFree(h->Wt64); cudaFree(h->b64); cudaFree(h->center); cudaFree(h->k_dev); cudaFree(h->staged);
    cudaFree(h->W32); cudaFree(h->Wt32); cudaFree(h->b32); cudaFree(h->fast_blobs);
    cudaFree(h->stream_dev); cudaFree(h->stream_statics);
    cudaFree(h->sel_atoms); cudaFree(h->sel_slots); cudaFree(h->inverse); cudaFree(h->pairs_local); cudaFree(h->dihedrals_local);
    cudaFree(h->h_coords); cudaFree(h->h_centers); cudaFree(h->h_forces); cudaFree(h->h_energies);
    if (h->staged_host) cudaFreeHost(h->staged_host);
    for (int i = 0; i < annf_context::kParamSlots; i++)
        if (h->param_ev[i]) cudaEventDestroy(h->param_ev[i]);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    if (h->host_stream2) cudaStreamDestroy(h->host_stream2);
    if (h->in_stream) cudaStreamDestroy(h->in_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    for (int i = 0; i < 8; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
    }
    delete h;
}

}  // namespace

extern "C" {

const char* annf_version(void) { return "annforce_b200 0.1 (sm_100a, abi 1)"; }

const char* annf_last_error(void) { return g_last_error.c_str(); }

int annf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

annf_status annf_create(const annf_desc* desc, annf_handle* out) {
    if (out == nullptr) return fail(ANNF_ERR_INVALID, "annf_create: out is NULL");
    *out = nullptr;
    if (desc == nullptr || desc->struct_size != sizeof(annf_desc))
        return fail(ANNF_ERR_INVALID, "annf_create: descriptor missing or built against another ABI (struct_size %u, expected %zu)",
                    desc ? desc->struct_size : 0u, sizeof(annf_desc));
    const int L = desc->num_layers;
    if (L < 2 || L > ANNF_MAX_LAYERS)
        return fail(ANNF_ERR_INVALID, "num_layers = %d, supported range is 2..%d", L, ANNF_MAX_LAYERS);
    if (!desc->num_of_nodes || !desc->layer_types || !desc->coeff || !desc->bias || !desc->potential_center)
        return fail(ANNF_ERR_INVALID, "annf_create: NULL array in descriptor");
    for (int l = 0; l < L; l++)
        if (desc->num_of_nodes[l] <= 0) return fail(ANNF_ERR_INVALID, "num_of_nodes[%d] = %d", l, desc->num_of_nodes[l]);
    for (int l = 0; l + 1 < L; l++) {
        const int t = desc->layer_types[l];
        if (t < ANNF_LINEAR || t > ANNF_SOFTMAX) return fail(ANNF_ERR_INVALID, "layer_types[%d] = %d is not a layer type", l, t);
        if (t == ANNF_CIRCULAR && desc->num_of_nodes[l + 1] % 2 != 0)
            return fail(ANNF_ERR_INVALID, "Circular layer %d needs an even number of nodes", l + 1);   // ANN_ReferenceKernels.cpp:211
        if (!desc->coeff[l] || !desc->bias[l]) return fail(ANNF_ERR_INVALID, "coeff[%d] / bias[%d] is NULL", l, l);
    }
    const int n_last = desc->num_of_nodes[L - 1];
    const int want_center = desc->layer_types[L - 2] == ANNF_CIRCULAR ? n_last / 2 : n_last;   // ANN_ReferenceKernels.cpp:61-66
    if (desc->num_center != want_center)
        return fail(ANNF_ERR_INVALID, "potential_center has %d values, the output layer needs %d", desc->num_center, want_center);
    if (desc->scaling_factor == 0.0 || !std::isfinite(desc->scaling_factor))
        return fail(ANNF_ERR_INVALID, "scaling_factor must be finite and non-zero");
    if (desc->num_system_atoms <= 0) return fail(ANNF_ERR_INVALID, "num_system_atoms must be positive");
    if (desc->data_type_in_input_layer == ANNF_INPUT_CARTESIAN) {
        if (desc->num_backbone_atoms <= 0 || 3 * desc->num_backbone_atoms != desc->num_of_nodes[0])
            return fail(ANNF_ERR_INVALID, "Cartesian input: 3 * %d selected atoms != %d input nodes",
                        desc->num_backbone_atoms, desc->num_of_nodes[0]);                          // CudaANN_Kernels.cpp:188
        std::vector<char> seen((size_t) desc->num_system_atoms + 1, 0);
        for (int i = 0; i < desc->num_backbone_atoms; i++) {
            const int a = desc->index_of_backbone_atoms[i];
            if (a < 1 || a > desc->num_system_atoms)
                return fail(ANNF_ERR_INVALID, "index_of_backbone_atoms[%d] = %d is outside 1..%d (indices are 1-based)", i, a,
                            desc->num_system_atoms);
            // The reference accumulates forces per list entry (ANN_ReferenceKernels.cpp:401), so a repeated atom would need
            // its contributions summed; the batched kernels write one force row per selected atom.  A repeated atom in a
            // backbone list is a user error in every known use, so it is rejected here rather than evaluated two ways.
            if (seen[a])
                return fail(ANNF_ERR_INVALID, "index_of_backbone_atoms lists atom %d more than once (entry %d); each selected atom must "
                            "appear once", a, i);
            seen[a] = 1;
        }
    } else if (desc->data_type_in_input_layer == ANNF_INPUT_DIHEDRAL_COS_SIN) {
        if (desc->num_dihedrals <= 0 || !desc->dihedrals || 2 * desc->num_dihedrals != desc->num_of_nodes[0])
            return fail(ANNF_ERR_INVALID, "dihedral input: 2 * %d dihedrals != %d input nodes", desc->num_dihedrals, desc->num_of_nodes[0]);
        for (int i = 0; i < 4 * desc->num_dihedrals; i++)
            if (desc->dihedrals[i] < 0 || desc->dihedrals[i] >= desc->num_system_atoms)
                return fail(ANNF_ERR_INVALID, "dihedral atom index %d is outside the %d-atom system", desc->dihedrals[i], desc->num_system_atoms);
    } else if (desc->data_type_in_input_layer == ANNF_INPUT_PAIR_DISTANCES) {
        if (desc->num_pairs <= 0 || !desc->pairs || desc->num_pairs != desc->num_of_nodes[0])
            return fail(ANNF_ERR_INVALID, "pair-distance input: %d pairs != %d input nodes", desc->num_pairs, desc->num_of_nodes[0]);   // CudaANN_Kernels.cpp:211
        for (int i = 0; i < 2 * desc->num_pairs; i++)
            if (desc->pairs[i] < 0 || desc->pairs[i] >= desc->num_system_atoms)
                return fail(ANNF_ERR_INVALID, "pair atom index %d is outside the %d-atom system", desc->pairs[i], desc->num_system_atoms);
    } else {
        return fail(ANNF_ERR_INVALID, "data_type_in_input_layer = %d", desc->data_type_in_input_layer);
    }
    if (desc->precision < ANNF_PREC_AUTO || desc->precision > ANNF_PREC_F16X3)
        return fail(ANNF_ERR_INVALID, "precision = %d", desc->precision);

    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(ANNF_ERR_NO_DEVICE, "no CUDA device is available; this library has no CPU fallback");
    }
    int device = desc->device;
    if (device < 0) ANNF_CUDA(cudaGetDevice(&device));
    if (device >= count) return fail(ANNF_ERR_INVALID, "device %d requested, %d present", device, count);
    cudaDeviceProp prop;
    ANNF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(ANNF_ERR_UNSUPPORTED, "device %d is sm_%d%d; this build contains sm_100a code only", device, prop.major, prop.minor);

    annf_context* h = new (std::nothrow) annf_context();
    if (!h) return fail(ANNF_ERR_INVALID, "out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->num_system_atoms = desc->num_system_atoms;
    DeviceGuard guard(device);

    // ---- flatten the network --------------------------------------------------------------------------
    NetView& net = h->net;
    net.L = L;
    std::vector<double> W64, Wt64, b64;
    std::vector<std::vector<double>> coeff_rows;
    net.node_off[0] = 0;
    for (int l = 0; l < L; l++) {
        net.nodes[l] = desc->num_of_nodes[l];
        net.node_off[l + 1] = net.node_off[l] + net.nodes[l];
        h->nodes.push_back(net.nodes[l]);
    }
    for (int l = 0; l + 1 < L; l++) {
        const int n_in = net.nodes[l], n_out = net.nodes[l + 1];
        net.types[l] = desc->layer_types[l];
        h->types.push_back(net.types[l]);
        net.w_off[l] = (long long) W64.size();
        net.b_off[l] = (int) b64.size();
        h->macs += (long long) n_in * n_out;
        const double* w = desc->coeff[l];
        for (size_t t = 0; t < (size_t) n_in * n_out; t++)
            if (!std::isfinite(w[t])) { delete h; return fail(ANNF_ERR_INVALID, "coeff[%d][%zu] is not finite", l, t); }
        W64.insert(W64.end(), w, w + (size_t) n_in * n_out);
        Wt64.resize(W64.size());
        double* wt = Wt64.data() + net.w_off[l];
        for (int j = 0; j < n_out; j++)
            for (int i = 0; i < n_in; i++) wt[(size_t) i * n_out + j] = w[(size_t) j * n_in + i];
        b64.insert(b64.end(), desc->bias[l], desc->bias[l] + n_out);
        coeff_rows.emplace_back(w, w + (size_t) n_in * n_out);
    }
    std::vector<float> W32(W64.begin(), W64.end()), Wt32(Wt64.begin(), Wt64.end()), b32(b64.begin(), b64.end());
    std::vector<double> center(desc->potential_center, desc->potential_center + desc->num_center);
    std::vector<double> kvec(1, desc->force_constant);
    // feature atoms: the Cartesian selection as given, or the sorted union of the atoms of all pairs / dihedrals, with the
    // pair / dihedral lists re-expressed as indices into that union
    std::vector<int> sel, pairs_local, dihedrals_local;
    if (desc->data_type_in_input_layer == ANNF_INPUT_CARTESIAN) {
        sel.resize(desc->num_backbone_atoms);
        for (int i = 0; i < desc->num_backbone_atoms; i++) sel[i] = desc->index_of_backbone_atoms[i] - 1;   // ANN_ReferenceKernels.cpp:132
    } else {
        const bool is_pairs = desc->data_type_in_input_layer == ANNF_INPUT_PAIR_DISTANCES;
        const int* list = is_pairs ? desc->pairs : desc->dihedrals;
        const int count = is_pairs ? 2 * desc->num_pairs : 4 * desc->num_dihedrals;
        sel.assign(list, list + count);
        std::sort(sel.begin(), sel.end());
        sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
        std::vector<int>& local = is_pairs ? pairs_local : dihedrals_local;
        local.resize(count);
        for (int i = 0; i < count; i++) local[i] = (int) (std::lower_bound(sel.begin(), sel.end(), list[i]) - sel.begin());
    }

#define ANNF_CUDA_H(call)                                                                            \
    do {                                                                                             \
        cudaError_t err__ = (call);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            release(h);                                                                              \
            return fail(ANNF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
        }                                                                                            \
    } while (0)

    ANNF_CUDA_H(upload(&h->W64, W64));
    ANNF_CUDA_H(upload(&h->Wt64, Wt64));
    ANNF_CUDA_H(upload(&h->b64, b64));
    ANNF_CUDA_H(upload(&h->W32, W32));
    ANNF_CUDA_H(upload(&h->Wt32, Wt32));
    ANNF_CUDA_H(upload(&h->b32, b32));
    ANNF_CUDA_H(upload(&h->center, center));
    ANNF_CUDA_H(upload(&h->k_dev, kvec));
    ANNF_CUDA_H(upload(&h->sel_atoms, sel));
    ANNF_CUDA_H(upload(&h->sel_slots, sel));
    ANNF_CUDA_H(upload(&h->pairs_local, pairs_local));
    ANNF_CUDA_H(upload(&h->dihedrals_local, dihedrals_local));
    ANNF_CUDA_H(cudaMalloc((void**) &h->inverse, sizeof(int) * h->num_system_atoms));
    ANNF_CUDA_H(cudaMalloc((void**) &h->staged, sizeof(double) * annf_context::kParamSlots * (desc->num_center + 1)));
    ANNF_CUDA_H(cudaMallocHost((void**) &h->staged_host, sizeof(double) * annf_context::kParamSlots * (desc->num_center + 1)));
    for (int i = 0; i < annf_context::kParamSlots; i++) ANNF_CUDA_H(cudaEventCreateWithFlags(&h->param_ev[i], cudaEventDisableTiming));
    net.W64 = h->W64; net.Wt64 = h->Wt64; net.b64 = h->b64;
    net.W32 = h->W32; net.Wt32 = h->Wt32; net.b32 = h->b32;
    net.center = h->center; net.k_dev = h->k_dev; net.num_center = desc->num_center;
    net.scale = desc->scaling_factor;
    net.input_type = desc->data_type_in_input_layer;
    net.n_sel = (int) sel.size();
    net.sel_atoms = h->sel_atoms; net.sel_slots = h->sel_slots;
    net.n_pairs = (int) pairs_local.size() / 2; net.pairs = h->pairs_local;
    net.n_dihedrals = (int) dihedrals_local.size() / 4; net.dihedrals = h->dihedrals_local;

    // ---- kernel selection ---------------------------------------------------------------------------------
    // single replica: fp64 always (latency-bound; B200 fp64 FMA is half rate, parity is ~1e-12)
    h->prec_single = ANNF_PREC_FP64;
    const double weight_bytes = 8.0 * (double) h->macs;
    int cl = 1;
    while (cl < 8 && weight_bytes / cl > 48.0 * 1024) cl *= 2;
    h->cluster_size = cl;
    if (single_smem_bytes(net, cl) > 226 * 1024) {
        release(h);
        return fail(ANNF_ERR_UNSUPPORTED, "network too large for the single-replica kernel's shared memory (%zu bytes)",
                    single_smem_bytes(net, cl));
    }
    ANNF_CUDA_H(prepare_single(net, cl));
    {
        std::vector<double> blobs;
        h->fast_single = fast_single_plan(net, cl, coeff_rows, b64, &h->fast_plan, &blobs);
        if (h->fast_single) {
            ANNF_CUDA_H(upload(&h->fast_blobs, blobs));
            ANNF_CUDA_H(prepare_single_fast());
        }
    }

    // batched: extra shared-memory slots for non-elementwise hidden layers
    h->extra_nodes = 0;
    for (int l = 0; l < kMaxLayers; l++) h->x_off[l] = -1;
    for (int l = 1; l + 1 < L; l++) {
        const int t = net.types[l - 1];
        if (t == ANNF_CIRCULAR || t == ANNF_SOFTMAX) {
            h->x_off[l] = h->extra_nodes;
            h->extra_nodes += 2 * net.nodes[l];
        }
    }
    auto pick_tile = [&](bool fp64) {
        for (int tr : {8, 4, 2, 1})
            if (batched_smem_bytes(net, h->extra_nodes, tr, fp64) <= 200 * 1024) return tr;
        return 0;
    };
    h->tile_fp64 = net.nodes[L - 1] <= 256 ? pick_tile(true) : 0;    // the fused kernel's head scratch holds 256 outputs
    h->tile_fp32 = net.nodes[L - 1] <= 256 ? pick_tile(false) : 0;
    int prec = desc->precision;
    std::string why_not;
    // AUTO takes the fp16-pair tensor path for wide networks: same 11 + 11-bit operand split as 3xTF32, half the operand
    // bytes and twice the MMA rate; ANNF_AUTO_TENSOR=tf32x3 switches AUTO back for comparisons.
    int auto_tensor = ANNF_PREC_F16X3;
    if (const char* env = getenv("ANNF_AUTO_TENSOR"))
        if (strcmp(env, "tf32x3") == 0) auto_tensor = ANNF_PREC_TF32X3;
    const bool small_net = h->macs <= (1 << 16) && h->tile_fp64 > 0;
    if (prec == ANNF_PREC_TF32X3 || prec == ANNF_PREC_F16X3 || (prec == ANNF_PREC_AUTO && !small_net)) {
        const int want = prec == ANNF_PREC_AUTO ? auto_tensor : prec;
        h->tensor = tensor_plan_create(net, coeff_rows, b64, want == ANNF_PREC_F16X3 ? 1 : 0, &why_not);
        if (h->tensor == nullptr && prec != ANNF_PREC_AUTO) {
            release(h);
            return fail(ANNF_ERR_UNSUPPORTED, "%s tensor-core path unavailable for this network: %s",
                        prec == ANNF_PREC_F16X3 ? "F16X3" : "TF32X3", why_not.c_str());
        }
        if (h->tensor != nullptr) prec = want;
    }
    if (prec == ANNF_PREC_AUTO) prec = small_net ? ANNF_PREC_FP64 : ANNF_PREC_FP32;
    const bool tensor_path = prec == ANNF_PREC_TF32X3 || prec == ANNF_PREC_F16X3;
    h->prec_batched = prec;
    h->batched_path = tensor_path ? 1 : 0;
    if (!tensor_path) {
        std::vector<double> s64;
        std::vector<float> s32;
        std::vector<unsigned char> statics;
        const bool fp64 = prec == ANNF_PREC_FP64;
        if (fp64 && build_dmma_stream(net, coeff_rows, b64, sel, &h->stream_plan, &s64, &statics)) {
            double* a = nullptr;
            ANNF_CUDA_H(upload(&a, s64)); h->stream_dev = a;
            ANNF_CUDA_H(upload(&h->stream_statics, statics));
            h->stream_static_bytes = (int) statics.size();
            h->dmma = true;
            h->batched_path = 3;
        } else if (build_stream(net, coeff_rows, b64, sel, fp64, &h->stream_plan, &s64, &s32, &statics) &&
            stream_pick_tile(net, h->stream_plan, fp64, 1, h->sm_count) > 0) {
            if (fp64) {
                double* a = nullptr;
                ANNF_CUDA_H(upload(&a, s64)); h->stream_dev = a;
            } else {
                float* a = nullptr;
                ANNF_CUDA_H(upload(&a, s32)); h->stream_dev = a;
            }
            ANNF_CUDA_H(upload(&h->stream_statics, statics));
            h->stream_static_bytes = (int) statics.size();
            h->streamed = true;
            h->batched_path = 2;
        }
    }
    if (!h->streamed && !h->dmma && ((prec == ANNF_PREC_FP64 && h->tile_fp64 == 0) || (prec == ANNF_PREC_FP32 && h->tile_fp32 == 0))) {
        release(h);
        return fail(ANNF_ERR_UNSUPPORTED, "network too large for the fused batched kernel's shared memory in this precision");
    }
    ANNF_CUDA_H(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    ANNF_CUDA_H(cudaStreamCreateWithFlags(&h->host_stream2, cudaStreamNonBlocking));
    ANNF_CUDA_H(cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking));
    ANNF_CUDA_H(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++) {
        ANNF_CUDA_H(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        ANNF_CUDA_H(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
    }
    ANNF_CUDA_H(cudaDeviceSynchronize());
#undef ANNF_CUDA_H
    *out = h;
    return ANNF_OK;
}

annf_status annf_destroy(annf_handle h) {
    if (h == nullptr) return ANNF_OK;
    release(h);
    return ANNF_OK;
}

annf_status annf_get_info(annf_handle h, annf_info* info) {
    if (!h || !info) return fail(ANNF_ERR_INVALID, "annf_get_info: NULL argument");
    info->precision_single = h->prec_single;
    info->precision_batched = h->prec_batched;
    info->batched_path = h->batched_path;
    info->cluster_size = h->cluster_size;
    info->macs_per_eval = h->macs;
    info->kernel_launches = h->launches;
    info->sm_count = h->sm_count;
    info->single_path = h->fast_single ? 1 : 0;
    return ANNF_OK;
}

annf_status annf_update_parameters(annf_handle h, const double* potential_center, int32_t num_center,
                                   double force_constant, void* stream) {
    if (!h || !potential_center) return fail(ANNF_ERR_INVALID, "annf_update_parameters: NULL argument");
    if (num_center != h->net.num_center)
        return fail(ANNF_ERR_INVALID, "updateParametersInContext: potential_center has %d values, the context was built with %d",
                    num_center, h->net.num_center);
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    const int slot = h->param_next;
    h->param_next = (slot + 1) % annf_context::kParamSlots;
    ANNF_CUDA(cudaEventSynchronize(h->param_ev[slot]));          // the update that last used this slot has left the host buffer
    double* host = h->staged_host + (size_t) slot * (num_center + 1);
    double* dev = h->staged + (size_t) slot * (num_center + 1);
    memcpy(host, potential_center, sizeof(double) * num_center);
    host[num_center] = force_constant;
    ANNF_CUDA(cudaMemcpyAsync(dev, host, sizeof(double) * (num_center + 1), cudaMemcpyHostToDevice, s));
    set_parameters_kernel<<<(num_center + 1 + 127) / 128, 128, 0, s>>>(h->center, dev, num_center, h->k_dev);
    ANNF_CUDA(cudaGetLastError());
    ANNF_CUDA(cudaEventRecord(h->param_ev[slot], s));
    h->launches++;
    return ANNF_OK;
}

annf_status annf_reserve(annf_handle h, int32_t max_replicas, void* stream) {
    if (!h) return fail(ANNF_ERR_INVALID, "annf_reserve: NULL handle");
    if (max_replicas < 0) return fail(ANNF_ERR_INVALID, "annf_reserve: max_replicas = %d", max_replicas);
    DeviceGuard guard(h->device);
    if (h->tensor != nullptr && max_replicas > 0)
        ANNF_CUDA(tensor_plan_reserve(h->tensor, h->net, max_replicas, (cudaStream_t) stream));
    return ANNF_OK;
}

annf_status annf_set_atom_order(annf_handle h, const int32_t* atom_index_dev, int32_t num_atoms, void* stream) {
    if (!h || !atom_index_dev) return fail(ANNF_ERR_INVALID, "annf_set_atom_order: NULL argument");
    if (num_atoms != h->num_system_atoms)
        return fail(ANNF_ERR_INVALID, "annf_set_atom_order: %d atoms, the handle was created for %d", num_atoms, h->num_system_atoms);
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    invert_order_kernel<<<(num_atoms + 255) / 256, 256, 0, s>>>(atom_index_dev, num_atoms, h->inverse);
    remap_kernel<<<(h->net.n_sel + 255) / 256, 256, 0, s>>>(h->sel_atoms, h->net.n_sel, h->inverse, h->sel_slots);
    ANNF_CUDA(cudaGetLastError());
    h->order_dirty = true;
    h->launches += 2;
    return ANNF_OK;
}

annf_status annf_eval_openmm(annf_handle h, const void* posq, const void* posq_correction,
                             unsigned long long* force_buffer, int32_t padded_num_atoms, void* energy_buffer,
                             uint32_t flags, void* stream) {
    if (!h || !posq) return fail(ANNF_ERR_INVALID, "annf_eval_openmm: NULL argument");
    if (!force_buffer && !(flags & ANNF_FLAG_SKIP_FORCES)) return fail(ANNF_ERR_INVALID, "annf_eval_openmm: force_buffer is NULL");
    if (padded_num_atoms < h->num_system_atoms)
        return fail(ANNF_ERR_INVALID, "annf_eval_openmm: padded_num_atoms %d < %d atoms", padded_num_atoms, h->num_system_atoms);
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    cudaError_t err;
    if (h->fast_single) {
        h->prof.begin("annf_single_fast_kernel", s);
        err = launch_single_fast(h->net, h->fast_plan, h->fast_blobs, posq, posq_correction, force_buffer, padded_num_atoms, energy_buffer, flags, s,
                                 !h->order_dirty);
        h->order_dirty = false;
    } else {
        h->prof.begin("annf_single_kernel", s);
        err = launch_single(h->net, h->cluster_size, posq, posq_correction, force_buffer, padded_num_atoms, energy_buffer, flags, s);
    }
    h->prof.end(s);
    if (err != cudaSuccess) return fail(ANNF_ERR_CUDA, "single-replica launch failed: %s", cudaGetErrorString(err));
    h->launches++;
    return ANNF_OK;
}

annf_status annf_eval_batched(annf_handle h, int32_t R, const float* coords, int32_t atoms_per_replica,
                              const float* centers, float* forces, double* energies, void* stream) {
    if (!h) return fail(ANNF_ERR_INVALID, "annf_eval_batched: NULL handle");
    if (R == 0) return ANNF_OK;                                    // an empty batch is a no-op (its buffers may be NULL)
    if (!coords || !forces || !energies) return fail(ANNF_ERR_INVALID, "annf_eval_batched: NULL argument");
    if (R < 0) return fail(ANNF_ERR_INVALID, "annf_eval_batched: num_replicas = %d", R);
    if (atoms_per_replica < h->num_system_atoms)
        return fail(ANNF_ERR_INVALID, "annf_eval_batched: atoms_per_replica %d < %d system atoms", atoms_per_replica, h->num_system_atoms);
    if (R == 0) return ANNF_OK;
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    cudaError_t err;
    if (h->batched_path == 1) {
        err = tensor_plan_run(h->tensor, &h->prof, h->net, R, coords, atoms_per_replica, centers, forces, energies, s, &h->launches,
                              h->tensor_slot);
    } else if (h->dmma) {
        h->prof.begin("annf_dmma_kernel<f64>", s);
        err = launch_dmma(h->net, h->stream_plan, h->stream_statics, h->stream_static_bytes, (const double*) h->stream_dev, R, coords,
                          atoms_per_replica, centers, forces, energies, s);
        h->prof.end(s);
        h->launches++;
    } else if (h->streamed) {
        const bool fp64 = h->prec_batched == ANNF_PREC_FP64;
        h->prof.begin(fp64 ? "annf_stream_kernel<f64>" : "annf_stream_kernel<f32>", s);
        err = launch_stream(h->net, h->stream_plan, h->stream_statics, h->stream_static_bytes, h->stream_dev, fp64,
                            stream_pick_tile(h->net, h->stream_plan, fp64, R, h->sm_count), R, coords, atoms_per_replica, centers,
                            forces, energies, s);
        h->prof.end(s);
        h->launches++;
    } else {
        const bool fp64 = h->prec_batched == ANNF_PREC_FP64;
        h->prof.begin(fp64 ? "annf_batched_kernel<f64>" : "annf_batched_kernel<f32>", s);
        err = launch_batched(h->net, h->x_off, h->extra_nodes, fp64 ? h->tile_fp64 : h->tile_fp32, fp64, R, coords,
                             atoms_per_replica, centers, forces, energies, s);
        h->prof.end(s);
        h->launches++;
    }
    if (err != cudaSuccess) return fail(ANNF_ERR_CUDA, "batched launch failed: %s", cudaGetErrorString(err));
    return ANNF_OK;
}

annf_status annf_eval_batched_host(annf_handle h, int32_t R, const float* coords_host, int32_t atoms_per_replica,
                                   const float* centers_host, float* forces_host, double* energies_host) {
    if (!h) return fail(ANNF_ERR_INVALID, "annf_eval_batched_host: NULL handle");
    if (R == 0) return ANNF_OK;
    if (!coords_host || !forces_host || !energies_host) return fail(ANNF_ERR_INVALID, "annf_eval_batched_host: NULL argument");
    if (R < 0) return fail(ANNF_ERR_INVALID, "annf_eval_batched_host: num_replicas = %d", R);
    DeviceGuard guard(h->device);
    const size_t n_coords = (size_t) R * atoms_per_replica * 3, n_centers = (size_t) R * h->net.num_center;
    if (n_coords > h->h_cap_coords) {
        cudaFree(h->h_coords); cudaFree(h->h_forces);
        h->h_coords = h->h_forces = nullptr; h->h_cap_coords = 0;
        ANNF_CUDA(cudaMalloc((void**) &h->h_coords, sizeof(float) * n_coords));
        ANNF_CUDA(cudaMalloc((void**) &h->h_forces, sizeof(float) * n_coords));
        h->h_cap_coords = n_coords;
    }
    if (centers_host && n_centers > h->h_cap_centers) {
        cudaFree(h->h_centers); h->h_centers = nullptr; h->h_cap_centers = 0;
        ANNF_CUDA(cudaMalloc((void**) &h->h_centers, sizeof(float) * n_centers));
        h->h_cap_centers = n_centers;
    }
    if ((size_t) R > h->h_cap_energies) {
        cudaFree(h->h_energies); h->h_energies = nullptr; h->h_cap_energies = 0;
        ANNF_CUDA(cudaMalloc((void**) &h->h_energies, sizeof(double) * R));
        h->h_cap_energies = R;
    }
    // Software pipeline over replica chunks: the copy-in of chunk c+1 and the copy-out of chunk c-1 overlap
    // the evaluation of chunk c (PCIe is full duplex; streams ordered by events).  Consecutive chunks evaluate on
    // two compute streams with separate activation buffers, so their latency-bound kernels overlap as well.
    int chunks = 1;
    {
        const size_t bytes = sizeof(float) * n_coords;
        if (bytes >= (4u << 20)) chunks = 3;
        else if (bytes >= (1u << 20)) chunks = 2;
        if (const char* env = getenv("ANNF_HOST_CHUNKS")) chunks = atoi(env);
        chunks = std::max(1, std::min(std::min(chunks, 8), R));
    }
    const size_t row = (size_t) atoms_per_replica * 3;
    const bool zero_fill = h->net.n_sel != atoms_per_replica;   // kernels write the selected atoms only
    if (chunks == 1) {
        // small batches: nothing to overlap, so one stream and one wait (the pipelined path costs ~6 stream / event calls more)
        cudaStream_t s = h->host_stream;
        ANNF_CUDA(cudaMemcpyAsync(h->h_coords, coords_host, sizeof(float) * n_coords, cudaMemcpyHostToDevice, s));
        if (centers_host) ANNF_CUDA(cudaMemcpyAsync(h->h_centers, centers_host, sizeof(float) * n_centers, cudaMemcpyHostToDevice, s));
        if (zero_fill) ANNF_CUDA(cudaMemsetAsync(h->h_forces, 0, sizeof(float) * n_coords, s));
        annf_status st = annf_eval_batched(h, R, h->h_coords, atoms_per_replica, centers_host ? h->h_centers : nullptr, h->h_forces,
                                           h->h_energies, s);
        if (st != ANNF_OK) { cudaStreamSynchronize(s); return st; }
        ANNF_CUDA(cudaMemcpyAsync(forces_host, h->h_forces, sizeof(float) * n_coords, cudaMemcpyDeviceToHost, s));
        ANNF_CUDA(cudaMemcpyAsync(energies_host, h->h_energies, sizeof(double) * R, cudaMemcpyDeviceToHost, s));
        ANNF_CUDA(cudaStreamSynchronize(s));
        return ANNF_OK;
    }
    // Chunk sizes DEcrease geometrically (each chunk `ratio` times smaller than the one before; ANNF_HOST_CHUNK_RATIO, 1 = equal).
    // What the schedule balances, measured with ANNF_HOST_TRACE=1 at BASELINE C5 (profiles/r02_host_trace.md): the link is NOT
    // full duplex in practice — a copy-in moves 49 GB/s alone and 28 GB/s while a copy-out is running — so the 36.9 MB of a step
    // cross at ~75 GB/s in aggregate once both directions are busy, and the schedule only decides how soon that state is
    // reached (small first chunk) and how short the tail after the last byte lands is (small last chunk), against ~35 us of
    // host submission per chunk.  3 chunks at ratio 1.5 measured best (0.647 ms; 4 equal chunks 0.75, 4 at ratio 2 0.69).
    int bounds[9];
    {
        double ratio = 1.5;
        if (const char* env = getenv("ANNF_HOST_CHUNK_RATIO")) ratio = std::max(1.0, atof(env));
        double weights[8], total_w = 0.0;
        for (int c = 0; c < chunks; c++) { weights[c] = std::pow(ratio, chunks - 1 - c); total_w += weights[c]; }
        double acc = 0.0;
        bounds[0] = 0;
        for (int c = 0; c < chunks; c++) {
            acc += weights[c];
            bounds[c + 1] = (int) std::llround((double) R * acc / total_w);
        }
        bounds[chunks] = R;
        for (int c = 0; c < chunks; c++) if (bounds[c + 1] <= bounds[c]) bounds[c + 1] = std::min(R, bounds[c] + 1);
    }
    // ANNF_HOST_TRACE=1 (diagnostic): device timeline of this call — when each chunk's copy-in, evaluation and copy-out
    // finished, in us since the first copy was enqueued — printed to stderr as one JSON line, plus host submission times.
    static const bool trace = getenv("ANNF_HOST_TRACE") != nullptr;
    cudaEvent_t tr0 = nullptr, tr_in[8] = {}, tr_comp[8] = {}, tr_out[8] = {};
    double host_us[8] = {};
    const auto host_t0 = std::chrono::steady_clock::now();
    if (trace) {
        cudaEventCreate(&tr0);
        for (int c = 0; c < chunks; c++) { cudaEventCreate(&tr_in[c]); cudaEventCreate(&tr_comp[c]); cudaEventCreate(&tr_out[c]); }
        cudaEventRecord(tr0, h->in_stream);
    }
    for (int c = 0; c < chunks; c++) {
        const int r0 = bounds[c], r1 = bounds[c + 1], n = r1 - r0;
        if (n <= 0) continue;
        ANNF_CUDA(cudaMemcpyAsync(h->h_coords + r0 * row, coords_host + r0 * row, sizeof(float) * n * row, cudaMemcpyHostToDevice, h->in_stream));
        if (centers_host)
            ANNF_CUDA(cudaMemcpyAsync(h->h_centers + (size_t) r0 * h->net.num_center, centers_host + (size_t) r0 * h->net.num_center,
                                      sizeof(float) * n * h->net.num_center, cudaMemcpyHostToDevice, h->in_stream));
        cudaStream_t compute = (c & 1) ? h->host_stream2 : h->host_stream;
        h->tensor_slot = c & 1;
        ANNF_CUDA(cudaEventRecord(h->ev_in[c], h->in_stream));
        if (trace) cudaEventRecord(tr_in[c], h->in_stream);
        ANNF_CUDA(cudaStreamWaitEvent(compute, h->ev_in[c], 0));
        if (zero_fill) ANNF_CUDA(cudaMemsetAsync(h->h_forces + r0 * row, 0, sizeof(float) * n * row, compute));
        annf_status st = annf_eval_batched(h, n, h->h_coords + r0 * row, atoms_per_replica,
                                           centers_host ? h->h_centers + (size_t) r0 * h->net.num_center : nullptr,
                                           h->h_forces + r0 * row, h->h_energies + r0, compute);
        h->tensor_slot = 0;
        if (st != ANNF_OK) {
            cudaStreamSynchronize(h->in_stream); cudaStreamSynchronize(h->host_stream); cudaStreamSynchronize(h->host_stream2);
            cudaStreamSynchronize(h->out_stream);
            return st;
        }
        ANNF_CUDA(cudaEventRecord(h->ev_comp[c], compute));
        if (trace) cudaEventRecord(tr_comp[c], compute);
        ANNF_CUDA(cudaStreamWaitEvent(h->out_stream, h->ev_comp[c], 0));
        ANNF_CUDA(cudaMemcpyAsync(forces_host + r0 * row, h->h_forces + r0 * row, sizeof(float) * n * row, cudaMemcpyDeviceToHost, h->out_stream));
        ANNF_CUDA(cudaMemcpyAsync(energies_host + r0, h->h_energies + r0, sizeof(double) * n, cudaMemcpyDeviceToHost, h->out_stream));
        if (trace) {
            cudaEventRecord(tr_out[c], h->out_stream);
            host_us[c] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
        }
    }
    ANNF_CUDA(cudaStreamSynchronize(h->out_stream));
    ANNF_CUDA(cudaStreamSynchronize(h->host_stream));
    ANNF_CUDA(cudaStreamSynchronize(h->host_stream2));
    ANNF_CUDA(cudaStreamSynchronize(h->in_stream));
    if (trace) {
        const double host_done = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
        fprintf(stderr, "{\"annf_host_trace\": {\"replicas\": %d, \"host_returned_us\": %.1f, \"chunks\": [", R, host_done);
        for (int c = 0; c < chunks; c++) {
            float a = 0.f, b = 0.f, d = 0.f;
            cudaEventElapsedTime(&a, tr0, tr_in[c]); cudaEventElapsedTime(&b, tr0, tr_comp[c]); cudaEventElapsedTime(&d, tr0, tr_out[c]);
            fprintf(stderr, "%s{\"rows\": %d, \"host_submitted_us\": %.1f, \"copied_in_us\": %.1f, \"evaluated_us\": %.1f, \"copied_out_us\": %.1f}",
                    c ? ", " : "", bounds[c + 1] - bounds[c], host_us[c], 1e3 * a, 1e3 * b, 1e3 * d);
            cudaEventDestroy(tr_in[c]); cudaEventDestroy(tr_comp[c]); cudaEventDestroy(tr_out[c]);
        }
        fprintf(stderr, "]}}\n");
        cudaEventDestroy(tr0);
    }
    return ANNF_OK;
}

#ifndef ANNF_PHASE_CLOCKS
#define ANNF_NEED_PHASE_CLOCKS(name) return fail(ANNF_ERR_UNSUPPORTED, name ": this build carries no phase clocks (make EXTRA=-DANNF_PHASE_CLOCKS)")
#else
#define ANNF_NEED_PHASE_CLOCKS(name) do { } while (0)
#endif

annf_status annf_debug_single_phases(annf_handle h, int64_t* clocks16) {
    if (!h || !clocks16) return fail(ANNF_ERR_INVALID, "annf_debug_single_phases: NULL argument");
    ANNF_NEED_PHASE_CLOCKS("annf_debug_single_phases");
    DeviceGuard guard(h->device);
    ANNF_CUDA(cudaDeviceSynchronize());
    ANNF_CUDA(read_single_phase_clocks(reinterpret_cast<long long*>(clocks16)));
    return ANNF_OK;
}

annf_status annf_debug_stream_phases(annf_handle h, int64_t* clocks32) {
    if (!h || !clocks32) return fail(ANNF_ERR_INVALID, "annf_debug_stream_phases: NULL argument");
    ANNF_NEED_PHASE_CLOCKS("annf_debug_stream_phases");
    DeviceGuard guard(h->device);
    ANNF_CUDA(cudaDeviceSynchronize());
    if (h->dmma) ANNF_CUDA(read_dmma_phase_clocks(reinterpret_cast<long long*>(clocks32)));
    else ANNF_CUDA(read_stream_phase_clocks(reinterpret_cast<long long*>(clocks32)));
    return ANNF_OK;
}

annf_status annf_debug_tensor_phases(annf_handle h, int64_t* clocks192, int32_t reset) {
    if (!h || (!clocks192 && !reset)) return fail(ANNF_ERR_INVALID, "annf_debug_tensor_phases: NULL argument");
    ANNF_NEED_PHASE_CLOCKS("annf_debug_tensor_phases");
    DeviceGuard guard(h->device);
    ANNF_CUDA(cudaDeviceSynchronize());
    ANNF_CUDA(tensor_phase_clocks(reinterpret_cast<unsigned long long*>(clocks192), reset != 0));
    return ANNF_OK;
}

annf_status annf_set_profiling(annf_handle h, int32_t enabled) {
    if (!h) return fail(ANNF_ERR_INVALID, "annf_set_profiling: NULL handle");
    DeviceGuard guard(h->device);
    if (!enabled) h->prof.drain();
    h->prof.enabled = enabled != 0;
    return ANNF_OK;
}

annf_status annf_get_kernel_times(annf_handle h, annf_kernel_time* out, int32_t capacity, int32_t* count, int32_t reset) {
    if (!h || !count) return fail(ANNF_ERR_INVALID, "annf_get_kernel_times: NULL argument");
    DeviceGuard guard(h->device);
    h->prof.drain();
    int n = 0;
    for (auto& kv : h->prof.totals) {
        if (out && n < capacity) {
            memset(&out[n], 0, sizeof(annf_kernel_time));
            strncpy(out[n].name, h->prof.names[kv.first].c_str(), sizeof(out[n].name) - 1);
            out[n].launches = kv.second.first;
            out[n].total_ms = kv.second.second;
        }
        n++;
    }
    *count = n;
    if (reset) h->prof.totals.clear();
    return ANNF_OK;
}

}  // extern "C"
