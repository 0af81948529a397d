// BENCH INFRASTRUCTURE, clearly labelled: a standalone RESTATEMENT of the kernel body the legacy CUDA plugin
// string-generates and hands to OpenMM's CudaBondedUtilities (/root/reference/platforms/cuda/src/CudaANN_Kernels.cpp:183-338),
// for the only network class it supports: 3 node layers, Tanh / Linear, Cartesian input (:84-86, :257-260, :270-289).
// The legacy plugin itself cannot be built here (OpenMM's CUDA platform is absent); this program
// exists so that the north-star target "per-step bias cost vs the legacy CUDA plugin at 1 replica" has a measured,
// honestly-labelled stand-in.  It reproduces the legacy execution scheme, not its text:
//   * 60 "bonds" = 60 threads of ONE block, loop stride 60 (:97, :265), __syncthreads() between stages (:261-306);
//   * activations live in GLOBAL-memory scratch arrays shared by the 60 threads (:133-137, :167-169), fp32 everywhere;
//   * every thread scales all inputs (:189-193), threads 0..2 centre one component each (:194-205);
//   * thread 0 alone forms the forces (:326-337); every thread then adds its (mostly zero) forces and energy to the
//     context buffers, as OpenMM's bonded kernel does for each "bond" [OpenMM API knowledge].
// It is favourable to the legacy scheme: no other bonded terms share the launch, and one 128-thread block is launched.
// `pairs` selects the legacy plugin's other input mode, pair distances (:210-256 features, :339-351 forces): the same 21-40-2
// network fed with the 21 distances between the 7 atoms; every thread computes every distance, thread 0 forms the forces.
// usage: legacy_restatement [steps] [pairs]   -> one JSON object
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kThreads = 60;          // num_of_parallel_threads (:97)
constexpr int kAtoms = 7, kN0 = 21, kN1 = 40, kN2 = 2;

struct LegacyArgs {
    const float4* posq;
    unsigned long long* force;
    int padded;
    float* energy;
    const float *coeff0, *coeff1, *bias0, *bias1, *center;
    float *in0, *in1, *in2;            // global scratch "INPUT_0..2"
    float k, scale;
    int atom[kAtoms];
    int pairs;                         // 0: Cartesian input, 1: pair distances
};
constexpr int kPairs = kAtoms * (kAtoms - 1) / 2;          // 21 = kN0

__global__ void legacy_style_kernel(LegacyArgs a) {
    const int index = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bond = index < kThreads;
    float energy = 0.f;
    float3 f[kAtoms];
    for (int i = 0; i < kAtoms; i++) f[i] = make_float3(0.f, 0.f, 0.f);
    float3 delta[kPairs];
    float dis[kPairs];
    if (bond && a.pairs) {
        int t = 0;
        for (int i = 0; i < kAtoms; i++)                         // every thread computes every distance (:239-256)
            for (int j = i + 1; j < kAtoms; j++, t++) {
                const float4 p = a.posq[a.atom[i]], q = a.posq[a.atom[j]];
                delta[t] = make_float3(p.x - q.x, p.y - q.y, p.z - q.z);
                dis[t] = sqrtf(delta[t].x * delta[t].x + delta[t].y * delta[t].y + delta[t].z * delta[t].z) / a.scale;
                a.in0[t] = dis[t];
            }
    } else if (bond) {
        for (int i = 0; i < kAtoms; i++) {                       // every thread writes every input (:189-193)
            const float4 p = a.posq[a.atom[i]];
            a.in0[3 * i] = p.x / a.scale; a.in0[3 * i + 1] = p.y / a.scale; a.in0[3 * i + 2] = p.z / a.scale;
        }
        if (index < 3) {                                         // one component per thread (:196-205)
            float com = 0.f;
            for (int i = index; i < kN0; i += 3) com += a.in0[i];
            com /= kAtoms;
            for (int i = index; i < kN0; i += 3) a.in0[i] -= com;
        }
    }
    __syncthreads();
    if (bond)
        for (int j = index; j < kN1; j += kThreads) {            // forward layer 1 (:265-276)
            float t = a.bias0[j];
            for (int i = 0; i < kN0; i++) t += a.coeff0[j * kN0 + i] * a.in0[i];
            a.in1[j] = tanhf(t);
        }
    __syncthreads();
    if (bond)
        for (int j = index; j < kN2; j += kThreads) {            // forward layer 2
            float t = a.bias1[j];
            for (int i = 0; i < kN1; i++) t += a.coeff1[j * kN1 + i] * a.in1[i];
            a.in2[j] = tanhf(t);
        }
    __syncthreads();
    if (bond)
        for (int j = index; j < kN2; j += kThreads) {            // energy + output derivative (:281-290)
            const float t = a.in2[j];
            energy += 0.5f * (t - a.center[j]) * (t - a.center[j]) * a.k;
            a.in2[j] = (t - a.center[j]) * a.k * (1.f - t * t);
        }
    __syncthreads();
    if (bond)
        for (int i = index; i < kN1; i += kThreads) {            // backward through connection 1 (:292-305)
            float t = 0.f;
            for (int j = 0; j < kN2; j++) t += a.coeff1[i + j * kN1] * a.in2[j];
            const float act = a.in1[i];
            a.in1[i] = t * (1.f - act * act);
        }
    __syncthreads();
    if (bond)
        for (int i = index; i < kN0; i += kThreads) {            // backward through connection 0
            float t = 0.f;
            for (int j = 0; j < kN1; j++) t += a.coeff0[i + j * kN0] * a.in1[j];
            a.in0[i] = t;
        }
    __syncthreads();
    if (bond && a.pairs) {
        if (index == 0) {                                        // thread 0 alone forms the forces (:343-350)
            int t = 0;
            for (int i = 0; i < kAtoms; i++)
                for (int j = i + 1; j < kAtoms; j++, t++) {
                    const float c = a.in0[t] / (a.scale * a.scale * dis[t]);
                    f[i].x -= delta[t].x * c; f[i].y -= delta[t].y * c; f[i].z -= delta[t].z * c;
                    f[j].x += delta[t].x * c; f[j].y += delta[t].y * c; f[j].z += delta[t].z * c;
                }
        }
        for (int i = 0; i < kAtoms; i++) {
            atomicAdd(&a.force[a.atom[i]], (unsigned long long) ((long long) (f[i].x * 0x100000000)));
            atomicAdd(&a.force[a.atom[i] + a.padded], (unsigned long long) ((long long) (f[i].y * 0x100000000)));
            atomicAdd(&a.force[a.atom[i] + 2 * a.padded], (unsigned long long) ((long long) (f[i].z * 0x100000000)));
        }
        a.energy[index] += energy;
    } else if (bond) {
        float3 avg = make_float3(0.f, 0.f, 0.f);                 // every thread averages (:313-320) ...
        for (int i = 0; i < kAtoms; i++) { avg.x += a.in0[3 * i]; avg.y += a.in0[3 * i + 1]; avg.z += a.in0[3 * i + 2]; }
        avg.x /= kAtoms; avg.y /= kAtoms; avg.z /= kAtoms;
        if (index == 0)                                          // ... thread 0 alone forms the forces (:326-332)
            for (int i = 0; i < kAtoms; i++)
                f[i] = make_float3(-(a.in0[3 * i] - avg.x) / a.scale, -(a.in0[3 * i + 1] - avg.y) / a.scale, -(a.in0[3 * i + 2] - avg.z) / a.scale);
        for (int i = 0; i < kAtoms; i++) {                       // every "bond" adds its forces to the context buffers
            atomicAdd(&a.force[a.atom[i]], (unsigned long long) ((long long) (f[i].x * 0x100000000)));
            atomicAdd(&a.force[a.atom[i] + a.padded], (unsigned long long) ((long long) (f[i].y * 0x100000000)));
            atomicAdd(&a.force[a.atom[i] + 2 * a.padded], (unsigned long long) ((long long) (f[i].z * 0x100000000)));
        }
        a.energy[index] += energy;
    }
}

// launch floor: what one dependent kernel launch per MD step costs when the kernel does nothing
__global__ void empty_kernel(int* sink) { if (sink != nullptr && threadIdx.x == 1024) *sink = 0; }

static double lcg(unsigned long long& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return ((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0; }

template <typename T> static T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    cudaMalloc((void**) &d, sizeof(T) * v.size());
    cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
    return d;
}

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 20000;
    const bool pairs = argc > 2 && argv[2][0] == 'p';
    unsigned long long seed = 1234;
    std::vector<float> c0(kN1 * kN0), c1(kN2 * kN1), b0(kN1), b1(kN2), cen = {0.3f, 0.3f};
    const double bound0 = sqrt(6.0 / (kN0 + kN1)), bound1 = sqrt(6.0 / (kN1 + kN2));
    for (auto& v : c0) v = (float) (bound0 * lcg(seed));
    for (auto& v : b0) v = (float) (0.1 * lcg(seed));
    for (auto& v : c1) v = (float) (bound1 * lcg(seed));
    for (auto& v : b1) v = (float) (0.1 * lcg(seed));
    const int num_atoms = 22, padded = 32;
    std::vector<float4> posq(padded);
    for (int i = 0; i < num_atoms; i++) posq[i] = make_float4((float) (0.5 * lcg(seed)), (float) (0.5 * lcg(seed)), (float) (0.5 * lcg(seed)), 0.f);
    LegacyArgs a;
    a.posq = upload(posq);
    cudaMalloc((void**) &a.force, sizeof(unsigned long long) * 3 * padded);
    cudaMemset(a.force, 0, sizeof(unsigned long long) * 3 * padded);
    cudaMalloc((void**) &a.energy, sizeof(float) * 128);
    cudaMemset(a.energy, 0, sizeof(float) * 128);
    a.padded = padded;
    a.coeff0 = upload(c0); a.coeff1 = upload(c1); a.bias0 = upload(b0); a.bias1 = upload(b1); a.center = upload(cen);
    cudaMalloc((void**) &a.in0, sizeof(float) * kN0); cudaMalloc((void**) &a.in1, sizeof(float) * kN1); cudaMalloc((void**) &a.in2, sizeof(float) * kN2);
    a.k = 500.f; a.scale = 0.5f;
    a.pairs = pairs ? 1 : 0;
    const int sel[kAtoms] = {1, 4, 6, 8, 14, 16, 18};
    for (int i = 0; i < kAtoms; i++) a.atom[i] = sel[i];
    cudaStream_t stream;
    cudaStreamCreate(&stream);
    for (int i = 0; i < 200; i++) legacy_style_kernel<<<1, 128, 0, stream>>>(a);
    cudaStreamSynchronize(stream);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
    for (int i = 0; i < steps; i++) legacy_style_kernel<<<1, 128, 0, stream>>>(a);
    cudaEventRecord(e1, stream);
    cudaStreamSynchronize(stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    // graph replay of 100 launches, like bench.py does for the launch-bound workloads
    cudaGraph_t graph; cudaGraphExec_t exec;
    cudaStreamBeginCapture(stream, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 100; i++) legacy_style_kernel<<<1, 128, 0, stream>>>(a);
    cudaStreamEndCapture(stream, &graph);
    cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphLaunch(exec, stream);
    cudaStreamSynchronize(stream);
    cudaEventRecord(e0, stream);
    for (int i = 0; i < steps / 100; i++) cudaGraphLaunch(exec, stream);
    cudaEventRecord(e1, stream);
    cudaStreamSynchronize(stream);
    float ms_graph = 0.f;
    cudaEventElapsedTime(&ms_graph, e0, e1);
    // the same two measurements for an empty kernel
    for (int i = 0; i < 200; i++) empty_kernel<<<1, 128, 0, stream>>>(nullptr);
    cudaStreamSynchronize(stream);
    cudaEventRecord(e0, stream);
    for (int i = 0; i < steps; i++) empty_kernel<<<1, 128, 0, stream>>>(nullptr);
    cudaEventRecord(e1, stream);
    cudaStreamSynchronize(stream);
    float ms_empty = 0.f;
    cudaEventElapsedTime(&ms_empty, e0, e1);
    cudaGraph_t graph2; cudaGraphExec_t exec2;
    cudaStreamBeginCapture(stream, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 100; i++) empty_kernel<<<1, 128, 0, stream>>>(nullptr);
    cudaStreamEndCapture(stream, &graph2);
    cudaGraphInstantiate(&exec2, graph2, 0);
    cudaGraphLaunch(exec2, stream);
    cudaStreamSynchronize(stream);
    cudaEventRecord(e0, stream);
    for (int i = 0; i < steps / 100; i++) cudaGraphLaunch(exec2, stream);
    cudaEventRecord(e1, stream);
    cudaStreamSynchronize(stream);
    float ms_empty_graph = 0.f;
    cudaEventElapsedTime(&ms_empty_graph, e0, e1);
    const cudaError_t err = cudaGetLastError();
    printf("{\"what\": \"restatement of the legacy plugin's generated kernel (NOT the legacy plugin)\", \"network\": \"21-40-2 Tanh/Tanh, 7 of 22 atoms, %s, fp32\", "
           "\"steps\": %d, \"us_per_step_stream\": %.4f, \"us_per_step_graph\": %.4f, \"empty_kernel_us_per_step_stream\": %.4f, "
           "\"empty_kernel_us_per_step_graph\": %.4f, \"cuda_error\": \"%s\"}\n",
           pairs ? "21 pair distances" : "Cartesian input", steps, 1e3 * ms / steps, 1e3 * ms_graph / (steps / 100 * 100), 1e3 * ms_empty / steps, 1e3 * ms_empty_graph / (steps / 100 * 100),
           cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : 1;
}
