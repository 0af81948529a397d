// Micro-benchmark (diagnostic, not product): how fast do the partial tiles of a split-K cluster cross between SMs?
// A cluster of S CTAs, every CTA holding a [S][slice] float buffer; every CTA needs slice r of every peer (a reduce-scatter).
//   pull   256 threads read the S-1 remote slices with ld.shared::cluster.v4 and sum them       (what cooperative_epilogue does)
//   push   one thread per CTA sends slice j to CTA j with cp.async.bulk.shared::cluster.shared::cta (S-1 bulk copies,
//          completion counted on the receiver's mbarrier), then 256 threads sum S local buffers
// All 148/S clusters run at once (the fabric is shared).  Prints cycles per exchange (median CTA) and bytes / clk / SM.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void wait_par(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <int S>
__global__ void __launch_bounds__(256) probe(int slice_floats, int mode, int reps, long long* cycles, float* sink) {
    extern __shared__ __align__(128) float smem[];            // own[S][slice] | recv[S][slice]
    __shared__ __align__(8) unsigned long long bar;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int) cluster.block_rank(), tid = threadIdx.x;
    float* own = smem;
    float* recv = smem + (size_t) S * slice_floats;
    for (int i = tid; i < S * slice_floats; i += 256) own[i] = (float) (rank + 1) * 0.001f * (float) (i & 1023);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cluster.sync();
    float acc = 0.f;
    long long total = 0;
    for (int it = 0; it < reps; it++) {
        const long long t0 = clock64();
        if (mode == 0) {
            for (int v = tid; v < slice_floats / 4; v += 256) {
                float4 sum = reinterpret_cast<const float4*>(own + (size_t) rank * slice_floats)[v];
#pragma unroll
                for (int d = 1; d < S; d++) {
                    const int peer = (rank + d) % S;
                    const uint32_t addr = map_to(s_u32(own + (size_t) rank * slice_floats + 4 * v), (uint32_t) peer);
                    float4 x;
                    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
                    sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
                }
                acc += sum.x + sum.y + sum.z + sum.w;
            }
        } else {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&bar)), "r"((uint32_t) ((S - 1) * slice_floats * 4)) : "memory");
                for (int d = 1; d < S; d++) {
                    const int peer = (rank + d) % S;
                    const uint32_t dst = map_to(s_u32(recv + (size_t) rank * slice_floats), (uint32_t) peer);
                    const uint32_t rbar = map_to(s_u32(&bar), (uint32_t) peer);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "r"(s_u32(own + (size_t) peer * slice_floats)), "r"((uint32_t) (slice_floats * 4)), "r"(rbar) : "memory");
                }
            }
            wait_par(s_u32(&bar), (uint32_t) it & 1u);
            for (int v = tid; v < slice_floats / 4; v += 256) {
                float4 sum = reinterpret_cast<const float4*>(own + (size_t) rank * slice_floats)[v];
#pragma unroll
                for (int d = 1; d < S; d++) {
                    const int peer = (rank + d) % S;
                    const float4 x = reinterpret_cast<const float4*>(recv + (size_t) peer * slice_floats)[v];
                    sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
                }
                acc += sum.x + sum.y + sum.z + sum.w;
            }
        }
        __syncthreads();
        total += clock64() - t0;
        cluster.sync();                                       // receive buffers free again
    }
    if (tid == 0) cycles[blockIdx.x] = total / reps;
    if (acc == 123.456f) sink[0] = acc;
}

template <int S>
static void run(int slice_floats, int mode) {
    const int ctas = (148 / S) * S;
    long long* cyc; float* sink;
    cudaMalloc(&cyc, sizeof(long long) * ctas); cudaMalloc(&sink, 4);
    const size_t smem = sizeof(float) * 2 * S * slice_floats;
    cudaFuncSetAttribute(probe<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(probe<S>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int reps = 50;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<S>, slice_floats, mode, reps, cyc, sink);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("S=%d slice=%d mode=%d: %s\n", S, slice_floats, mode, cudaGetErrorString(e)); cudaGetLastError(); return; }
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double bytes = (double) (S - 1) * slice_floats * 4;
    printf("S=%d slice=%6.1f KB  %s  cycles median %6lld max %6lld   remote bytes per CTA %6.1f KB = %5.1f B/clk/SM\n", S, slice_floats * 4 / 1024.0,
           mode == 0 ? "pull ld.shared::cluster" : "push cp.async.bulk     ", h[ctas / 2], h[ctas - 1], bytes / 1024.0, bytes / (double) h[ctas / 2]);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int mode = 0; mode < 2; mode++) {
        run<4>(4096, mode);      // 16 KB slices: a 128 x 128 fp32 tile split 4 ways
        run<4>(2048, mode);
        run<2>(8192, mode);      // 32 KB: split 2
        run<8>(2048, mode);      // 8 KB: split 8
    }
    return 0;
}
