// Micro-benchmark (diagnostic, not product): per-chunk overhead of a producer / consumer ring built from cp.async.bulk
// and full / empty mbarriers when the consumers do NO work: W consumer warps wait for "full", one lane arrives on "empty"
// (count W); one producer thread waits for "empty" and re-issues.  mode 0: every lane polls; mode 1: lane 0 polls, then
// __syncwarp; mode 2: warp 0 polls, the others follow through a named barrier.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_par(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int S = 4;

__global__ void ring(const unsigned char* src, int bytes, int copies, int W, int mode, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[S], empty[S];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&empty[s])), "r"(W) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == W) {
        if (lane == 0) {
            size_t off = 0;
            for (int c = 0; c < copies; c++) {
                const int s = c % S;
                if (c >= S) wait_par(s_u32(&empty[s]), ((c / S) - 1) & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&full[s])), "r"((uint32_t) bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s_u32(smem + (size_t) s * bytes)), "l"(__cvta_generic_to_global(src + off)), "r"((uint32_t) bytes), "r"(s_u32(&full[s])) : "memory");
                off += bytes;
                if (off + bytes > 524288) off = 0;
            }
        }
        return;
    }
    for (int c = 0; c < copies; c++) {
        const int s = c % S;
        const uint32_t parity = (c / S) & 1;
        if (mode == 0) {
            wait_par(s_u32(&full[s]), parity);
        } else if (mode == 1) {
            if (lane == 0) wait_par(s_u32(&full[s]), parity);
            __syncwarp();
        } else {
            if (warp == 0) wait_par(s_u32(&full[s]), parity);
            asm volatile("bar.sync 1, %0;" ::"r"(W * 32) : "memory");
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&empty[s])) : "memory");
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    unsigned char* src; cudaMalloc(&src, 1 << 20); cudaMemset(src, 1, 1 << 20);
    long long* out; cudaMalloc(&out, sizeof(long long) * 256);
    cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int copies = 512;
    printf("%8s %5s %5s %12s\n", "bytes", "W", "mode", "cyc/chunk");
    for (int bytes : {4096, 16384})
        for (int W : {1, 4, 8, 16})
            for (int mode : {0, 1, 2}) {
                for (int rep = 0; rep < 2; rep++) ring<<<128, (W + 1) * 32, (size_t) S * bytes>>>(src, bytes, copies, W, mode, out);
                cudaDeviceSynchronize();
                long long h[128];
                cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < 128; i++) mx = h[i] > mx ? h[i] : mx;
                printf("%8d %5d %5d %12.1f\n", bytes, W, mode, (double) mx / copies);
            }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
