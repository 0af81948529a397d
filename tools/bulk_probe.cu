// Micro-benchmark (diagnostic, not product): sustained rate of 1-D cp.async.bulk (global -> shared, mbarrier completion)
// per SM, as a function of copy size, copies in flight, CTAs in flight and whether all CTAs read the same addresses.
// One thread per CTA issues and waits; nothing consumes the data.  Prints cycles per copy and bytes/clk/SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void probe(const unsigned char* src, size_t span, int bytes, int stages, int copies, int same, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned char* base = src + (same ? 0 : ((size_t) blockIdx.x * 1048576) % span);
    size_t off = same ? 0 : 0;
    const long long t0 = clock64();
    for (int c = 0; c < copies + stages; c++) {
        const int s = c % stages;
        if (c >= stages) {
            const uint32_t parity = ((c / stages) - 1) & 1;
            asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s_u32(&bars[s])), "r"(parity) : "memory");
        }
        if (c < copies) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&bars[s])), "r"((uint32_t) bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s_u32(smem + (size_t) s * bytes)), "l"(__cvta_generic_to_global(base + off)), "r"((uint32_t) bytes), "r"(s_u32(&bars[s])) : "memory");
            off += bytes;
            if (off + bytes > 524288) off = 0;       // walk a 512 KB window (the size of a weight stream)
        }
    }
    out[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t span = 256u << 20;
    unsigned char* src; cudaMalloc(&src, span + (1 << 20)); cudaMemset(src, 1, span + (1 << 20));
    long long* out; cudaMalloc(&out, sizeof(long long) * 1024);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int copies = 256;
    printf("%8s %7s %6s %5s %12s %12s\n", "bytes", "stages", "ctas", "same", "cyc/copy", "B/clk/SM");
    for (int same = 1; same >= 0; same--)
        for (int ctas : {1, 128})
            for (int stages : {1, 4})
                for (int bytes : {2048, 4096, 8192, 16384, 32768}) {
                    if ((size_t) stages * bytes > 160 * 1024) continue;
                    for (int rep = 0; rep < 2; rep++) probe<<<ctas, 32, (size_t) stages * bytes>>>(src, span, bytes, stages, copies, same, out);
                    cudaDeviceSynchronize();
                    long long h[1024];
                    cudaMemcpy(h, out, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int i = 0; i < ctas; i++) mx = h[i] > mx ? h[i] : mx;
                    printf("%8d %7d %6d %5d %12.1f %12.2f\n", bytes, stages, ctas, same, (double) mx / copies, (double) bytes * copies / mx);
                }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
