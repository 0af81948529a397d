// Micro-benchmark (diagnostic, not product): issue rate and dependent-issue latency of DFMA and of the FP64 tensor
// instruction mma.sync.m8n8k4.f64 (DMMA) on one SM, as a function of independent accumulator chains per warp and warps per
// scheduler.  Prints cycles per instruction per scheduler and the implied FMA/clk/SM.
#include <cuda_runtime.h>
#include <cstdio>

template <int CH, bool MMA>
__global__ void probe(int iters, double seed, double* sink, long long* cycles) {
    double c[CH][2];
    for (int i = 0; i < CH; i++) { c[i][0] = seed * (i + 1); c[i][1] = seed * (i + 2); }
    const double a = seed * 1.0000001, b = seed * 0.9999999;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MMA) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            } else {
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[i][0]) : "d"(a), "d"(b));
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CH, bool MMA>
static void run(int warps, double* sink, long long* cyc) {
    const int iters = 2000;
    probe<CH, MMA><<<1, warps * 32>>>(iters, 1.0, sink, cyc);
    probe<CH, MMA><<<1, warps * 32>>>(iters, 1.0, sink, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_sched_instr = (double) iters * CH * (warps / 4.0 < 1 ? 1 : warps / 4.0);   // warp-instructions per scheduler
    const double cyc_per = (double) h / per_sched_instr;
    const double fma_per_instr = MMA ? 256.0 : 32.0;
    const int sched = warps < 4 ? warps : 4;
    printf("%5s chains %d warps %2d : %8.2f cycles / instr / scheduler  -> %7.1f FMA/clk/SM\n", MMA ? "DMMA" : "DFMA", CH, warps, cyc_per,
           fma_per_instr * sched / cyc_per);
}

int main() {
    double* sink; long long* cyc;
    cudaMalloc(&sink, sizeof(double) * 4096); cudaMalloc(&cyc, sizeof(long long) * 16);
    for (int warps : {1, 4, 8, 16}) {
        run<1, false>(warps, sink, cyc); run<2, false>(warps, sink, cyc); run<4, false>(warps, sink, cyc); run<8, false>(warps, sink, cyc); run<16, false>(warps, sink, cyc);
    }
    for (int warps : {1, 4, 8, 16}) {
        run<1, true>(warps, sink, cyc); run<2, true>(warps, sink, cyc); run<4, true>(warps, sink, cyc); run<8, true>(warps, sink, cyc);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
