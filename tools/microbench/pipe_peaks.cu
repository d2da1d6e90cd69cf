// pipe_peaks.cu — measures the instruction throughput of the pipes the matcher's roofline is quoted against
// ("POPC pipe peak ... to be confirmed by microbenchmark"): POPC (XU pipe), LOP3 (integer ALU) and IMAD
// (FMA pipe), in lane-operations per clock per SM, on the device it runs on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peaks tools/microbench/pipe_peaks.cu && ./pipe_peaks
// Every thread runs long chains of independent operations (8 accumulators) so that only issue / pipe throughput
// limits the loop; 148 x 8 CTAs of 256 threads keep every SM full.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, UNROLL = 8;

template <int OP>
__global__ void __launch_bounds__(256) pipe_kernel(unsigned *out, unsigned seed) {
    unsigned a[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) a[k] = seed + threadIdx.x * 2654435761u + k * 40503u;
    unsigned b = seed ^ 0x9e3779b9u, c = seed * 3u + 1u;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            if (OP == 0) a[k] = __popc(a[k]) + b;          // POPC + IADD: the add keeps the chain value-dependent
            if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < UNROLL; k++) r ^= a[k];
    if (r == 0x12345678u) out[0] = r;  // keeps the loop alive
}

template <int OP>
double run(const char *name, int sms, double mhz) {
    unsigned *d;
    cudaMalloc(&d, 4);
    dim3 grid(sms * 8), block(256);
    pipe_kernel<OP><<<grid, block>>>(d, 1u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) pipe_kernel<OP><<<grid, block>>>(d, 1u + r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 5.0 * (double)grid.x * 256 * ITERS * UNROLL;
    const double per_clk_sm = ops / (ms * 1e-3) / (mhz * 1e6) / sms;
    printf("\"%s\": {\"lane_ops_per_s\": %.4e, \"per_clk_per_sm\": %.2f}", name, ops / (ms * 1e-3), per_clk_sm);
    cudaFree(d);
    return per_clk_sm;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_mhz_max\": %.0f, ", p.name, p.multiProcessorCount, mhz);
    run<0>("popc_plus_iadd", p.multiProcessorCount, mhz);
    printf(", ");
    run<1>("lop3", p.multiProcessorCount, mhz);
    printf(", ");
    run<2>("imad", p.multiProcessorCount, mhz);
    printf("}\n");
    return 0;
}
