// Rate of mma.sync.aligned.m8n8k4.f64 against DFMA on one GPU: independent accumulator chains per warp, no memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_dmma scripts/microbench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters) {
    double c[CHAINS][2];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = (double)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters) {
    double c[CHAINS];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = (double)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<8><<<p.multiProcessorCount, threads>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * (threads / 32) * p.multiProcessorCount;
            if (rep) printf("DMMA m8n8k4, %4d threads/SM, 8 chains: %.2f TFLOP/s (%.3f ms)\n", threads, flop / ms * 1e-9, ms);
            cudaEventRecord(e0);
            dfma_kernel<8><<<p.multiProcessorCount, threads>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double flop2 = 2.0 * 8.0 * iters * threads * p.multiProcessorCount;
            if (rep) printf("DFMA,         %4d threads/SM, 8 chains: %.2f TFLOP/s (%.3f ms)\n", threads, flop2 / ms * 1e-9, ms);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
