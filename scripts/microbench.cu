// Pipe-rate micro-benchmarks for the integer work of the local-energy kernels (B200, sm_100a):
//   POPC.32, LOP3, DADD, shared-memory ATOMS.OR (random words), LDS.128 with 16/64-byte strides.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench scripts/microbench.cu ; run: ./microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

__global__ void k_popc(uint32_t *out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u, c = a + 77u, d = b + 1234567u;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < ITER; ++i) {
        s0 += __popc(a ^ s3); s1 += __popc(b ^ s0); s2 += __popc(c ^ s1); s3 += __popc(d ^ s2);
        s0 += __popc(a + s2); s1 += __popc(b + s3); s2 += __popc(c + s0); s3 += __popc(d + s1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}
__global__ void k_popc_indep(uint32_t *out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += __popc(a ^ (uint32_t)(i * 8 + k) * 0x85ebca6bu);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
}
__global__ void k_lop(uint32_t *out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u, c = a + 77u, d = b + 1234567u;
    for (int i = 0; i < ITER; ++i) {
        a = (a & b) ^ c; b = (b | c) ^ d; c = (c & d) ^ a; d = (d | a) ^ b;
        a = (a & c) ^ d; b = (b | d) ^ a; c = (c & a) ^ b; d = (d | b) ^ c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
__global__ void k_dadd(double *out, double seed) {
    double a = threadIdx.x + seed, b = a * 0.5, c = a * 0.25, d = a * 0.125, e = 1.0 + seed;
    for (int i = 0; i < ITER; ++i) {
        a += e; b += e; c += e; d += e; a += b; b += c; c += d; d += a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
__global__ void k_atoms(uint32_t *out, uint32_t seed, int active_lanes) {
    extern __shared__ uint32_t bm[];
    for (int i = threadIdx.x; i < 32 * 736; i += blockDim.x) bm[i] = 0;
    __syncthreads();
    uint32_t *my = bm + (threadIdx.x >> 5) * 736;
    uint32_t r = threadIdx.x * 2654435761u + seed;
    const bool act = (threadIdx.x & 31) < active_lanes;
    for (int i = 0; i < ITER; ++i) {
        r = r * 1664525u + 1013904223u;
        const uint32_t u = (r >> 8) % 23157u;
        if (act) atomicOr(my + (u >> 5), 1u << (u & 31));
    }
    __syncthreads();
    uint32_t s = 0;
    for (int i = threadIdx.x; i < 32 * 736; i += blockDim.x) s += bm[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lds128(uint32_t *out, uint32_t seed, int stride_recs, int random_start) {
    extern __shared__ uint4 recs[];
    const int nrec = 8192;
    for (int i = threadIdx.x; i < nrec; i += blockDim.x) recs[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    uint32_t r = threadIdx.x * 2654435761u + seed;
    uint32_t s = 0;
    uint32_t base = random_start ? ((r >> 7) % 1024u) * 4u : (threadIdx.x & 31) * stride_recs;
    for (int i = 0; i < ITER; ++i) {
        const uint4 v = recs[(base + (i & 3)) & (nrec - 1)];
        s += v.x ^ v.y ^ v.z ^ v.w;
        if ((i & 3) == 3) base = random_start ? (base * 5u + 4u * 77u) & (nrec - 1) & ~3u : base + 32 * stride_recs;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    printf("%s: %d SMs, max clock %.0f MHz\n", p.name, sms, clk / 1e6);
    uint32_t *out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
    const int blocks = sms * 2, threads = 1024;
    const double lanes = (double)blocks * threads;
    float ms;
    ms = time_ms([&] { k_popc<<<blocks, threads>>>(out, 1); });
    printf("POPC (dependent mix, +1 IADD each): %.2f lane-ops/clk/SM\n", lanes * ITER * 8 / (ms * 1e-3) / clk / sms);
    ms = time_ms([&] { k_popc_indep<<<blocks, threads>>>(out, 1); });
    printf("POPC (independent, + LOP/IMAD/IADD): %.2f lane-ops/clk/SM\n", lanes * ITER * 8 / (ms * 1e-3) / clk / sms);
    ms = time_ms([&] { k_lop<<<blocks, threads>>>(out, 1); });
    printf("LOP3 (2-input pairs fuse to 1 LOP3): %.2f lane-ops/clk/SM\n", lanes * ITER * 8 / (ms * 1e-3) / clk / sms);
    ms = time_ms([&] { k_dadd<<<blocks, threads>>>((double *)out, 1.0); });
    printf("DADD: %.2f lane-ops/clk/SM\n", lanes * ITER * 8 / (ms * 1e-3) / clk / sms);
    cudaFuncSetAttribute(k_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 736 * 4);
    for (int act : {32, 8, 4, 1}) {
        ms = time_ms([&] { k_atoms<<<sms, threads, 32 * 736 * 4>>>(out, 1, act); });
        printf("ATOMS.OR random word of a per-warp 736-word row, %2d active lanes: %.2f lane-atomics/clk/SM, %.3f warp-instr/clk/SM\n", act,
               (double)sms * threads * ITER * act / 32 / (ms * 1e-3) / clk / sms, (double)sms * (threads / 32) * ITER / (ms * 1e-3) / clk / sms);
    }
    cudaFuncSetAttribute(k_lds128, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    for (int stride : {1, 4, 5, 6}) {
        ms = time_ms([&] { k_lds128<<<sms, threads, 8192 * 16>>>(out, 1, stride, 0); });
        printf("LDS.128 lane stride %d records (%3d B): %.1f B/clk/SM\n", stride, stride * 16, (double)sms * threads * ITER * 16 / (ms * 1e-3) / clk / sms);
    }
    ms = time_ms([&] { k_lds128<<<sms, threads, 8192 * 16>>>(out, 1, 0, 1); });
    printf("LDS.128 random 64-byte-aligned group starts: %.1f B/clk/SM\n", (double)sms * threads * ITER * 16 / (ms * 1e-3) / clk / sms);
    cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
