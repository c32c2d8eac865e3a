// Write-pattern microbenchmark for the emit kernel: how fast can 148 x 32 warps stream (dest 4 B, x' 8 B, H 8 B) rows to HBM
// when every warp writes chunks of CH consecutive rows and the chunks are handed out in different orders?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_write scripts/microbench_write.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE, bool CS>
__global__ void __launch_bounds__(1024, 1) wr(int32_t *dest, long long *xp, double *H, long long n_samples, int per_sample, int n_tiles,
                                              unsigned *counters) {
    const int lane = threadIdx.x & 31;
    const int ch = per_sample / n_tiles;
    for (;;) {
        long long s; int ti;
        if (MODE == 0) {  // as the emit kernel: the CTA stays on tile blockIdx % n_tiles, samples by a per-tile counter
            ti = blockIdx.x % n_tiles;
            unsigned t = 0; if (lane == 0) t = atomicAdd(counters + ti, 1u); t = __shfl_sync(~0u, t, 0);
            if (t >= n_samples) { // help the next tiles
                bool found = false;
                for (int k = 1; k < n_tiles && !found; ++k) { ti = (blockIdx.x + k) % n_tiles; if (lane == 0) t = atomicAdd(counters + ti, 1u); t = __shfl_sync(~0u, t, 0); found = t < n_samples; }
                if (!found) return;
            }
            s = t;
        } else if (MODE == 1) {  // one global counter, chunk-major: consecutive tickets = consecutive chunks of one sample
            unsigned long long t = 0; if (lane == 0) t = atomicAdd((unsigned long long *)counters, 1ull); t = __shfl_sync(~0u, t, 0);
            if (t >= (unsigned long long)n_samples * n_tiles) return;
            s = t / n_tiles; ti = (int)(t % n_tiles);
        } else {  // whole sample per warp
            unsigned t = 0; if (lane == 0) t = atomicAdd(counters, 1u); t = __shfl_sync(~0u, t, 0);
            if (t >= n_samples) return;
            s = t; ti = -1;
        }
        const long long o0 = s * per_sample + (ti < 0 ? 0 : (long long)ti * ch);
        const int len = ti < 0 ? per_sample : ch;
        for (int k = lane; k < len; k += 32) {
            const long long r = o0 + k;
            if (CS) { __stcs(dest + r, (int)s); __stcs(xp + r, r ^ 0x5555); __stcs(H + r, (double)k); }
            else { dest[r] = (int)s; xp[r] = r ^ 0x5555; H[r] = (double)k; }
        }
    }
}
int main() {
    const long long n = 16384; const int per = 3840, nt = 12;
    const long long M = n * per;
    int32_t *d; long long *x; double *h; unsigned *c;
    cudaMalloc(&d, M * 4); cudaMalloc(&x, M * 8); cudaMalloc(&h, M * 8); cudaMalloc(&c, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto kern) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaMemset(c, 0, 4096);
            cudaEventRecord(e0);
            kern<<<148, 1024>>>(d, x, h, n, per, nt, c);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
        }
        printf("%-48s %.3f ms  %.0f GB/s  (%s)\n", name, best, 20.0 * M / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    run("per-tile counters (emit kernel order), st.cs", wr<0, true>);
    run("per-tile counters (emit kernel order), plain", wr<0, false>);
    run("chunk-major global counter, st.cs", wr<1, true>);
    run("chunk-major global counter, plain", wr<1, false>);
    run("whole sample per warp, st.cs", wr<2, true>);
    run("whole sample per warp, plain", wr<2, false>);
    return 0;
}
