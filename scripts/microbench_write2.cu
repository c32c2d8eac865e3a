// Second write-pattern microbenchmark: what makes chunked row writes slow?  (see microbench_write.cu)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// MODE 0: a warp takes one chunk of `ch` rows per ticket (global counter, chunks in address order)
// MODE 1: a CTA takes 32 consecutive chunks per ticket, one per warp
// MODE 2: a CTA takes 32 samples per ticket (one per warp) and walks the n_tiles chunks of its sample in turn with a
//         __syncthreads() between chunks (what cycling the resident tile per 32 samples would look like)
// MODE 3: as MODE 0 but the chunk index is scrambled (chunks in random order)
// MODE 4: the emit kernel's pattern without its per-warp tickets: a CTA stays on tile blockIdx % n_tiles and takes 32 consecutive
//         samples per ticket (one per warp); warp w writes the chunk (sample, tile): 32 chunks at a stride of one sample row
template <int MIS>
__global__ void __launch_bounds__(1024, 1) wr4(int32_t *dest, long long *xp, double *H, long long n_samples, int ch, int n_tiles, unsigned *counters) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned s_t;
    for (int tried = 0; tried < n_tiles; ++tried) {
        const int ti = (blockIdx.x + tried) % n_tiles;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_t = atomicAdd(counters + 32 * ti, 1u);
            __syncthreads();
            const long long s = (long long)s_t * 32 + warp;
            if ((long long)s_t * 32 >= n_samples) break;
            if (s >= n_samples) continue;
            // MIS: every chunk starts MIS rows later and is MIS... rows shorter, so no warp store is 128-byte aligned
            const long long o0 = (s * n_tiles + ti) * ch + (MIS ? (MIS + (s * 7 + ti * 3) % 9) : 0);
            for (int k = lane; k < ch - 16; k += 32) {
                const long long r = o0 + k;
                __stcs(dest + r, (int)s); __stcs(xp + r, r ^ 0x5555); __stcs(H + r, (double)k);
            }
        }
    }
}
template <int MODE>
__global__ void __launch_bounds__(1024, 1) wr(int32_t *dest, long long *xp, double *H, long long n_chunks, int ch, int n_tiles, unsigned *counters) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned s_t;
    for (;;) {
        long long c0; int reps = 1;
        if (MODE == 0 || MODE == 3) {
            unsigned t = 0; if (lane == 0) t = atomicAdd(counters, 1u); t = __shfl_sync(~0u, t, 0);
            if (t >= n_chunks) return;
            c0 = t;
            if (MODE == 3) c0 = (long long)(((unsigned long long)t * 2654435761ull) % (unsigned long long)n_chunks);
        } else {
            __syncthreads();
            if (threadIdx.x == 0) s_t = atomicAdd(counters, 1u);
            __syncthreads();
            const unsigned t = s_t;
            if (MODE == 1) { c0 = (long long)t * 32 + warp; if ((long long)t * 32 >= n_chunks) return; if (c0 >= n_chunks) continue; }
            else { c0 = ((long long)t * 32 + warp) * n_tiles; reps = n_tiles; if ((long long)t * 32 * n_tiles >= n_chunks) return; }
        }
        for (int rr = 0; rr < reps; ++rr) {
            const long long o0 = (c0 + rr) * ch;
            if (c0 + rr < n_chunks)
                for (int k = lane; k < ch; k += 32) {
                    const long long r = o0 + k;
                    __stcs(dest + r, (int)c0); __stcs(xp + r, r ^ 0x5555); __stcs(H + r, (double)k);
                }
            if (MODE == 2) __syncthreads();
        }
    }
}
int main() {
    const long long M = 16384ll * 3840;
    int32_t *d; long long *x; double *h; unsigned *c;
    cudaMalloc(&d, M * 4); cudaMalloc(&x, M * 8); cudaMalloc(&h, M * 8); cudaMalloc(&c, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto kern, int ch, int nt) {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            cudaMemset(c, 0, 4096);
            cudaEventRecord(e0);
            kern<<<148, 1024>>>(d, x, h, M / ch, ch, nt, c);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
        }
        printf("%-40s chunk %5d rows: %.3f ms  %.0f GB/s  (%s)\n", name, ch, best, 20.0 * M / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int ch : {160, 320, 640, 1280, 1920, 3840, 7680}) run("warp per chunk, address order", wr<0>, ch, 1);
    for (int ch : {320, 1280, 3840}) run("warp per chunk, scrambled order", wr<3>, ch, 1);
    for (int ch : {160, 320, 640}) run("CTA takes 32 consecutive chunks", wr<1>, ch, 1);
    for (int nt : {6, 12, 24}) run("CTA: 32 samples, chunks in turn + sync", wr<2>, 3840 / nt, nt);
    for (int mis : {0, 1}) for (int nt : {6, 12, 24}) {
        const int ch = 3840 / nt;
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            cudaMemset(c, 0, 4096);
            cudaEventRecord(e0);
            if (mis) wr4<5><<<148, 1024>>>(d, x, h, M / 3840, ch, nt, c); else wr4<0><<<148, 1024>>>(d, x, h, M / 3840, ch, nt, c);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
        }
        printf("%-40s chunk %5d rows: %.3f ms  %.0f GB/s  (%s)\n", mis ? "same, chunk starts misaligned" : "CTA on one tile, 32 samples at a stride", ch, best, 20.0 * (M / 3840) * nt * (ch - 16) / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
