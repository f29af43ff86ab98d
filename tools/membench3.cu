// How many in-place write streams per input vector does the expansion pattern tolerate?
// One tile per CTA in launch order; S images of each input vector, image s at x + s*n_in.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int S, int U>
__global__ void __launch_bounds__(512) k_exp(float4 *p, uint64_t n_in) {
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = (i0 + (uint64_t)u * blockDim.x < n_in) ? __ldcs(p + i0 + (uint64_t)u * blockDim.x) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (i0 + (uint64_t)u * blockDim.x >= n_in) continue;
#pragma unroll 16
        for (int s = 0; s < S; ++s) {
            const float c = 0.5f + 0.001f * s;
            __stcs(p + i0 + (uint64_t)u * blockDim.x + (uint64_t)s * n_in,
                   make_float4(c * x[u].x - 0.1f * x[u].y, c * x[u].y + 0.1f * x[u].x, c * x[u].z - 0.1f * x[u].w, c * x[u].w + 0.1f * x[u].z));
        }
    }
}

template <typename F> float timeit(F f, int reps = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

template <int S, int U>
void run(float4 *p, uint64_t total_vecs, int thr) {
    const uint64_t n_in = total_vecs / S;
    const uint64_t grid = (n_in + (uint64_t)thr * U - 1) / ((uint64_t)thr * U);
    const double moved = 16.0 * (double)total_vecs * (1.0 + 1.0 / S);
    float ms = timeit([&] { k_exp<S, U><<<(unsigned)grid, thr>>>(p, n_in); });
    printf("S=%3d U=%d thr=%d  %8.3f ms  %8.1f GB/s\n", S, U, thr, ms, moved / ms / 1e6); fflush(stdout);
}

int main() {
    const uint64_t bytes = 32ull << 30;
    float4 *p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 0, bytes));
    const uint64_t total = bytes / 16;
    for (int thr : {256, 512}) {
        run<2, 4>(p, total, thr); run<4, 4>(p, total, thr); run<8, 2>(p, total, thr); run<16, 2>(p, total, thr);
        run<32, 2>(p, total, thr); run<32, 1>(p, total, thr); run<64, 1>(p, total, thr); run<128, 1>(p, total, thr); run<256, 1>(p, total, thr);
    }
    return 0;
}
