// Ceiling of the expansion-pass access pattern: read in[x] (n_in vectors), write S images in place
// (image s at x + s*n_in).  Variants: cache hints, vectors per thread, block size.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int S, int U, int HINT>
__global__ void __launch_bounds__(1024) k_exp(float4 *p, uint64_t n_in) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n_in; i0 += stride) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(p + i0 + (uint64_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float c = 0.5f + 0.01f * s;
                const float4 o = make_float4(c * x[u].x - 0.1f * x[u].y, c * x[u].y + 0.1f * x[u].x, c * x[u].z - 0.1f * x[u].w, c * x[u].w + 0.1f * x[u].z);
                float4 *q = p + i0 + (uint64_t)u * blockDim.x + (uint64_t)s * n_in;
                if (HINT == 0) *q = o; else if (HINT == 1) __stcs(q, o); else __stcg(q, o);
            }
        }
    }
}
// stream-major order: for each s, all U vectors (so a thread's consecutive stores go to one stream)
template <int S, int U>
__global__ void __launch_bounds__(1024) k_exp_sm(float4 *p, uint64_t n_in) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n_in; i0 += stride) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(p + i0 + (uint64_t)u * blockDim.x);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float c = 0.5f + 0.01f * s;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4 o = make_float4(c * x[u].x - 0.1f * x[u].y, c * x[u].y + 0.1f * x[u].x, c * x[u].z - 0.1f * x[u].w, c * x[u].w + 0.1f * x[u].z);
                __stcs(p + i0 + (uint64_t)u * blockDim.x + (uint64_t)s * n_in, o);
            }
        }
    }
}
// out-of-place: read from a, write S images to b (no in-place overwrite)
template <int S, int U>
__global__ void __launch_bounds__(1024) k_exp_oop(const float4 *a, float4 *p, uint64_t n_in) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n_in; i0 += stride) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(a + i0 + (uint64_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float c = 0.5f + 0.01f * s;
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

int main() {
    const uint64_t bytes = 32ull << 30;            // 32 GiB result, S = 16 -> 2 GiB input
    float4 *p, *a; CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&a, bytes / 16)); CK(cudaMemset(p, 0, bytes)); CK(cudaMemset(a, 0, bytes / 16));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint64_t n_in = bytes / 16 / 16;
    const double moved = (double)bytes * (1.0 + 1.0 / 16);
    auto rep = [&](const char *name, float ms) { printf("%-58s %8.3f ms  %8.1f GB/s\n", name, ms, moved / ms / 1e6); fflush(stdout); };
    for (int thr : {256, 512}) for (int occ : {2, 4, 8}) {
        const int g = sms * occ;
        printf("-- %d CTAs/SM x %d threads\n", occ, thr);
        rep("in place S=16 U=1 .cs", timeit([&] { k_exp<16, 1, 1><<<g, thr>>>(p, n_in); }));
        rep("in place S=16 U=2 .cs", timeit([&] { k_exp<16, 2, 1><<<g, thr>>>(p, n_in); }));
        rep("in place S=16 U=4 .cs", timeit([&] { k_exp<16, 4, 1><<<g, thr>>>(p, n_in); }));
        rep("in place S=16 U=2 plain", timeit([&] { k_exp<16, 2, 0><<<g, thr>>>(p, n_in); }));
        rep("in place S=16 U=2 .cg", timeit([&] { k_exp<16, 2, 2><<<g, thr>>>(p, n_in); }));
        rep("in place S=16 U=4 stream-major .cs", timeit([&] { k_exp_sm<16, 4><<<g, thr>>>(p, n_in); }));
        rep("out of place S=16 U=2 .cs", timeit([&] { k_exp_oop<16, 2><<<g, thr>>>(a, p, n_in); }));
    }
    return 0;
}
