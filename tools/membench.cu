// HBM pattern ceilings on this GPU: what a write-only, read-only and copy stream reach,
// and what the strided multi-stream store pattern of the lazily materialising block pass
// reaches.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k_store(float4 *p, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (MODE == 0) p[i] = v;
        else if (MODE == 1) __stcs(p + i, v);
        else if (MODE == 2) __stwt(p + i, v);
        else __stcg(p + i, v);
    }
}

// each thread writes S vectors at stride n/S (S far-apart streams), like k_block's stores
template <int S, int MODE>
__global__ void __launch_bounds__(256) k_store_streams(float4 *p, uint64_t n) {
    const uint64_t per = n / S;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += stride) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (MODE == 0) p[i + s * per] = v; else __stcs(p + i + s * per, v);
        }
    }
}

template <int U>
__global__ void __launch_bounds__(256) k_read(const float4 *p, uint64_t n, float *out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    float acc = 0.f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i < n; i += stride) {
        float4 t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = (i + u * 256 < n) ? __ldcs(p + i + u * 256) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += t[u].x + t[u].y + t[u].z + t[u].w;
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <int U>
__global__ void __launch_bounds__(256) k_copy(const float4 *a, float4 *b, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i < n; i += stride) {
        float4 t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * 256 < n) t[u] = __ldcs(a + i + u * 256);
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * 256 < n) __stcs(b + i + u * 256, t[u]);
    }
}

// in-place read-modify-write (the dense gate pass pattern, two streams half the state apart)
template <int U>
__global__ void __launch_bounds__(256) k_rmw_pair(float4 *p, uint64_t n) {
    const uint64_t half = n / 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; i < half; i += stride) {
        float4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * 256 < half) { a[u] = __ldcs(p + i + u * 256); b[u] = __ldcs(p + half + i + u * 256); }
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * 256 < half) {
            float4 x = a[u], y = b[u];
            __stcs(p + i + u * 256, make_float4(0.6f * x.x + 0.8f * y.y, 0.6f * x.y - 0.8f * y.x, 0.6f * x.z + 0.8f * y.w, 0.6f * x.w - 0.8f * y.z));
            __stcs(p + half + i + u * 256, make_float4(0.6f * y.x + 0.8f * x.y, 0.6f * y.y - 0.8f * x.x, 0.6f * y.z + 0.8f * x.w, 0.6f * y.w - 0.8f * x.z));
        }
    }
}

template <typename F>
float timeit(F f, int reps = 3) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    const uint64_t bytes = (argc > 1 ? strtoull(argv[1], 0, 10) : 16ull) << 30;
    const uint64_t n = bytes / 16;
    float4 *a, *b; float *out;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char *name, double moved, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, moved / ms / 1e6); fflush(stdout); };
    rep("cudaMemsetAsync", bytes, timeit([&] { cudaMemsetAsync(a, 1, bytes); }));
    rep("cudaMemcpyAsync D2D (r+w)", 2.0 * bytes, timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }));
    for (int occ : {4, 8, 16}) {
        const int g = sms * occ;
        printf("-- grid = %d x %d\n", sms, occ);
        rep("store plain", bytes, timeit([&] { k_store<0><<<g, 256>>>(a, n); }));
        rep("store .cs", bytes, timeit([&] { k_store<1><<<g, 256>>>(a, n); }));
        rep("store .wt", bytes, timeit([&] { k_store<2><<<g, 256>>>(a, n); }));
        rep("store .cg", bytes, timeit([&] { k_store<3><<<g, 256>>>(a, n); }));
        rep("store 16 streams plain", bytes, timeit([&] { k_store_streams<16, 0><<<g, 256>>>(a, n); }));
        rep("store 16 streams .cs", bytes, timeit([&] { k_store_streams<16, 1><<<g, 256>>>(a, n); }));
        rep("store 32 streams .cs", bytes, timeit([&] { k_store_streams<32, 1><<<g, 256>>>(a, n); }));
        rep("store 4 streams .cs", bytes, timeit([&] { k_store_streams<4, 1><<<g, 256>>>(a, n); }));
        rep("read U=4", bytes, timeit([&] { k_read<4><<<g, 256>>>(a, n, out); }));
        rep("read U=8", bytes, timeit([&] { k_read<8><<<g, 256>>>(a, n, out); }));
        rep("copy U=4 (r+w)", 2.0 * bytes, timeit([&] { k_copy<4><<<g, 256>>>(a, b, n); }));
        rep("copy U=8 (r+w)", 2.0 * bytes, timeit([&] { k_copy<8><<<g, 256>>>(a, b, n); }));
        rep("rmw pair in place U=2 (r+w)", 2.0 * bytes, timeit([&] { k_rmw_pair<2><<<g, 256>>>(a, n); }));
        rep("rmw pair in place U=4 (r+w)", 2.0 * bytes, timeit([&] { k_rmw_pair<4><<<g, 256>>>(a, n); }));
    }
    return 0;
}
