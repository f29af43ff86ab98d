// What keeps the rotated final expansion pass (k_expand_low: ONE sequential write stream, 7.1 TB/s) below the
// 7.4-7.6 TB/s that memset and k_init reach?  Variants of its store pattern, from a bare sequential writer to
// the kernel's own shape (8 warps per CTA, a warp emits 2 KiB per input -- four 512-byte warp stores -- inputs
// interleaved over the warps at a granularity of two), with the kernel's other shared-memory and issue load
// switched on one by one.  Not product code: a tuning aid for the next round (build: nvcc -O3 -arch=sm_100a).
//   ./membench4 [GiB written per launch, default 32]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// V0: the k_init shape -- a CTA covers 32 KiB, warps adjacent, 8 stores per thread
__global__ void __launch_bounds__(256) k_seq(float4 *p, uint64_t nvec) {
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * 8 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const uint64_t i = i0 + (uint64_t)u * blockDim.x;
        if (i < nvec) __stcs(p + i, make_float4(1.f, 2.f, 3.f, (float)u));
    }
}

// V0b: the same shape with a CTA covering ITERS x 4 KiB (a loop of ITERS stores per thread): does the size of a CTA's
// region matter, or only the order inside it?  (round 2)
template <int ITERS>
__global__ void __launch_bounds__(256) k_seq_n(float4 *p, uint64_t nvec) {
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * ITERS + threadIdx.x;
#pragma unroll 4
    for (int u = 0; u < ITERS; ++u) {
        const uint64_t i = i0 + (uint64_t)u * blockDim.x;
        if (i < nvec) __stcs(p + i, make_float4(1.f, 2.f, 3.f, (float)u));
    }
}

// V0c: the cooperative k_expand_low shape (round 2): a CTA owns 2^TB inputs of 2 KiB; at every step its 8 warps write
// 8 consecutive 512-byte pieces (4 KiB contiguous, like k_init), each piece = one broadcast LDS.128 + two 8-byte LDS
// and 12 fp32 operations; before that a store-free phase (thread t prepares input t, __syncthreads).
template <int TB, int PAUSE>
__global__ void __launch_bounds__(256) k_coop(float4 *p, const float4 *in) {
    __shared__ float4 us[4][256];
    __shared__ float2 as[4][257], bs[8][257];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int tile = 0; tile < (1 << TB) / 256; ++tile) {
        const uint64_t x0 = ((uint64_t)blockIdx.x << TB) + (uint64_t)tile * 256;
        const float4 mine = in[(x0 + threadIdx.x) & 0xfffff];
        float a = mine.x;
#pragma unroll 1
        for (int k = 0; k < PAUSE; ++k) a = a * 1.0001f + 0.5f;
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 4; ++s) us[s][threadIdx.x] = make_float4(a, mine.y, mine.z, (float)s);
#pragma unroll
        for (int t = 0; t < 4; ++t) as[t][threadIdx.x] = make_float2(a, mine.y + t);
#pragma unroll
        for (int t = 0; t < 8; ++t) bs[t][threadIdx.x] = make_float2(a, mine.z + t);
        __syncthreads();
#pragma unroll 4
        for (int g = warp; g < 4 * 256; g += 8) {
            const int i = g >> 2, s = g & 3;
            const float2 fa = as[lane & 3][i], fb = bs[lane >> 2][i];
            const float4 u = us[s][i];
            const float lr = fa.x * fb.x - fa.y * fb.y, li = fa.x * fb.y + fa.y * fb.x;
            __stcs(p + ((x0 + i) << 7) + (uint64_t)s * 32 + lane,
                   make_float4(lr * u.x - li * u.y, lr * u.y + li * u.x, lr * u.z - li * u.w, lr * u.w + li * u.z));
        }
        acc += a;
    }
    if (acc == 12345.678f) p[0] = make_float4(acc, 0, 0, 0);
}

// V1..V4: the k_expand_low shape.  TB: log2 inputs per CTA; LDS: shared-memory broadcast loads per input (the
// kernel does 2 + 4); FMA: dependent fp32 operations per store (the kernel: ~8 per 16 bytes); PAUSE: a phase
// without stores per 32 inputs (the kernel's phase A), in dependent fp32 operations.
template <int TB, int LDS, int FMA, int PAUSE>
__global__ void __launch_bounds__(256) k_low(float4 *p, const float4 *in) {
    __shared__ float4 sh[8][32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kPerWarp = (1 << TB) / 8;
    float acc = 0.f;
    for (int batch = 0; batch < kPerWarp / 32; ++batch) {
        const uint64_t x0 = ((uint64_t)blockIdx.x << TB) + (uint64_t)batch * 256 + 2u * warp;
        const float4 mine = in[(x0 + (uint64_t)(lane >> 1) * 16 + (lane & 1)) & 0xfffff];
        float a = mine.x;
#pragma unroll 1
        for (int k = 0; k < PAUSE; ++k) a = a * 1.0001f + 0.5f;
#pragma unroll
        for (int s = 0; s < 4; ++s) sh[warp][s * 32 + lane] = make_float4(a, mine.y, mine.z, (float)s);
        __syncwarp();
#pragma unroll 2
        for (int i = 0; i < 32; ++i) {
            const uint64_t x = x0 + (uint64_t)(i >> 1) * 16 + (i & 1);
            float4 u[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) u[s] = (s < LDS) ? sh[warp][s * 32 + i] : make_float4(1.f, 2.f, 3.f, 4.f);
            float l = 1.f + 1e-3f * lane;
#pragma unroll
            for (int k = 0; k < FMA; ++k) l = l * 1.0001f + u[k & 3].x;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                __stcs(p + (x << 7) + (uint64_t)s * 32 + lane, make_float4(l * u[s].x, l * u[s].y, l * u[s].z, l * u[s].w));
        }
        __syncwarp();
        acc += a;
    }
    if (acc == 12345.678f) p[0] = make_float4(acc, 0, 0, 0);
}

template <typename F> float timeit(F f, int reps = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

template <int TB, int LDS, int FMA, int PAUSE>
void run_low(float4 *p, const float4 *in, uint64_t nvec, const char *what) {
    const uint64_t inputs = nvec >> 7;                       // 128 16-byte vectors (2 KiB) per input
    const unsigned grid = (unsigned)(inputs >> TB);
    float ms = timeit([&] { k_low<TB, LDS, FMA, PAUSE><<<grid, 256>>>(p, in); });
    printf("low  TB=%2d LDS=%d FMA=%2d PAUSE=%4d  %8.3f ms  %8.1f GB/s   %s\n", TB, LDS, FMA, PAUSE, ms, 16.0 * nvec / ms / 1e6, what);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const uint64_t gib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 32;
    const uint64_t bytes = gib << 30;
    float4 *p, *in;
    CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 0, bytes));
    CK(cudaMalloc(&in, 16ull << 20)); CK(cudaMemset(in, 0, 16ull << 20));
    const uint64_t nvec = bytes / 16;
    float ms = timeit([&] { CK(cudaMemsetAsync(p, 1, bytes)); });
    printf("memset                                %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { k_seq<<<(unsigned)((nvec + 2047) / 2048), 256>>>(p, nvec); });
    printf("seq  (k_init shape, 32 KiB per CTA)   %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { k_seq_n<8><<<(unsigned)(nvec / (256 * 8)), 256>>>(p, nvec); });
    printf("seq_n ITERS=   8 (32 KiB per CTA)      %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { k_seq_n<32><<<(unsigned)(nvec / (256 * 32)), 256>>>(p, nvec); });
    printf("seq_n ITERS=  32 (128 KiB per CTA)     %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { k_seq_n<128><<<(unsigned)(nvec / (256 * 128)), 256>>>(p, nvec); });
    printf("seq_n ITERS= 128 (512 KiB per CTA)     %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { k_seq_n<512><<<(unsigned)(nvec / (256 * 512)), 256>>>(p, nvec); });
    printf("seq_n ITERS= 512 (2 MiB per CTA)       %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    {
        const uint64_t inputs = nvec >> 7;
        ms = timeit([&] { k_coop<8, 300><<<(unsigned)(inputs >> 8), 256>>>(p, in); });
        printf("coop TB= 8 PAUSE=300 (cooperative)    %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
        ms = timeit([&] { k_coop<10, 300><<<(unsigned)(inputs >> 10), 256>>>(p, in); });
        printf("coop TB=10 PAUSE=300 (4 tiles/CTA)    %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    }
    run_low<8, 0, 0, 0>(p, in, nvec, "bare store pattern, 256 inputs per CTA");
    run_low<10, 0, 0, 0>(p, in, nvec, "bare store pattern, 1024 inputs per CTA");
    run_low<8, 0, 8, 0>(p, in, nvec, "+ the multiplies");
    run_low<8, 4, 8, 0>(p, in, nvec, "+ 4 broadcast LDS.128 per input");
    run_low<8, 4, 8, 300>(p, in, nvec, "+ a store-free phase per 32 inputs (phase A)");
    run_low<8, 4, 8, 1200>(p, in, nvec, "+ a long store-free phase");
    run_low<10, 4, 8, 300>(p, in, nvec, "same, 1024 inputs per CTA");
    return 0;
}
