// Follow-up of membench4: is it the WINDOW of concurrently written addresses that costs k_expand_low its last 5 %?
// A CTA keeps its 512 KiB of output, but the pieces are interleaved with those of the G - 1 other CTAs of its "gang"
// at a granularity of R x 4 KiB: gang g, CTA c, step u -> ((g * ITERS / R + u / R) * G + c) * R + u % R   (4 KiB units).
// CTAs of a gang start together and advance at the same rate, so the gang writes one compact moving window of
// G x R x 4 KiB -- the k_init picture -- although every CTA still lives for 512 KiB.
//   nvcc -O3 -arch=sm_100a -o membench5 tools/membench5.cu && ./membench5 [GiB]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int ITERS>
__global__ void __launch_bounds__(256) k_seq_n(float4 *p, uint64_t nvec) {
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * ITERS + threadIdx.x;
#pragma unroll 4
    for (int u = 0; u < ITERS; ++u) {
        const uint64_t i = i0 + (uint64_t)u * blockDim.x;
        if (i < nvec) __stcs(p + i, make_float4(1.f, 2.f, 3.f, (float)u));
    }
}

template <int ITERS, int R>
__global__ void __launch_bounds__(256) k_seq_gang(float4 *p, uint64_t nvec, unsigned G) {
    const uint64_t g = blockIdx.x / G, c = blockIdx.x % G;
#pragma unroll 4
    for (int u = 0; u < ITERS; ++u) {
        const uint64_t piece = ((g * (ITERS / R) + (uint64_t)(u / R)) * G + c) * R + (u % R);
        const uint64_t i = piece * 256 + threadIdx.x;
        if (i < nvec) __stcs(p + i, make_float4(1.f, 2.f, 3.f, (float)u));
    }
}

// the k_expand_low shape of membench4 (TB = 8: 256 inputs of 2 KiB per CTA, a warp owns one batch of 32 inputs, pairs
// interleaved over the 8 warps) with the gang interleave: the CTA's k-th block of 16 inputs (32 KiB) is block k * G + c of
// its gang.  G = 0: the plain layout.
template <int LDS, int FMA, int PAUSE>
__global__ void __launch_bounds__(256) k_low_gang(float4 *p, const float4 *in, unsigned G) {
    __shared__ float4 sh[8][32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t base, stride;                                   // input index of the CTA's block 0, distance between its blocks
    if (G) { const uint64_t g = blockIdx.x / G, c = blockIdx.x % G; base = (g * 16 * G + c) * 16; stride = 16ull * G; }
    else { base = (uint64_t)blockIdx.x << 8; stride = 16; }
    const uint64_t x0 = base + 2u * warp;
    const float4 mine = in[(x0 + (uint64_t)(lane >> 1) * stride + (lane & 1)) & 0xfffff];
    float a = mine.x;
#pragma unroll 1
    for (int k = 0; k < PAUSE; ++k) a = a * 1.0001f + 0.5f;
#pragma unroll
    for (int s = 0; s < 4; ++s) sh[warp][s * 32 + lane] = make_float4(a, mine.y, mine.z, (float)s);
    __syncwarp();
#pragma unroll 2
    for (int i = 0; i < 32; ++i) {
        const uint64_t x = x0 + (uint64_t)(i >> 1) * stride + (i & 1);
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
    if (a == 12345.678f) p[0] = make_float4(a, 0, 0, 0);
}

template <typename F> float timeit(F f, int reps = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char **argv) {
    const uint64_t gib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 32;
    const uint64_t bytes = gib << 30;
    float4 *p, *in;
    CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 0, bytes));
    CK(cudaMalloc(&in, 16ull << 20)); CK(cudaMemset(in, 0, 16ull << 20));
    const uint64_t nvec = bytes / 16;
    float ms = timeit([&] { k_seq_n<8><<<(unsigned)(nvec / (256 * 8)), 256>>>(p, nvec); });
    printf("seq_n ITERS=8 (32 KiB per CTA)                      %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { k_seq_n<128><<<(unsigned)(nvec / (256 * 128)), 256>>>(p, nvec); });
    printf("seq_n ITERS=128 (512 KiB per CTA)                   %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    const unsigned ctas = (unsigned)(nvec / (256 * 128));
    for (unsigned G : {128u, 256u, 512u, 1024u, 2048u}) {
        ms = timeit([&] { k_seq_gang<128, 8><<<ctas, 256>>>(p, nvec, G); });
        printf("seq_gang 512 KiB per CTA, 32 KiB pieces, G=%4u      %8.3f ms  %8.1f GB/s\n", G, ms, bytes / ms / 1e6);
    }
    for (unsigned G : {256u, 1024u}) {
        ms = timeit([&] { k_seq_gang<128, 1><<<ctas, 256>>>(p, nvec, G); });
        printf("seq_gang 512 KiB per CTA,  4 KiB pieces, G=%4u      %8.3f ms  %8.1f GB/s\n", G, ms, bytes / ms / 1e6);
        ms = timeit([&] { k_seq_gang<128, 2><<<ctas, 256>>>(p, nvec, G); });
        printf("seq_gang 512 KiB per CTA,  8 KiB pieces, G=%4u      %8.3f ms  %8.1f GB/s\n", G, ms, bytes / ms / 1e6);
    }
    const unsigned lctas = (unsigned)((nvec >> 7) >> 8);
    for (unsigned G : {0u, 64u, 128u, 256u, 512u, 1024u, 2048u}) {
        ms = timeit([&] { k_low_gang<0, 0, 0><<<lctas, 256>>>(p, in, G); });
        printf("low bare store pattern, G=%4u                        %8.3f ms  %8.1f GB/s\n", G, ms, bytes / ms / 1e6);
        ms = timeit([&] { k_low_gang<4, 8, 300><<<lctas, 256>>>(p, in, G); });
        printf("low + LDS + multiplies + phase A, G=%4u              %8.3f ms  %8.1f GB/s\n", G, ms, bytes / ms / 1e6);
    }
    return 0;
}
