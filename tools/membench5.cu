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
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

extern __shared__ float4 dyn_pad[];                      // occupancy knob: dynamic shared memory nobody touches

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

// Warp-specialised shape: NP producer warps do the store-free phase A (input load, PAUSE dependent operations, the per-
// input tables into a shared-memory ring of S stages of 32 inputs), NC consumer warps do nothing but phase B (broadcast
// loads, two multiplies, 512-byte stores): the stores of an SM come from NC warps that never wait for anything else --
// the occupancy sweep above says few resident store warps are what a writer wants.  One mbarrier pair per stage.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

template <int TB, int NP, int NC, int S, int PAUSE>
__global__ void __launch_bounds__((NP + NC) * 32) k_spec(float4 *p, const float4 *in) {
    __shared__ float4 us[S][4][32];
    __shared__ float2 as[S][4][33], bs[S][8][33];
    __shared__ uint64_t bars[2 * S];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[S + s]), NC); }
    }
    __syncthreads();
    constexpr int NB = (1 << TB) / 32;                       // batches of 32 inputs per CTA
    const uint64_t xb = (uint64_t)blockIdx.x << TB;
    if (warp < NP) {
        float4 nxt = in[(xb + (uint64_t)warp * 32 + lane) & 0xfffff];
        for (int b = warp; b < NB; b += NP) {
            const int st = b % S, k = b / S;
            const float4 mine = nxt;
            if (b + NP < NB) nxt = in[(xb + (uint64_t)(b + NP) * 32 + lane) & 0xfffff];
            float a = mine.x;
#pragma unroll 1
            for (int q = 0; q < PAUSE; ++q) a = a * 1.0001f + 0.5f;
            mbar_wait(smem_u32(&bars[S + st]), (k & 1) ^ 1);
#pragma unroll
            for (int t = 0; t < 4; ++t) us[st][t][lane] = make_float4(a, mine.y, mine.z, (float)t);
#pragma unroll
            for (int t = 0; t < 4; ++t) as[st][t][lane] = make_float2(a, mine.y + t);
#pragma unroll
            for (int t = 0; t < 8; ++t) bs[st][t][lane] = make_float2(a, mine.z + t);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[st]));
        }
    } else {
        const int cw = warp - NP;
        for (int b = 0; b < NB; ++b) {
            const int st = b % S, k = b / S;
            mbar_wait(smem_u32(&bars[st]), k & 1);
#pragma unroll 2
            for (int i = cw; i < 32; i += NC) {              // the consumer warps cover NC consecutive inputs per step
                const float2 fa = as[st][lane & 3][i], fb = bs[st][lane >> 2][i];
                const float lr = fa.x * fb.x - fa.y * fb.y, li = fa.x * fb.y + fa.y * fb.x;
                const uint64_t x = xb + (uint64_t)b * 32 + i;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 u = us[st][t][i];
                    __stcs(p + (x << 7) + (uint64_t)t * 32 + lane,
                           make_float4(lr * u.x - li * u.y, lr * u.y + li * u.x, lr * u.z - li * u.w, lr * u.w + li * u.z));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[S + st]));
        }
    }
}

template <int TB, int NP, int NC, int S, int PAUSE>
void run_spec(float4 *p, const float4 *in, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)((nvec >> 7) >> TB);
    const int dyn = per_sm <= 0 ? 0 : std::max(0, (227 * 1024) / per_sm - 1024 - (int)(S * (2048 + 12 * 33 * 8) + 1024));
    CK(cudaFuncSetAttribute(k_spec<TB, NP, NC, S, PAUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spec<TB, NP, NC, S, PAUSE>, (NP + NC) * 32, dyn));
    float ms = timeit([&] { k_spec<TB, NP, NC, S, PAUSE><<<grid, (NP + NC) * 32, dyn>>>(p, in); });
    printf("spec TB=%2d producers=%d consumers=%d stages=%d PAUSE=%4d, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", TB, NP, NC, S, PAUSE, occ, ms,
           16.0 * nvec / ms / 1e6);
    fflush(stdout);
}

// Output through shared memory and the bulk-copy engine (cp.async.bulk.global.shared::cta): a warp parks the 2 KiB of
// an input (IPB inputs: WB = IPB x 2 KiB) in its own shared-memory buffer, one lane issues ONE bulk store for them.  The
// memory system then sees WB-byte bursts from 1 thread instead of 4 x IPB 512-byte warp stores.  k_low's shape otherwise.
__device__ __forceinline__ void bulk_s2g(void *g, uint32_t s, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int IPB, int NBUF, int LDS, int FMA, int PAUSE>
__global__ void __launch_bounds__(256) k_tma_low(float4 *p, const float4 *in) {
    extern __shared__ __align__(128) float4 obuf[];          // [8 warps][NBUF][IPB * 128]
    __shared__ float4 sh[8][32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *mybuf = obuf + (size_t)warp * NBUF * IPB * 128;
    const uint64_t x0 = ((uint64_t)blockIdx.x << 8) + 2u * warp;
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
        float4 *buf = mybuf + (size_t)((i / IPB) % NBUF) * IPB * 128 + (i % IPB) * 128;
        if (i % IPB == 0) {
            if (lane == 0) bulk_wait_read<NBUF - 1>();       // the buffer's previous bulk store has read it
            __syncwarp();
        }
        float4 u[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) u[s] = (s < LDS) ? sh[warp][s * 32 + i] : make_float4(1.f, 2.f, 3.f, 4.f);
        float l = 1.f + 1e-3f * lane;
#pragma unroll
        for (int k = 0; k < FMA; ++k) l = l * 1.0001f + u[k & 3].x;
#pragma unroll
        for (int s = 0; s < 4; ++s) buf[s * 32 + lane] = make_float4(l * u[s].x, l * u[s].y, l * u[s].z, l * u[s].w);
        if (i % IPB == IPB - 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                const uint64_t xf = x - (IPB - 1);           // IPB <= 2: the inputs of a pair are adjacent
                bulk_s2g(p + (xf << 7), smem_u32(buf - (IPB - 1) * 128), IPB * 2048);
                bulk_commit();
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    if (a == 12345.678f) p[0] = make_float4(a, 0, 0, 0);
}

// the whole CTA fills CH bytes (CH / 2 KiB consecutive inputs), one thread issues one bulk store for them
template <int CH, int NBUF, int ITERS>
__global__ void __launch_bounds__(256) k_tma_cta(float4 *p) {
    extern __shared__ __align__(128) float4 obuf[];          // [NBUF][CH / 16]
    const uint64_t base = (uint64_t)blockIdx.x * ITERS * (CH / 16);
    for (int it = 0; it < ITERS; ++it) {
        float4 *buf = obuf + (size_t)(it % NBUF) * (CH / 16);
        if (threadIdx.x == 0) bulk_wait_read<NBUF - 1>();
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CH / 4096; ++u) buf[u * 256 + threadIdx.x] = make_float4(1.f, 2.f, 3.f, (float)it);
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) { bulk_s2g(p + base + (uint64_t)it * (CH / 16), smem_u32(buf), CH); bulk_commit(); }
    }
    if (threadIdx.x == 0) bulk_wait_read<0>();
}

template <int IPB, int NBUF, int LDS, int FMA, int PAUSE>
void run_tma_low(float4 *p, const float4 *in, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)((nvec >> 7) >> 8);
    const int need = 8 * NBUF * IPB * 2048;
    const int dyn = std::max(need, per_sm <= 0 ? 0 : (227 * 1024) / per_sm - 1024 - 17 * 1024);
    CK(cudaFuncSetAttribute(k_tma_low<IPB, NBUF, LDS, FMA, PAUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma_low<IPB, NBUF, LDS, FMA, PAUSE>, 256, dyn));
    float ms = timeit([&] { k_tma_low<IPB, NBUF, LDS, FMA, PAUSE><<<grid, 256, dyn>>>(p, in); });
    printf("tma_low %d KiB per bulk store, %d buffers per warp, LDS=%d FMA=%d PAUSE=%d, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", IPB * 2, NBUF,
           LDS, FMA, PAUSE, occ, ms, 16.0 * nvec / ms / 1e6);
    fflush(stdout);
}

template <int CH, int NBUF, int ITERS>
void run_tma_cta(float4 *p, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)(nvec * 16 / ((uint64_t)CH * ITERS));
    const int need = NBUF * CH;
    const int dyn = std::max(need, per_sm <= 0 ? 0 : (227 * 1024) / per_sm - 2048);
    CK(cudaFuncSetAttribute(k_tma_cta<CH, NBUF, ITERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma_cta<CH, NBUF, ITERS>, 256, dyn));
    float ms = timeit([&] { k_tma_cta<CH, NBUF, ITERS><<<grid, 256, dyn>>>(p); });
    printf("tma_cta %2d KiB per bulk store, %d buffers, %4d KiB per CTA, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", CH / 1024, NBUF,
           CH / 1024 * ITERS, occ, ms, 16.0 * nvec / ms / 1e6);
    fflush(stdout);
}


// k_low's shape with the output through a CTA-wide shared-memory ring: at step k of a batch the 8 compute warps write
// their input pairs of chunk k (16 consecutive inputs = 32 KiB) into ring buffer c % NBUF, a ninth warp waits for the
// chunk (mbarrier, one arrive per compute warp) and issues ONE 32 KiB bulk store; it frees a buffer as soon as the
// bulk engine has read it (wait_group.read).  An SM then emits one sequential stream of 32 KiB bursts per CTA.
template <int TB, int NBUF, int LDS, int FMA, int PAUSE>
__global__ void __launch_bounds__(288) k_ring(float4 *p, const float4 *in) {
    extern __shared__ __align__(128) float4 ring[];          // [NBUF][16 * 128]
    __shared__ float4 sh[8][32 * 4];
    __shared__ uint64_t bars[2 * NBUF];                      // full[NBUF] (8 arrivals), empty[NBUF] (1 arrival)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        for (int b = 0; b < NBUF; ++b) { mbar_init(smem_u32(&bars[b]), 8); mbar_init(smem_u32(&bars[NBUF + b]), 1); }
    __syncthreads();
    constexpr int kBatches = (1 << TB) / 256, kChunks = kBatches * 16;
    const uint64_t xt = (uint64_t)blockIdx.x << TB;
    if (warp == 8) {
        if (lane == 0) {
            for (int c = 0; c < kChunks; ++c) {
                const int b = c % NBUF;
                mbar_wait(smem_u32(&bars[b]), (c / NBUF) & 1);
                bulk_s2g(p + ((xt + (uint64_t)c * 16) << 7), smem_u32(ring + (size_t)b * 2048), 32768);
                bulk_commit();
                if (c >= 1) { bulk_wait_read<1>(); mbar_arrive(smem_u32(&bars[NBUF + (c - 1) % NBUF])); }
            }
            bulk_wait_read<0>();
        }
        return;
    }
    float acc = 0.f;
    float4 nxt = in[(xt + 2u * warp + (uint64_t)(lane >> 1) * 16 + (lane & 1)) & 0xfffff];
    for (int batch = 0; batch < kBatches; ++batch) {
        const uint64_t x0 = xt + (uint64_t)batch * 256 + 2u * warp;
        const float4 mine = nxt;
        if (batch + 1 < kBatches) nxt = in[(x0 + 256 + (uint64_t)(lane >> 1) * 16 + (lane & 1)) & 0xfffff];
        float a = mine.x;
#pragma unroll 1
        for (int k = 0; k < PAUSE; ++k) a = a * 1.0001f + 0.5f;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 4; ++s) sh[warp][s * 32 + lane] = make_float4(a, mine.y, mine.z, (float)s);
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            const int c = batch * 16 + k, b = c % NBUF;
            mbar_wait(smem_u32(&bars[NBUF + b]), ((c / NBUF) & 1) ^ 1);
            float4 *dst = ring + (size_t)b * 2048 + (2 * warp) * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 2 * k + h;
                float4 u[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) u[s] = (s < LDS) ? sh[warp][s * 32 + i] : make_float4(1.f, 2.f, 3.f, 4.f);
                float l = 1.f + 1e-3f * lane;
#pragma unroll
                for (int q = 0; q < FMA; ++q) l = l * 1.0001f + u[q & 3].x;
#pragma unroll
                for (int s = 0; s < 4; ++s) dst[h * 128 + s * 32 + lane] = make_float4(l * u[s].x, l * u[s].y, l * u[s].z, l * u[s].w);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[b]));
        }
        acc += a;
    }
    if (acc == 12345.678f) p[0] = make_float4(acc, 0, 0, 0);
}

template <int TB, int NBUF, int LDS, int FMA, int PAUSE>
void run_ring(float4 *p, const float4 *in, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)((nvec >> 7) >> TB);
    const int need = NBUF * 32768;
    const int dyn = std::max(need, per_sm <= 0 ? 0 : (227 * 1024) / per_sm - 1024 - 17 * 1024);
    CK(cudaFuncSetAttribute(k_ring<TB, NBUF, LDS, FMA, PAUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ring<TB, NBUF, LDS, FMA, PAUSE>, 288, dyn));
    float ms = timeit([&] { k_ring<TB, NBUF, LDS, FMA, PAUSE><<<grid, 288, dyn>>>(p, in); });
    printf("ring TB=%2d, %d x 32 KiB buffers, LDS=%d FMA=%d PAUSE=%4d, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", TB, NBUF, LDS, FMA, PAUSE, occ, ms,
           16.0 * nvec / ms / 1e6);
    fflush(stdout);
}

// Software-pipelined k_low: phase A of batch b + 1 (the PAUSE dependent operations and the table stores) is spread over
// the 32 store iterations of batch b (double-buffered tables), so a warp never stops storing between batches -- only the
// CTA's first batch has a store-free phase.  NB batches per warp (TB = 8 + log2 NB).
template <int NB, int PAUSE>
__global__ void __launch_bounds__(256) k_low_sp(float4 *p, const float4 *in) {
    __shared__ float4 sh[2][8][32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t xt = (uint64_t)blockIdx.x * (256 * NB);
    auto in_of = [&](int b) { return in[(xt + (uint64_t)b * 256 + 2u * warp + (uint64_t)(lane >> 1) * 16 + (lane & 1)) & 0xfffff]; };
    float4 mine = in_of(0);
    float a = mine.x;
#pragma unroll 1
    for (int k = 0; k < PAUSE; ++k) a = a * 1.0001f + 0.5f;
#pragma unroll
    for (int s = 0; s < 4; ++s) sh[0][warp][s * 32 + lane] = make_float4(a, mine.y, mine.z, (float)s);
    __syncwarp();
    float acc = 0.f;
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
        const int cur = b & 1;
        const uint64_t x0 = xt + (uint64_t)b * 256 + 2u * warp;
        const bool more = b + 1 < NB;
        float4 nx = more ? in_of(b + 1) : make_float4(0, 0, 0, 0);
        float an = nx.x;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint64_t x = x0 + (uint64_t)(i >> 1) * 16 + (i & 1);
            float4 u[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) u[s] = sh[cur][warp][s * 32 + i];
            float l = 1.f + 1e-3f * lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) l = l * 1.0001f + u[k & 3].x;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                __stcs(p + (x << 7) + (uint64_t)s * 32 + lane, make_float4(l * u[s].x, l * u[s].y, l * u[s].z, l * u[s].w));
            // this iteration's slice of the next batch's phase A
#pragma unroll
            for (int k = 0; k < (PAUSE + 31) / 32; ++k) an = an * 1.0001f + 0.5f;
        }
        if (more) {
#pragma unroll
            for (int s = 0; s < 4; ++s) sh[cur ^ 1][warp][s * 32 + lane] = make_float4(an, nx.y, nx.z, (float)s);
        }
        __syncwarp();
        acc += an;
    }
    if (acc == 12345.678f) p[0] = make_float4(acc, 0, 0, 0);
}

template <int NB, int PAUSE>
void run_low_sp(float4 *p, const float4 *in, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)((nvec >> 7) / (256 * NB));
    const int dyn = per_sm <= 0 ? 0 : std::max(0, (227 * 1024) / per_sm - 1024 - 33 * 1024);
    CK(cudaFuncSetAttribute(k_low_sp<NB, PAUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_low_sp<NB, PAUSE>, 256, dyn));
    float ms = timeit([&] { k_low_sp<NB, PAUSE><<<grid, 256, dyn>>>(p, in); });
    printf("low_sp (phase A inside the store loop) %2d batches per warp, PAUSE=%4d, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", NB, PAUSE, occ, ms,
           16.0 * nvec / ms / 1e6);
    fflush(stdout);
}

// Bare store patterns of a long-lived 8-warp CTA over its 512 KiB (1024 pieces of 512 bytes = one warp store each):
// P = 0: at step u the 8 warps write the 8 consecutive pieces 8u .. 8u+7 (4 KiB contiguous per step, k_init's order);
// P = 1: k_expand_low's order -- warp w writes the 4 pieces of input 16k + 2w, then those of input 16k + 2w + 1;
// P = 2: every warp its own contiguous 64 KiB.
template <int P>
__global__ void __launch_bounds__(256) k_pat(float4 *p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *base = p + (uint64_t)blockIdx.x * 1024 * 32;
#pragma unroll 8
    for (int u = 0; u < 128; ++u) {
        int piece;
        if (P == 0) piece = u * 8 + warp;
        else if (P == 1) piece = ((u >> 3) * 16 + 2 * warp + ((u >> 2) & 1)) * 4 + (u & 3);
        else piece = warp * 128 + u;
        __stcs(base + (size_t)piece * 32 + lane, make_float4(1.f, 2.f, 3.f, (float)u));
    }
}

// Cooperative + software-pipelined model: the CTA's 256 threads prepare 256 inputs (thread = input, tables shared by the
// CTA, double-buffered), phase B walks the 1024 pieces in address order with the 8 warps on 8 consecutive pieces (P = 0's
// order); the next batch's phase A is spread over the store loop; one __syncthreads per 512 KiB.
template <int NB, int PAUSE>
__global__ void __launch_bounds__(256) k_coop_sp(float4 *p, const float4 *in) {
    extern __shared__ __align__(16) unsigned char dynraw[];
    float4 (*us)[4][256] = reinterpret_cast<float4 (*)[4][256]>(dynraw);                       // [2][4][256]
    float2 (*as)[4][257] = reinterpret_cast<float2 (*)[4][257]>(dynraw + 2 * 4 * 256 * 16);     // [2][4][257]
    float2 (*bs)[8][257] = reinterpret_cast<float2 (*)[8][257]>(dynraw + 2 * 4 * 256 * 16 + 2 * 4 * 257 * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t xt = (uint64_t)blockIdx.x * (256 * NB);
    float4 mine = in[(xt + threadIdx.x) & 0xfffff];
    float a = mine.x;
#pragma unroll 1
    for (int k = 0; k < PAUSE; ++k) a = a * 1.0001f + 0.5f;
    auto park = [&](int buf, float av, float4 m) {
#pragma unroll
        for (int t = 0; t < 4; ++t) us[buf][t][threadIdx.x] = make_float4(av, m.y, m.z, (float)t);
#pragma unroll
        for (int t = 0; t < 4; ++t) as[buf][t][threadIdx.x] = make_float2(av, m.y + t);
#pragma unroll
        for (int t = 0; t < 8; ++t) bs[buf][t][threadIdx.x] = make_float2(av, m.z + t);
    };
    park(0, a, mine);
    __syncthreads();
    float acc = 0.f;
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
        const int cur = b & 1;
        const uint64_t x0 = xt + (uint64_t)b * 256;
        const bool more = b + 1 < NB;
        float4 nx = more ? in[(x0 + 256 + threadIdx.x) & 0xfffff] : make_float4(0, 0, 0, 0);
        float an = nx.x;
#pragma unroll 8
        for (int k = 0; k < 128; ++k) {
            const int g = k * 8 + warp, i = g >> 2, t = g & 3;
            const float2 fa = as[cur][lane & 3][i], fb = bs[cur][lane >> 2][i];
            const float4 u = us[cur][t][i];
            const float lr = fa.x * fb.x - fa.y * fb.y, li = fa.x * fb.y + fa.y * fb.x;
            __stcs(p + ((x0 + i) << 7) + (uint64_t)t * 32 + lane,
                   make_float4(lr * u.x - li * u.y, lr * u.y + li * u.x, lr * u.z - li * u.w, lr * u.w + li * u.z));
#pragma unroll
            for (int q = 0; q < (PAUSE + 127) / 128; ++q) an = an * 1.0001f + 0.5f;
        }
        if (more) park(cur ^ 1, an, nx);
        __syncthreads();
        acc += an;
    }
    if (acc == 12345.678f) p[0] = make_float4(acc, 0, 0, 0);
}

template <int P> void run_pat(float4 *p, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)(nvec / (1024 * 32));
    const int dyn = per_sm >= 8 ? 0 : (227 * 1024) / per_sm - 2048;
    CK(cudaFuncSetAttribute(k_pat<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pat<P>, 256, dyn));
    float ms = timeit([&] { k_pat<P><<<grid, 256, dyn>>>(p); });
    printf("bare pattern %d, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", P, occ, ms, 16.0 * nvec / ms / 1e6);
    fflush(stdout);
}

template <int NB, int PAUSE> void run_coop_sp(float4 *p, const float4 *in, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)((nvec >> 7) / (256 * NB));
    const int need = 2 * 4 * 256 * 16 + 2 * 4 * 257 * 8 + 2 * 8 * 257 * 8;
    const int dyn = std::max(need, (227 * 1024) / per_sm - 2048);
    CK(cudaFuncSetAttribute(k_coop_sp<NB, PAUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_coop_sp<NB, PAUSE>, 256, dyn));
    float ms = timeit([&] { k_coop_sp<NB, PAUSE><<<grid, 256, dyn>>>(p, in); });
    printf("coop_sp %2d x 512 KiB per CTA, PAUSE=%4d, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", NB, PAUSE, occ, ms, 16.0 * nvec / ms / 1e6);
    fflush(stdout);
}

// Is it the READS?  k_low_sp with the per-batch input load (a) as it is, (b) replaced by arithmetic (no global load at
// all), (c) all of the CTA's loads issued once at the start (registers).
template <int NB, int PAUSE, int LOADS>
__global__ void __launch_bounds__(256) k_low_rd(float4 *p, const float4 *in) {
    __shared__ float4 sh[2][8][32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t xt = (uint64_t)blockIdx.x * (256 * NB);
    auto in_of = [&](int b) -> float4 {
        if (LOADS == 0) return make_float4(1.f + lane, 2.f + b, 3.f + warp, 4.f);
        return in[(xt + (uint64_t)b * 256 + 2u * warp + (uint64_t)(lane >> 1) * 16 + (lane & 1)) & 0xfffff];
    };
    float4 all[NB];
    if (LOADS == 2) {
#pragma unroll
        for (int b = 0; b < NB; ++b) all[b] = in_of(b);
    }
    float4 mine = LOADS == 2 ? all[0] : in_of(0);
    float a = mine.x;
#pragma unroll 1
    for (int k = 0; k < PAUSE; ++k) a = a * 1.0001f + 0.5f;
#pragma unroll
    for (int s = 0; s < 4; ++s) sh[0][warp][s * 32 + lane] = make_float4(a, mine.y, mine.z, (float)s);
    __syncwarp();
    float acc = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int cur = b & 1;
        const uint64_t x0 = xt + (uint64_t)b * 256 + 2u * warp;
        const bool more = b + 1 < NB;
        float4 nx = more ? (LOADS == 2 ? all[(b + 1) % NB] : in_of(b + 1)) : make_float4(0, 0, 0, 0);
        float an = nx.x;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint64_t x = x0 + (uint64_t)(i >> 1) * 16 + (i & 1);
            float4 u[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) u[s] = sh[cur][warp][s * 32 + i];
            float l = 1.f + 1e-3f * lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) l = l * 1.0001f + u[k & 3].x;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                __stcs(p + (x << 7) + (uint64_t)s * 32 + lane, make_float4(l * u[s].x, l * u[s].y, l * u[s].z, l * u[s].w));
#pragma unroll
            for (int k = 0; k < (PAUSE + 31) / 32; ++k) an = an * 1.0001f + 0.5f;
        }
        if (more) {
#pragma unroll
            for (int s = 0; s < 4; ++s) sh[cur ^ 1][warp][s * 32 + lane] = make_float4(an, nx.y, nx.z, (float)s);
        }
        __syncwarp();
        acc += an;
    }
    if (acc == 12345.678f) p[0] = make_float4(acc, 0, 0, 0);
}

template <int NB, int PAUSE, int LOADS>
void run_low_rd(float4 *p, const float4 *in, uint64_t nvec, int per_sm) {
    const unsigned grid = (unsigned)((nvec >> 7) / (256 * NB));
    const int dyn = per_sm <= 0 ? 0 : std::max(0, (227 * 1024) / per_sm - 1024 - 33 * 1024);
    CK(cudaFuncSetAttribute(k_low_rd<NB, PAUSE, LOADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_low_rd<NB, PAUSE, LOADS>, 256, dyn));
    float ms = timeit([&] { k_low_rd<NB, PAUSE, LOADS><<<grid, 256, dyn>>>(p, in); });
    printf("low_rd %2d batches per warp, PAUSE=%4d, loads: %s, %d CTAs per SM   %8.3f ms  %8.1f GB/s\n", NB, PAUSE,
           LOADS == 0 ? "none        " : LOADS == 1 ? "per batch   " : "all at start", occ, ms, 16.0 * nvec / ms / 1e6);
    fflush(stdout);
}

void earlier(float4 *p, const float4 *in, uint64_t nvec, uint64_t bytes) {
    float ms;
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
    // occupancy: the same long-lived writers with fewer resident CTAs per SM (dynamic shared memory as the limiter)
    for (int per_sm : {8, 6, 5, 4, 3, 2, 1}) {
        const int dyn = per_sm >= 8 ? 0 : (227 * 1024) / per_sm - 1024 - 1024;
        CK(cudaFuncSetAttribute(k_seq_n<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        ms = timeit([&] { k_seq_n<128><<<ctas, 256, dyn>>>(p, nvec); });
        int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_seq_n<128>, 256, dyn));
        printf("seq_n 512 KiB per CTA, %d CTAs per SM                 %8.3f ms  %8.1f GB/s\n", occ, ms, bytes / ms / 1e6);
        const int dyn2 = std::max(0, (227 * 1024) / per_sm - 1024 - 17 * 1024);
        CK(cudaFuncSetAttribute(k_low_gang<4, 8, 300>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn2));
        ms = timeit([&] { k_low_gang<4, 8, 300><<<lctas, 256, dyn2>>>(p, in, 0); });
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_low_gang<4, 8, 300>, 256, dyn2));
        printf("low + LDS + multiplies + phase A, %d CTAs per SM      %8.3f ms  %8.1f GB/s\n", occ, ms, bytes / ms / 1e6);
    }
    for (int per_sm : {1, 2, 3}) {
        run_spec<10, 4, 8, 8, 300>(p, in, nvec, per_sm);
        run_spec<10, 2, 8, 8, 300>(p, in, nvec, per_sm);
        run_spec<10, 4, 8, 8, 1200>(p, in, nvec, per_sm);
        run_spec<10, 4, 4, 8, 300>(p, in, nvec, per_sm);
        run_spec<10, 4, 16, 8, 300>(p, in, nvec, per_sm);
        run_spec<8, 4, 8, 8, 300>(p, in, nvec, per_sm);
        run_spec<12, 4, 8, 8, 300>(p, in, nvec, per_sm);
    }
}

void tma_runs(float4 *p, const float4 *in, uint64_t nvec) {
    for (int per_sm : {0, 4, 2, 1}) {
        run_tma_cta<16384, 2, 32>(p, nvec, per_sm);
        run_tma_cta<32768, 2, 16>(p, nvec, per_sm);
        run_tma_cta<32768, 2, 1>(p, nvec, per_sm);
        run_tma_cta<65536, 2, 8>(p, nvec, per_sm);
        run_tma_cta<8192, 4, 64>(p, nvec, per_sm);
    }
    for (int per_sm : {0, 3, 2}) {
        run_tma_low<1, 2, 0, 0, 0>(p, in, nvec, per_sm);
        run_tma_low<2, 2, 0, 0, 0>(p, in, nvec, per_sm);
        run_tma_low<1, 2, 4, 8, 300>(p, in, nvec, per_sm);
        run_tma_low<2, 2, 4, 8, 300>(p, in, nvec, per_sm);
        run_tma_low<1, 4, 4, 8, 300>(p, in, nvec, per_sm);
    }
}

void ring_runs(float4 *p, const float4 *in, uint64_t nvec) {
    for (int per_sm : {1, 2, 3}) {
        run_ring<8, 2, 0, 0, 0>(p, in, nvec, per_sm);
        run_ring<8, 3, 0, 0, 0>(p, in, nvec, per_sm);
        run_ring<8, 2, 4, 8, 300>(p, in, nvec, per_sm);
        run_ring<8, 3, 4, 8, 300>(p, in, nvec, per_sm);
        run_ring<10, 2, 4, 8, 300>(p, in, nvec, per_sm);
        run_ring<10, 3, 4, 8, 300>(p, in, nvec, per_sm);
        run_ring<12, 3, 4, 8, 300>(p, in, nvec, per_sm);
        run_ring<10, 3, 4, 8, 1200>(p, in, nvec, per_sm);
    }
}

void sp_runs(float4 *p, const float4 *in, uint64_t nvec) {
    for (int per_sm : {1, 2, 3, 4, 6}) {
        run_low_sp<4, 300>(p, in, nvec, per_sm);
        run_low_sp<8, 300>(p, in, nvec, per_sm);
        run_low_sp<16, 300>(p, in, nvec, per_sm);
        run_low_sp<8, 1200>(p, in, nvec, per_sm);
        run_low_sp<16, 1200>(p, in, nvec, per_sm);
    }
}

void pat_runs(float4 *p, const float4 *in, uint64_t nvec) {
    for (int per_sm : {8, 4, 2, 1}) {
        run_pat<0>(p, nvec, per_sm);
        run_pat<1>(p, nvec, per_sm);
        run_pat<2>(p, nvec, per_sm);
    }
    for (int per_sm : {1, 2, 3}) {
        run_coop_sp<1, 300>(p, in, nvec, per_sm);
        run_coop_sp<4, 300>(p, in, nvec, per_sm);
        run_coop_sp<16, 300>(p, in, nvec, per_sm);
        run_coop_sp<4, 1200>(p, in, nvec, per_sm);
    }
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
    (void)ctas;
    if (argc > 2) {                                            // ./membench5 GiB all: the earlier experiments too
        earlier(p, in, nvec, bytes);
    }
    if (argc > 3) {                                            // ./membench5 GiB all tma: the bulk-store writers
        tma_runs(p, in, nvec);
    }
    if (argc > 4) ring_runs(p, in, nvec);                      // ./membench5 GiB all tma ring
    if (argc > 5) sp_runs(p, in, nvec);                        // ./membench5 GiB all tma ring sp
    if (argc > 6) pat_runs(p, in, nvec);                       // ./membench5 GiB all tma ring sp pat
    for (int per_sm : {1, 2, 3}) {
        run_low_rd<4, 300, 1>(p, in, nvec, per_sm);
        run_low_rd<4, 300, 0>(p, in, nvec, per_sm);
        run_low_rd<4, 300, 2>(p, in, nvec, per_sm);
        run_low_rd<8, 300, 1>(p, in, nvec, per_sm);
        run_low_rd<8, 300, 0>(p, in, nvec, per_sm);
        run_low_rd<8, 300, 2>(p, in, nvec, per_sm);
        run_low_rd<1, 300, 1>(p, in, nvec, per_sm);
        run_low_rd<1, 300, 0>(p, in, nvec, per_sm);
    }
    return 0;
}
