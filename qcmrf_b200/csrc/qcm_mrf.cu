// Exact MRF inference by enumeration on the GPU: what the reference gets from the proprietary
// `kiopto_native` ("px") module in /root/reference/eval.py:84-93 --
//     lnZ = px.infer(b, task='partition');  p[xid] = exp(px.logpot(b, xid) - lnZ)
// -- for binary variables: energy(x) = sum_C w[off_C + y_C(x)], state id xid with x_0 as its most
// significant bit (eval.py:100-101), weights clique-major in itertools.product order.
// This is SURVEY.md App. E.3 (iii), "one diagonal write of 2^n entries", as a service for the evaluation
// scripts: eval.py's ground truth then scales to the sizes the simulator handles (n ~ 26-30) instead of
// stopping where a Python loop over 2^n states does.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "qcmrf_b200.h"

namespace {

constexpr int kMrfThreads = 256;
constexpr int kMrfMaxCliques = 128;
constexpr int kMrfMaxClique = 10;

struct MrfArgs {
    int32_t n, n_cliques;
    int32_t off[kMrfMaxCliques];                 // start of clique c's 2^m weights
    int8_t m[kMrfMaxCliques];
    int8_t bit[kMrfMaxCliques][kMrfMaxClique];   // xid bit feeding weight-index bit j of clique c
    int32_t dim;
};

__device__ __forceinline__ double mrf_energy(const MrfArgs &a, const double *w, uint64_t x) {
    double e = 0.0;
    for (int c = 0; c < a.n_cliques; ++c) {
        uint32_t y = 0;
        for (int j = 0; j < a.m[c]; ++j) y |= (uint32_t)((x >> a.bit[c][j]) & 1ull) << j;
        e += w[a.off[c] + y];
    }
    return e;
}

// mode 0: per-block max of the energies; mode 1: per-block sum of exp(e - shift) (+ optional pmf / energies)
// Fixed association: thread-strided partials, xor-shuffle tree, warps in order => deterministic for a fixed grid.
__global__ void __launch_bounds__(kMrfThreads) k_mrf(const __grid_constant__ MrfArgs a, const double *__restrict__ weights, int mode,
                                                     double shift, double *partial, double *out, int out_kind) {
    extern __shared__ double w[];
    for (int i = threadIdx.x; i < a.dim; i += blockDim.x) w[i] = weights[i];
    __syncthreads();
    const uint64_t count = 1ull << a.n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = mode == 0 ? -INFINITY : 0.0;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < count; x += stride) {
        const double e = mrf_energy(a, w, x);
        if (mode == 0) {
            acc = fmax(acc, e);
        } else {
            const double t = exp(e - shift);
            acc += t;
            if (out) out[x] = out_kind == 0 ? e : t;     // energies, or exp(e - lnZ) when shift == lnZ
        }
    }
    __shared__ double ws[kMrfThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, acc, o);
        acc = mode == 0 ? fmax(acc, t) : acc + t;
    }
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = ws[0];
        for (int i = 1; i < kMrfThreads / 32; ++i) t = mode == 0 ? fmax(t, ws[i]) : t + ws[i];
        partial[blockIdx.x] = t;
    }
}

thread_local char g_mrf_err[256];

}  // namespace

extern "C" {

const char *qcm_mrf_last_error(void) { return g_mrf_err; }

int qcm_mrf_exact(int device, int n, int n_cliques, const int32_t *clique_size, const int32_t *clique_vars,
                  const double *weights, double *log_z_out, double *pmf_out, double *energies_out, double *device_ms_out) {
    auto fail = [&](int code, const char *msg) { snprintf(g_mrf_err, sizeof g_mrf_err, "%s", msg); return code; };
    if (!clique_size || !clique_vars || !weights || !log_z_out) return fail(QCM_ERR_INVALID, "NULL argument");
    if (n < 1 || n > 34) return fail(QCM_ERR_INVALID, "n out of range [1, 34]");
    if (n_cliques < 1 || n_cliques > kMrfMaxCliques) return fail(QCM_ERR_INVALID, "number of cliques out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(QCM_ERR_NO_DEVICE, "no CUDA device: qcmrf_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(QCM_ERR_INVALID, "device out of range");
    MrfArgs a{};
    a.n = n;
    a.n_cliques = n_cliques;
    int dim = 0, pos = 0;
    for (int c = 0; c < n_cliques; ++c) {
        const int m = clique_size[c];
        if (m < 1 || m > kMrfMaxClique) return fail(QCM_ERR_INVALID, "clique size out of range [1, 10]");
        a.m[c] = (int8_t)m;
        a.off[c] = dim;
        for (int j = 0; j < m; ++j) {
            const int v = clique_vars[pos + j];
            if (v < 0 || v >= n) return fail(QCM_ERR_INVALID, "clique vertex out of range");
            // weight index y = sum_j x_{v_j} 2^(m-1-j) (itertools.product order); x_v is xid bit n-1-v
            a.bit[c][m - 1 - j] = (int8_t)(n - 1 - v);
        }
        pos += m;
        dim += 1 << m;
    }
    a.dim = dim;
    const size_t smem = (size_t)dim * sizeof(double);
    if (smem > 200 * 1024) return fail(QCM_ERR_UNSUPPORTED, "weights do not fit shared memory");
#define MRF_CUDA(call)                                                          \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            snprintf(g_mrf_err, sizeof g_mrf_err, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            cudaFree(d_w); cudaFree(d_part); cudaFree(d_out);                    \
            return e_ == cudaErrorMemoryAllocation ? QCM_ERR_NOMEM : QCM_ERR_CUDA; \
        }                                                                        \
    } while (0)
    double *d_w = nullptr, *d_part = nullptr, *d_out = nullptr;
    MRF_CUDA(cudaSetDevice(device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const uint64_t count = 1ull << n;
    const int blocks = (int)std::min<uint64_t>((count + kMrfThreads - 1) / kMrfThreads, (uint64_t)sms * 8);
    if (smem > 48 * 1024) MRF_CUDA(cudaFuncSetAttribute(k_mrf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MRF_CUDA(cudaMalloc(&d_w, smem));
    MRF_CUDA(cudaMalloc(&d_part, sizeof(double) * blocks));
    if (pmf_out || energies_out) MRF_CUDA(cudaMalloc(&d_out, sizeof(double) * count));
    MRF_CUDA(cudaMemcpy(d_w, weights, smem, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    MRF_CUDA(cudaEventCreate(&e0));
    MRF_CUDA(cudaEventCreate(&e1));
    std::vector<double> part(blocks);
    MRF_CUDA(cudaEventRecord(e0, 0));
    // pass 1: the largest energy (the shift of the log-sum-exp)
    k_mrf<<<blocks, kMrfThreads, smem>>>(a, d_w, 0, 0.0, d_part, nullptr, 0);
    MRF_CUDA(cudaGetLastError());
    MRF_CUDA(cudaMemcpy(part.data(), d_part, sizeof(double) * blocks, cudaMemcpyDeviceToHost));
    double emax = part[0];
    for (int i = 1; i < blocks; ++i) emax = std::max(emax, part[i]);
    // pass 2: sum exp(e - emax), block partials added in block order
    k_mrf<<<blocks, kMrfThreads, smem>>>(a, d_w, 1, emax, d_part, energies_out ? d_out : nullptr, 0);
    MRF_CUDA(cudaGetLastError());
    MRF_CUDA(cudaMemcpy(part.data(), d_part, sizeof(double) * blocks, cudaMemcpyDeviceToHost));
    double sum = 0.0;
    for (int i = 0; i < blocks; ++i) sum += part[i];
    const double log_z = emax + log(sum);
    *log_z_out = log_z;
    if (energies_out) MRF_CUDA(cudaMemcpy(energies_out, d_out, sizeof(double) * count, cudaMemcpyDeviceToHost));
    if (pmf_out) {
        // pass 3: p[xid] = exp(e - lnZ), written once
        k_mrf<<<blocks, kMrfThreads, smem>>>(a, d_w, 1, log_z, d_part, d_out, 1);
        MRF_CUDA(cudaGetLastError());
        MRF_CUDA(cudaMemcpy(pmf_out, d_out, sizeof(double) * count, cudaMemcpyDeviceToHost));
    }
    MRF_CUDA(cudaEventRecord(e1, 0));
    MRF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (device_ms_out) *device_ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_w);
    cudaFree(d_part);
    cudaFree(d_out);
#undef MRF_CUDA
    return QCM_OK;
}

}  // extern "C"
