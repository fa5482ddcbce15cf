// Counter-based resample indices (device-RNG mode of the bootstrap, K5 of SURVEY.md §2.2).
//
// The reference draws `np.random.choice(range(P), P, replace=True)` from numpy's unseeded global
// MT19937 (app/src/pl1_graphs.py:81-82); that stream is reproduced on the host and arrives here
// as weights.  This file is the second mode: Philox4x32-10 (Salmon et al., SC'11) keyed by the
// seed, counter = (draw >> 2, replicate, 0, 0), index = mulhi(u32, P).  Its CPU restatement lives
// in oracle/philox_oracle.py and must agree bit for bit.
#include "common.cuh"

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// grid = (ceil(ceil(p/4)/256), nb)
__global__ void philox_indices_kernel(uint64_t seed, long long b0, long long p, long long* __restrict__ idx,
                                      int* __restrict__ weights) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // counter = 4 draws
    const long long b = blockIdx.y;
    if (q * 4 >= p) return;
    uint32_t r[4];
    philox4x32_10((uint32_t)q, (uint32_t)(b0 + b), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int i = 0; i < 4; ++i) {
        const long long k = q * 4 + i;
        if (k >= p) break;
        const long long v = (long long)(((uint64_t)r[i] * (uint64_t)p) >> 32);
        if (idx) idx[b * p + k] = v;
        if (weights) atomicAdd(&weights[b * p + v], 1);
    }
}

extern "C" int gv2_dev_philox_weights(gv2_ctx* ctx, uint64_t seed, int64_t b0, int64_t nb, int64_t p, int32_t* weights) {
    if (!ctx || !weights || nb < 0 || p < 0 || p >= (1LL << 32)) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_philox_weights: bad arguments");
    if (nb == 0 || p == 0) return GV2_OK;
    GV2_CUDA(ctx, cudaMemsetAsync(weights, 0, (size_t)nb * p * sizeof(int32_t), ctx->stream));
    dim3 grid((unsigned)(((p + 3) / 4 + 255) / 256), (unsigned)nb);
    philox_indices_kernel<<<grid, 256, 0, ctx->stream>>>(seed, b0, p, nullptr, weights);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

extern "C" int gv2_philox_indices_host(gv2_ctx* ctx, uint64_t seed, int64_t nb, int64_t p, int64_t* out) {
    if (!ctx || !out || nb < 0 || p < 0 || p >= (1LL << 32)) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_philox_indices_host: bad arguments");
    if (nb == 0 || p == 0) return GV2_OK;
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    long long* d = nullptr;
    GV2_CUDA(ctx, cudaMalloc((void**)&d, (size_t)nb * p * sizeof(long long)));
    dim3 grid((unsigned)(((p + 3) / 4 + 255) / 256), (unsigned)nb);
    philox_indices_kernel<<<grid, 256, 0, ctx->stream>>>(seed, 0, p, d, nullptr);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, d, (size_t)nb * p * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "philox indices: %s", cudaGetErrorString(e));
    return GV2_OK;
}
