// Counter-based resample indices (device-RNG mode of the bootstrap, K5 of SURVEY.md §2.2).
//
// The reference draws `np.random.choice(range(P), P, replace=True)` from numpy's unseeded global
// MT19937 (app/src/pl1_graphs.py:81-82); that stream is reproduced on the host and arrives here
// as weights.  This file is the second mode: Philox4x32-10 (Salmon et al., SC'11) keyed by the
// seed, counter = (draw >> 2, replicate, 0, 0), index = mulhi(u32, P).  Its CPU restatement lives
// in oracle/philox_oracle.py and must agree bit for bit.
#include "common.cuh"

#include "philox_kernels.cuh"

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
