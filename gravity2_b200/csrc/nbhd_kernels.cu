// PPHMM-neighbourhood parity mode of SimilarityMat_Constructor (PphmmNeighbourhoodWeight != 0),
// app/utils/similarity_matrix_constructor.py:112-169, SURVEY.md quirk Q1 + Appendix A.
//
// The reference scales the flagged entries of a signature row IN PLACE every time the pair loop
// touches the row: negative-list entries by (1 - w), positive-list entries by (1 + w), then
// `row[row < threshold] = 0`.  When pair (i, j >= i) is evaluated row i has been through j + 2
// applications and row j through i + 1 (i == j: the one row through i + 2).  One application maps
// x -> (x * (1 - w)) * (1 + w) for an entry on both lists, so after k applications an entry holds
// x * f^k (f = its per-application factor) provided it never dropped below the threshold on the way;
// x * f^t is monotone in t (0 < f), so "never dropped" == "neither the first nor the k-th value is
// below the threshold".  Everything is float64; the pair sums sum(min), sum(max) are formed directly.
#include "common.cuh"

#include <algorithm>

#include "nbhd_kernels.cuh"

// Cleaned, optionally pre-multiplied PPHMM generalised-Jaccard matrix under the neighbourhood weights.
//   sig [n][p] float64 (host), cls [n][p] uint8 (host; bit 0 negative list, bit 1 positive list),
//   mul [n][n] float64 or NULL, out [n][n] float64 (host)
extern "C" int gv2_similarity_nbhd_host(gv2_ctx* ctx, const double* sig, int64_t n, int64_t p, int64_t ld_sig,
                                        const uint8_t* cls, double weight, double threshold, const double* mul,
                                        double* out) {
    if (!ctx || !sig || !cls || !out || n < 0 || p < 0) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_similarity_nbhd_host: bad arguments");
    if (!(weight > -1.0 && weight < 1.0))
        return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "gv2_similarity_nbhd_host: |PphmmNeighbourhoodWeight| must be < 1 (got %g)", weight);
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return GV2_OK;
    double *d_tab = nullptr, *d_mul = nullptr, *d_out = nullptr;
    uint8_t* d_cls = nullptr;
    auto cleanup = [&]() { cudaFree(d_tab); cudaFree(d_cls); cudaFree(d_mul); cudaFree(d_out); };
    cudaError_t e = cudaMalloc((void**)&d_tab, (size_t)std::max<int64_t>(1, n * p) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_cls, (size_t)std::max<int64_t>(1, n * p));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, (size_t)n * n * sizeof(double));
    if (e == cudaSuccess && mul) e = cudaMalloc((void**)&d_mul, (size_t)n * n * sizeof(double));
    if (e != cudaSuccess) { cleanup(); return gv2_fail(ctx, GV2_ERR_NOMEM, "gv2_similarity_nbhd_host: %s", cudaGetErrorString(e)); }
    if (p > 0) {
        e = cudaMemcpy2DAsync(d_tab, p * sizeof(double), sig, ld_sig * sizeof(double), p * sizeof(double), n,
                              cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cls, cls, (size_t)n * p, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess && mul) e = cudaMemcpyAsync(d_mul, mul, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const unsigned nt = (unsigned)((n + NB_T - 1) / NB_T);
        gj_nbhd_kernel<<<dim3(nt, nt), NB_T * NB_T, 0, ctx->stream>>>(d_tab, d_cls, n, p, p, weight, threshold, d_mul, d_out);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "gv2_similarity_nbhd_host: %s", cudaGetErrorString(e));
    return GV2_OK;
}
