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

#define NB_T 16        // pair tile edge: 16 x 16 pairs per CTA, one pair per thread
#define NB_KC 32       // columns per shared-memory stage

__device__ __forceinline__ double nbhd_value(double x, unsigned cls, double thr, double f1, double fk) {
    // cls: bit 0 = on the negative list, bit 1 = on the positive list (f1 / fk chosen by the caller)
    if (cls == 0) return (x < thr) ? 0.0 : x;
    const double x1 = x * f1, xk = x * fk;
    return (x1 < thr || xk < thr) ? 0.0 : xk;
}

// grid = (ceil(n/16) j-tiles, ceil(n/16) i-tiles); tiles below the diagonal exit.
// out [n][n]: cleaned GJ (NaN -> 0, negative -> 0, :201-205), times mul[i][j] when mul != NULL, mirrored.
__global__ void __launch_bounds__(NB_T* NB_T)
gj_nbhd_kernel(const double* __restrict__ tab, const uint8_t* __restrict__ cls, long long n, long long p, long long ld,
               double w, double thr, const double* __restrict__ mul, double* __restrict__ out) {
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bj < bi) return;
    __shared__ double sI[NB_T][NB_KC + 1], sJ[NB_T][NB_KC + 1];
    __shared__ uint8_t cI[NB_T][NB_KC], cJ[NB_T][NB_KC];
    const int tid = threadIdx.x, ti = tid / NB_T, tj = tid % NB_T;
    const long long i = (long long)bi * NB_T + ti, j = (long long)bj * NB_T + tj;
    // per-application factors by class, first application and the pair's application counts
    const double neg = 1.0 - w, pos = 1.0 + w;
    const double f1[4] = {1.0, neg, pos, neg * pos};
    // row i carries j + 2 applications, row j carries i + 1 (pairs below the diagonal of a diagonal tile are
    // computed with the roles swapped, so that (i, j) and (j, i) agree bit for bit)
    const long long lo = i < j ? i : j, hi = i < j ? j : i;
    const double k_lo = (double)(hi + 2), k_hi = (double)(lo + 1);
    double fa[4], fb[4];      // fa: factors of the row with the smaller index, fb: of the larger
    fa[0] = fb[0] = 1.0;
    for (int c = 1; c < 4; ++c) {
        fa[c] = pow(f1[c], k_lo);
        fb[c] = pow(f1[c], lo == hi ? k_lo : k_hi);
    }
    const bool swap = i > j;   // thread's "I" row is the larger index
    double smin = 0.0, smax = 0.0;
    for (long long c0 = 0; c0 < p; c0 += NB_KC) {
        for (int q = tid; q < 2 * NB_T * NB_KC; q += NB_T * NB_T) {
            const int which = q / (NB_T * NB_KC), r = (q / NB_KC) % NB_T, c = q % NB_KC;
            const long long row = (long long)(which ? bj : bi) * NB_T + r, col = c0 + c;
            double v = 0.0;
            uint8_t k = 0;
            if (row < n && col < p) { v = tab[row * ld + col]; k = cls[row * p + col] & 3; }
            if (which) { sJ[r][c] = v; cJ[r][c] = k; } else { sI[r][c] = v; cI[r][c] = k; }
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < NB_KC; ++c) {
            const unsigned ka = cI[ti][c], kb = cJ[tj][c];
            const double a = nbhd_value(sI[ti][c], ka, thr, f1[ka], swap ? fb[ka] : fa[ka]);
            const double b = nbhd_value(sJ[tj][c], kb, thr, f1[kb], swap ? fa[kb] : fb[kb]);
            // np.minimum / np.maximum propagate NaN
            smin += (a != a || b != b) ? (a + b) : fmin(a, b);
            smax += (a != a || b != b) ? (a + b) : fmax(a, b);
        }
        __syncthreads();
    }
    if (i >= n || j >= n) return;
    // `... / sum(max) if not sum(max) == 0 else 0` (:167-168); NaN and negatives are cleaned to 0 (:201-205)
    double g = (smax == 0.0) ? 0.0 : smin / smax;
    if (g != g || g < 0.0) g = 0.0;
    out[i * n + j] = mul ? g * mul[i * n + j] : g;
    if (bi != bj) out[j * n + i] = mul ? g * mul[j * n + i] : g;     // mirror (:169); diagonal tiles hold both threads
}

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
