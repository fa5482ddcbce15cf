// Classification epilogue of pipeline II on a similarity matrix that stays in HBM:
//   * row maximum and first arg-maximum over a column block: the 1-nearest-neighbour proposal of
//     TaxonomicAssignmentProposerAndEvaluator (app/utils/classification_utils.py:63-68: np.max / np.argmax of
//     SimMat[-N_ucf:, :N_ref] along axis 1);
//   * gather of listed cells: the intra- / inter-class similarity lists of
//     PairwiseSimilarityScore_Cutoff_Dict_Constructor (classification_utils.py:10-43); the host computes WHICH cells
//     (list order and the np.random.choice subsample are replicate independent), the device only reads them.
#include "common.cuh"

// np.max / np.argmax semantics: a NaN wins (and the first NaN is the arg-maximum), else the first largest value
__device__ __forceinline__ bool cl_better(double v, long long i, double bv, long long bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (vn) return i < bi;
    return v > bv || (v == bv && i < bi);
}

// grid = rows r1 - r0, block = 256
__global__ void __launch_bounds__(256)
rowmax_block_kernel(const double* __restrict__ mat, long long ld, long long r0, long long c0, long long c1,
                    double* __restrict__ out_max, long long* __restrict__ out_arg) {
    const double* row = mat + (r0 + blockIdx.x) * ld;
    double bv = -INFINITY;
    long long bi = 0x7fffffffffffffffLL;
    for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
        const double v = row[c];
        if (cl_better(v, c, bv, bi)) { bv = v; bi = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (cl_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    __shared__ double sv[8];
    __shared__ long long si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (cl_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        out_max[blockIdx.x] = bv;
        out_arg[blockIdx.x] = bi - c0;               // index within the column block, as np.argmax of the slice
    }
}

__global__ void gather_cells_kernel(const double* __restrict__ mat, long long ld, const long long* __restrict__ ii,
                                    const long long* __restrict__ jj, long long count, double* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) out[k] = mat[ii[k] * ld + jj[k]];
}

extern "C" int gv2_dev_rowmax_block(gv2_ctx* ctx, const double* mat, int64_t ld, int64_t row_begin, int64_t row_end,
                                    int64_t col_begin, int64_t col_end, double* out_max, int64_t* out_arg) {
    if (!ctx || !mat || !out_max || !out_arg || row_end < row_begin || col_end <= col_begin)
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_rowmax_block: bad arguments (an empty column block has no maximum)");
    if (row_end == row_begin) return GV2_OK;
    rowmax_block_kernel<<<(unsigned)(row_end - row_begin), 256, 0, ctx->stream>>>(mat, ld, row_begin, col_begin, col_end, out_max,
                                                                                  (long long*)out_arg);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

extern "C" int gv2_dev_gather_cells(gv2_ctx* ctx, const double* mat, int64_t ld, const int64_t* rows, const int64_t* cols,
                                    int64_t count, double* out) {
    if (!ctx || !mat || count < 0 || (count && (!rows || !cols || !out))) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_gather_cells: bad arguments");
    if (count == 0) return GV2_OK;
    gather_cells_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(mat, ld, (const long long*)rows, (const long long*)cols,
                                                                                 count, out);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}
