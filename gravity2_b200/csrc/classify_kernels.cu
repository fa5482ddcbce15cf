// Classification epilogue of pipeline II on a similarity matrix that stays in HBM:
//   * row maximum and first arg-maximum over a column block: the 1-nearest-neighbour proposal of
//     TaxonomicAssignmentProposerAndEvaluator (app/utils/classification_utils.py:63-68: np.max / np.argmax of
//     SimMat[-N_ucf:, :N_ref] along axis 1);
//   * gather of listed cells: the intra- / inter-class similarity lists of
//     PairwiseSimilarityScore_Cutoff_Dict_Constructor (classification_utils.py:10-43); the host computes WHICH cells
//     (list order and the np.random.choice subsample are replicate independent), the device only reads them.
#include "common.cuh"

#include "classify_kernels.cuh"

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
