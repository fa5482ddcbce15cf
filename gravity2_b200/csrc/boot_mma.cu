// Batched bootstrap Jaccard on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// What it computes: for every genome pair (i <= j) and every replicate b of a block,
//     S[b][i][j] = sum_c w[b][c] * min(T[i][c], T[j][c])
// i.e. the numerator of the generalised Jaccard of a column-resampled signature table
// (app/src/pl1_graphs.py:81-86 + app/utils/similarity_matrix_constructor.py:167): resampling P
// columns with replacement == integer column multiplicities w = bincount(idx).
//
// Why tensor cores: min(T_i, T_j) is not a contraction, but once the vector m_ij[c] = min(..) is
// formed, all replicates share it:  S[pair][b] = sum_c m_pair[c] * W[b][c]  is a GEMM with
// M = pairs, N = replicates, K = PPHMM columns.  CUDA cores produce the A operand (the min tile),
// the tensor cores do the pairs x replicates x columns contraction.
//
// Exactness: fp32 tensor-core accumulation truncates when it aligns small products to a large
// accumulator (measured here: 2.5e-6 relative after only 130 accumulation steps with a bf16
// hi/mid/lo split), so the contraction is done in INTEGERS instead.  The table is held as 32-bit
// fixed point x_fix = rint(x * (2^32-1)/max(T)) (min commutes with the monotone map), the four
// bytes of min(x_i, x_j) are four uint8 digit planes, the multiplicities are uint8, and
// tcgen05.mma.kind::i8 accumulates each plane exactly in int32 (digit <= 255, sum_c w_c = P, so
// |sum| <= 255 P << 2^31).  The planes are recombined in fp64: sum = q * sum_d 256^d * plane_d.
// The only error is the input quantisation q/2 = max(T) * 2^-33 per element.
//
// CTA = 128 pairs (8 i-rows x 16 j-rows of one 128x128 genome tile) x N <= 256 replicates x TWO digit planes;
// TMEM = 2 planes x 256 columns of int32 (all 512).  The four planes take two launches (low half-word, then
// high half-word, which adds 65536 x its sums to the workspace): the wider N amortises the A-operand reads of
// the tensor core and the producers' shared-memory traffic over twice the replicates (the SM's shared-memory
// data pipe, not the tensor pipe, is what bounds this kernel).  One pipeline stage = 128 PPHMM columns.
//   warps 0-15 producers: stream the 24 fixed-point table rows through a cp.async ring, form the
//              u32 min, byte-transpose into the four K-major SWIZZLE_128B digit tiles; afterwards
//              drain TMEM (epilogue)
//   warp 16    TMA: W tile [N rows x 128 columns] uint8, SWIZZLE_128B, one elected thread
//   warp 17    MMA issuer: 2 planes x 4 (K=32) tcgen05.mma (M128 N256) per stage, one elected thread
#include "common.cuh"

#include <cuda.h>
#include <math.h>
#include <string.h>

#include "boot_mma_kernels.cuh"

// ---------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

int boot_mma_available() {
    static int state = -1;
    if (state < 0) {
        const char* env = getenv("GV2_DISABLE_MMA");
        state = (env && env[0] == '1') ? 0 : (get_encode() ? 1 : 0);
    }
    return state;
}

// Largest entry (ignoring NaN) and "has a negative entry" of a float64 table.  The exact integer replicate paths
// need a non-negative table: min does not commute with the unsigned fixed-point map otherwise.
int boot_table_stats(gv2_ctx* ctx, const double* src, int64_t n, int64_t k, int64_t ld_src, double* max_value, int* has_negative) {
    unsigned long long* stats = nullptr;
    GV2_CUDA(ctx, cudaMalloc((void**)&stats, 2 * sizeof(unsigned long long)));
    GV2_CUDA(ctx, cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (n > 0 && k > 0) {
        table_stats_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(src, n, k, ld_src, stats);
        ctx->launches++;
    }
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h, stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(stats);
    if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "table stats: %s", cudaGetErrorString(e));
    memcpy(max_value, &h[0], sizeof(double));
    *has_negative = h[1] != 0;
    return GV2_OK;
}

// Builds the 32-bit fixed-point table of the tensor-core path: x -> rint(x * (2^32 - 1) / max_value), zero padded.
int boot_mma_prepare_table(gv2_ctx* ctx, const double* src, int64_t n, int64_t k, int64_t ld_src, uint32_t* dst,
                           int64_t n_pad, int64_t k_pad, double max_value, double* quantum) {
    const double scale = max_value > 0.0 ? 4294967295.0 / max_value : 0.0;
    *quantum = max_value > 0.0 ? max_value / 4294967295.0 : 0.0;
    table_to_fixed_kernel<<<(unsigned)n_pad, 256, 0, ctx->stream>>>(src, n, k, ld_src, scale, dst, k_pad);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

// ws must hold ALL tiles [0, ntiles) of every replicate (tile_begin == 0)
int boot_mma_rowsums(gv2_ctx* ctx, const double* ws, int64_t ntl, int64_t n, int64_t nb, double* out) {
    if (n == 0 || nb == 0) return GV2_OK;
    const int nt = (int)((n + GV2_TILE - 1) / GV2_TILE);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)nb);
    diag_rowsums_kernel<<<grid, 256, 0, ctx->stream>>>(ws, (long long)ntl * GV2_TILE * GV2_TILE, nt, n, out);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

static void bm_blocks(int64_t nb, int64_t* nblocks, int64_t* nblk) {
    *nblocks = (nb + BM_NMAX - 1) / BM_NMAX;
    *nblk = ((nb + *nblocks - 1) / *nblocks + 15) / 16 * 16;
}

size_t boot_mma_scratch_bytes(int64_t k_pad, int64_t nb) {
    int64_t nblocks, nblk;
    bm_blocks(nb, &nblocks, &nblk);
    return (size_t)(nblocks * nblk) * k_pad;
}

// tab: fixed point [n_pad][k_pad] (k_pad % 128 == 0); w: int32 [nb][k] multiplicities (device);
// ws: fp64 [nb][ntl][128*128] in the gj_ws_index layout, tiles [tile_begin, tile_end)
int boot_mma_launch(gv2_ctx* ctx, const uint32_t* tab, double quantum, int64_t n_pad, int64_t k_pad, int64_t tile_begin,
                    int64_t tile_end, const int32_t* w, int64_t k, int64_t nb, void* w8_scratch, int* overflow_flag,
                    double* ws) {
    if (k_pad % BM_KC) return gv2_fail(ctx, GV2_ERR_ARG, "boot_mma_launch: k_pad %% 128 != 0");
    const int64_t ntl = tile_end - tile_begin;
    if (ntl <= 0 || nb <= 0) return GV2_OK;
    PFN_encodeTiled enc = get_encode();
    if (!enc) return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
    int64_t nblocks, nblk;
    bm_blocks(nb, &nblocks, &nblk);
    const int64_t nb_pad = nblocks * nblk;
    uint8_t* wb = (uint8_t*)w8_scratch;                       // [nb_pad][k_pad]
    dim3 gw((unsigned)((k_pad + 255) / 256), (unsigned)nb_pad);
    weights_to_u8_kernel<<<gw, 256, 0, ctx->stream>>>(w, nb, k, k_pad, wb, overflow_flag);
    GV2_LAUNCH_CHECK(ctx);
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)nb_pad};
    cuuint64_t strides[1] = {(cuuint64_t)k_pad};
    cuuint32_t box[2] = {BM_KC, (cuuint32_t)nblk};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, wb, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gv2_fail(ctx, GV2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    if (!(ctx->smem_attr_done & GV2_ATTR_MMA)) {
        GV2_CUDA(ctx, cudaFuncSetAttribute(boot_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BM_SMEM));
        GV2_CUDA(ctx, cudaFuncSetAttribute(boot_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BM_SMEM));
        ctx->smem_attr_done |= GV2_ATTR_MMA;
    }
    if (ntl > 65535) return gv2_fail(ctx, GV2_ERR_ARG, "boot_mma_launch: more than 65535 genome tiles per launch");
    dim3 grid(128, (unsigned)ntl, (unsigned)nblocks);
    Gv2Timer timer(ctx, GV2_TIME_PAIRSUMS);
    timer.begin();
    // low half-word (digit planes 0, 1) stores, high half-word (planes 2, 3) accumulates on top
    boot_mma_kernel<false><<<grid, BM_THREADS, BM_SMEM, ctx->stream>>>(tab, (int)k_pad, (int)(k_pad / BM_KC), (int)(n_pad / GV2_TILE),
                                                                       tile_begin, map, (int)nb, (int)nblk,
                                                                       (long long)ntl * GV2_TILE * GV2_TILE, quantum, ws);
    GV2_LAUNCH_CHECK(ctx);
    boot_mma_kernel<true><<<grid, BM_THREADS, BM_SMEM, ctx->stream>>>(tab, (int)k_pad, (int)(k_pad / BM_KC), (int)(n_pad / GV2_TILE),
                                                                      tile_begin, map, (int)nb, (int)nblk,
                                                                      (long long)ntl * GV2_TILE * GV2_TILE, quantum, ws);
    GV2_LAUNCH_CHECK(ctx);
    timer.end();
    return GV2_OK;
}
