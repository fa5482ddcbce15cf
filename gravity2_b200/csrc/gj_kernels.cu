// Generalised-Jaccard kernels (K1/K2/K3 of SURVEY.md §2.2) for sm_100a.
//
// Reference semantics: app/utils/similarity_matrix_constructor.py:167-177 computes, per
// genome pair, sum(min(a,b)) / sum(max(a,b)) over full float64 rows.  Here the all-pairs
// problem is a (min,+) "GEMM": 128x128 genome-pair tiles, PPHMM columns staged 32 at a
// time through a 3-stage cp.async shared-memory ring, 8x8 register tile per thread,
// one FMNMX (ALU pipe) + one FADD/FFMA (FMA pipe) per (pair, column) cell.
//   sum(max) is never formed per cell: sum(max) = sum(a) + sum(b) - sum(min).
// Precision: values are rounded once to fp32 (6e-8 relative), min is exact, accumulation
// is fp32 within a slab of <= GJ_SLAB columns and each slab total is promoted to fp64 in
// the pair-tile workspace (direct-min + slab promotion passes the 1e-6
// gate; the sum|a-b| identity does not).
#include "common.cuh"

#include <math.h>

#include <type_traits>

#include "gj_kernels.cuh"

// ---------------------------------------------------------------------------------------
// host launchers (C ABI; declared in include/gravity2_b200.h)
static size_t gj_smem_bytes() {
    return (size_t)(2 * GJ_STAGES * GV2_TILE * GJ_LDS + GJ_STAGES * GV2_BK) * sizeof(float);
}

extern "C" int64_t gv2_num_tiles(int64_t n) {
    int64_t nt = (n + GV2_TILE - 1) / GV2_TILE;
    return nt * (nt + 1) / 2;
}

extern "C" int gv2_dev_prepare_table(gv2_ctx* ctx, const double* src, int64_t n, int64_t k,
                                     int64_t ld_src, double threshold, int apply_threshold,
                                     float* dst, int64_t n_pad, int64_t k_pad, double* rowsum,
                                     uint8_t* nanflag) {
    if (!ctx || !src || !dst || n < 0 || k < 0 || n_pad < n || k_pad < k || (k_pad % GV2_BK) ||
        (n_pad % GV2_TILE))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_prepare_table: bad arguments");
    if (n_pad == 0) return GV2_OK;
    prepare_table_kernel<<<(unsigned)n_pad, 256, 0, ctx->stream>>>(src, n, k, ld_src, threshold,
                                                                    apply_threshold, dst, k_pad, rowsum, nanflag);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

extern "C" int gv2_dev_weights_to_float(gv2_ctx* ctx, const int32_t* w, int64_t nb, int64_t k,
                                        int64_t k_pad, float* out) {
    if (!ctx || !w || !out || nb <= 0) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_weights_to_float: bad arguments");
    dim3 grid((unsigned)((k_pad + 255) / 256), (unsigned)nb);
    weights_to_float_kernel<<<grid, 256, 0, ctx->stream>>>(w, k, k_pad, out);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

extern "C" int gv2_dev_weighted_rowsums(gv2_ctx* ctx, const float* tab, int64_t n, int64_t k_pad,
                                        const float* w, int64_t nb, double* out) {
    if (!ctx || !tab || !w || !out) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_weighted_rowsums: bad arguments");
    if (n == 0 || nb == 0) return GV2_OK;
    dim3 grid((unsigned)n, (unsigned)nb);
    weighted_rowsum_kernel<<<grid, 256, 0, ctx->stream>>>(tab, k_pad, w, k_pad, n, out);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

int gj_min_sums_launch(gv2_ctx* ctx, const float* tab, int64_t tab_stride, int64_t n_pad, int64_t k_pad,
                       int64_t tile_begin, int64_t tile_end, const float* w, int64_t nb, int ksplit,
                       double* ws) {
    if (!ctx || !tab || !ws || nb <= 0 || ksplit <= 0 || tile_end < tile_begin || (k_pad % GV2_BK) ||
        (n_pad % GV2_TILE))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_gj_min_sums: bad arguments");
    const int64_t ntl = tile_end - tile_begin;
    if (ntl == 0) return GV2_OK;
    if (tile_end > gv2_num_tiles(n_pad)) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_gj_min_sums: tile range");
    if (nb > 65535 || ksplit > 65535) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_gj_min_sums: nb/ksplit > 65535");
    size_t smem = gj_smem_bytes();
    if (!(ctx->smem_attr_done & GV2_ATTR_GJ)) {
        GV2_CUDA(ctx, cudaFuncSetAttribute(gj_min_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GV2_CUDA(ctx, cudaFuncSetAttribute(gj_min_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smem_attr_done |= GV2_ATTR_GJ;
    }
    dim3 grid((unsigned)ntl, (unsigned)ksplit, (unsigned)nb);
    const int nt = (int)(n_pad / GV2_TILE);
    const int kblocks = (int)(k_pad / GV2_BK);
    if (w)
        gj_min_tile_kernel<true><<<grid, GJ_THREADS, smem, ctx->stream>>>(tab, tab_stride, (int)k_pad, kblocks, nt, tile_begin, w, k_pad, ws);
    else
        gj_min_tile_kernel<false><<<grid, GJ_THREADS, smem, ctx->stream>>>(tab, tab_stride, (int)k_pad, kblocks, nt, tile_begin, nullptr, 0, ws);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

static GjComp to_comp(const gv2_gj_component* c);

// GOM signature tables of a job, fp64 [nb][n][g] -> unsigned fixed point [nb][n_pad][g_pad] (zero padded) for the FIXED fused
// tail: q = rint(x * 2^k), k the largest exponent for which every row sum of the launch still fits 32 bits
// (2^k * max row sum + g_pad / 2 < 2^32, k <= 30; GOM signatures are distance correlations in [0, 1], so k >= 24 for
// G <= 224 and typically 26-27).  The scale cancels in sum(min) / sum(max).  Negative entries count as 0; NaN entries are
// stored as 0 and flag their row (as in the fp32 image).  Row sums of the quantised values are exact integers, as doubles.
// scratch: 8 bytes of device memory (the launch's largest row sum)
int gom_quantize_launch(gv2_ctx* ctx, const double* tab64, int64_t n, int64_t g, int64_t nb, unsigned* dst, int64_t n_pad, int64_t g_pad,
                        double* rowsum, uint8_t* nanflag, unsigned long long* scratch) {
    if (nb <= 0 || n_pad == 0) return GV2_OK;
    if (nb > 65535) return gv2_fail(ctx, GV2_ERR_ARG, "gom_quantize_launch: nb > 65535");
    GV2_CUDA(ctx, cudaMemsetAsync(scratch, 0, sizeof(unsigned long long), ctx->stream));
    if (n > 0 && g > 0) {
        dim3 g1((unsigned)n, (unsigned)nb);
        gom_rowmax_kernel<<<g1, 128, 0, ctx->stream>>>(tab64, n, g, scratch);
        GV2_LAUNCH_CHECK(ctx);
    }
    dim3 grid((unsigned)n_pad, (unsigned)nb);
    gom_quantize_kernel<<<grid, 128, 0, ctx->stream>>>(tab64, n, g, scratch, dst, n_pad, g_pad, rowsum, nanflag);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

// Fused GOM pair sums + epilogue for replicates [0, nb): one table per replicate (tab_stride), all tiles, schemes G and PG.
// fixed != 0: `tab` holds the unsigned fixed-point image written by gom_quantize_launch (exact integer cell arithmetic).
int gj_fused_launch(gv2_ctx* ctx, const float* tab, int64_t tab_stride, int64_t n, int64_t n_pad, int64_t k, int64_t k_pad, int64_t nb,
                    const gv2_gj_component* comp_p, const gv2_gj_component* comp_g, int scheme, double power, int out_kind,
                    double* out, int64_t out_replicate_stride, int fixed, double p_scale) {
    if (scheme != GV2_SCHEME_G && scheme != GV2_SCHEME_PG) return gv2_fail(ctx, GV2_ERR_ARG, "gj_fused_launch: scheme %d", scheme);
    if (!tab || !comp_g || !comp_g->rowsums || !out || (k_pad % GV2_BK) || (n_pad % GV2_TILE) || k < 0 || k > k_pad ||
        (scheme == GV2_SCHEME_PG && (!comp_p || !comp_p->min_sums || !comp_p->rowsums || comp_p->ksplit != 1)))
        return gv2_fail(ctx, GV2_ERR_ARG, "gj_fused_launch: bad arguments");
    const int64_t ntl = gv2_num_tiles(n);
    if (ntl == 0 || nb <= 0) return GV2_OK;
    if (nb > 65535) return gv2_fail(ctx, GV2_ERR_ARG, "gj_fused_launch: nb > 65535");
    const size_t smem = gf_smem_bytes();
    static_assert(GF_STAGES * GF_STAGE_FLOATS * sizeof(float) >= 64 * GJ_STAGE_LD * sizeof(double), "transpose stage must fit the ring");
    typedef void (*fused_fn)(const float*, long long, int, int, int, int, const GjFuse, unsigned);
    static const fused_fn table[8] = {gj_fused_kernel<false, false, false>, gj_fused_kernel<false, false, true>,
                                      gj_fused_kernel<false, true, false>,  gj_fused_kernel<false, true, true>,
                                      gj_fused_kernel<true, false, false>,  gj_fused_kernel<true, false, true>,
                                      gj_fused_kernel<true, true, false>,   gj_fused_kernel<true, true, true>};
    if (!(ctx->smem_attr_done & GV2_ATTR_FUSED)) {
        for (int v = 0; v < 8; ++v) GV2_CUDA(ctx, cudaFuncSetAttribute((const void*)table[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smem_attr_done |= GV2_ATTR_FUSED;
    }
    GjFuse fz{};
    if (scheme == GV2_SCHEME_PG) fz.P = to_comp(comp_p);
    fz.g_rowsum = comp_g->rowsums; fz.g_nan = comp_g->nanflags;
    fz.g_stride_row = comp_g->replicate_stride_rows; fz.g_stride_nan = comp_g->replicate_stride_nan;
    fz.n = n; fz.scheme = scheme; fz.out_kind = out_kind; fz.pw = power; fz.out = out; fz.out_b_stride = out_replicate_stride;
    fz.p_scale = p_scale > 0.0 ? p_scale : 1.0;
    dim3 grid((unsigned)(2 * ntl), 1, (unsigned)nb);
    Gv2Timer timer(ctx, GV2_TIME_TAIL);
    timer.begin();
    const int k_groups = (int)((k + 3) / 4);                     // 4-column groups that hold real columns
    const int nkb = (int)((k + GV2_BK - 1) / GV2_BK);            // stages that hold real columns
    const int variant = (fixed ? 4 : 0) + (scheme == GV2_SCHEME_PG ? 2 : 0) + (power != 1.0 ? 1 : 0);
    table[variant]<<<grid, GJ_THREADS, smem, ctx->stream>>>(tab, tab_stride, (int)k_pad, nkb, k_groups, (int)(n_pad / GV2_TILE), fz, 1u);
    GV2_LAUNCH_CHECK(ctx);
    timer.end();
    return GV2_OK;
}

extern "C" int gv2_dev_gj_fused(gv2_ctx* ctx, const float* tab, int64_t tab_replicate_stride, int64_t n, int64_t n_pad,
                                int64_t k, int64_t k_pad, int64_t nb, const gv2_gj_component* comp_p,
                                const gv2_gj_component* comp_g, int scheme, double p, int out_kind, double* out,
                                int64_t out_replicate_stride) {
    if (!ctx) return GV2_ERR_ARG;
    return gj_fused_launch(ctx, tab, tab_replicate_stride, n, n_pad, k, k_pad, nb, comp_p, comp_g, scheme, p, out_kind, out,
                           out_replicate_stride, 0, 1.0);
}

extern "C" int gv2_dev_gj_min_sums(gv2_ctx* ctx, const float* tab, int64_t n_pad, int64_t k_pad,
                                   int64_t tile_begin, int64_t tile_end, const float* w,
                                   int64_t nb, int ksplit, double* ws) {
    if (!w && nb != 1) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_gj_min_sums: nb must be 1 without weights");
    return gj_min_sums_launch(ctx, tab, 0, n_pad, k_pad, tile_begin, tile_end, w, nb, ksplit, ws);
}

static GjComp to_comp(const gv2_gj_component* c) {
    GjComp r{};
    if (c) {
        r.ws = c->min_sums; r.rowsum = c->rowsums; r.nan = c->nanflags; r.ksplit = c->ksplit;
        r.b_stride_ws = c->replicate_stride_sums; r.b_stride_row = c->replicate_stride_rows;
        r.b_stride_nan = c->replicate_stride_nan;
    }
    return r;
}

extern "C" int gv2_dev_gj_epilogue(gv2_ctx* ctx, int64_t n, int64_t tile_begin, int64_t tile_end,
                                   int64_t nb, const gv2_gj_component* comp_p,
                                   const gv2_gj_component* comp_g, const double* spr, int scheme,
                                   double p, int out_kind, int packed, double* out,
                                   int64_t out_replicate_stride) {
    if (!ctx || !out || n < 0 || nb <= 0 || tile_end < tile_begin)
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_gj_epilogue: bad arguments");
    const bool need_p = scheme == GV2_SCHEME_P || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_PR || scheme == GV2_SCHEME_PL;
    const bool need_g = scheme == GV2_SCHEME_G || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_RG;
    const bool need_r = scheme == GV2_SCHEME_R || scheme == GV2_SCHEME_RG || scheme == GV2_SCHEME_PR ||
                        scheme == GV2_SCHEME_L || scheme == GV2_SCHEME_PL;
    if (scheme < GV2_SCHEME_P || scheme > GV2_SCHEME_PL)
        return gv2_fail(ctx, GV2_ERR_ARG, "unknown similarity scheme %d", scheme);
    if ((need_p && (!comp_p || !comp_p->min_sums || !comp_p->rowsums)) ||
        (need_g && (!comp_g || !comp_g->min_sums || !comp_g->rowsums)) || (need_r && !spr))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_gj_epilogue: scheme %d is missing an input", scheme);
    const int64_t ntl = tile_end - tile_begin;
    if (ntl == 0 || n == 0) return GV2_OK;
    const int nt = (int)((n + GV2_TILE - 1) / GV2_TILE);
    dim3 grid((unsigned)(ntl * 16), (unsigned)nb), block(32, 8);
    gj_epilogue_kernel<<<grid, block, 0, ctx->stream>>>(n, nt, tile_begin, ntl, to_comp(comp_p), to_comp(comp_g),
                                                        spr, scheme, p, out_kind, packed, out, out_replicate_stride);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

extern "C" int gv2_dev_unpack_tiles(gv2_ctx* ctx, int64_t n, int64_t tile_begin, int64_t tile_end,
                                    const double* packed, double* out) {
    if (!ctx || !packed || !out || tile_end < tile_begin) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_unpack_tiles: bad arguments");
    const int64_t ntl = tile_end - tile_begin;
    if (ntl == 0 || n == 0) return GV2_OK;
    const int nt = (int)((n + GV2_TILE - 1) / GV2_TILE);
    dim3 block(32, 8);
    unpack_tiles_kernel<<<(unsigned)(ntl * 16), block, 0, ctx->stream>>>(n, nt, tile_begin, packed, out);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}
