// Batched bootstrap Jaccard numerators on SPARSE signature tables (work ~ co-present PPHMMs), sm_100a.
//
// Same contract as boot_mma.cu: for every genome pair (i <= j) and every replicate b
//     S[b][i][j] = sum_c w[b][c] * min(T[i][c], T[j][c])
// (app/src/pl1_graphs.py:81-86 + app/utils/similarity_matrix_constructor.py:167, multiplicities
// w = bincount(idx)), written to the same fp64 pair-sum workspace.
//
// Real signature tables are ~99 % zeros (a genome hits a few hundred of 30 000 PPHMMs), and
// min(T[i][c], T[j][c]) is zero unless BOTH genomes hit c.  So per pair only the co-present columns
// matter: ~4 on average at the 1 % density of the BASELINE configs, against 30 000 columns in the
// dense contraction.  The kernel walks CSR rows.
//
// Exact and deterministic: values are 48-bit fixed point f = rint(x * 2^48 (1 - 2^-30) / max(T)) (a
// monotone map, so min commutes), held as a 32-bit high and a 16-bit low limb; the multiplicities are
// uint8; sum_c w_c * hi_c accumulates in 64-bit and sum_c w_c * lo_c in 32-bit integers (the launch
// checks sum_c w_c <= 2^(32-16)), so the sums do not depend on the order the lists are built in.  The
// only error is the input quantisation max(T) * 2^-49 per element (~1e-15 relative on S).
//
// Two kernels share the match logic (a 16 x 16 block of a 128 x 128 genome-pair tile per CTA):
//   build  : the 16 j-rows are scattered into a shared-memory table mask[c] = bit set of the j-rows
//            that hit column c (uint16 per column, 60 KB for 30 000 columns); their column lists are
//            copied to shared memory;
//   match  : the entries of all 16 i-rows, one per thread: mask[c] names the j-rows sharing the
//            column (usually none).  Hits are counted per pair, the pair lists are laid out back to
//            back (each padded to 4 entries), and a second pass writes (column, min value), the j value
//            found by binary search in the shared-memory copy of the row;
//   reduce : the threads become replicate lanes: per list entry a load of the transposed multiplicities
//            wT[c][replicates], IMAD.WIDE (high limb) and IMAD (low limb) into integer accumulators.
// * bs_lists_build_kernel + boot_sparse_lists_kernel (the default): the match does not depend on the
//   replicate, so it runs ONCE per job into global memory (2.4 GB of lists at the 10k x 30k config); the
//   per-launch kernel only streams the lists: no shared memory, no barriers, four replicates per lane
//   (one 4-byte multiplicity load serves four multiply-adds), 4 pairs x 4 replicates of accumulators per
//   thread, results leave as whole 32-byte sectors.
// * boot_sparse_kernel (lists too large for a tenth of the memory, or more than 65 536 columns): match
//   on the fly inside every launch, pool of 2048 entries in shared memory (a batch of i-rows that would
//   overflow it is halved, down to a slice of one row), column windows of 32 768, one replicate per lane,
//   the 16 pairs of an i-row leave as one whole 128-byte line.
// The row sums sum_c w_c x_ic are the diagonal pairs, exactly as in the tensor-core path.
#include "common.cuh"

#include <cub/cub.cuh>

#include "boot_sparse_kernels.cuh"

// ---------------------------------------------------------------------------------- host
struct BootSparse {
    long long* ptr = nullptr;
    int* col = nullptr;
    uint32_t* hi = nullptr;
    uint16_t* lo = nullptr;
    long long nnz = 0;
    unsigned long long cells = 0;     // co-present (pair incl. diagonal, column) cells
    double quantum = 0.0;             // value of one unit of the 48-bit fixed point
    size_t bytes = 0;
    BsLists lists;                    // precomputed pair lists (rowptr == nullptr: match on the fly)
    int64_t n_pad = 0, k_pad = 0;
};

void boot_sparse_free(BootSparse* s) {
    if (!s) return;
    cudaFree(s->ptr); cudaFree(s->col); cudaFree(s->hi); cudaFree(s->lo);
    cudaFree(s->lists.rowptr); cudaFree(s->lists.pend); cudaFree(s->lists.ent);
    delete s;
}

size_t boot_sparse_bytes(const BootSparse* s) { return s ? s->bytes : 0; }
int boot_sparse_has_lists(const BootSparse* s) { return s && s->lists.rowptr ? 1 : 0; }
int boot_sparse_rowsums(gv2_ctx* ctx, const BootSparse* s, int64_t n, double* out) {
    if (n == 0) return GV2_OK;
    BsCsr csr{s->ptr, s->col, s->hi, s->lo};
    bs_rowsums_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(csr, n, s->quantum, out);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}
long long boot_sparse_nnz(const BootSparse* s) { return s ? s->nnz : 0; }
unsigned long long boot_sparse_cells(const BootSparse* s) { return s ? s->cells : 0; }

// CSR (48-bit fixed point) of the float64 table [n][k] (row stride ld), padded with empty rows to n_pad;
// max_value = largest entry of the table (> 0, finite; the caller has checked there are no negative entries)
int boot_sparse_build(gv2_ctx* ctx, const double* tab, int64_t n, int64_t k, int64_t ld, int64_t n_pad, int64_t k_pad,
                      double max_value, BootSparse** out) {
    *out = nullptr;
    BootSparse* s = new BootSparse();
    long long* cnt = nullptr;
    int* colcnt = nullptr;
    unsigned long long* cells = nullptr;
    void* tmp = nullptr;
    auto fail = [&](cudaError_t e, const char* what) {
        cudaFree(cnt); cudaFree(colcnt); cudaFree(cells); cudaFree(tmp);
        boot_sparse_free(s);
        return gv2_fail(ctx, e == cudaErrorMemoryAllocation ? GV2_ERR_NOMEM : GV2_ERR_CUDA, "boot_sparse_build (%s): %s", what, cudaGetErrorString(e));
    };
    // x -> rint(x * scale) < 2^48: the (1 - 2^-30) keeps the largest entry below 2^48 after rounding
    const double scale = max_value > 0.0 ? 281474976710656.0 * (1.0 - 9.313225746154785e-10) / max_value : 0.0;
    s->quantum = scale > 0.0 ? 1.0 / scale : 0.0;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&cnt, (size_t)(n_pad + 1) * sizeof(long long))) != cudaSuccess) return fail(e, "counts");
    if ((e = cudaMalloc((void**)&s->ptr, (size_t)(n_pad + 1) * sizeof(long long))) != cudaSuccess) return fail(e, "row pointers");
    if ((e = cudaMemsetAsync(cnt, 0, (size_t)(n_pad + 1) * sizeof(long long), ctx->stream)) != cudaSuccess) return fail(e, "memset");
    bs_count_kernel<<<(unsigned)n_pad, 256, 0, ctx->stream>>>(tab, n, k, ld, scale, cnt);
    ctx->launches++;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, s->ptr, (int)(n_pad + 1), ctx->stream);
    if ((e = cudaMalloc(&tmp, tb ? tb : 1)) != cudaSuccess) return fail(e, "scan scratch");
    cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, s->ptr, (int)(n_pad + 1), ctx->stream);
    ctx->launches++;
    if ((e = cudaMemcpyAsync(&s->nnz, s->ptr + n_pad, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return fail(e, "nnz");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "scan");
    const size_t nz = (size_t)std::max<long long>(1, s->nnz);
    if ((e = cudaMalloc((void**)&s->col, nz * sizeof(int))) != cudaSuccess) return fail(e, "columns");
    if ((e = cudaMalloc((void**)&s->hi, nz * sizeof(uint32_t))) != cudaSuccess) return fail(e, "values");
    if ((e = cudaMalloc((void**)&s->lo, nz * sizeof(uint16_t))) != cudaSuccess) return fail(e, "values");
    if (n > 0) {
        bs_fill_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(tab, n, k, ld, scale, s->ptr, s->col, s->hi, s->lo);
        ctx->launches++;
    }
    if ((e = cudaMalloc((void**)&colcnt, (size_t)k_pad * sizeof(int))) != cudaSuccess) return fail(e, "column counts");
    if ((e = cudaMalloc((void**)&cells, sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cells");
    cudaMemsetAsync(colcnt, 0, (size_t)k_pad * sizeof(int), ctx->stream);
    cudaMemsetAsync(cells, 0, sizeof(unsigned long long), ctx->stream);
    if (s->nnz) {
        bs_colcount_kernel<<<(unsigned)std::min<long long>(4096, (s->nnz + 255) / 256), 256, 0, ctx->stream>>>(s->col, s->nnz, colcnt);
        ctx->launches++;
    }
    bs_cells_kernel<<<(unsigned)std::min<long long>(1024, (k_pad + 255) / 256), 256, 0, ctx->stream>>>(colcnt, k_pad, cells);
    ctx->launches++;
    if ((e = cudaMemcpyAsync(&s->cells, cells, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return fail(e, "cells");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "column statistics");
    cudaFree(cnt); cudaFree(colcnt); cudaFree(cells); cudaFree(tmp);
    s->bytes = (size_t)(n_pad + 1) * sizeof(long long) + nz * (sizeof(int) + sizeof(uint32_t) + sizeof(uint16_t));
    s->n_pad = n_pad; s->k_pad = k_pad;
    *out = s;
    return GV2_OK;
}

// Pair lists for all tiles: two passes (count, fill) around one scan.  Skipped (the launch then matches on the fly) when
// the table is wider than 65536 columns, the lists would take more than `budget` bytes, or GV2_SPARSE_LISTS=0.
int boot_sparse_build_lists(gv2_ctx* ctx, BootSparse* s, size_t budget) {
    const char* env = getenv("GV2_SPARSE_LISTS");
    if (env && env[0] == '0') return GV2_OK;
    const int64_t n_pad = s->n_pad, k_pad = s->k_pad;
    if (k_pad > 65536 || BS_LIST_SMEM(k_pad) > 200 * 1024) return GV2_OK;
    const int64_t nt = n_pad / GV2_TILE, ntl = nt * (nt + 1) / 2;
    if (ntl > 65535) return GV2_OK;
    // padded entries <= co-present cells + 3 per pair; a cheap upper bound decides before anything is allocated
    const double upper = ((double)s->cells + 1.5 * 0.5 * (double)n_pad * (double)n_pad) * 8.0 + (double)ntl * 64 * 16 * 72.0;
    if (upper > (double)budget) return GV2_OK;
    BsLists L;
    L.nrows = ntl * 64 * 16;
    long long* rowtot = nullptr;
    void* tmp = nullptr;
    auto fail = [&](cudaError_t e, const char* what) {
        cudaFree(rowtot); cudaFree(tmp); cudaFree(L.rowptr); cudaFree(L.pend); cudaFree(L.ent);
        if (e == cudaErrorMemoryAllocation) {       // the lists are an optimisation: without them every launch matches on the fly
            cudaGetLastError();
            return GV2_OK;
        }
        return gv2_fail(ctx, GV2_ERR_CUDA, "boot_sparse_build_lists (%s): %s", what, cudaGetErrorString(e));
    };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&rowtot, (size_t)(L.nrows + 1) * sizeof(long long))) != cudaSuccess) return fail(e, "row totals");
    if ((e = cudaMalloc((void**)&L.rowptr, (size_t)(L.nrows + 1) * sizeof(long long))) != cudaSuccess) return fail(e, "row pointers");
    if ((e = cudaMalloc((void**)&L.pend, (size_t)L.nrows * 16 * sizeof(uint32_t))) != cudaSuccess) return fail(e, "pair ends");
    if ((e = cudaMemsetAsync(rowtot + L.nrows, 0, sizeof(long long), ctx->stream)) != cudaSuccess) return fail(e, "memset");
    const size_t smem = BS_LIST_SMEM(k_pad);
    if ((e = cudaFuncSetAttribute(bs_lists_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return fail(e, "attr");
    if ((e = cudaFuncSetAttribute(bs_lists_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return fail(e, "attr");
    BsCsr csr{s->ptr, s->col, s->hi, s->lo};
    dim3 grid(64, (unsigned)ntl);
    bs_lists_build_kernel<false><<<grid, BS_THREADS, smem, ctx->stream>>>(csr, (int)nt, (int)k_pad, rowtot, nullptr, L.pend, nullptr);
    ctx->launches++;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, rowtot, L.rowptr, (int)(L.nrows + 1), ctx->stream);
    if ((e = cudaMalloc(&tmp, tb ? tb : 1)) != cudaSuccess) return fail(e, "scan scratch");
    cub::DeviceScan::ExclusiveSum(tmp, tb, rowtot, L.rowptr, (int)(L.nrows + 1), ctx->stream);
    ctx->launches++;
    if ((e = cudaMemcpyAsync(&L.nent, L.rowptr + L.nrows, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return fail(e, "total");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "count pass");
    if ((double)L.nent * 8.0 > (double)budget) {        // cannot happen given the bound above; kept as a guard
        cudaFree(rowtot); cudaFree(tmp); cudaFree(L.rowptr); cudaFree(L.pend);
        return GV2_OK;
    }
    if ((e = cudaMalloc((void**)&L.ent, (size_t)std::max<long long>(4, L.nent) * 8)) != cudaSuccess) return fail(e, "entries");
    if ((e = cudaMemsetAsync(L.ent, 0, (size_t)std::max<long long>(4, L.nent) * 8, ctx->stream)) != cudaSuccess) return fail(e, "entries");
    bs_lists_build_kernel<true><<<grid, BS_THREADS, smem, ctx->stream>>>(csr, (int)nt, (int)k_pad, nullptr, L.rowptr, L.pend, L.ent);
    ctx->launches++;
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "fill pass");
    cudaFree(rowtot); cudaFree(tmp);
    L.bytes = (size_t)(L.nrows + 1) * 8 + (size_t)L.nrows * 64 + (size_t)std::max<long long>(1, L.nent) * 8;
    s->lists = L;
    s->bytes += L.bytes;
    return GV2_OK;
}

static int bs_lanes(int64_t nb) { return nb <= 64 ? 64 : nb <= 128 ? 128 : 256; }
static int64_t bs_wld(int64_t nb) {
    const int64_t per = bs_lanes(nb);
    return (nb + per - 1) / per * per;
}
// the streaming (precomputed lists) kernel: four replicates per lane
// the streaming (precomputed lists) kernel: RW warps of 128 replicates per block
static int bss_rw(int64_t nb) { return nb <= 128 ? 1 : nb <= 256 ? 2 : nb <= 512 ? 4 : 8; }
static int64_t bsl_wld(int64_t nb) {
    const int64_t per = 128 * bss_rw(nb);
    return (nb + per - 1) / per * per;
}
size_t boot_sparse_scratch_bytes(int64_t k_pad, int64_t nb) { return (size_t)k_pad * std::max(bs_wld(nb), bsl_wld(nb)); }

// ws: fp64 [nb][ntl][128*128] (gj_ws_index layout), tiles [tile_begin, tile_end); w: int32 [nb][k] (device).
// *overflow_flag is set to 1 when a multiplicity exceeds 255 and to 2 when a replicate's multiplicities sum to
// more than 65536 (neither can happen for a bootstrap resample of k <= 65536 columns).
int boot_sparse_launch(gv2_ctx* ctx, const BootSparse* s, int64_t n_pad, int64_t k_pad, int64_t tile_begin,
                       int64_t tile_end, const int32_t* w, int64_t k, int64_t nb, void* wT_scratch, int* overflow_flag,
                       double* ws) {
    const int64_t ntl = tile_end - tile_begin;
    if (ntl <= 0 || nb <= 0) return GV2_OK;
    const bool use_lists = s->lists.rowptr != nullptr;
    const int rl = bs_lanes(nb);
    const bool pairlane = use_lists && nb <= 12;     // few replicates: lanes = pairs, compact multiplicity rows
    const int64_t w_ld = pairlane ? (nb <= 4 ? 4 : nb <= 8 ? 8 : 16) : use_lists ? bsl_wld(nb) : bs_wld(nb);
    uint8_t* wT = (uint8_t*)wT_scratch;
    dim3 gt((unsigned)((k_pad + 31) / 32), (unsigned)((w_ld + 31) / 32)), bt(32, 8);
    bs_weights_transpose_kernel<<<gt, bt, 0, ctx->stream>>>(w, nb, k, k_pad, w_ld, wT, overflow_flag);
    GV2_LAUNCH_CHECK(ctx);
    if (!(ctx->smem_attr_done & GV2_ATTR_SPARSE)) {
        const int mx = (int)bs_smem_bytes(BS_WIN);
        GV2_CUDA(ctx, cudaFuncSetAttribute(boot_sparse_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        GV2_CUDA(ctx, cudaFuncSetAttribute(boot_sparse_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        GV2_CUDA(ctx, cudaFuncSetAttribute(boot_sparse_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        ctx->smem_attr_done |= GV2_ATTR_SPARSE;
    }
    Gv2Timer timer(ctx, GV2_TIME_PAIRSUMS);
    timer.begin();
    const long long rep_stride_l = (long long)ntl * GV2_TILE * GV2_TILE;
    if (use_lists) {                                 // (the lists exist only when all tiles fit one grid: ntl <= 65535)
        const BsLists& L = s->lists;
        const int rwp = bss_rw(nb);
        const size_t ssm = (size_t)8 * BSS_WARP_BYTES;
        if (!(ctx->smem_attr_done & GV2_ATTR_STREAM)) {
            GV2_CUDA(ctx, cudaFuncSetAttribute(boot_sparse_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
            GV2_CUDA(ctx, cudaFuncSetAttribute(boot_sparse_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
            GV2_CUDA(ctx, cudaFuncSetAttribute(boot_sparse_stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
            GV2_CUDA(ctx, cudaFuncSetAttribute(boot_sparse_stream_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
            ctx->smem_attr_done |= GV2_ATTR_STREAM;
        }
        dim3 gl(64, (unsigned)ntl, (unsigned)std::max<int64_t>(1, w_ld / (128 * rwp)));
        const int ntt = (int)(n_pad / GV2_TILE);
        if (pairlane) {
            dim3 gp(64, (unsigned)ntl);
            if (nb <= 4)
                boot_sparse_pairlane_kernel<1><<<gp, BS_THREADS, 0, ctx->stream>>>(L.rowptr, L.pend, L.ent, ntt, tile_begin, wT, w_ld, (int)nb, rep_stride_l, s->quantum, ws);
            else if (nb <= 8)
                boot_sparse_pairlane_kernel<2><<<gp, BS_THREADS, 0, ctx->stream>>>(L.rowptr, L.pend, L.ent, ntt, tile_begin, wT, w_ld, (int)nb, rep_stride_l, s->quantum, ws);
            else
                boot_sparse_pairlane_kernel<3><<<gp, BS_THREADS, 0, ctx->stream>>>(L.rowptr, L.pend, L.ent, ntt, tile_begin, wT, w_ld, (int)nb, rep_stride_l, s->quantum, ws);
        } else if (rwp == 1)
            boot_sparse_stream_kernel<1><<<gl, BS_THREADS, ssm, ctx->stream>>>(L.rowptr, L.pend, L.ent, ntt, tile_begin, wT, w_ld, (int)nb, rep_stride_l, s->quantum, ws);
        else if (rwp == 2)
            boot_sparse_stream_kernel<2><<<gl, BS_THREADS, ssm, ctx->stream>>>(L.rowptr, L.pend, L.ent, ntt, tile_begin, wT, w_ld, (int)nb, rep_stride_l, s->quantum, ws);
        else if (rwp == 4)
            boot_sparse_stream_kernel<4><<<gl, BS_THREADS, ssm, ctx->stream>>>(L.rowptr, L.pend, L.ent, ntt, tile_begin, wT, w_ld, (int)nb, rep_stride_l, s->quantum, ws);
        else
            boot_sparse_stream_kernel<8><<<gl, BS_THREADS, ssm, ctx->stream>>>(L.rowptr, L.pend, L.ent, ntt, tile_begin, wT, w_ld, (int)nb, rep_stride_l, s->quantum, ws);
        GV2_LAUNCH_CHECK(ctx);
        timer.end();
        return GV2_OK;
    }
    BsCsr csr{s->ptr, s->col, s->hi, s->lo};
    const int nt = (int)(n_pad / GV2_TILE), nwin = (int)((k_pad + BS_WIN - 1) / BS_WIN);
    const int kwin = (int)std::min<int64_t>(k_pad, BS_WIN);            // k_pad is a multiple of 128
    const size_t smem = bs_smem_bytes(kwin);
    const long long rep_stride = (long long)ntl * GV2_TILE * GV2_TILE;
    // grid.y holds at most 65535 tiles: larger tile ranges (n > 46 000 genomes) go in several launches
    for (int64_t t0 = 0; t0 < ntl; t0 += 65535) {
        const int64_t tc = std::min<int64_t>(65535, ntl - t0);
        dim3 grid(64, (unsigned)tc, (unsigned)(w_ld / rl));
        double* wsc = ws + (size_t)t0 * (GV2_TILE * GV2_TILE);
        if (rl == 64)
            boot_sparse_kernel<64><<<grid, BS_THREADS, smem, ctx->stream>>>(csr, nt, tile_begin + t0, wT, w_ld, kwin, nwin, (int)nb, rep_stride, s->quantum, wsc);
        else if (rl == 128)
            boot_sparse_kernel<128><<<grid, BS_THREADS, smem, ctx->stream>>>(csr, nt, tile_begin + t0, wT, w_ld, kwin, nwin, (int)nb, rep_stride, s->quantum, wsc);
        else
            boot_sparse_kernel<256><<<grid, BS_THREADS, smem, ctx->stream>>>(csr, nt, tile_begin + t0, wT, w_ld, kwin, nwin, (int)nb, rep_stride, s->quantum, wsc);
        GV2_LAUNCH_CHECK(ctx);
    }
    timer.end();
    return GV2_OK;
}
