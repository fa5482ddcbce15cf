// GOM signature scoring (K4 of SURVEY.md §2.2) for sm_100a: distance correlation between each
// genome's PPHMM location vector and each taxon's genomic-organisation model.
//
// Reference semantics (app/utils/parallel_gom_sig_generator.py:30-37, app/utils/dcor.py:4-62):
// for genome v and GOM g, R = {c : some member of g has loc != 0 at c, or v has loc != 0 at c};
// X = GOM[:, R].T (|R| points in m_g dims), Y = loc_v[R] (|R| points in 1 dim);
// dcor = sqrt(sum A.B)/n / sqrt( sqrt(sum A^2)/n * sqrt(sum B^2)/n ), A, B the double-centred
// Euclidean distance matrices; 0.0 unless both distance variances are > 0.
//
// Device algorithm (all fp64).  The centred sums are expanded into un-centred ones
//   sum A.B = S_ab - (2/n) sum_k a_k. b_k. + a.. b../n^2          
// and R is split into Z = R_g \ R_v (y = 0), I = R_g & R_v, E = R_v \ R_g (X = 0).  Then every
// term except three double sums over I x I is a product of per-GOM and per-genome aggregates:
//   a_cd = |X_c| for d in E, 0 inside E;  b_cd = |y_d| for c in Z, 0 inside Z.
// |I| is the handful of PPHMMs a genome shares with a taxon, so the |R|^2 work of the
// reference collapses to |I|^2 lookups into a per-GOM distance cache A_g = (|X_c - X_d|).
// A bootstrap column resample (pl1_graphs.py:81-100, virus_classification.py:295-308) is the
// same computation with integer point multiplicities w_c (n = sum w): every sum carries w_c
// (and w_d), so w_c = 0 removes a column.  Replicates are the SIMT lanes: one warp scores one
// (genome, GOM) for 32 replicates at once.
#include "common.cuh"

#include <cub/cub.cuh>
#include <math.h>

#include "gom.h"

#include "gom_score_kernels.cuh"

// ------------------------------------------------------------------------------ host side
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return gv2_fail(ctx, GV2_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,     \
                            cudaGetErrorString(e__));                                         \
    } while (0)

template <typename T>
static cudaError_t dalloc(T** p, size_t count, size_t* total) {
    size_t bytes = (count ? count : 1) * sizeof(T);
    if (total) *total += bytes;
    return cudaMalloc((void**)p, bytes);
}

void gom_db_free(GomDb* db) {
    if (!db) return;
    cudaFree(db->g_ptr); cudaFree(db->g_col); cudaFree(db->g_val); cudaFree(db->g_ord);
    cudaFree(db->mem_ptr); cudaFree(db->mem_row); cudaFree(db->pt_ptr); cudaFree(db->pt_col);
    cudaFree(db->pos); cudaFree(db->x_off); cudaFree(db->a_off); cudaFree(db->X); cudaFree(db->nrm); cudaFree(db->nrm2);
    cudaFree(db->A); cudaFree(db->wT); cudaFree(db->arT); cudaFree(db->rT);
    cudaFree(db->gagg); cudaFree(db->cpart); cudaFree(db->vagg); cudaFree(db->wT8); cudaFree(db->wG); cudaFree(db->flag); cudaFree(db->heavy_list);
    cudaFree(db->pr_ni); cudaFree(db->pr_ptr); cudaFree(db->rec);
    delete db;
}

// loc: device [n][p] (ld_loc); rowtab: device table the GOM member rows live in (ld_row);
// mem_row_h[M] rows of rowtab, GOM k owns mem_ptr_h[k]..mem_ptr_h[k+1].
int gom_db_build(gv2_ctx* ctx, const double* loc, int64_t n, int64_t p, int64_t ld_loc,
                 const double* rowtab, int64_t ld_row, const int64_t* mem_row_h,
                 const int32_t* mem_ptr_h, int64_t G, int flags, GomDb** out) {
    GomDb* db = new GomDb();
    db->n = n; db->p = p; db->G = G;
    cudaStream_t st = ctx->stream;
    const int64_t M = mem_ptr_h[G];
    if (G * p >= (int64_t)INT32_MAX) { delete db; return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "GOM index map G*P = %lld exceeds 2^31", (long long)(G * p)); }
    // ---- genome CSR
    int* cnt = nullptr;
    CK(dalloc(&cnt, n, nullptr));
    CK(dalloc(&db->g_ptr, n + 1, &db->bytes));
    if (n) { csr_count_kernel<<<(unsigned)n, 256, 0, st>>>(loc, n, p, ld_loc, cnt); ctx->launches++; }
    {
        void* tmp = nullptr; size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, db->g_ptr, (int)n, st);
        CK(cudaMalloc(&tmp, tb ? tb : 1));
        cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, db->g_ptr, (int)n, st);
        ctx->launches++;
        CK(cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    long long last_ptr = 0; int last_cnt = 0;
    if (n) {
        CK(cudaMemcpy(&last_ptr, db->g_ptr + (n - 1), sizeof(long long), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&last_cnt, cnt + (n - 1), sizeof(int), cudaMemcpyDeviceToHost));
    }
    db->nnz = last_ptr + last_cnt;
    CK(cudaMemcpy(db->g_ptr + n, &db->nnz, sizeof(long long), cudaMemcpyHostToDevice));
    cudaFree(cnt);
    CK(dalloc(&db->g_col, db->nnz, &db->bytes));
    CK(dalloc(&db->g_val, db->nnz, &db->bytes));
    if (n) { csr_fill_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(loc, n, p, ld_loc, db->g_ptr, db->g_col, db->g_val); ctx->launches++; }
    // ---- GOM members
    CK(dalloc(&db->mem_ptr, G + 1, &db->bytes));
    CK(dalloc(&db->mem_row, M, &db->bytes));
    CK(cudaMemcpy(db->mem_ptr, mem_ptr_h, (G + 1) * sizeof(int), cudaMemcpyHostToDevice));
    if (M) CK(cudaMemcpy(db->mem_row, mem_row_h, M * sizeof(long long), cudaMemcpyHostToDevice));
    int *flag = nullptr, *scan = nullptr;
    CK(dalloc(&flag, G * p, nullptr));
    CK(dalloc(&scan, G * p + 1, nullptr));
    CK(dalloc(&db->pos, G * p, &db->bytes));
    dim3 gridc((unsigned)((p + 255) / 256), (unsigned)G);
    std::vector<int> pt_ptr_h(G + 1, 0);
    if (G && p) {
        gom_colany_kernel<<<gridc, 256, 0, st>>>(rowtab, ld_row, p, db->mem_row, db->mem_ptr, (flags & GV2_GOM_ALL_COLUMNS) ? 1 : 0, flag);
        ctx->launches++;
        void* tmp = nullptr; size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, scan, (int)(G * p), st);
        CK(cudaMalloc(&tmp, tb ? tb : 1));
        cub::DeviceScan::ExclusiveSum(tmp, tb, flag, scan, (int)(G * p), st);
        ctx->launches++;
        CK(cudaStreamSynchronize(st));
        cudaFree(tmp);
        for (int64_t g = 0; g < G; ++g)
            CK(cudaMemcpy(&pt_ptr_h[g], scan + g * p, sizeof(int), cudaMemcpyDeviceToHost));
        int lastf = 0, lasts = 0;
        CK(cudaMemcpy(&lastf, flag + (G * p - 1), sizeof(int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&lasts, scan + (G * p - 1), sizeof(int), cudaMemcpyDeviceToHost));
        pt_ptr_h[G] = lasts + lastf;
    }
    db->npt = pt_ptr_h[G];
    std::vector<long long> x_off_h(G + 1, 0), a_off_h(G + 1, 0);
    int max_ng = 0;
    for (int64_t g = 0; g < G; ++g) {
        long long ng = pt_ptr_h[g + 1] - pt_ptr_h[g], mg = mem_ptr_h[g + 1] - mem_ptr_h[g];
        x_off_h[g + 1] = x_off_h[g] + ng * mg;
        a_off_h[g + 1] = a_off_h[g] + ng * ng;
        if (ng > max_ng) max_ng = (int)ng;
    }
    db->max_ng = max_ng;
    db->sum_ng2 = a_off_h[G];
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t need = (size_t)(a_off_h[G] + x_off_h[G]) * sizeof(double);
    if (need > free_b * 0.6) {
        cudaFree(flag); cudaFree(scan);
        int rc = gv2_fail(ctx, GV2_ERR_NOMEM, "GOM distance cache needs %.1f GB (sum n_g^2 doubles), %.1f GB free: "
                          "location tables this dense are outside the GOM kernel's design range", need / 1e9, free_b / 1e9);
        gom_db_free(db);
        return rc;
    }
    CK(dalloc(&db->pt_ptr, G + 1, &db->bytes));
    CK(dalloc(&db->x_off, G + 1, &db->bytes));
    CK(dalloc(&db->a_off, G + 1, &db->bytes));
    CK(cudaMemcpy(db->pt_ptr, pt_ptr_h.data(), (G + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db->x_off, x_off_h.data(), (G + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db->a_off, a_off_h.data(), (G + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    CK(dalloc(&db->pt_col, db->npt, &db->bytes));
    CK(dalloc(&db->nrm, db->npt, &db->bytes));
    CK(dalloc(&db->nrm2, db->npt, &db->bytes));
    CK(dalloc(&db->X, x_off_h[G], &db->bytes));
    CK(dalloc(&db->A, a_off_h[G], &db->bytes));
    if (G && p) {
        gom_fill_kernel<<<gridc, 256, 0, st>>>(rowtab, ld_row, p, db->mem_row, db->mem_ptr, flag, scan,
                                               db->x_off, db->pos, db->pt_col, db->X, db->nrm, db->nrm2);
        ctx->launches++;
    }
    if (max_ng) {
        dim3 g2((unsigned)((max_ng + 15) / 16), (unsigned)((max_ng + 15) / 16), (unsigned)G), b2(16, 16);
        gom_dist_cache_kernel<<<g2, b2, 0, st>>>(db->pt_ptr, db->mem_ptr, db->x_off, db->a_off, db->X, db->A);
        ctx->launches++;
    }
    // ---- per-chunk (32 replicates) scratch
    CK(dalloc(&db->wT, p * NBC, &db->bytes));
    CK(dalloc(&db->arT, (size_t)2 * db->npt * NBC, &db->bytes));      // two chunks of 32 replicates per pass
    CK(dalloc(&db->rT, (size_t)db->nnz * NBC, &db->bytes));
    CK(dalloc(&db->gagg, (size_t)G * GAGG * NBC, &db->bytes));
    {
        int max_mg = 0;
        for (int64_t g = 0; g < G; ++g) max_mg = std::max<int>(max_mg, mem_ptr_h[g + 1] - mem_ptr_h[g]);
        db->n_ctile = (max_mg + 7) / 8;
        CK(dalloc(&db->cpart, (size_t)G * std::max(1, db->n_ctile) * NBC, &db->bytes));
    }
    CK(dalloc(&db->vagg, (size_t)n * VAGG * NBC, &db->bytes));
    // intersection scratch: each (genome, GOM) warp needs 2 ints per hit of the genome
    int max_t = 0;
    {
        std::vector<long long> ptr_h(n + 1);
        if (n) CK(cudaMemcpy(ptr_h.data(), db->g_ptr, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
        for (int64_t v = 0; v < n; ++v) max_t = std::max<int>(max_t, (int)(ptr_h[v + 1] - ptr_h[v]));
    }
    db->max_t = max_t;
    CK(dalloc(&db->g_ord, db->nnz, &db->bytes));
    if (n && max_t) {
        const size_t vo = (size_t)max_t * 8;
        if (vo > 200 * 1024) { gom_db_free(db); return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "a genome has %d location hits", max_t); }
        CK(cudaFuncSetAttribute(genome_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(vo, 1024)));
        genome_order_kernel<<<(unsigned)n, 128, vo, st>>>(db->g_ptr, db->g_val, db->g_ord);
        ctx->launches++;
    }
    CK(dalloc(&db->wT8, 2 * p * NBC, &db->bytes));
    CK(dalloc(&db->wG, (size_t)2 * db->npt * NBC, &db->bytes));
    CK(cudaFuncSetAttribute(gom_rowsums_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ga_smem_bytes(1)));
    CK(cudaFuncSetAttribute(gom_rowsums_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ga_smem_bytes(2)));
    CK(dalloc(&db->flag, 2, &db->bytes));          // [0] weight overflow, [1] heavy-list length
    CK(cudaMemsetAsync(db->flag, 0, 2 * sizeof(int), st));
    // ---- pair records (replicate independent): intersection entries + M triangle of every (genome, GOM) pair, genome-major
    {
        const long long total_w = n * G;
        long long* pbytes = nullptr;
        // offsets: [0, total) light records (0 bytes for heavy pairs), [total, 2 total) heavy records, one scan over both
        if (2 * total_w + 1 >= (long long)INT32_MAX) { gom_db_free(db); cudaFree(flag); cudaFree(scan);
            return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "GOM pair table: %lld (genome, GOM) pairs exceed 2^30", (long long)total_w); }
        CK(dalloc(&pbytes, 2 * total_w + 1, nullptr));
        CK(dalloc(&db->pr_ni, total_w + 1, &db->bytes));
        CK(dalloc(&db->pr_ptr, 2 * total_w + 1, &db->bytes));
        CK(cudaMemsetAsync(pbytes, 0, (size_t)(2 * total_w + 1) * sizeof(long long), st));
        if (total_w) {
            gom_pair_count_kernel<<<(unsigned)((total_w + 7) / 8), 256, 0, st>>>(n, (int)G, p, db->g_ptr, db->g_col, db->pos, VG_LIGHT_CAP, db->pr_ni, pbytes);
            ctx->launches++;
        }
        void* tmp = nullptr; size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, pbytes, db->pr_ptr, (int)(2 * total_w + 1), st);
        CK(cudaMalloc(&tmp, tb ? tb : 1));
        cub::DeviceScan::ExclusiveSum(tmp, tb, pbytes, db->pr_ptr, (int)(2 * total_w + 1), st);
        ctx->launches++;
        long long rec_bytes = 0;
        CK(cudaMemcpyAsync(&rec_bytes, db->pr_ptr + 2 * total_w, sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        cudaFree(tmp); cudaFree(pbytes);
        size_t free_b2 = 0, total_b2 = 0;
        cudaMemGetInfo(&free_b2, &total_b2);
        if ((double)rec_bytes > 0.5 * (double)free_b2) {
            int rc = gv2_fail(ctx, GV2_ERR_NOMEM, "GOM pair records need %.1f GB (sum over (genome, GOM) pairs of |I|^2 / 2 doubles), %.1f GB free: "
                              "location tables this dense are outside the GOM kernel's design range", rec_bytes / 1e9, free_b2 / 1e9);
            gom_db_free(db);
            cudaFree(flag); cudaFree(scan);
            return rc;
        }
        CK(dalloc(&db->rec, (size_t)std::max<long long>(16, rec_bytes), &db->bytes));
        CK(dalloc(&db->heavy_list, (size_t)std::max<long long>(1, total_w), &db->bytes));
        if (total_w) {
            gom_pair_fill_kernel<<<(unsigned)((total_w + 7) / 8), 256, 0, st>>>(n, (int)G, p, db->g_ptr, db->g_col, db->g_val, db->pos, db->pt_ptr,
                                                                                 db->nrm, db->a_off, db->A, db->pr_ni, db->pr_ptr, VG_LIGHT_CAP, db->rec);
            ctx->launches++;
            gom_heavy_list_kernel<<<(unsigned)((total_w + 255) / 256), 256, 0, st>>>(total_w, db->pr_ni, VG_LIGHT_CAP, db->heavy_list, db->flag + 1);
            ctx->launches++;
            CK(cudaMemcpyAsync(&db->n_heavy, db->flag + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            unsigned long long* d_me = nullptr;
            CK(dalloc(&d_me, 2, nullptr));
            CK(cudaMemsetAsync(d_me, 0, 2 * sizeof(unsigned long long), st));
            gom_mentries_kernel<<<(unsigned)std::min<long long>(2048, (total_w + 255) / 256), 256, 0, st>>>(total_w, db->pr_ni, d_me);
            ctx->launches++;
            unsigned long long me[2] = {0, 0};
            CK(cudaMemcpyAsync(me, d_me, sizeof me, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            cudaFree(d_me);
            db->n_mentries = (long long)me[0];
            db->n_isect = (long long)me[1];
        }
    }
    const size_t vg_smem = (size_t)((max_t + 3) & ~3) * (VG_ENTRY_BYTES + 8 * sizeof(double)) + 128 + 3 * VG_NSUM * 32 * sizeof(double);
    if (vg_smem > 200 * 1024) {
        gom_db_free(db);
        return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "a genome has %d location hits; the GOM scoring kernels stage at most %d per genome "
                        "in shared memory", max_t, (int)(190 * 1024 / VG_ENTRY_BYTES));
    }
    db->vg_smem = vg_smem;
    CK(cudaFuncSetAttribute(gom_score_heavy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(vg_smem, 1024)));
    CK(cudaStreamSynchronize(st));
    cudaFree(flag); cudaFree(scan);
    CK(cudaGetLastError());
    *out = db;
    return GV2_OK;
}

// weights: device int32 [nb][p] or nullptr (nb must be 1); out: device [nb][n][G]
int gom_db_score(gv2_ctx* ctx, GomDb* db, const int32_t* weights, int64_t nb, double* out) {
    cudaStream_t st = ctx->stream;
    const int64_t n = db->n, p = db->p, G = db->G;
    if (n == 0 || G == 0) return GV2_OK;
    if (weights && !db->weights_validated) {       // the flag starts clean per call (jobs validate up front instead)
        cudaError_t e0 = cudaMemsetAsync(db->flag, 0, sizeof(int), st);
        if (e0 != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "gom_db_score: %s", cudaGetErrorString(e0));
    }
    // Two chunks of 32 replicates (the lanes of a warp) per pass: the GOM-side row sums of both go through ONE sweep of the
    // distance caches; everything else runs per chunk on that chunk's half of the scratch.
    for (int64_t b0 = 0; b0 < nb; b0 += 2 * NBC) {
        const int nh = nb - b0 > NBC ? 2 : 1;
        Gv2Timer timer(ctx, GV2_TIME_GOM);
        timer.begin();
        for (int h = 0; h < nh; ++h) {
            const int cb = (int)std::min<int64_t>(NBC, nb - b0 - h * NBC);
            if (p) {
                weights_transpose_kernel<<<(unsigned)((p * NBC + 255) / 256), 256, 0, st>>>(
                    weights ? weights + (b0 + h * NBC) * p : nullptr, p, cb, db->wT, db->wT8 + (size_t)h * p * NBC, db->flag);
                ctx->launches++;
            }
            if (db->max_ng) {
                gom_gather_weights_kernel<<<(unsigned)(((long long)db->npt * NBC + 255) / 256), 256, 0, st>>>(db->pt_col, db->npt, db->wT,
                                                                                                              db->wG + (size_t)h * db->npt * NBC);
                ctx->launches++;
            }
        }
        double* wG1 = db->wG + (size_t)db->npt * NBC;
        double* arT1 = db->arT + (size_t)db->npt * NBC;
        if (db->max_ng) {
            dim3 g1((unsigned)((db->max_ng + GA_ROWS - 1) / GA_ROWS), (unsigned)G), b1(32, 8);
            if (nh == 2) gom_rowsums_kernel<2><<<g1, b1, ga_smem_bytes(2), st>>>(db->pt_ptr, db->a_off, db->A, db->wG, wG1, db->arT, arT1);
            else gom_rowsums_kernel<1><<<g1, b1, ga_smem_bytes(1), st>>>(db->pt_ptr, db->a_off, db->A, db->wG, nullptr, db->arT, nullptr);
            ctx->launches++;
        }
        for (int h = 0; h < nh; ++h) {
            const int cb = (int)std::min<int64_t>(NBC, nb - b0 - h * NBC);
            const double* wG = db->wG + (size_t)h * db->npt * NBC;
            const double* arT = db->arT + (size_t)h * db->npt * NBC;
            const unsigned char* wT8 = db->wT8 + (size_t)h * p * NBC;
            double* outh = out + (b0 + h * NBC) * n * G;
            if (db->n_ctile) {
                dim3 gc((unsigned)G, (unsigned)db->n_ctile);
                gom_centroid_kernel<<<gc, 256, 0, st>>>(db->pt_ptr, db->mem_ptr, db->x_off, db->X, wG, db->n_ctile, db->cpart);
                ctx->launches++;
            }
            gom_aggregates_kernel<<<(unsigned)G, 256, 0, st>>>(db->pt_ptr, db->nrm, db->nrm2, wG, arT, db->cpart, db->n_ctile, db->gagg);
            ctx->launches++;
            genome_rowsums_sorted_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(n, db->g_ptr, db->g_col, db->g_val, db->g_ord, wT8, db->rT, db->vagg);
            ctx->launches++;
            gom_score_light_kernel<<<(unsigned)n, 128, 0, st>>>(n, (int)G, db->g_ptr, db->pt_ptr, db->pr_ni, db->pr_ptr, db->rec, wT8, arT,
                                                                db->rT, db->gagg, db->vagg, cb, outh);
            ctx->launches++;
            if (db->n_heavy) {
                gom_score_heavy_kernel<<<(unsigned)std::min<long long>(db->n_heavy, ctx->sm_count * 8), 128, db->vg_smem, st>>>(
                    n, (int)G, db->max_t, db->g_ptr, db->pt_ptr, db->pr_ni, db->pr_ptr, db->rec, wT8, arT, db->rT, db->gagg, db->vagg, cb,
                    outh, db->heavy_list, db->n_heavy);
                ctx->launches++;
            }
        }
        timer.end();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "gom_db_score: %s", cudaGetErrorString(e));
    }
    if (weights && !db->weights_validated) {      // multiplicities are staged as uint8 in shared memory: refuse anything above 255
        int flag = 0;
        cudaError_t e = cudaMemcpyAsync(&flag, db->flag, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "gom_db_score: %s", cudaGetErrorString(e));
        if (flag) return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "a column multiplicity exceeds 255 (uint8 weights of the GOM kernels)");
    }
    return GV2_OK;
}
