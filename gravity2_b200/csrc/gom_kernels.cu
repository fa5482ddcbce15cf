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

#define NBC 32   // replicates per chunk = lanes per warp

// ------------------------------------------------------------------------------ genome CSR
__global__ void csr_count_kernel(const double* __restrict__ tab, long long n, long long p, long long ld,
                                 int* __restrict__ cnt) {
    const long long r = blockIdx.x;
    const double* row = tab + r * ld;
    int c = 0;
    for (long long k = threadIdx.x; k < p; k += blockDim.x) c += (row[k] != 0.0);
    __shared__ int red[32];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        cnt[r] = t;
    }
}

// one warp per row, column order preserved
__global__ void csr_fill_kernel(const double* __restrict__ tab, long long n, long long p, long long ld,
                                const long long* __restrict__ ptr, int* __restrict__ col,
                                double* __restrict__ val) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const int lane = threadIdx.x & 31;
    const double* row = tab + r * ld;
    long long o = ptr[r];
    for (long long k0 = 0; k0 < p; k0 += 32) {
        long long k = k0 + lane;
        double v = k < p ? row[k] : 0.0;
        unsigned m = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) {
            long long w = o + __popc(m & ((1u << lane) - 1u));
            col[w] = (int)k;
            val[w] = v;
        }
        o += __popc(m);
    }
}

// ------------------------------------------------------------------------------- GOM side
// flag[g*P + c] = any member of g has rowtab != 0 at c.   grid = (ceil(P/256), G)
__global__ void gom_colany_kernel(const double* __restrict__ rowtab, long long ld, long long p,
                                  const long long* __restrict__ mem_row, const int* __restrict__ mem_ptr,
                                  int all_columns, int* __restrict__ flag) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= p) return;
    int any = all_columns;
    for (int j = mem_ptr[g]; j < mem_ptr[g + 1]; ++j) any |= (rowtab[mem_row[j] * ld + c] != 0.0);
    flag[(long long)g * p + c] = any;
}

// scan[] = exclusive prefix of flag[] over the flattened (g, c) index; point id = scan value.
__global__ void gom_fill_kernel(const double* __restrict__ rowtab, long long ld, long long p,
                                const long long* __restrict__ mem_row, const int* __restrict__ mem_ptr,
                                const int* __restrict__ flag, const int* __restrict__ scan,
                                const long long* __restrict__ x_off, int* __restrict__ pos,
                                int* __restrict__ pt_col, double* __restrict__ X,
                                double* __restrict__ nrm) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= p) return;
    const long long f = (long long)g * p + c;
    if (!flag[f]) {
        pos[f] = -1;
        return;
    }
    const int pt = scan[f];
    const int k = pt - scan[(long long)g * p];
    pos[f] = k;
    pt_col[pt] = (int)c;
    const int m0 = mem_ptr[g], mg = mem_ptr[g + 1] - m0;
    double* x = X + x_off[g] + (long long)k * mg;
    double s = 0.0;
    for (int j = 0; j < mg; ++j) {
        double v = rowtab[mem_row[m0 + j] * ld + c];
        x[j] = v;
        s += v * v;
    }
    nrm[pt] = sqrt(s);
}

// A_g[k][l] = |X_k - X_l|.  grid = (ceil(ng/16), ceil(ng/16), G), block (16,16)
__global__ void gom_dist_cache_kernel(const int* __restrict__ pt_ptr, const int* __restrict__ mem_ptr,
                                      const long long* __restrict__ x_off, const long long* __restrict__ a_off,
                                      const double* __restrict__ X, double* __restrict__ A) {
    const int g = blockIdx.z;
    const int ng = pt_ptr[g + 1] - pt_ptr[g];
    const int k0 = blockIdx.y * 16, l0 = blockIdx.x * 16;
    if (k0 >= ng || l0 >= ng) return;
    const int mg = mem_ptr[g + 1] - mem_ptr[g];
    const double* x = X + x_off[g];
    __shared__ double sk[16][17], sl[16][17];
    const int tx = threadIdx.x, ty = threadIdx.y;
    double acc = 0.0;
    for (int j0 = 0; j0 < mg; j0 += 16) {
        const int j = j0 + tx;
        sk[ty][tx] = (k0 + ty < ng && j < mg) ? x[(long long)(k0 + ty) * mg + j] : 0.0;
        sl[ty][tx] = (l0 + ty < ng && j < mg) ? x[(long long)(l0 + ty) * mg + j] : 0.0;
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            double d = sk[ty][jj] - sl[tx][jj];
            acc += d * d;
        }
        __syncthreads();
    }
    const int k = k0 + ty, l = l0 + tx;
    if (k < ng && l < ng) A[a_off[g] + (long long)k * ng + l] = sqrt(acc);
}

// wT[c][b] = (double) w[b0 + b][c] for b < nb, else 0.   grid = ceil(P*NBC/256)
__global__ void weights_transpose_kernel(const int* __restrict__ w, long long p, int nb,
                                         double* __restrict__ wT) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * NBC) return;
    const long long c = i / NBC;
    const int b = (int)(i % NBC);
    wT[i] = (w == nullptr) ? (b < nb ? 1.0 : 0.0) : (b < nb ? (double)w[(long long)b * p + c] : 0.0);
}

// GA: arT[pt][b] = sum_l w_l A[k][l],  sqT[pt][b] = sum_l w_l A[k][l]^2.
// grid = (ceil(max_ng/8), G), block = (32 b, 8 k)
__global__ void gom_rowsums_kernel(const int* __restrict__ pt_ptr, const int* __restrict__ pt_col,
                                   const long long* __restrict__ a_off, const double* __restrict__ A,
                                   const double* __restrict__ wT, double* __restrict__ arT,
                                   double* __restrict__ sqT) {
    const int g = blockIdx.y;
    const int p0 = pt_ptr[g], ng = pt_ptr[g + 1] - p0;
    const int k = blockIdx.x * 8 + threadIdx.y;
    if (k >= ng) return;
    const int b = threadIdx.x;
    const double* a = A + a_off[g] + (long long)k * ng;
    double s1 = 0.0, s2 = 0.0;
    for (int l = 0; l < ng; ++l) {
        const double w = wT[(long long)pt_col[p0 + l] * NBC + b];
        const double d = a[l];
        s1 = fma(w, d, s1);
        s2 = fma(w * d, d, s2);
    }
    arT[(long long)(p0 + k) * NBC + b] = s1;
    sqT[(long long)(p0 + k) * NBC + b] = s2;
}

// GB: per-GOM aggregates, lane = replicate.  gagg[g][q][b], q in 0..GAGG-1.  grid = G, block = 32
#define GAGG 7   // n_g, N1, N2, AR1, AR2, ARN, SAA
__global__ void gom_aggregates_kernel(const int* __restrict__ pt_ptr, const int* __restrict__ pt_col,
                                      const double* __restrict__ nrm, const double* __restrict__ wT,
                                      const double* __restrict__ arT, const double* __restrict__ sqT,
                                      double* __restrict__ gagg) {
    const int g = blockIdx.x, b = threadIdx.x;
    double n = 0, N1 = 0, N2 = 0, AR1 = 0, AR2 = 0, ARN = 0, SAA = 0;
    for (int pt = pt_ptr[g]; pt < pt_ptr[g + 1]; ++pt) {
        const double w = wT[(long long)pt_col[pt] * NBC + b];
        const double nr = nrm[pt], ar = arT[(long long)pt * NBC + b], sq = sqT[(long long)pt * NBC + b];
        n += w;
        N1 = fma(w, nr, N1);
        N2 = fma(w * nr, nr, N2);
        AR1 = fma(w, ar, AR1);
        AR2 = fma(w * ar, ar, AR2);
        ARN = fma(w * ar, nr, ARN);
        SAA = fma(w, sq, SAA);
    }
    double* o = gagg + (long long)g * GAGG * NBC + b;
    o[0 * NBC] = n; o[1 * NBC] = N1; o[2 * NBC] = N2; o[3 * NBC] = AR1;
    o[4 * NBC] = AR2; o[5 * NBC] = ARN; o[6 * NBC] = SAA;
}

// VA: per-genome weighted row sums r_c = sum_{d in R_v} w_d |y_c - y_d| and aggregates.
// One block per genome; threadIdx.x = replicate lane, threadIdx.y strides over the genome's hits.
#define VAGG 7   // t, Y1, Y2, Ys, RR1, RR2, RY
__global__ void genome_rowsums_kernel(const long long* __restrict__ ptr, const int* __restrict__ col,
                                      const double* __restrict__ val, const double* __restrict__ wT,
                                      double* __restrict__ rT, double* __restrict__ vagg) {
    const long long v = blockIdx.x;
    const long long o = ptr[v];
    const int t = (int)(ptr[v + 1] - o);
    const int b = threadIdx.x, ny = blockDim.y;
    double T = 0, Y1 = 0, Y2 = 0, Ys = 0, RR1 = 0, RR2 = 0, RY = 0;
    for (int c = threadIdx.y; c < t; c += ny) {
        const double yc = val[o + c];
        double r = 0.0;
        for (int d = 0; d < t; ++d) r = fma(wT[(long long)col[o + d] * NBC + b], fabs(yc - val[o + d]), r);
        rT[(o + c) * NBC + b] = r;
        const double w = wT[(long long)col[o + c] * NBC + b], ay = fabs(yc);
        T += w;
        Y1 = fma(w, ay, Y1);
        Y2 = fma(w * yc, yc, Y2);
        Ys = fma(w, yc, Ys);
        RR1 = fma(w, r, RR1);
        RR2 = fma(w * r, r, RR2);
        RY = fma(w * r, ay, RY);
    }
    __shared__ double red[8][VAGG][NBC];
    double q[VAGG] = {T, Y1, Y2, Ys, RR1, RR2, RY};
    for (int i = 0; i < VAGG; ++i) red[threadIdx.y][i][b] = q[i];
    __syncthreads();
    if (threadIdx.y == 0) {
        for (int i = 0; i < VAGG; ++i) {
            double s = 0.0;
            for (int y = 0; y < ny; ++y) s += red[y][i][b];
            vagg[(v * VAGG + i) * NBC + b] = s;
        }
    }
}

// VG: one warp per (genome, GOM); lane = replicate.
// out[b][v][g].  grid = ceil(n*G / warps_per_block)
__global__ void __launch_bounds__(128)
gom_score_kernel(long long n, int G, long long p, const long long* __restrict__ ptr,
                 const int* __restrict__ col, const double* __restrict__ val,
                 const int* __restrict__ pt_ptr, const int* __restrict__ pos,
                 const double* __restrict__ nrm, const long long* __restrict__ a_off,
                 const double* __restrict__ A, const double* __restrict__ wT,
                 const double* __restrict__ arT, const double* __restrict__ rT,
                 const double* __restrict__ gagg, const double* __restrict__ vagg, int nb,
                 double* __restrict__ out, int* __restrict__ ilist, long long ilist_stride,
                 long long w_base, long long w_end) {
    const int lane = threadIdx.x & 31;
    const long long wrel = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long wid = w_base + wrel;
    if (wid >= w_end) return;
    // GOM-major order keeps one GOM's distance cache hot in L2 while all genomes are scored
    const int g = (int)(wid / n);
    const long long v = wid % n;
    const long long o = ptr[v];
    const int t = (int)(ptr[v + 1] - o);
    const int p0 = pt_ptr[g], ng = pt_ptr[g + 1] - p0;
    const int* posg = pos + (long long)g * p;
    const double* Ag = A + a_off[g];

    // intersection list I0 = R_v & R_g (replicate independent): entries (hit index, point index)
    int* il = ilist + wrel * ilist_stride;   // [2 * t] scratch in global memory
    int ni = 0;
    for (int c0 = 0; c0 < t; c0 += 32) {
        const int c = c0 + lane;
        int k = -1;
        if (c < t) k = posg[col[o + c]];
        const unsigned m = __ballot_sync(0xffffffffu, k >= 0);
        if (k >= 0) {
            const int w = ni + __popc(m & ((1u << lane) - 1u));
            il[2 * w] = c;
            il[2 * w + 1] = k;
        }
        ni += __popc(m);
    }
    __syncwarp();

    const int b = lane;
    double SW = 0, S_ar = 0, S_nrm = 0, S_r = 0, S_y = 0, C_arr = 0, C_ary = 0, C_nr = 0, C_ny = 0;
    double Q1 = 0, Q2 = 0, Q3 = 0;
    for (int a = 0; a < ni; ++a) {
        const int c = il[2 * a], k = il[2 * a + 1];
        const double yc = val[o + c], ayc = fabs(yc);
        const double wc = wT[(long long)col[o + c] * NBC + b];
        const double ar = arT[(long long)(p0 + k) * NBC + b];
        const double nr = nrm[p0 + k];
        const double r = rT[(o + c) * NBC + b];
        SW += wc;
        S_ar = fma(wc, ar, S_ar);
        S_nrm = fma(wc, nr, S_nrm);
        S_r = fma(wc, r, S_r);
        S_y = fma(wc, ayc, S_y);
        C_arr = fma(wc * ar, r, C_arr);
        C_ary = fma(wc * ar, ayc, C_ary);
        C_nr = fma(wc * nr, r, C_nr);
        C_ny = fma(wc * nr, ayc, C_ny);
        const double* arow = Ag + (long long)k * ng;
        double q1 = 0, q2 = 0, q3 = 0;
        for (int e = 0; e < a; ++e) {           // unordered pairs e < a, doubled below
            const int d = il[2 * e], l = il[2 * e + 1];
            const double yd = val[o + d];
            const double wd = wT[(long long)col[o + d] * NBC + b];
            const double dist = arow[l];
            const double dy = fabs(yc - yd);
            q1 = fma(wd * dist, dy, q1);
            q2 = fma(wd * dist, ayc + fabs(yd), q2);
            q3 = fma(wd * (nr + nrm[p0 + l]), dy, q3);
        }
        Q1 = fma(2.0 * wc, q1, Q1);
        Q2 = fma(wc, q2, Q2);
        Q3 = fma(wc, q3, Q3);
    }
    if (b >= nb) return;
    const double* ga = gagg + (long long)g * GAGG * NBC + b;
    const double n_g = ga[0], N1 = ga[NBC], N2 = ga[2 * NBC], AR1 = ga[3 * NBC], AR2 = ga[4 * NBC],
                 ARN = ga[5 * NBC], SAA = ga[6 * NBC];
    const double* va = vagg + v * VAGG * NBC + b;
    const double T = va[0], Y1 = va[NBC], Y2 = va[2 * NBC], Ys = va[3 * NBC], RR1 = va[4 * NBC],
                 RR2 = va[5 * NBC], RY = va[6 * NBC];
    const double e = T - SW, z = n_g - SW, nn = n_g + e;
    double res = 0.0;
    if (nn > 0.0) {
        const double Sab = 2.0 * (C_ary - Q2) + 2.0 * (N1 - S_nrm) * (Y1 - S_y) + Q1 + 2.0 * (C_nr - Q3);
        const double cross = ((AR1 - S_ar) + e * (N1 - S_nrm)) * Y1 +
                             (C_arr + z * C_ary + e * C_nr + e * z * C_ny) +
                             N1 * ((RR1 - S_r) + z * (Y1 - S_y));
        const double a_tot = AR1 + 2.0 * e * N1, b_tot = RR1 + 2.0 * z * Y1;
        const double rowA = AR2 + 2.0 * e * ARN + e * e * N2 + e * N1 * N1;
        const double rowB = RR2 + 2.0 * z * RY + z * z * Y2 + z * Y1 * Y1;
        const double Saa = SAA + 2.0 * e * N2;
        const double Sbb = 2.0 * T * Y2 - 2.0 * Ys * Ys + 2.0 * z * Y2;
        const double sAB = Sab - 2.0 / nn * cross + a_tot * b_tot / (nn * nn);
        const double sAA = Saa - 2.0 / nn * rowA + a_tot * a_tot / (nn * nn);
        const double sBB = Sbb - 2.0 / nn * rowB + b_tot * b_tot / (nn * nn);
        // dcor.py:10,17,54-60
        const double dcov = sqrt(sAB) / nn;            // NaN if the sum rounds below zero, as numpy
        const double dvA = sqrt(sAA / (nn * nn));
        const double dvB = sqrt(sBB / (nn * nn));
        if (dvA > 0.0 && dvB > 0.0) res = dcov / sqrt(dvA * dvB);
    }
    out[((long long)b * n + v) * G + g] = res;
}

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
    cudaFree(db->g_ptr); cudaFree(db->g_col); cudaFree(db->g_val);
    cudaFree(db->mem_ptr); cudaFree(db->mem_row); cudaFree(db->pt_ptr); cudaFree(db->pt_col);
    cudaFree(db->pos); cudaFree(db->x_off); cudaFree(db->a_off); cudaFree(db->X); cudaFree(db->nrm);
    cudaFree(db->A); cudaFree(db->wT); cudaFree(db->arT); cudaFree(db->sqT); cudaFree(db->rT);
    cudaFree(db->gagg); cudaFree(db->vagg); cudaFree(db->ilist);
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
    CK(dalloc(&db->X, x_off_h[G], &db->bytes));
    CK(dalloc(&db->A, a_off_h[G], &db->bytes));
    if (G && p) {
        gom_fill_kernel<<<gridc, 256, 0, st>>>(rowtab, ld_row, p, db->mem_row, db->mem_ptr, flag, scan,
                                               db->x_off, db->pos, db->pt_col, db->X, db->nrm);
        ctx->launches++;
    }
    if (max_ng) {
        dim3 g2((unsigned)((max_ng + 15) / 16), (unsigned)((max_ng + 15) / 16), (unsigned)G), b2(16, 16);
        gom_dist_cache_kernel<<<g2, b2, 0, st>>>(db->pt_ptr, db->mem_ptr, db->x_off, db->a_off, db->X, db->A);
        ctx->launches++;
    }
    // ---- per-chunk (32 replicates) scratch
    CK(dalloc(&db->wT, p * NBC, &db->bytes));
    CK(dalloc(&db->arT, (size_t)db->npt * NBC, &db->bytes));
    CK(dalloc(&db->sqT, (size_t)db->npt * NBC, &db->bytes));
    CK(dalloc(&db->rT, (size_t)db->nnz * NBC, &db->bytes));
    CK(dalloc(&db->gagg, (size_t)G * GAGG * NBC, &db->bytes));
    CK(dalloc(&db->vagg, (size_t)n * VAGG * NBC, &db->bytes));
    // intersection scratch: each (genome, GOM) warp needs 2 ints per hit of the genome
    int max_t = 0;
    {
        std::vector<long long> ptr_h(n + 1);
        if (n) CK(cudaMemcpy(ptr_h.data(), db->g_ptr, (n + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
        for (int64_t v = 0; v < n; ++v) max_t = std::max<int>(max_t, (int)(ptr_h[v + 1] - ptr_h[v]));
    }
    db->max_t = max_t;
    db->ilist_stride = 2LL * max_t;
    // scored in waves of `wave` (genome, GOM) warps so that the scratch stays bounded
    long long total_w = n * G;
    long long cap = (long long)(1ULL << 30) / (sizeof(int) * (db->ilist_stride ? db->ilist_stride : 1));
    db->wave = std::max<long long>(128, std::min<long long>(total_w, cap));
    CK(dalloc(&db->ilist, (size_t)db->wave * db->ilist_stride, &db->bytes));
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
    for (int64_t b0 = 0; b0 < nb; b0 += NBC) {
        const int cb = (int)std::min<int64_t>(NBC, nb - b0);
        if (p) {
            weights_transpose_kernel<<<(unsigned)((p * NBC + 255) / 256), 256, 0, st>>>(
                weights ? weights + b0 * p : nullptr, p, cb, db->wT);
            ctx->launches++;
        }
        if (db->max_ng) {
            dim3 g1((unsigned)((db->max_ng + 7) / 8), (unsigned)G), b1(32, 8);
            gom_rowsums_kernel<<<g1, b1, 0, st>>>(db->pt_ptr, db->pt_col, db->a_off, db->A, db->wT, db->arT, db->sqT);
            ctx->launches++;
        }
        gom_aggregates_kernel<<<(unsigned)G, 32, 0, st>>>(db->pt_ptr, db->pt_col, db->nrm, db->wT, db->arT, db->sqT, db->gagg);
        ctx->launches++;
        dim3 bv(32, 8);
        genome_rowsums_kernel<<<(unsigned)n, bv, 0, st>>>(db->g_ptr, db->g_col, db->g_val, db->wT, db->rT, db->vagg);
        ctx->launches++;
        // score in waves; a wave covers a contiguous range of the GOM-major (g, v) index
        const long long total_w = n * G;
        for (long long w0 = 0; w0 < total_w; w0 += db->wave) {
            const long long wn = std::min<long long>(db->wave, total_w - w0);
            gom_score_kernel<<<(unsigned)((wn + 3) / 4), 128, 0, st>>>(
                n, (int)G, p, db->g_ptr, db->g_col, db->g_val, db->pt_ptr, db->pos, db->nrm, db->a_off,
                db->A, db->wT, db->arT, db->rT, db->gagg, db->vagg, cb, out + b0 * n * G,
                db->ilist, db->ilist_stride, w0, w0 + wn);
            ctx->launches++;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "gom_db_score: %s", cudaGetErrorString(e));
    }
    return GV2_OK;
}
