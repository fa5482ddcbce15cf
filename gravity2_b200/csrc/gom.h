// Device-side GOM database (internal).  See gom_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

struct gv2_ctx;

struct GomDb {
    int64_t n = 0, p = 0, G = 0;
    long long nnz = 0;          // non-zero location entries of the scored genomes
    int npt = 0;                // sum over GOMs of |R_g|
    int max_ng = 0, max_t = 0;
    long long sum_ng2 = 0;      // sum over GOMs of |R_g|^2 (entries of the distance caches)
    long long n_isect = 0;      // entries of all intersection lists
    long long n_mentries = 0;   // sum over (GOM, genome) pairs of |I| (|I| - 1) / 2
    size_t vg_smem = 0;
    size_t bytes = 0;
    // genome CSR
    long long* g_ptr = nullptr; int* g_col = nullptr; double* g_val = nullptr;
    int* g_ord = nullptr;       // per genome: its hit indices ordered by location value (replicate independent)
    // GOM members and point sets
    int* mem_ptr = nullptr; long long* mem_row = nullptr;
    int* pt_ptr = nullptr; int* pt_col = nullptr; int* pos = nullptr;
    long long* x_off = nullptr; long long* a_off = nullptr;
    double* X = nullptr; double* nrm = nullptr; double* nrm2 = nullptr; double* A = nullptr;
    // per-chunk scratch (32 replicates)
    double* wT = nullptr; double* arT = nullptr; double* rT = nullptr;
    double* gagg = nullptr; double* vagg = nullptr;
    double* cpart = nullptr; int n_ctile = 0;   // partial |centroid|^2 per (GOM, tile of eight members)
    unsigned char* wT8 = nullptr; double* wG = nullptr; int* flag = nullptr;   // uint8 weights, per-GOM gathered weights, flags
    long long* heavy_list = nullptr;   // (genome, GOM) pairs whose intersection list overflows the light kernel
    int n_heavy = 0;
    bool weights_validated = false;    // the caller has checked 0 <= w <= 255 for every call (device jobs): no flag read-back per call
    // replicate-independent pair records, genome-major (pair id = v * G + g): record of pair w at rec + pr_ptr[w] =
    // pr_ni[w] entries of 32 bytes (GOM point, genome hit, column, |X|, y), then the lower triangle of M (doubles)
    int* pr_ni = nullptr;
    long long* pr_ptr = nullptr;
    unsigned char* rec = nullptr;
};

int gom_db_build(gv2_ctx* ctx, const double* loc, int64_t n, int64_t p, int64_t ld_loc,
                 const double* rowtab, int64_t ld_row, const int64_t* mem_row_h,
                 const int32_t* mem_ptr_h, int64_t G, int flags, GomDb** out);
int gom_db_score(gv2_ctx* ctx, GomDb* db, const int32_t* weights, int64_t nb, double* out);
void gom_db_free(GomDb* db);
