/* libgravity_b200 -- C ABI of the B200 (sm_100a) pairwise genome-similarity engine.
 *
 * Drop-in boundary for the hot path of GRAViTy-V2 (Mayne941/gravity2): PPHMM / GOM
 * signature tables -> generalised-Jaccard similarity matrices -> composite (PG) distance
 * -> bootstrap replicate matrices, plus the GOM distance-correlation scoring and the
 * presence/absence (popcount) matrices.  Citations are file:line in the reference checkout.
 *
 * Conventions
 *   - plain C: pointers + int64 sizes; no C++ / torch types cross this boundary;
 *   - every entry point returns int: GV2_OK (0) or a negative GV2_ERR_* code; the message is
 *     read with gv2_last_error(ctx);
 *   - "_host" entry points take HOST pointers owned by the caller (numpy arrays in the
 *     reference's process) and return results in caller-owned host buffers;
 *   - "gv2_dev_" entry points take DEVICE pointers on the context's device and enqueue work on
 *     the context's stream without synchronising (the caller synchronises with gv2_synchronize
 *     or its own stream);
 *   - a context is single-caller (not re-entrant), like the reference's call sites
 *     (SURVEY.md section 8b);
 *   - there is NO CPU fallback: without a usable CUDA device gv2_create fails.
 */
#ifndef GRAVITY2_B200_H
#define GRAVITY2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GV2_OK 0
#define GV2_ERR_ARG (-1)
#define GV2_ERR_CUDA (-2)
#define GV2_ERR_NOMEM (-3)
#define GV2_ERR_UNSUPPORTED (-4)
#define GV2_ERR_ZERODIV (-5) /* shared_norm_pphmm_ratio 0/0, shared_pphmm_graphs.py:86 */

/* SimilarityMeasurementScheme, similarity_matrix_constructor.py:207-222 */
#define GV2_SCHEME_P 0
#define GV2_SCHEME_G 1
#define GV2_SCHEME_PG 2
#define GV2_SCHEME_R 3
#define GV2_SCHEME_RG 4
#define GV2_SCHEME_PR 5
#define GV2_SCHEME_L 6  /* `spr` then carries the location dCor matrix (gv2_gom_table_host, one GOM per genome) */
#define GV2_SCHEME_PL 7

#define GV2_OUT_SIMILARITY 0 /* SimilarityMat ** p, similarity_matrix_constructor.py:226 */
#define GV2_OUT_DISTANCE 1   /* D = 1 - S; D[D<0] = 0, app/src/pl1_graphs.py:52-53 */

#define GV2_TILE_EDGE 128 /* genome-pair tile edge; unit of multi-GPU sharding */

typedef struct gv2_ctx gv2_ctx;

/* ---- context ------------------------------------------------------------------------ */
int gv2_version(void);
int gv2_device_count(void);
/* Binds to CUDA device `device` and creates a private stream.  Lazy: the first CUDA call in
 * the process happens here, never at library load (fork safety). */
int gv2_create(int device, gv2_ctx** out);
/* Same, but enqueue on an existing cudaStream_t (e.g. torch's current stream). */
int gv2_create_on_stream(int device, void* cuda_stream, gv2_ctx** out);
void gv2_destroy(gv2_ctx* ctx);
const char* gv2_last_error(const gv2_ctx* ctx); /* ctx may be NULL: last creation error */
int gv2_synchronize(gv2_ctx* ctx);
uint64_t gv2_launch_count(const gv2_ctx* ctx); /* kernels launched so far through ctx */
/* CUDA-event timing of the hot kernels, on the context's stream, by group: enable, run, then read the summed
 * duration and the number of timed launch groups of one kind (reading synchronises and resets that kind). */
#define GV2_TIME_PAIRSUMS 0 /* replicate pair sums: CSR walk (sparse tables) or tensor-core contraction */
#define GV2_TIME_TAIL 1     /* fused GOM Jaccard + epilogue kernel */
#define GV2_TIME_GOM 2      /* GOM dCor scoring kernels of one 32-replicate chunk */
#define GV2_TIME_LINKAGE 3  /* device UPGMA of one sub-chunk of replicate matrices */
#define GV2_TIME_BASE 4     /* one gv2_job_base call (un-resampled matrix) */
int gv2_timing_enable(gv2_ctx* ctx, int on);
int gv2_timing_read_kind(gv2_ctx* ctx, int kind, double* total_ms, int64_t* launches);
int gv2_timing_read(gv2_ctx* ctx, double* total_ms, int64_t* launches); /* = kind GV2_TIME_PAIRSUMS */

/* device / pinned-host memory for hosts without their own CUDA binding (ctypes, cgo, JNI ...) */
int gv2_dev_alloc(gv2_ctx* ctx, size_t bytes, void** out);
int gv2_dev_free(gv2_ctx* ctx, void* ptr);
int gv2_host_alloc(gv2_ctx* ctx, size_t bytes, void** out); /* page-locked */
int gv2_host_free(gv2_ctx* ctx, void* ptr);
int gv2_memcpy_h2d(gv2_ctx* ctx, void* dst, const void* src, size_t bytes); /* synchronous */
int gv2_memcpy_d2h(gv2_ctx* ctx, void* dst, const void* src, size_t bytes);
int gv2_mem_info(gv2_ctx* ctx, size_t* free_bytes, size_t* total_bytes); /* cudaMemGetInfo of the context's device */

/* ---- host-pointer entry points (what the reference-side binding calls) ----------------- */

/* Replaces SimilarityMat_Constructor's pair loop + clean/combine/power for
 * PphmmNeighbourhoodWeight == 0 (app/utils/similarity_matrix_constructor.py:112-226).
 *   sig  [n][p] float64, row stride ld_sig   (NULL if the scheme has no "P")
 *   gom  [n][g] float64, row stride ld_gom   (NULL if the scheme has no "G")
 *   spr  [n][n] float64: SPRSignatureTable for R / RG / PR; the PPHMM-location dCor matrix
 *        (:179-186, computed with gv2_gom_table_host) for L / PL; else NULL
 *   apply_threshold != 0: entries `x < threshold` count as 0 (:139-140); the caller mirrors
 *   that mutation onto its own array.
 *   out  [n][n] float64, fresh symmetric matrix (S**p or D per out_kind). */
int gv2_similarity_host(gv2_ctx* ctx, const double* sig, int64_t n, int64_t p, int64_t ld_sig,
                        const double* gom, int64_t g, int64_t ld_gom, const double* spr,
                        int scheme, double power, double threshold, int apply_threshold,
                        int out_kind, double* out);

/* PphmmNeighbourhoodWeight != 0 parity mode of the same function (:112-169 with the in-place, visit-count
 * dependent row scaling of :120-140): the cleaned (:201-205) PPHMM generalised-Jaccard
 * matrix in which, at pair (i, j >= i), row i has been through j + 2 and row j through i + 1 applications of
 * "negative-list entries x (1 - weight), positive-list entries x (1 + weight), then x < threshold -> 0".
 *   cls [n][p] uint8 (dense, row stride p): bit 0 = column on the genome's negative list, bit 1 = on its
 *       positive list (the lists of :79-110, built host-side from the two CSVs by the binding);
 *   mul [n][n] float64 or NULL: element-wise factor applied to the cleaned matrix (SPR / location dCor
 *       matrix of schemes PR / PL);
 *   out [n][n] float64.  The caller combines it through gv2_similarity_host (as `spr`, schemes R / RG) and
 *   mirrors the reference's N + 1 applications onto its own table. */
int gv2_similarity_nbhd_host(gv2_ctx* ctx, const double* sig, int64_t n, int64_t p, int64_t ld_sig,
                             const uint8_t* cls, double weight, double threshold, const double* mul,
                             double* out);

/* Replaces GOMSignatureTable_Constructor / generate_gom_sigs
 * (app/utils/parallel_gom_sig_generator.py:10-37, gom_signature_table_constructor.py:34-61).
 *   loc       [n][p] float64 PPHMM location table of the genomes to score
 *   gom_rows  [gom_offsets[g]][p] float64: the GOMDB members' location rows, GOM after GOM in
 *             GOMIDList order; GOM k owns rows gom_offsets[k] .. gom_offsets[k+1]
 *   weights   [nb][p] int32 column multiplicities (bootstrap resample as bincount of the
 *             reference's index list) or NULL (nb = 1, all ones)
 *   out       [nb][n][g] float64; NaN where the reference yields NaN (dcor.py:10).
 * With one single-member GOM per genome (gom_rows = loc, gom_offsets = 0..n) the result is the
 * PPHMM-location dCor matrix of scheme "L" (similarity_matrix_constructor.py:179-186). */
int gv2_gom_table_host(gv2_ctx* ctx, const double* loc, int64_t n, int64_t p, int64_t ld_loc,
                       const double* gom_rows, int64_t ld_gom_rows, const int64_t* gom_offsets,
                       int64_t g, const int32_t* weights, int64_t nb, int flags, double* out);
#define GV2_GOM_ALL_COLUMNS 1 /* flags: score over every column instead of the relevant set R
                                 (plain dcor(X, Y) of app/utils/dcor.py:34-62) */

/* Replaces the clustering of DistMat2Tree (app/utils/dist_mat_to_tree.py:7-12): scipy.cluster.hierarchy.linkage of the
 * upper triangle of `dist` [n][n] (host), method "average" | "complete" | "weighted" | "ward" | "single" | "centroid" |
 * "median" (all seven of scipy, api_classes.py:123); Z [n-1][4] float64 as scipy returns it (sorted and relabelled; for
 * centroid / median in merge order, heights not monotone), bit for bit.  centroid / median: n <= 12 000. */
int gv2_linkage_host(gv2_ctx* ctx, const double* dist, int64_t n, const char* method, double* Z);
/* to_tree + GetNewick (app/utils/get_newick.py:1-13) of a linkage matrix: host only, no context.  *out_len = length
 * without the terminator; with out == NULL only the length is returned. */
int gv2_newick_from_linkage(const double* Z, int64_t n, const char* const* leaf_names, char* out, size_t out_cap,
                            size_t* out_len);

/* Replaces shared_pphmm_ratio / the counts of shared_norm_pphmm_ratio
 * (app/utils/shared_pphmm_graphs.py:59-97) and the binary Jaccard of
 * RefVirusAnnotator.sort_pphmm (app/src/ref_virus_annotator.py:148-158): bit-packs
 * `tab != 0` and counts co-present columns.
 *   inter [n][n] int32 (popcount(a & b)), nnz [n] int32 (popcount(a)); bit-exact. */
int gv2_shared_counts_host(gv2_ctx* ctx, const double* tab, int64_t n, int64_t p, int64_t ld,
                           int32_t* inter, int32_t* nnz);

/* Bootstrap replicate distance (or similarity) matrices: the loop bodies of
 * app/src/pl1_graphs.py:78-107 and app/src/virus_classification.py:290-315 for all
 * replicates [0, nb) in batched launches.
 *   sig, loc   [n][p] float64 (PL1 passes the signature table as loc)
 *   gom_of_row [n_gom_rows] int32: GOM index (0..g-1) of loc row r, or -1 if that row is not
 *              a GOMDB member; only rows [0, n_gom_rows) can be members (PL2: the reference rows)
 *   weights    [nb][p] int32 = bincount of each replicate's resample index list
 *   out        [nb][n][n] float64 */
int gv2_bootstrap_host(gv2_ctx* ctx, const double* sig, const double* loc, int64_t n, int64_t p,
                       const int32_t* gom_of_row, int64_t n_gom_rows, int64_t g,
                       const int32_t* weights, int64_t nb, int scheme, double power,
                       int out_kind, double* out);

/* Counter-based (Philox4x32-10) resample indices, the device-RNG mode: out[b][k] in [0, p).
 * The reference-identical mode draws the indices on the host with numpy's legacy MT19937
 *  and passes their bincount as `weights`. */
int gv2_philox_indices_host(gv2_ctx* ctx, uint64_t seed, int64_t nb, int64_t p, int64_t* out);
int gv2_dev_philox_weights(gv2_ctx* ctx, uint64_t seed, int64_t b0, int64_t nb, int64_t p,
                           int32_t* weights /* [nb][p] */);

/* ---- device-resident job: tables stay in HBM across calls ------------------------------ */
typedef struct gv2_job gv2_job;
typedef int (*gv2_replicate_cb)(int64_t replicate, const double* matrix /* [n][n] host */, void* user);
/* Uploads (or, with *_is_device != 0, reads in place) the float64 tables, converts them to the
 * kernels' layouts and builds the GOM side structures.  loc / gom_of_row may be NULL when the
 * scheme has no "G". */
int gv2_job_create(gv2_ctx* ctx, const double* sig, const double* loc, int tables_are_device,
                   int64_t n, int64_t p, const int32_t* gom_of_row, int64_t n_gom_rows, int64_t g,
                   int scheme, double power, int out_kind, gv2_job** out);
/* The same job from COO triplets in HOST memory (rows / cols int32, vals float64, no duplicate cells): the form the reference
 * pickles beside the dense tables (PPHMMSignatureTable_coo / PPHMMLocationTable_coo, app/src/ref_virus_annotator.py:395-400).
 * 16 bytes per hit cross PCIe instead of 8 bytes per cell; the dense images are rebuilt in HBM.  loc_rows == NULL: the
 * signature table doubles as location table (PL1, quirk Q3). */
int gv2_job_create_coo(gv2_ctx* ctx, int64_t n, int64_t p, const int32_t* sig_rows, const int32_t* sig_cols, const double* sig_vals,
                       int64_t sig_nnz, const int32_t* loc_rows, const int32_t* loc_cols, const double* loc_vals, int64_t loc_nnz,
                       const int32_t* gom_of_row, int64_t n_gom_rows, int64_t g, int scheme, double power, int out_kind,
                       gv2_job** out);
void gv2_job_destroy(gv2_job* job);
/* Base (un-resampled) matrix for pair tiles [tile_begin, tile_end) of the upper triangle:
 *   packed == 0: writes the tiles (and their mirrors) into dev_out [n][n];
 *   packed != 0: dev_out is [tile_end-tile_begin][128*128] tile blocks (multi-GPU all-gather
 *   payload; scatter with gv2_dev_unpack_tiles). */
int gv2_job_base(gv2_job* job, int64_t tile_begin, int64_t tile_end, int packed, double* dev_out);
/* Replicates [0, nb) given their device-resident weights; dev_out [nb][n][n]. */
int gv2_job_replicates(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_out);
/* Streaming form.  host_out == NULL (device-only): replicate b goes to slot (b % ring_len) of dev_ring [ring_len][n][n],
 * so after the call the last ring_len replicates are resident.  With host_out != NULL ([nb][n][n] host memory, ideally
 * pinned) the ring is used as two halves (one computing, one draining; ring_len >= 2) and every replicate is also copied
 * out on a second stream while the next ones are computed, so the device never holds more than ring_len matrices.
 * callback == NULL: host_out is linear, replicate b lands in host_out[b].
 * callback != NULL: host_out is a HOST RING [ring_len][n][n]; as soon as replicate b is in host memory the
 * library calls callback(b, matrix, user) on the calling thread (in replicate order) while the GPU already
 * computes the next sub-chunk; the matrix is valid until the callback returns; non-zero return stops. */
int gv2_job_replicates_stream(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring,
                              int64_t ring_len, double* host_out, gv2_replicate_cb callback, void* user);
/* The same with every replicate delivered in scipy's condensed form -- the n (n - 1) / 2 entries above the diagonal in
 * squareform order, what app/utils/dist_mat_to_tree.py:8-9 hands to linkage() -- so half the bytes cross PCIe: host_out is
 * [nb][n(n-1)/2] (callback == NULL) or a host ring [ring_len][n(n-1)/2]; the callback receives the condensed vector. */
int gv2_job_replicates_condensed(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring,
                                 int64_t ring_len, double* host_out, gv2_replicate_cb callback, void* user);
/* Device consumer (classification epilogue of pipeline II): callback(b, matrix, user) receives
 * the DEVICE pointer of replicate b's matrix in the ring, on the calling thread, in replicate order, after the stream has
 * been synchronised; valid until the callback returns (gv2_dev_rowmax_block / gv2_dev_gather_cells read it in place). */
int gv2_job_replicates_device(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring, int64_t ring_len,
                              gv2_replicate_cb callback, void* user);
/* Replicate dendrograms without shipping the matrices: each replicate's distance matrix is
 * clustered on the device (scipy.cluster.hierarchy.linkage(..., method): the nearest-neighbour chain behind "average"
 * (UPGMA, the default), "complete", "weighted" and "ward", the minimum spanning tree behind "single", the heap-based generic
 * algorithm behind "centroid" / "median" (n <= 12 000); scipy's tie order) and printed as DistMat2Tree + GetNewick print it (app/utils/dist_mat_to_tree.py:5-21,
 * app/utils/get_newick.py:1-13); callback(b, newick, length, user) runs on the calling thread in replicate order.
 * The job must have been created with GV2_OUT_DISTANCE.  dev_ring [ring_len][n][n] is scratch: with ring_len >= 8 it is used
 * as four parts, one being computed while the others are clustered on their own streams (ring_len / 4 CTAs each; the
 * clustering of a matrix takes ~2x as long as computing it); as many matrices as memory allows, up to ~128 on a 148-SM part. */
typedef int (*gv2_newick_cb)(int64_t replicate, const char* newick, size_t length, void* user);
int gv2_job_replicates_newick(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring,
                              int64_t ring_len, const char* method, const char* const* leaf_names,
                              gv2_newick_cb callback, void* user);
/* GOM signature table of the job's loc table under the given weights (NULL = base):
 * dev_out [nb][n][g]. */
int gv2_job_gom_table(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_out);
size_t gv2_job_device_bytes(const gv2_job* job);
/* Sizes of the job's GOM database (roofline bookkeeping): stats[0] = sum_g |R_g| (GOM points), [1] = sum_g |R_g|^2
 * (distance-cache entries), [2] = non-zero location entries, [3] = (genome, GOM) pairs on the heavy list, [4] = entries of the
 * intersection lists, [5] = sum over pairs of |I| (|I| - 1) / 2 (pair-sum terms per replicate), [6] = max |R_g|, [7] = max hits. */
int gv2_job_gom_stats(const gv2_job* job, int64_t* stats /* [8] */);
/* Replicate pair-sum path chosen for the job's signature table: exact integer arithmetic on a fixed-point image
 * (CSR walk of sparse tables / tensor-core contraction of dense ones) or the fp32 tile kernel (tables with negative
 * entries).  The environment variable GV2_BOOT_PATH = sparse | mma | fp32 overrides the cost model. */
#define GV2_PATH_NONE 0
#define GV2_PATH_FP32 1
#define GV2_PATH_MMA 2
#define GV2_PATH_SPARSE 3
int gv2_job_replicate_path(const gv2_job* job);
uint64_t gv2_job_copresent_cells(const gv2_job* job); /* sum over column c of cnt_c (cnt_c + 1) / 2 */

/* ---- device-pointer primitives ----------------------------------------------------------- */
int64_t gv2_num_tiles(int64_t n); /* upper-triangular 128x128 tiles incl. diagonal */

typedef struct {
    const double* min_sums;        /* [nb][ksplit][tiles][128*128] from gv2_dev_gj_min_sums */
    const double* rowsums;         /* [nb][n] sum_c w_c * x_c */
    const uint8_t* nanflags;       /* [nb][n] row contains NaN, or NULL */
    int32_t ksplit;
    int64_t replicate_stride_sums; /* elements between replicates in min_sums (0 = shared) */
    int64_t replicate_stride_rows; /* elements between replicates in rowsums (0 = shared) */
    int64_t replicate_stride_nan;  /* elements between replicates in nanflags (0 = shared) */
} gv2_gj_component;

/* float64 [n][k] -> zero-padded float32 [n_pad][k_pad] (n_pad % 128 == 0, k_pad % 32 == 0),
 * per-row sum and NaN flag. */
int gv2_dev_prepare_table(gv2_ctx* ctx, const double* src, int64_t n, int64_t k, int64_t ld_src,
                          double threshold, int apply_threshold, float* dst, int64_t n_pad,
                          int64_t k_pad, double* rowsum, uint8_t* nanflag);
int gv2_dev_weights_to_float(gv2_ctx* ctx, const int32_t* w, int64_t nb, int64_t k, int64_t k_pad,
                             float* out);
int gv2_dev_weighted_rowsums(gv2_ctx* ctx, const float* tab, int64_t n, int64_t k_pad,
                             const float* w, int64_t nb, double* out);
/* K1: sum_c w_c * min(x_ic, x_jc) for pair tiles [tile_begin, tile_end), nb weight rows
 * (w == NULL: nb must be 1, unweighted). */
int gv2_dev_gj_min_sums(gv2_ctx* ctx, const float* tab, int64_t n_pad, int64_t k_pad,
                        int64_t tile_begin, int64_t tile_end, const float* w, int64_t nb,
                        int ksplit, double* ws);
/* K3: clean / combine / clamp / power / (1 - S) / mirror. */
int gv2_dev_gj_epilogue(gv2_ctx* ctx, int64_t n, int64_t tile_begin, int64_t tile_end, int64_t nb,
                        const gv2_gj_component* comp_p, const gv2_gj_component* comp_g,
                        const double* spr, int scheme, double power, int out_kind, int packed,
                        double* out, int64_t out_replicate_stride);
/* K1 + K3 in one kernel for the per-replicate tail (schemes G and PG, all tiles): pair sums of the fp32 table(s)
 * `tab` (one [n_pad][k_pad] table per replicate, k real columns, tab_replicate_stride elements apart; the GOM signature tables)
 * are formed in registers and combined at once with the PPHMM component comp_p (its min_sums from
 * gv2_dev_gj_min_sums / the replicate kernels, ksplit == 1; NULL for scheme G); comp_g carries only the row sums and
 * NaN flags of `tab`.  out [nb][n][n], symmetric. */
int gv2_dev_gj_fused(gv2_ctx* ctx, const float* tab, int64_t tab_replicate_stride, int64_t n, int64_t n_pad,
                     int64_t k, int64_t k_pad, int64_t nb, const gv2_gj_component* comp_p,
                     const gv2_gj_component* comp_g, int scheme, double p, int out_kind, double* out,
                     int64_t out_replicate_stride);
int gv2_dev_unpack_tiles(gv2_ctx* ctx, int64_t n, int64_t tile_begin, int64_t tile_end,
                         const double* packed, double* out);
/* K6: bit-pack `tab != 0` into [n_pad][words] uint64 rows; counts per pair tile range. */
int gv2_dev_pack_presence(gv2_ctx* ctx, const double* tab, int64_t n, int64_t p, int64_t ld,
                          uint64_t* bits, int64_t n_pad, int64_t words, int32_t* nnz);
int gv2_dev_shared_counts(gv2_ctx* ctx, const uint64_t* bits, int64_t n, int64_t n_pad,
                          int64_t words, int64_t tile_begin, int64_t tile_end, int32_t* inter);
/* Same counts as [tile_end - tile_begin][128*128] int32 tile blocks (multi-GPU all-gather payload) and their scatter
 * (+ mirror) into inter [n][n]. */
int gv2_dev_shared_counts_packed(gv2_ctx* ctx, const uint64_t* bits, int64_t n, int64_t n_pad, int64_t words,
                                 int64_t tile_begin, int64_t tile_end, int32_t* packed);
int gv2_dev_unpack_tiles_i32(gv2_ctx* ctx, int64_t n, int64_t tile_begin, int64_t tile_end, const int32_t* packed,
                             int32_t* out);
/* POPC-pipe peak of the context's device, measured (register-resident AND + POPC + add chains on every SM):
 * the roofline denominator of the presence / absence path. */
int gv2_dev_popc_peak(gv2_ctx* ctx, int iters, double* popc_per_s);

/* ---- every GPU of the node behind one caller (SURVEY.md section 8b "Threading / processes", 8e) ------------------------
 * Replaces nothing the reference has -- its only parallelism on this path is multiprocessing.Pool(N_CPUs) inside
 * GOMSignatureTable_Constructor (app/utils/parallel_gom_sig_generator.py:18-24) around the single-threaded bootstrap loops of
 * app/src/pl1_graphs.py:74-117 and app/src/virus_classification.py:290-344 -- and keeps that caller single-process:
 * one process, one host thread + one context per GPU inside the library.  Tables: one upload, then a broadcast over NVLink.
 * Replicates: by index (replicate b on GPU b mod N), no exchange.  Un-resampled matrix: one upper-triangular tile range per
 * GPU, packed blocks all-gathered with NCCL (ncclCommInitAll in this process; libnccl.so.2 is dlopen'ed, and without it the
 * blocks travel by cudaMemcpyPeerAsync).  Callbacks run on the CALLING thread, in replicate order. */
typedef struct gv2_multi gv2_multi;
typedef struct gv2_multi_job gv2_multi_job;
int gv2_multi_create(const int* devices /* NULL: all visible */, int ndev, gv2_multi** out);
void gv2_multi_destroy(gv2_multi* m);
const char* gv2_multi_last_error(const gv2_multi* m);
int gv2_multi_device_count(const gv2_multi* m);
int gv2_multi_uses_nccl(const gv2_multi* m);
gv2_ctx* gv2_multi_context(gv2_multi* m, int i); /* the i-th GPU's context (launch counts, timing) */
/* Arguments as gv2_job_create with HOST tables (loc == sig: one table). */
int gv2_multi_job_create(gv2_multi* m, const double* sig, const double* loc, int64_t n, int64_t p, const int32_t* gom_of_row,
                         int64_t n_gom_rows, int64_t g, int scheme, double power, int out_kind, gv2_multi_job** out);
void gv2_multi_job_destroy(gv2_multi_job* job);
int gv2_multi_job_base_host(gv2_multi_job* job, double* host_out /* [n][n] */);
/* host_weights [nb][p]; ring_len <= 0: chosen from the free memory.  Semantics of gv2_job_replicates_newick /
 * gv2_job_replicates_stream (callback mode) respectively. */
int gv2_multi_job_replicates_newick(gv2_multi_job* job, const int32_t* host_weights, int64_t nb, const char* method,
                                    const char* const* leaf_names, int64_t ring_len, gv2_newick_cb callback, void* user);
int gv2_multi_job_replicates_host(gv2_multi_job* job, const int32_t* host_weights, int64_t nb, int64_t ring_len,
                                  gv2_replicate_cb callback, void* user);

/* ---- classification epilogue of pipeline II on a device-resident matrix ------------ */
/* np.max / np.argmax along axis 1 of mat[row_begin:row_end, col_begin:col_end] (first occurrence of the maximum, index
 * relative to col_begin): the 1-nearest-neighbour proposal of TaxonomicAssignmentProposerAndEvaluator
 * (app/utils/classification_utils.py:63-68) on SimMat[-N_ucf:, :N_ref] without shipping the matrix. */
int gv2_dev_rowmax_block(gv2_ctx* ctx, const double* mat, int64_t ld, int64_t row_begin, int64_t row_end,
                         int64_t col_begin, int64_t col_end, double* out_max, int64_t* out_arg);
/* out[k] = mat[rows[k]][cols[k]] (device index arrays): the cells of the intra- / inter-class lists of
 * PairwiseSimilarityScore_Cutoff_Dict_Constructor (classification_utils.py:10-43), in the order the host lists them. */
int gv2_dev_gather_cells(gv2_ctx* ctx, const double* mat, int64_t ld, const int64_t* rows, const int64_t* cols,
                         int64_t count, double* out);

#ifdef __cplusplus
}
#endif
#endif /* GRAVITY2_B200_H */
