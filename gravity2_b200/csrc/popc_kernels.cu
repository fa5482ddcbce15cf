// Presence/absence path (K6 of SURVEY.md §2.2): bit-packed co-occurrence counts.
//
// Reference: shared_pphmm_ratio / shared_norm_pphmm_ratio (app/utils/shared_pphmm_graphs.py:59-97)
// count, with a triple Python loop, the PPHMMs two genomes both hit; the binary Jaccard of
// RefVirusAnnotator.sort_pphmm (app/src/ref_virus_annotator.py:148-158) is the same count over
// the transposed table.  Here `tab != 0` is packed 32 columns per word and the all-pairs count
// is a tiled AND + POPC reduction with the same 128x128 / 8x8-per-thread shape as the Jaccard
// kernel.  Integer arithmetic throughout: results are bit-exact.
#include "common.cuh"

#include <algorithm>

#include "popc_kernels.cuh"

// POPC-pipe peak of this device, measured: register-resident LOP3 + POPC + IADD chains (8 independent per thread), no
// memory traffic.  The roofline denominator of the presence / absence path (SURVEY.md section 8d asks for a micro-benchmark
// instead of the programming guide's table figure).
__global__ void __launch_bounds__(256)
popc_peak_kernel(int iters, unsigned seed, unsigned* __restrict__ sink) {
    unsigned a[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = seed * (threadIdx.x + 1 + 977 * k) + blockIdx.x; acc[k] = 0; }
    unsigned m = seed | 0x55aa55aau;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            acc[k] += __popc(a[k] & m);
            a[k] += acc[k];                      // keeps the AND operand changing: nothing hoists out of the loop
        }
        m = (m << 1) | (m >> 31);
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 0x12345u) sink[0] = s;
}

static int pc_launch(gv2_ctx* ctx, const uint64_t* bits, int64_t n, int64_t n_pad, int64_t words, int64_t tile_begin,
                     int64_t tile_end, int packed, int32_t* out) {
    const int64_t ntl = tile_end - tile_begin;
    if (ntl == 0) return GV2_OK;
    // stage ring, reused as the [128][129] int transpose of the mirror
    const size_t smem = std::max((size_t)2 * PC_STAGES * GV2_TILE * PC_LDS * sizeof(unsigned), (size_t)GV2_TILE * 129 * sizeof(int));
    if (!(ctx->smem_attr_done & GV2_ATTR_POPC)) {
        GV2_CUDA(ctx, cudaFuncSetAttribute(shared_count_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GV2_CUDA(ctx, cudaFuncSetAttribute(shared_count_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smem_attr_done |= GV2_ATTR_POPC;
    }
    const int nt = (int)(n_pad / GV2_TILE), ld = (int)(2 * words), kb = (int)(2 * words / PC_BK);
    for (int64_t t0 = 0; t0 < ntl; t0 += 1 << 30) {                  // grid.x holds 2^31 - 1 blocks
        const int64_t tc = std::min<int64_t>(1 << 30, ntl - t0);
        if (packed)
            shared_count_tile_kernel<true><<<(unsigned)tc, PC_THREADS, smem, ctx->stream>>>((const unsigned*)bits, ld, kb, nt, tile_begin + t0, n,
                                                                                          out + t0 * (GV2_TILE * GV2_TILE));
        else
            shared_count_tile_kernel<false><<<(unsigned)tc, PC_THREADS, smem, ctx->stream>>>((const unsigned*)bits, ld, kb, nt, tile_begin + t0, n, out);
        GV2_LAUNCH_CHECK(ctx);
    }
    return GV2_OK;
}

extern "C" int gv2_dev_pack_presence(gv2_ctx* ctx, const double* tab, int64_t n, int64_t p, int64_t ld,
                                     uint64_t* bits, int64_t n_pad, int64_t words, int32_t* nnz) {
    // `words` counts 64-bit words per row; rows are addressed as 2*words 32-bit words
    if (!ctx || !tab || !bits || n_pad < n || (n_pad % GV2_TILE) || words * 64 < p || ((2 * words) % PC_BK))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_pack_presence: bad arguments (words*2 must be a multiple of %d)", PC_BK);
    if (n_pad == 0) return GV2_OK;
    pack_presence_kernel<<<(unsigned)((n_pad + 7) / 8), 256, 0, ctx->stream>>>(tab, n, p, ld, (unsigned*)bits, 2 * words, nnz);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}

extern "C" int gv2_dev_shared_counts(gv2_ctx* ctx, const uint64_t* bits, int64_t n, int64_t n_pad, int64_t words,
                                     int64_t tile_begin, int64_t tile_end, int32_t* inter) {
    if (!ctx || !bits || !inter || tile_end < tile_begin || (n_pad % GV2_TILE) || ((2 * words) % PC_BK))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_shared_counts: bad arguments");
    const int64_t ntl = tile_end - tile_begin;
    if (ntl == 0) return GV2_OK;
    return pc_launch(ctx, bits, n, n_pad, words, tile_begin, tile_end, 0, inter);
}

extern "C" int gv2_dev_shared_counts_packed(gv2_ctx* ctx, const uint64_t* bits, int64_t n, int64_t n_pad, int64_t words,
                                            int64_t tile_begin, int64_t tile_end, int32_t* packed) {
    if (!ctx || !bits || !packed || tile_end < tile_begin || (n_pad % GV2_TILE) || ((2 * words) % PC_BK))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_shared_counts_packed: bad arguments");
    return pc_launch(ctx, bits, n, n_pad, words, tile_begin, tile_end, 1, packed);
}

extern "C" int gv2_dev_unpack_tiles_i32(gv2_ctx* ctx, int64_t n, int64_t tile_begin, int64_t tile_end, const int32_t* packed,
                                        int32_t* out) {
    if (!ctx || !packed || !out || tile_end < tile_begin) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_unpack_tiles_i32: bad arguments");
    const int64_t ntl = tile_end - tile_begin;
    if (ntl == 0 || n == 0) return GV2_OK;
    dim3 block(32, 8);
    for (int64_t t0 = 0; t0 < ntl; t0 += 1 << 20) {                  // grid.x limit: 2^31 - 1 blocks of 16 per tile
        const int64_t tc = std::min<int64_t>(1 << 20, ntl - t0);
        unpack_tiles_i32_kernel<<<(unsigned)(tc * 16), block, 0, ctx->stream>>>(n, (int)((n + GV2_TILE - 1) / GV2_TILE), tile_begin + t0,
                                                                                 packed + t0 * (GV2_TILE * GV2_TILE), out);
        GV2_LAUNCH_CHECK(ctx);
    }
    return GV2_OK;
}

extern "C" int gv2_dev_popc_peak(gv2_ctx* ctx, int iters, double* popc_per_s) {
    if (!ctx || !popc_per_s || iters <= 0) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_popc_peak: bad arguments");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned* sink = nullptr;
    GV2_CUDA(ctx, cudaMalloc((void**)&sink, sizeof(unsigned)));
    cudaEvent_t a, b;
    GV2_CUDA(ctx, cudaEventCreate(&a));
    GV2_CUDA(ctx, cudaEventCreate(&b));
    const unsigned blocks = (unsigned)ctx->sm_count * 8;             // 8 CTAs of 256 threads per SM: every SMSP saturated
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {                              // first pass warms up; best of the rest
        cudaEventRecord(a, ctx->stream);
        popc_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(iters, 0x9e3779b9u + rep, sink);
        cudaEventRecord(b, ctx->stream);
        ctx->launches++;
        GV2_CUDA(ctx, cudaEventSynchronize(b));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    *popc_per_s = (double)blocks * 256.0 * 8.0 * iters / (best * 1e-3);
    return GV2_OK;
}
