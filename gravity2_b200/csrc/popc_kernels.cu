// Presence/absence path (K6 of SURVEY.md §2.2): bit-packed co-occurrence counts.
//
// Reference: shared_pphmm_ratio / shared_norm_pphmm_ratio (app/utils/shared_pphmm_graphs.py:59-97)
// count, with a triple Python loop, the PPHMMs two genomes both hit; the binary Jaccard of
// RefVirusAnnotator.sort_pphmm (app/src/ref_virus_annotator.py:148-158) is the same count over
// the transposed table.  Here `tab != 0` is packed 32 columns per word and the all-pairs count
// is a tiled AND + POPC reduction with the same 128x128 / 8x8-per-thread shape as the Jaccard
// kernel.  Integer arithmetic throughout: results are bit-exact.
#include "common.cuh"

#define PC_THREADS 256
#define PC_BK 32                  // 32-bit words per stage (1024 PPHMM columns)
#define PC_LDS (PC_BK + 4)
#define PC_STAGES 3

__device__ __forceinline__ void pc_cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

__host__ __device__ inline void pc_tile_coords(long long t, int nt, int* bi, int* bj) {
    double disc = (2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)t;
    int r = (int)floor(((2.0 * nt + 1.0) - sqrt(disc)) * 0.5);
    if (r < 0) r = 0;
    while (r > 0 && (long long)r * nt - (long long)r * (r - 1) / 2 > t) --r;
    while ((long long)(r + 1) * nt - (long long)(r + 1) * r / 2 <= t) ++r;
    *bi = r;
    *bj = r + (int)(t - ((long long)r * nt - (long long)r * (r - 1) / 2));
}

// one warp per row: ballot packs 32 columns into one word; words32 is the padded row length
__global__ void pack_presence_kernel(const double* __restrict__ tab, long long n, long long p, long long ld,
                                     unsigned* __restrict__ bits, long long words32, int* __restrict__ nnz) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    unsigned* out = bits + r * words32;
    int cnt = 0;
    if (r < n) {
        const double* row = tab + r * ld;
        for (long long w = 0; w < words32; ++w) {
            long long c = w * 32 + lane;
            double v = c < p ? row[c] : 0.0;
            unsigned m = __ballot_sync(0xffffffffu, v != 0.0);   // NaN != 0 is true, as in numpy
            if (lane == 0) out[w] = m;
            cnt += __popc(m);
        }
        if (lane == 0 && nnz) nnz[r] = cnt;
    } else {
        for (long long w = lane; w < words32; w += 32) out[w] = 0u;
    }
}

__global__ void __launch_bounds__(PC_THREADS, 2)
shared_count_tile_kernel(const unsigned* __restrict__ bits, int ld, int kblocks, int nt, long long tile_begin,
                         long long n, int* __restrict__ inter) {
    extern __shared__ __align__(16) unsigned psm[];
    unsigned* sA = psm;
    unsigned* sB = sA + PC_STAGES * GV2_TILE * PC_LDS;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int bi, bj;
    pc_tile_coords(tile_begin + blockIdx.x, nt, &bi, &bj);
    const unsigned* gA = bits + (size_t)bi * GV2_TILE * ld;
    const unsigned* gB = bits + (size_t)bj * GV2_TILE * ld;
    auto load_stage = [&](int stage, int kb) {
        const int k = kb * PC_BK;
        unsigned* a = sA + stage * GV2_TILE * PC_LDS;
        unsigned* b = sB + stage * GV2_TILE * PC_LDS;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int q = tid + s * PC_THREADS;
            int row = q >> 3, cc = (q & 7) * 4;
            pc_cp_async16(a + row * PC_LDS + cc, gA + (size_t)row * ld + k + cc);
            pc_cp_async16(b + row * PC_LDS + cc, gB + (size_t)row * ld + k + cc);
        }
    };
    int acc[8][8];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) acc[m][nn] = 0;
#pragma unroll
    for (int s = 0; s < PC_STAGES - 1; ++s) {
        if (s < kblocks) load_stage(s, s);
        asm volatile("cp.async.commit_group;\n" ::);
    }
    for (int it = 0; it < kblocks; ++it) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(PC_STAGES - 2));
        __syncthreads();
        {
            int nxt = it + PC_STAGES - 1;
            if (nxt < kblocks) load_stage(nxt % PC_STAGES, nxt);
            asm volatile("cp.async.commit_group;\n" ::);
        }
        const int stage = it % PC_STAGES;
        const unsigned* a = sA + stage * GV2_TILE * PC_LDS + ty * PC_LDS;
        const unsigned* b = sB + stage * GV2_TILE * PC_LDS + tx * PC_LDS;
#pragma unroll
        for (int kg = 0; kg < PC_BK / 4; ++kg) {
            uint4 bf[8];
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) bf[nn] = *reinterpret_cast<const uint4*>(b + nn * 16 * PC_LDS + kg * 4);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                uint4 af = *reinterpret_cast<const uint4*>(a + m * 16 * PC_LDS + kg * 4);
#pragma unroll
                for (int nn = 0; nn < 8; ++nn) {
                    acc[m][nn] += __popc(af.x & bf[nn].x) + __popc(af.y & bf[nn].y) +
                                  __popc(af.z & bf[nn].z) + __popc(af.w & bf[nn].w);
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) {
            long long gi = (long long)bi * GV2_TILE + ty + 16 * m;
            long long gj = (long long)bj * GV2_TILE + tx + 16 * nn;
            if (gi < n && gj < n) {
                inter[gi * n + gj] = acc[m][nn];
                inter[gj * n + gi] = acc[m][nn];
            }
        }
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
    const size_t smem = (size_t)2 * PC_STAGES * GV2_TILE * PC_LDS * sizeof(unsigned);
    if (!(ctx->smem_attr_done & GV2_ATTR_POPC)) {
        GV2_CUDA(ctx, cudaFuncSetAttribute(shared_count_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->smem_attr_done |= GV2_ATTR_POPC;
    }
    shared_count_tile_kernel<<<(unsigned)ntl, PC_THREADS, smem, ctx->stream>>>(
        (const unsigned*)bits, (int)(2 * words), (int)(2 * words / PC_BK), (int)(n_pad / GV2_TILE), tile_begin, n, inter);
    GV2_LAUNCH_CHECK(ctx);
    return GV2_OK;
}
