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

// PACKED == false: counts (and their mirrors) go into inter [n][n];  PACKED == true: inter is [tiles][128*128] row-major
// tile blocks (multi-GPU all-gather payload; scatter with unpack_tiles_i32_kernel)
template <bool PACKED>
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
    if (PACKED) {
        int* o = inter + (size_t)blockIdx.x * (GV2_TILE * GV2_TILE);
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) o[(ty + 16 * m) * GV2_TILE + tx + 16 * nn] = acc[m][nn];
        return;
    }
    // direct rows: 16 consecutive ints per (m, nn); mirror through shared memory (the stage ring is dead) so that the
    // transposed tile also leaves as whole 64-byte runs instead of one 4-byte store per 4 n bytes
    __syncthreads();
    int* tr = reinterpret_cast<int*>(psm);                           // [128][129]
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) {
            const int r = ty + 16 * m, c = tx + 16 * nn;
            const long long gi = (long long)bi * GV2_TILE + r, gj = (long long)bj * GV2_TILE + c;
            if (gi < n && gj < n) inter[gi * n + gj] = acc[m][nn];
            tr[c * 129 + r] = acc[m][nn];
        }
    if (bi == bj) return;                                            // diagonal tiles hold both triangles already
    __syncthreads();
    for (int q = tid; q < GV2_TILE * GV2_TILE; q += PC_THREADS) {
        const int c = q >> 7, r = q & 127;
        const long long gi = (long long)bi * GV2_TILE + r, gj = (long long)bj * GV2_TILE + c;
        if (gi < n && gj < n) inter[gj * n + gi] = tr[c * 129 + r];
    }
}

// packed int32 tiles [ntl][128*128] -> symmetric n x n.  grid = tiles * 16, block (32, 8)
__global__ void __launch_bounds__(256)
unpack_tiles_i32_kernel(long long n, int nt, long long tile_begin, const int* __restrict__ packed, int* __restrict__ out) {
    __shared__ int tr[32][33];
    const long long t = blockIdx.x >> 4;
    const int sub = blockIdx.x & 15, sr = sub >> 2, sc = sub & 3;
    int bi, bj;
    pc_tile_coords(tile_begin + t, nt, &bi, &bj);
    const int lx = threadIdx.x;
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const int ly = threadIdx.y + yy * 8;
        const int r = sr * 32 + ly, c = sc * 32 + lx;
        const long long gi = (long long)bi * GV2_TILE + r, gj = (long long)bj * GV2_TILE + c;
        const int v = packed[t * (long long)(GV2_TILE * GV2_TILE) + r * GV2_TILE + c];
        if (gi < n && gj < n) out[gi * n + gj] = v;
        tr[ly][lx] = v;
    }
    if (bi == bj) return;
    __syncthreads();
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const int ly = threadIdx.y + yy * 8;
        const long long gj = (long long)bj * GV2_TILE + sc * 32 + ly;
        const long long gi = (long long)bi * GV2_TILE + sr * 32 + lx;
        if (gi < n && gj < n) out[gj * n + gi] = tr[lx][ly];
    }
}

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
