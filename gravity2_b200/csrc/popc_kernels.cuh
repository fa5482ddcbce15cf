// The presence / absence kernels (pack, all-pairs AND + POPC tiles, tile unpacking).  popc_kernels.cu includes this file; so
// does tests/emul/popc_emul.cpp, which runs the SAME kernel source on host threads (cuda_block_emul.h) so that the kernels'
// logic -- tile coordinates, ragged edges, the stage ring, the mirror through shared memory -- is checked against the
// CPU restatement of the reference where there is no GPU.
#pragma once
#ifndef GV2_TILE
#define GV2_TILE 128
#endif

#define PC_THREADS 256
#define PC_BK 32                  // 32-bit words per stage (1024 PPHMM columns)
#define PC_LDS (PC_BK + 4)
#define PC_STAGES 3

#ifdef PC_EMULATED              // tests/emul: cp.async as cuda_block_emul.h models it
static inline void pc_cp_async16(void* smem, const void* gmem) { emu_cp_async(smem, gmem, 16); }
#define PC_CP_COMMIT() emu_cp_commit()
#define PC_CP_WAIT(n) emu_cp_wait(n)
#define PC_DYNAMIC_SMEM(name) unsigned* name = reinterpret_cast<unsigned*>(emu_dynamic_smem)
#else
__device__ __forceinline__ void pc_cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
#define PC_CP_COMMIT() asm volatile("cp.async.commit_group;\n" ::)
#define PC_CP_WAIT(n) asm volatile("cp.async.wait_group %0;\n" ::"n"(n))
#define PC_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned name[]
#endif

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
    PC_DYNAMIC_SMEM(psm);
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
        PC_CP_COMMIT();
    }
    for (int it = 0; it < kblocks; ++it) {
        PC_CP_WAIT(PC_STAGES - 2);
        __syncthreads();
        {
            int nxt = it + PC_STAGES - 1;
            if (nxt < kblocks) load_stage(nxt % PC_STAGES, nxt);
            PC_CP_COMMIT();
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

