// The sparse replicate pair-sum kernels: CSR of the 48-bit fixed-point table, pair lists (count + fill), transposed
// multiplicities, the streaming and pair-lane reduce kernels, row sums (and the on-the-fly kernel for tables whose lists do
// not fit).  boot_sparse.cu includes this file; so does tests/emul/bs_emul.cpp, which runs the SAME kernel source of the
// default (list) path on host threads (cuda_block_emul.h) and compares the pair sums BIT FOR BIT with integer arithmetic
// where there is no GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef GV2_TILE
#define GV2_TILE 128
#endif
#ifdef BS_EMULATED
#define BS_DYNAMIC_SMEM(name) uint8_t* name = reinterpret_cast<uint8_t*>(emu_dynamic_smem)
#else
#define BS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

#define BS_THREADS 256
#define BS_WIN 32768                    // columns per mask window (uint16 masks: <= 64 KB)
#define BS_JCAP 256                     // j-row columns kept in shared memory (longer rows: tail searched in global memory)
#define BS_POOL 2048                    // list entries per match pass (a 125-entry slice of one row can never overflow it)
#define BS_LO_BITS 16
static inline size_t bs_smem_bytes(int kwin) {
    return (size_t)kwin * 2 + 16 * BS_JCAP * 4 + BS_POOL * 8 + (256 + 256 + 260 + 16) * sizeof(int);
}

__host__ __device__ inline void bs_tile_coords(long long t, int nt, int* bi, int* bj) {
    double disc = (2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)t;
    int r = (int)floor(((2.0 * nt + 1.0) - sqrt(disc)) * 0.5);
    if (r < 0) r = 0;
    while (r > 0 && (long long)r * nt - (long long)r * (r - 1) / 2 > t) --r;
    while ((long long)(r + 1) * nt - (long long)(r + 1) * r / 2 <= t) ++r;
    *bi = r;
    *bj = r + (int)(t - ((long long)r * nt - (long long)r * (r - 1) / 2));
}

// exact integer limb sums -> fp64, with explicitly rounded operations so that every kernel (on-the-fly, streaming, row sums)
// produces the same bits from the same integers (duplicate genomes: pair sum == row sum exactly)
__device__ __forceinline__ double bs_to_double(unsigned long long hi, unsigned long long lo, double q_hi, double quantum) {
    return __fma_rn((double)hi, q_hi, __dmul_rn((double)lo, quantum));
}

struct BsCsr {
    const long long* ptr;     // [n_pad + 1]
    const int* col;           // [nnz] ascending within a row
    const uint32_t* hi;       // [nnz] bits 47..16 of the 48-bit fixed-point value
    const uint16_t* lo;       // [nnz] bits 15..0
};

// position of column c in a sorted row whose first `slen` columns are in shared memory, the rest in global memory
__device__ __forceinline__ int bs_find(const int* scol, int slen, const int* gcol, int len, int c) {
    int lo = 0, hi = slen - 1;
    const int* col = scol;
    int off = 0;
    if (slen == 0 || c > scol[slen - 1]) { col = gcol + slen; off = slen; hi = len - slen - 1; }
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (col[mid] < c) lo = mid + 1; else hi = mid;
    }
    return off + lo;
}

#ifndef BS_EMULATED            // (the on-the-fly kernel uses cub::BlockScan; the emulation covers the default list path)
// grid = (64 sub-blocks, genome tiles in [tile_begin, tile_end), replicate blocks of RL)
// RL = replicate lanes per slot (one replicate per lane), PP = 256 / RL slots sharing the i-rows of a batch
template <int RL>
__global__ void __launch_bounds__(BS_THREADS, 2)
boot_sparse_kernel(BsCsr csr, int nt, long long tile_begin, const uint8_t* __restrict__ wT, long long w_ld, int kwin, int nwin,
                   int nb_total, long long ws_rep_stride, double quantum, double* __restrict__ ws) {
    constexpr int PP = BS_THREADS / RL;
    BS_DYNAMIC_SMEM(smem);
    uint32_t* mask32 = (uint32_t*)smem;                              // [kwin / 2]: two uint16 j-row masks per word
    int* jcol = (int*)(smem + (size_t)kwin * 2);                     // [16][BS_JCAP]
    uint2* pool = (uint2*)(jcol + 16 * BS_JCAP);                     // [BS_POOL] (window column | lo << 16, hi)
    int* cnt = (int*)(pool + BS_POOL);                               // [256] hits per pair of the batch
    int* fillc = cnt + 256;                                          // [256]
    int* base = fillc + 256;                                         // [257] list offsets (multiples of 4)
    int* jlen = base + 260;                                          // [16] j-row lengths
    __shared__ typename cub::BlockScan<int, BS_THREADS>::TempStorage scan_tmp;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = blockIdx.x, si = sub >> 3, sj = sub & 7;
    int bi, bj;
    bs_tile_coords(tile_begin + blockIdx.y, nt, &bi, &bj);
    if (bi == bj && si > sj) return;                                 // block entirely below the diagonal
    const bool diag_block = bi == bj && si == sj;
    const int slot = tid / RL, rl = tid % RL;
    const int rep0 = blockIdx.z * RL + rl;                           // this lane's replicate
    const long long rowI0 = (long long)bi * GV2_TILE + si * 16, rowJ0 = (long long)bj * GV2_TILE + sj * 16;
    double* wsblk = ws + (size_t)blockIdx.y * (GV2_TILE * GV2_TILE) + (size_t)(((si << 3) + sj) << 8);
    const double q_hi = quantum * 65536.0;

    for (int win = 0; win < nwin; ++win) {
        const int c_lo = win * BS_WIN;
        // ---- build: j-row masks of this column window (+ the shared-memory copy of the column lists)
        __syncthreads();
        for (int q = tid; q < kwin / 2; q += BS_THREADS) mask32[q] = 0u;
        __syncthreads();
        for (int h = 0; h < 2; ++h) {
            const int jj = warp + 8 * h;
            const long long j0 = csr.ptr[rowJ0 + jj], j1 = csr.ptr[rowJ0 + jj + 1];
            if (lane == 0) jlen[jj] = (int)(j1 - j0);
            for (long long e = j0 + lane; e < j1; e += 32) {
                const int c = csr.col[e];
                if (e - j0 < BS_JCAP) jcol[jj * BS_JCAP + (int)(e - j0)] = c;
                const int cw = c - c_lo;
                if ((unsigned)cw < (unsigned)kwin) atomicOr(&mask32[cw >> 1], 1u << (jj + 16 * (cw & 1)));
            }
        }
        __syncthreads();

        // ---- batches of i-rows [ra, rb); a single row may be sliced further to entries [cs0, cs1)
        int ra = 0, batch = 16;
        long long cs0 = -1, cs1 = -1;                                  // slice of row ra (cs0 < 0: whole rows)
        bool row_started = false;                                      // an earlier slice of row ra has been stored
        while (ra < 16) {
            const int rb = cs0 >= 0 ? ra + 1 : min(16, ra + batch);
            const int nrows = rb - ra;
            cnt[tid] = 0;
            fillc[tid] = 0;
            __syncthreads();
            for (int r = 0; r < nrows; ++r) {
                const int ii = ra + r;
                long long e0 = csr.ptr[rowI0 + ii], e1 = csr.ptr[rowI0 + ii + 1];
                if (cs0 >= 0) { e0 = cs0; e1 = cs1; }
                const unsigned keep = diag_block ? (0xffffu << ii) & 0xffffu : 0xffffu;   // pairs below the diagonal are never read
                for (long long e = e0 + tid; e < e1; e += BS_THREADS) {
                    const int cw = csr.col[e] - c_lo;
                    if ((unsigned)cw < (unsigned)kwin) {
                        unsigned m = (mask32[cw >> 1] >> (16 * (cw & 1))) & keep;
                        while (m) { const int jj = __ffs(m) - 1; m &= m - 1; atomicAdd(&cnt[r * 16 + jj], 1); }
                    }
                }
            }
            __syncthreads();
            {
                const int padded = tid < nrows * 16 ? (cnt[tid] + 3) & ~3 : 0;
                int off, total;
                cub::BlockScan<int, BS_THREADS>(scan_tmp).ExclusiveSum(padded, off, total);
                base[tid] = off;
                if (tid == 0) base[256] = total;
            }
            __syncthreads();
            if (base[256] > BS_POOL) {                                 // uniform decision: shrink the batch and retry
                if (nrows > 1) batch = nrows / 2;
                else {
                    if (cs0 < 0) { cs0 = csr.ptr[rowI0 + ra]; cs1 = csr.ptr[rowI0 + ra + 1]; }
                    cs1 = cs0 + (cs1 - cs0) / 2;
                }
                __syncthreads();
                continue;
            }
            for (int r = 0; r < nrows; ++r) {
                const int ii = ra + r;
                long long e0 = csr.ptr[rowI0 + ii], e1 = csr.ptr[rowI0 + ii + 1];
                if (cs0 >= 0) { e0 = cs0; e1 = cs1; }
                const unsigned keep = diag_block ? (0xffffu << ii) & 0xffffu : 0xffffu;
                for (long long e = e0 + tid; e < e1; e += BS_THREADS) {
                    const int c = csr.col[e];
                    const int cw = c - c_lo;
                    if ((unsigned)cw < (unsigned)kwin) {
                        unsigned m = (mask32[cw >> 1] >> (16 * (cw & 1))) & keep;
                        if (m) {
                            const unsigned long long fi = ((unsigned long long)csr.hi[e] << BS_LO_BITS) | csr.lo[e];
                            while (m) {
                                const int jj = __ffs(m) - 1;
                                m &= m - 1;
                                const long long j0 = csr.ptr[rowJ0 + jj];
                                const int pos = bs_find(jcol + jj * BS_JCAP, min(jlen[jj], BS_JCAP), csr.col + j0, jlen[jj], c);
                                const unsigned long long fj = ((unsigned long long)csr.hi[j0 + pos] << BS_LO_BITS) | csr.lo[j0 + pos];
                                const unsigned long long f = min(fi, fj);
                                const int q = r * 16 + jj;
                                pool[base[q] + atomicAdd(&fillc[q], 1)] =
                                    make_uint2((uint32_t)cw | ((uint32_t)(f & 0xffffu) << 16), (uint32_t)(f >> BS_LO_BITS));
                            }
                        }
                    }
                }
            }
            // pad every list to a multiple of 4 entries with (column 0, value 0): the reduce loop runs 4 entries a step
            if (tid < nrows * 16) {
                const int end = tid + 1 < 256 ? base[tid + 1] : base[256];
                for (int k = base[tid] + cnt[tid]; k < end; ++k) pool[k] = make_uint2(0u, 0u);
            }
            __syncthreads();
            // ---- reduce: lane = one replicate, slot s takes i-rows s, s + PP, ... of the batch; per i-row the 16 lists
            // are walked 4 entries a step (one byte of the transposed multiplicities per entry) into 16 accumulator
            // pairs, then the 16 sums leave as one whole 128-byte line of the replicate's workspace
            const bool first = win == 0 && !row_started;
            const uint8_t* wTl = wT + (size_t)c_lo * w_ld + rep0;       // this lane's column of the window's multiplicities
            const uint32_t wld32 = (uint32_t)w_ld;                     // window columns x row stride < 2^32 (checked at launch)
            for (int r = slot; r < nrows; r += PP) {
                unsigned long long ahi[16];
                uint32_t alo[16];
                const int beg = base[r * 16], end = r * 16 + 16 < 256 ? base[r * 16 + 16] : base[256];
                // the multiplicity bytes of step e + 4 are in flight while step e multiplies; the entries themselves are
                // re-read from shared memory (two LDS.128 a step) rather than carried in registers
                uint32_t wb[4];
                if (beg < end) {
                    const uint4 p0 = *reinterpret_cast<const uint4*>(pool + beg), p1 = *reinterpret_cast<const uint4*>(pool + beg + 2);
                    wb[0] = __ldg(wTl + (p0.x & 0xffffu) * wld32);
                    wb[1] = __ldg(wTl + (p0.z & 0xffffu) * wld32);
                    wb[2] = __ldg(wTl + (p1.x & 0xffffu) * wld32);
                    wb[3] = __ldg(wTl + (p1.z & 0xffffu) * wld32);
                }
                int e = beg;
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    const int ne = r * 16 + a + 1 < 256 ? base[r * 16 + a + 1] : base[256];
                    unsigned long long h = 0ull;
                    uint32_t l = 0u;
                    for (; e < ne; e += 4) {                           // lists are padded: a whole step belongs to one pair
                        const uint32_t w0 = wb[0], w1 = wb[1], w2 = wb[2], w3 = wb[3];
                        if (e + 4 < end) {
                            const uint4 n0 = *reinterpret_cast<const uint4*>(pool + e + 4), n1 = *reinterpret_cast<const uint4*>(pool + e + 6);
                            wb[0] = __ldg(wTl + (n0.x & 0xffffu) * wld32);
                            wb[1] = __ldg(wTl + (n0.z & 0xffffu) * wld32);
                            wb[2] = __ldg(wTl + (n1.x & 0xffffu) * wld32);
                            wb[3] = __ldg(wTl + (n1.z & 0xffffu) * wld32);
                        }
                        const uint4 c0 = *reinterpret_cast<const uint4*>(pool + e), c1 = *reinterpret_cast<const uint4*>(pool + e + 2);
                        h += (unsigned long long)c0.y * w0;
                        h += (unsigned long long)c0.w * w1;
                        h += (unsigned long long)c1.y * w2;
                        h += (unsigned long long)c1.w * w3;
                        l += (c0.x >> 16) * w0 + (c0.z >> 16) * w1 + (c1.x >> 16) * w2 + (c1.z >> 16) * w3;
                    }
                    ahi[a] = h;
                    alo[a] = l;
                }
                // exact integers -> fp64; the first pass over a row stores, later ones (row slices, further column
                // windows) add: same thread, program order, deterministic
                if ((first || beg < end) && rep0 < nb_total) {
                    double* o = wsblk + (size_t)rep0 * ws_rep_stride + ((ra + r) << 4);
#pragma unroll
                    for (int a = 0; a < 16; a += 2) {
                        double v0 = bs_to_double(ahi[a], alo[a], q_hi, quantum);
                        double v1 = bs_to_double(ahi[a + 1], alo[a + 1], q_hi, quantum);
                        if (!first) {
                            const double2 p = *reinterpret_cast<const double2*>(o + a);
                            v0 += p.x; v1 += p.y;
                        }
                        *reinterpret_cast<double2*>(o + a) = make_double2(v0, v1);
                    }
                }
            }
            __syncthreads();
            // ---- advance
            if (cs0 >= 0 && cs1 < csr.ptr[rowI0 + ra + 1]) {            // more slices of this row
                cs0 = cs1;
                cs1 = csr.ptr[rowI0 + ra + 1];
                row_started = true;
            } else {
                ra = rb;
                batch = 16 - ra;
                cs0 = cs1 = -1;
                row_started = false;
            }
        }
    }
}

#endif  // BS_EMULATED

// ---------------------------------------------------------------------------------- precomputed pair lists
// The match phase does not depend on the replicate, so for tables whose lists fit in memory (cfg3: 2.2e8 entries,
// 2.4 GB) it runs ONCE per job: the co-present (column, min value) lists of every genome pair are written to global
// memory, and the per-launch kernel only streams them.  Without the 60 KB mask the streaming kernel holds three CTAs
// per SM instead of two and has no barriers at all.
//   rows:   one per (16 x 16 block, i-row): rowptr[row] = first entry; pend[row][jj] = end of pair (i, jj)'s list
//           relative to rowptr[row] (lists back to back, each padded to a multiple of 4 entries with zero entries)
//   entries: 8 bytes each, stored as planar GROUPS of four (32 bytes): uint16 column x 4 (needs k_pad <= 65536), then
//           byte k (k = 0..5) of the four 48-bit values as one 32-bit word each
struct BsLists {
    long long* rowptr = nullptr;   // [nrows + 1], in entries (multiples of 4)
    uint32_t* pend = nullptr;      // [nrows][16]
    uint8_t* ent = nullptr;        // [rowptr[nrows] * 8]
    long long nrows = 0, nent = 0;
    size_t bytes = 0;
};

#define BS_LIST_SMEM(KP) ((size_t)(KP) * 2 + 16 * BS_JCAP * 4 + (256 + 256 + 16) * sizeof(int))

// FILL == false: pend (cumulative, padded) and rowtot per row.  FILL == true: entries at rowptr.  grid = (64, tiles)
template <bool FILL>
__global__ void __launch_bounds__(BS_THREADS, 1)
bs_lists_build_kernel(BsCsr csr, int nt, int k_pad, long long* __restrict__ rowtot, const long long* __restrict__ rowptr,
                      uint32_t* __restrict__ pend, uint8_t* __restrict__ ent) {
    BS_DYNAMIC_SMEM(smem);
    uint32_t* mask32 = (uint32_t*)smem;                              // [k_pad / 2]
    int* jcol = (int*)(smem + (size_t)k_pad * 2);                    // [16][BS_JCAP]
    int* cnt = jcol + 16 * BS_JCAP;                                  // [256]
    int* fillc = cnt + 256;                                          // [256]
    int* jlen = fillc + 256;                                         // [16]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = blockIdx.x, si = sub >> 3, sj = sub & 7;
    int bi, bj;
    bs_tile_coords(blockIdx.y, nt, &bi, &bj);
    const long long row0 = ((long long)blockIdx.y * 64 + sub) * 16;
    if (bi == bj && si > sj) {                                       // block entirely below the diagonal: empty rows
        if (!FILL && tid < 16) rowtot[row0 + tid] = 0;
        if (!FILL) pend[row0 * 16 + tid] = 0;
        return;
    }
    const bool diag_block = bi == bj && si == sj;
    const long long rowI0 = (long long)bi * GV2_TILE + si * 16, rowJ0 = (long long)bj * GV2_TILE + sj * 16;
    for (int q = tid; q < k_pad / 2; q += BS_THREADS) mask32[q] = 0u;
    cnt[tid] = 0;
    fillc[tid] = 0;
    __syncthreads();
    for (int h = 0; h < 2; ++h) {
        const int jj = warp + 8 * h;
        const long long j0 = csr.ptr[rowJ0 + jj], j1 = csr.ptr[rowJ0 + jj + 1];
        if (lane == 0) jlen[jj] = (int)(j1 - j0);
        for (long long e = j0 + lane; e < j1; e += 32) {
            const int c = csr.col[e];
            if (e - j0 < BS_JCAP) jcol[jj * BS_JCAP + (int)(e - j0)] = c;
            atomicOr(&mask32[c >> 1], 1u << (jj + 16 * (c & 1)));
        }
    }
    __syncthreads();
    for (int ii = 0; ii < 16; ++ii) {
        const long long e0 = csr.ptr[rowI0 + ii], e1 = csr.ptr[rowI0 + ii + 1];
        const unsigned keep = diag_block ? (0xffffu << ii) & 0xffffu : 0xffffu;   // pairs below the diagonal are never read
        const long long rbase = FILL ? rowptr[row0 + ii] : 0;
        for (long long e = e0 + tid; e < e1; e += BS_THREADS) {
            const int c = csr.col[e];
            unsigned m = (mask32[c >> 1] >> (16 * (c & 1))) & keep;
            if (!m) continue;
            if (!FILL) {
                while (m) { const int jj = __ffs(m) - 1; m &= m - 1; atomicAdd(&cnt[ii * 16 + jj], 1); }
            } else {
                const unsigned long long fi = ((unsigned long long)csr.hi[e] << BS_LO_BITS) | csr.lo[e];
                while (m) {
                    const int jj = __ffs(m) - 1;
                    m &= m - 1;
                    const long long j0 = csr.ptr[rowJ0 + jj];
                    const int pos = bs_find(jcol + jj * BS_JCAP, min(jlen[jj], BS_JCAP), csr.col + j0, jlen[jj], c);
                    const unsigned long long fj = ((unsigned long long)csr.hi[j0 + pos] << BS_LO_BITS) | csr.lo[j0 + pos];
                    const unsigned long long f = min(fi, fj);
                    const int q = ii * 16 + jj;
                    const long long at = rbase + (jj ? pend[(row0 + ii) * 16 + jj - 1] : 0u) + atomicAdd(&fillc[q], 1);
                    // planar group of four entries (32 bytes): four uint16 columns, then six words holding byte k of the
                    // four 48-bit values (the operand layout of the dp4a in the streaming kernel)
                    uint8_t* gp = ent + (at >> 2) * 32;
                    const int slot = (int)(at & 3);
                    reinterpret_cast<uint16_t*>(gp)[slot] = (uint16_t)c;
#pragma unroll
                    for (int k = 0; k < 6; ++k) gp[8 + 4 * k + slot] = (uint8_t)(f >> (8 * k));
                }
            }
        }
    }
    __syncthreads();
    if (!FILL) {
        // thread (ii, jj): padded cumulative end of its pair's list within the row (16-wide inclusive scan by shuffles)
        const int ii = tid >> 4, jj = tid & 15;
        int v = (cnt[tid] + 3) & ~3;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o, 16);
            if (jj >= o) v += u;
        }
        pend[(row0 + ii) * 16 + jj] = (uint32_t)v;
        if (jj == 15) rowtot[row0 + ii] = v;
    }
    // (FILL: the entry buffer was zeroed before the launch, so the padding of every list is already (column 0, value 0))
}

// Streams the precomputed lists: S[b][i][j] = sum over the pair's list of value x multiplicity, for 128 replicates per
// warp (lane = four replicates).  Everything the arithmetic waits for arrives asynchronously, per warp, with no block
// barrier anywhere:
//   lists    cp.async (LDGSTS) 8 bytes per lane, chunks of four groups (16 entries, 128 bytes), two chunks ahead;
//   weights  per entry the 128 multiplicity bytes of the warp's replicates, gathered by column from the transposed
//            table wT[c][replicate] with cp.async 16 bytes per lane (eight lanes per entry, four entries per
//            instruction), one chunk ahead -- the dependent list -> column -> multiplicity chain of the previous
//            kernel (two L2 latencies per step and warp) is off the critical path;
//   math     per group of four entries: the 4 x 4 multiplicity bytes are transposed in registers (8 PRMT) and each of the
//            six value-byte planes is one dp4a per replicate (24 dp4a for 16 products); plane sums stay in 32-bit integers
//            (<= 255 * sum_c w_c <= 2^24) and are recombined into the same high / low limbs as before, so the results are
//            bit-identical to the on-the-fly kernel's.
// A warp owns 16 / RG consecutive i-rows of a 16 x 16 pair block (their lists are contiguous) for its 128 replicates;
// block = RW replicate warps x RG row groups (RW * RG = 8).  grid = (64 sub-blocks, tiles, replicate blocks of 128 RW)
#define BSS_CH 4                                     // groups per chunk
#define BSS_LIST_BYTES (BSS_CH * 32)                 // 128
#define BSS_W_BYTES (BSS_CH * 4 * 128)               // 2048
#define BSS_OUT_LD 132                               // doubles per staged pair: 128 replicates + 4 (bank spread)
#define BSS_OUT_PAIRS 8                              // pairs staged per flush: half an i-row, 64 contiguous bytes per replicate
                                                     // (4 pairs / 32-byte requests / three CTAs per SM measured 1.65x slower)
#define BSS_OUT_BYTES (BSS_OUT_PAIRS * BSS_OUT_LD * 8)
#define BSS_WARP_BYTES (3 * BSS_LIST_BYTES + 2 * BSS_W_BYTES + BSS_OUT_BYTES)

#ifdef BS_EMULATED             // tests/emul: cp.async as cuda_block_emul.h models it (the __syncwarp after every wait hands the data to the warp)
static inline void bss_cp_async8(void* smem, const void* gmem) { emu_cp_async(smem, gmem, 8); }
static inline void bss_cp_async16(void* smem, const void* gmem) { emu_cp_async(smem, gmem, 16); }
static inline void bss_commit() { emu_cp_commit(); }
static inline void bss_wait2() { emu_cp_wait(2); }
static inline void bss_wait1() { emu_cp_wait(1); }
#else
__device__ __forceinline__ void bss_cp_async8(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void bss_cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void bss_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void bss_wait2() { asm volatile("cp.async.wait_group 2;\n" ::: "memory"); }
__device__ __forceinline__ void bss_wait1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
#endif

template <int RW>
__global__ void __launch_bounds__(BS_THREADS, 2)
boot_sparse_stream_kernel(const long long* __restrict__ rowptr, const uint32_t* __restrict__ pend, const uint8_t* __restrict__ ent,
                          int nt, long long tile_begin, const uint8_t* __restrict__ wT, long long w_ld, int nb_total,
                          long long ws_rep_stride, double quantum, double* __restrict__ ws) {
    constexpr int RG = 8 / RW;                       // row groups per block
    constexpr int NR = 16 / RG;                      // i-rows per warp
    BS_DYNAMIC_SMEM(smem);   // 8 * BSS_WARP_BYTES
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rw = warp % RW, rg = warp / RW;
    const int sub = blockIdx.x, si = sub >> 3, sj = sub & 7;
    int bi, bj;
    bs_tile_coords(tile_begin + blockIdx.y, nt, &bi, &bj);
    if (bi == bj && si > sj) return;                 // block entirely below the diagonal (never read)
    uint8_t* lbuf = smem + warp * BSS_WARP_BYTES;    // three list chunks
    uint8_t* wbuf = lbuf + 3 * BSS_LIST_BYTES;       // two chunks of multiplicities: [entry][128 replicates]
    double* obuf = reinterpret_cast<double*>(wbuf + 2 * BSS_W_BYTES);   // [pair & 7][replicate]: results of eight pairs
    const long long row_first = ((tile_begin + blockIdx.y) * 64 + sub) * 16 + rg * NR;
    const long long e_begin = rowptr[row_first];
    const int ngroups = (int)((rowptr[row_first + NR] - e_begin) >> 2);
    const uint8_t* gsrc = ent + e_begin * 8;
    const long long repw = ((long long)blockIdx.z * RW + rw) * 128;      // first replicate of this warp
    const uint8_t* wsrc = wT + repw + 16 * (lane & 7);
    const int esub = lane >> 3;                      // entry of a group this lane gathers for
    double* wsblk = ws + (size_t)blockIdx.y * (GV2_TILE * GV2_TILE) + (size_t)(((si << 3) + sj) << 8);
    const double q_hi = quantum * 65536.0;
    const int nchunks = (ngroups + BSS_CH - 1) / BSS_CH;

    auto issue_list = [&](int c) {
        if (lane < 16 && c * BSS_CH + (lane >> 2) < ngroups)
            bss_cp_async8(lbuf + (c % 3) * BSS_LIST_BYTES + 8 * lane, gsrc + (size_t)c * BSS_LIST_BYTES + 8 * lane);
        bss_commit();
    };
    auto issue_weights = [&](int c) {
        const uint8_t* ls = lbuf + (c % 3) * BSS_LIST_BYTES;
        uint8_t* wd = wbuf + (c & 1) * BSS_W_BYTES + esub * 128 + 16 * (lane & 7);
#pragma unroll
        for (int k = 0; k < BSS_CH; ++k)
            if (c * BSS_CH + k < ngroups) {
                const unsigned col = *reinterpret_cast<const uint16_t*>(ls + 32 * k + 2 * esub);
                bss_cp_async16(wd + k * 512, wsrc + (size_t)col * (size_t)w_ld);
            }
        bss_commit();
    };

    // pair bookkeeping: pair f = row * 16 + j of the warp's NR rows ends (in groups, relative to the warp's stream) at
    // ends[f >> 5] of lane f & 31 -- loaded once, so that no global load sits on the flush path
    constexpr int NE = NR >= 2 ? NR / 2 : 1;
    int ends[NE];
#pragma unroll
    for (int q = 0; q < NE; ++q) {
        const int f = 32 * q + lane;
        const long long rr = row_first + (f >> 4);
        ends[q] = f < NR * 16 ? (int)((rowptr[rr] - e_begin) >> 2) + (int)(pend[rr * 16 + (f & 15)] >> 2) : 0x7fffffff;
    }
    auto end_of = [&](int f) -> int {
        int v = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < NE; ++q) {
            const int u = __shfl_sync(0xffffffffu, ends[q], f & 31);
            if ((f >> 5) == q) v = u;
        }
        return v;
    };
    int pf = 0;                                      // current pair of the warp: row = pf >> 4, j = pf & 15
    int pair_end = end_of(0);
    unsigned acc[6][4];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[k][r] = 0u;
    // Results leave through shared memory: a lane owns four REPLICATES, i.e. four different 0.4 GB planes of the workspace, so
    // storing from registers costs one memory request per lane and instruction (32 per warp store) -- the load/store unit, not
    // HBM, was the limit.  Eight pairs (half an i-row: 64 contiguous bytes per replicate) are staged as [pair][replicate] and
    // written by lanes (pair, replicate): four 64-byte requests per warp store.
    auto flush_pair = [&]() {
        double v4[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            // planes 2..5 are the high 32-bit limb, planes 0..1 the low 16-bit limb of the 48-bit values
            unsigned long long hi = (unsigned long long)acc[2][r] + ((unsigned long long)acc[3][r] << 8) +
                                    ((unsigned long long)acc[4][r] << 16) + ((unsigned long long)acc[5][r] << 24);
            const unsigned lo = acc[0][r] + (acc[1][r] << 8);            // <= 255 * 65536 * 257 < 2^32
            v4[r] = bs_to_double(hi, lo, q_hi, quantum);
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[k][r] = 0u;
        }
        double* od = obuf + (pf & (BSS_OUT_PAIRS - 1)) * BSS_OUT_LD + 4 * lane;
        *reinterpret_cast<double2*>(od) = make_double2(v4[0], v4[1]);
        *reinterpret_cast<double2*>(od + 2) = make_double2(v4[2], v4[3]);
        if ((pf & (BSS_OUT_PAIRS - 1)) == BSS_OUT_PAIRS - 1) {
            __syncwarp();
            const int ii = rg * NR + (pf >> 4);
            const int pj = lane & (BSS_OUT_PAIRS - 1);                   // pair of the staged group
            constexpr int RPI = 32 / BSS_OUT_PAIRS;                      // replicates per store instruction
            double* o = wsblk + (ii << 4) + (pf & 15 & ~(BSS_OUT_PAIRS - 1)) + pj;
            const double* os = obuf + pj * BSS_OUT_LD + lane / BSS_OUT_PAIRS;
#pragma unroll 4
            for (int it = 0; it < 128 / RPI; ++it) {
                const long long rep = repw + RPI * it + lane / BSS_OUT_PAIRS;
                if (rep < nb_total) o[(size_t)rep * ws_rep_stride] = os[RPI * it];
            }
            __syncwarp();
        }
        ++pf;
        pair_end = pf < NR * 16 ? end_of(pf) : 0x7fffffff;
    };

    if (nchunks > 0) {
        issue_list(0);
        issue_list(1);
        bss_wait1();
        __syncwarp();
        issue_weights(0);
    }
    for (int c = 0; c < nchunks; ++c) {
        issue_list(c + 2);
        bss_wait2();                                 // list chunk c + 1 has landed
        __syncwarp();
        issue_weights(c + 1);
        bss_wait2();                                 // multiplicities of chunk c have landed
        __syncwarp();
        const uint8_t* ls = lbuf + (c % 3) * BSS_LIST_BYTES;
        const uint8_t* wl = wbuf + (c & 1) * BSS_W_BYTES + 4 * lane;
#pragma unroll
        for (int k = 0; k < BSS_CH; ++k) {
            const int g = c * BSS_CH + k;
            if (g >= ngroups) break;
            while (g >= pair_end) flush_pair();      // warp-uniform: lists are padded, a group belongs to one pair
            const uint4 A = *reinterpret_cast<const uint4*>(ls + 32 * k);         // columns (x, y), planes 0, 1
            const uint4 B = *reinterpret_cast<const uint4*>(ls + 32 * k + 16);    // planes 2..5
            const unsigned x0 = *reinterpret_cast<const unsigned*>(wl + (4 * k + 0) * 128);
            const unsigned x1 = *reinterpret_cast<const unsigned*>(wl + (4 * k + 1) * 128);
            const unsigned x2 = *reinterpret_cast<const unsigned*>(wl + (4 * k + 2) * 128);
            const unsigned x3 = *reinterpret_cast<const unsigned*>(wl + (4 * k + 3) * 128);
            // 4 x 4 byte transpose: y[r] = multiplicities of the four entries for replicate r of this lane
            const unsigned t0 = __byte_perm(x0, x1, 0x5140), t1 = __byte_perm(x2, x3, 0x5140);
            const unsigned t2 = __byte_perm(x0, x1, 0x7362), t3 = __byte_perm(x2, x3, 0x7362);
            unsigned y[4];
            y[0] = __byte_perm(t0, t1, 0x5410); y[1] = __byte_perm(t0, t1, 0x7632);
            y[2] = __byte_perm(t2, t3, 0x5410); y[3] = __byte_perm(t2, t3, 0x7632);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[0][r] = __dp4a(A.z, y[r], acc[0][r]);
                acc[1][r] = __dp4a(A.w, y[r], acc[1][r]);
                acc[2][r] = __dp4a(B.x, y[r], acc[2][r]);
                acc[3][r] = __dp4a(B.y, y[r], acc[3][r]);
                acc[4][r] = __dp4a(B.z, y[r], acc[4][r]);
                acc[5][r] = __dp4a(B.w, y[r], acc[5][r]);
            }
        }
        __syncwarp();                                // every lane is done with the slots the next iteration overwrites
    }
    while (pf < NR * 16) flush_pair();               // the remaining (possibly empty) pairs of the warp's rows
}

// Few replicates (<= 12: the base matrix is ONE all-ones replicate, the PL2 loop of the reference draws 10): with replicates
// on the lanes most of a warp would idle and every pair would cost a warp a flush.  Here the lanes are the 32 PAIRS of two
// i-rows of a 16 x 16 block with their replicates in registers: same byte planes, same dp4a, same limb recombination --
// bit-identical to the streaming kernel.  List lengths are very uneven (two genomes of one taxon share hundreds of columns,
// two of different taxa a handful), so
//   long lists (more than PL_SHORT groups) are walked by the whole warp, 32 groups per step, and their plane sums meet in
//                REDUX additions (exact integers: order does not matter);
//   short lists are walked by their own lane.
// Stores: the 16 pairs of an i-row are 128 contiguous bytes of a replicate's plane.  The multiplicity table of these launches
// is compact (4 / 8 / 16 bytes per column: L1 resident).  NQ = replicate quads per lane (nb <= 4 NQ).
// grid = (64 sub-blocks, tiles), block = 8 warps x 2 i-rows
#define PL_SHORT 3
template <int NQ>
__global__ void __launch_bounds__(BS_THREADS, NQ == 1 ? 3 : 2)
boot_sparse_pairlane_kernel(const long long* __restrict__ rowptr, const uint32_t* __restrict__ pend, const uint8_t* __restrict__ ent,
                            int nt, long long tile_begin, const uint8_t* __restrict__ wT, long long w_ld, int nb_total,
                            long long ws_rep_stride, double quantum, double* __restrict__ ws) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = blockIdx.x, si = sub >> 3, sj = sub & 7;
    int bi, bj;
    bs_tile_coords(tile_begin + blockIdx.y, nt, &bi, &bj);
    if (bi == bj && si > sj) return;                 // block entirely below the diagonal (never read)
    const int ii = 2 * warp + (lane >> 4), jj = lane & 15;
    const long long rr = ((tile_begin + blockIdx.y) * 64 + sub) * 16 + ii;
    const uint32_t pe = pend[rr * 16 + jj];
    uint32_t pb = __shfl_up_sync(0xffffffffu, pe, 1);
    if (jj == 0) pb = 0u;
    const uint8_t* gp = ent + (size_t)(rowptr[rr] + pb) * 8;
    const int ng = (int)((pe - pb) >> 2);
    const double q_hi = quantum * 65536.0;
    double* o = ws + (size_t)blockIdx.y * (GV2_TILE * GV2_TILE) + (size_t)(((si << 3) + sj) << 8) + (ii << 4) + jj;
    unsigned acc[NQ][6][4];
    auto clear = [&]() {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int k = 0; k < 6; ++k)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][k][r] = 0u;
    };
    // one group of four entries: gather the multiplicities of its columns, transpose 4 x 4 bytes, six dp4a per replicate
    auto consume = [&](const uint4 A, const uint4 B) {
        const uint8_t* wp[4] = {wT + (size_t)(A.x & 0xffffu) * (size_t)w_ld, wT + (size_t)(A.x >> 16) * (size_t)w_ld,
                                wT + (size_t)(A.y & 0xffffu) * (size_t)w_ld, wT + (size_t)(A.y >> 16) * (size_t)w_ld};
        unsigned x[NQ][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (NQ == 1) {
                x[0][e] = __ldg(reinterpret_cast<const unsigned*>(wp[e]));
            } else if (NQ == 2) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(wp[e]));
                x[0][e] = v.x; x[NQ > 1 ? 1 : 0][e] = v.y;
            } else {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(wp[e]));
                x[0][e] = v.x; x[NQ > 1 ? 1 : 0][e] = v.y; x[NQ > 2 ? 2 : 0][e] = v.z;
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const unsigned t0 = __byte_perm(x[q][0], x[q][1], 0x5140), t1 = __byte_perm(x[q][2], x[q][3], 0x5140);
            const unsigned t2 = __byte_perm(x[q][0], x[q][1], 0x7362), t3 = __byte_perm(x[q][2], x[q][3], 0x7362);
            unsigned y[4];
            y[0] = __byte_perm(t0, t1, 0x5410); y[1] = __byte_perm(t0, t1, 0x7632);
            y[2] = __byte_perm(t2, t3, 0x5410); y[3] = __byte_perm(t2, t3, 0x7632);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[q][0][r] = __dp4a(A.z, y[r], acc[q][0][r]);
                acc[q][1][r] = __dp4a(A.w, y[r], acc[q][1][r]);
                acc[q][2][r] = __dp4a(B.x, y[r], acc[q][2][r]);
                acc[q][3][r] = __dp4a(B.y, y[r], acc[q][3][r]);
                acc[q][4][r] = __dp4a(B.z, y[r], acc[q][4][r]);
                acc[q][5][r] = __dp4a(B.w, y[r], acc[q][5][r]);
            }
        }
    };
    auto store = [&](double* dst) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int rep = 4 * q + r;
                if (rep >= nb_total) continue;
                const unsigned long long hi = (unsigned long long)acc[q][2][r] + ((unsigned long long)acc[q][3][r] << 8) +
                                              ((unsigned long long)acc[q][4][r] << 16) + ((unsigned long long)acc[q][5][r] << 24);
                const unsigned lo = acc[q][0][r] + (acc[q][1][r] << 8);
                __stcs(dst + (size_t)rep * ws_rep_stride, bs_to_double(hi, lo, q_hi, quantum));
            }
    };
    // ---- long lists: the whole warp on one pair at a time
    unsigned longs = __ballot_sync(0xffffffffu, ng > PL_SHORT);
    while (longs) {
        const int src = __ffs(longs) - 1;
        longs &= longs - 1;
        const uint8_t* gps = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, (unsigned long long)gp, src));
        const int ngs = __shfl_sync(0xffffffffu, ng, src);
        clear();
        for (int g = lane; g < ngs; g += 32)
            consume(__ldg(reinterpret_cast<const uint4*>(gps + 32 * (size_t)g)), __ldg(reinterpret_cast<const uint4*>(gps + 32 * (size_t)g + 16)));
        // plane sums of a pair stay below 2^32 whatever the number of partial sums (255 * sum of multiplicities <= 2^24)
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int k = 0; k < 6; ++k)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][k][r] = __reduce_add_sync(0xffffffffu, acc[q][k][r]);
        if (lane == src) store(o);
    }
    // ---- short lists: every lane its own
    clear();
    if (ng <= PL_SHORT) {
        for (int g = 0; g < ng; ++g)
            consume(__ldg(reinterpret_cast<const uint4*>(gp + 32 * g)), __ldg(reinterpret_cast<const uint4*>(gp + 32 * g + 16)));
        store(o);
    }
}

// ---------------------------------------------------------------------------------- CSR of the fixed-point table
// entries: x > 0 (NaN rows are flagged separately by prepare_table; negative tables do not come here)
__device__ __forceinline__ unsigned long long bs_fixed48(double x, double scale) {
    return x > 0.0 ? (unsigned long long)__double2ull_rn(fmin(x * scale, 281474976710655.0)) : 0ull;
}

__global__ void bs_count_kernel(const double* __restrict__ tab, long long n, long long k, long long ld, double scale,
                                long long* __restrict__ cnt) {
    const long long r = blockIdx.x;
    int c = 0;
    if (r < n) {
        const double* row = tab + r * ld;
        for (long long q = threadIdx.x; q < k; q += blockDim.x) c += (bs_fixed48(row[q], scale) != 0ull);
    }
    __shared__ int red[8];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        cnt[r] = t;
    }
}

// one warp per row, ballot compaction keeps the columns ascending.  grid = ceil(n / 8), block = 256
__global__ void bs_fill_kernel(const double* __restrict__ tab, long long n, long long k, long long ld, double scale,
                               const long long* __restrict__ ptr, int* __restrict__ col, uint32_t* __restrict__ hi,
                               uint16_t* __restrict__ lo) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double* row = tab + r * ld;
    long long o = ptr[r];
    for (long long k0 = 0; k0 < k; k0 += 32) {
        const unsigned long long f = (k0 + lane < k) ? bs_fixed48(row[k0 + lane], scale) : 0ull;
        const unsigned hit = __ballot_sync(0xffffffffu, f != 0ull);
        if (f) {
            const long long d = o + __popc(hit & ((1u << lane) - 1));
            col[d] = (int)(k0 + lane);
            hi[d] = (uint32_t)(f >> BS_LO_BITS);
            lo[d] = (uint16_t)(f & 0xffffu);
        }
        o += __popc(hit);
    }
}

// per-column hit counts (postings lengths): sum_c cnt_c (cnt_c + 1) / 2 = total co-present (pair, column) cells
__global__ void bs_colcount_kernel(const int* __restrict__ col, long long nnz, int* __restrict__ colcnt) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x)
        atomicAdd(&colcnt[col[e]], 1);
}
__global__ void bs_cells_kernel(const int* __restrict__ colcnt, long long k, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < k; c += (long long)gridDim.x * blockDim.x) {
        const unsigned long long v = (unsigned long long)colcnt[c];
        s += v * (v + 1) / 2;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// int32 multiplicities [nb][k] -> transposed uint8 [k_pad][w_ld] (zero padded).  32 x 32 tiles through shared memory.
__global__ void bs_weights_transpose_kernel(const int* __restrict__ w, long long nb, long long k, long long k_pad,
                                            long long w_ld, uint8_t* __restrict__ out, int* __restrict__ overflow) {
    __shared__ uint8_t t[32][33];
    const long long c0 = (long long)blockIdx.x * 32, b0 = (long long)blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const long long b = b0 + y, c = c0 + threadIdx.x;
        int v = (b < nb && c < k) ? w[b * k + c] : 0;
        if (v < 0 || v > 255) { atomicExch(overflow, 1); v = 0; }
        t[y][threadIdx.x] = (uint8_t)v;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const long long c = c0 + y, b = b0 + threadIdx.x;
        if (c < k_pad && b < w_ld) out[c * w_ld + b] = t[threadIdx.x][y];
    }
}

// sum_c w[b][c] must stay <= 2^(32 - BS_LO_BITS): the low limbs accumulate in 32 bits.  grid = nb
__global__ void bs_wsum_check_kernel(const int* __restrict__ w, long long k, int* __restrict__ overflow) {
    const int* row = w + (long long)blockIdx.x * k;
    long long s = 0;
    for (long long c = threadIdx.x; c < k; c += blockDim.x) s += row[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ long long red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        if (t > (1ll << (32 - BS_LO_BITS))) atomicExch(overflow, 2);
    }
}

// Row sums sum_c x_ic of the fixed-point image = the diagonal pair sums under all-ones multiplicities, from the CSR:
// one warp per row.  grid = ceil(n / 8), block = 256
__global__ void bs_rowsums_kernel(BsCsr csr, long long n, double quantum, double* __restrict__ out) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    unsigned long long hi = 0ull, lo = 0ull;
    for (long long e = csr.ptr[r] + lane; e < csr.ptr[r + 1]; e += 32) { hi += csr.hi[e]; lo += csr.lo[e]; }
    for (int o = 16; o > 0; o >>= 1) {
        hi += __shfl_xor_sync(0xffffffffu, hi, o);
        lo += __shfl_xor_sync(0xffffffffu, lo, o);
    }
    if (lane == 0) out[r] = bs_to_double(hi, lo, quantum * 65536.0, quantum);
}

