// The linkage kernels (nearest-neighbour chain, minimum spanning tree, heap-based generic algorithm) and their device
// helpers.  linkage.cu includes this file; so does tests/emul/lk_emul.cpp, which runs the SAME kernel source on host
// threads -- one per CUDA thread, barriers for __syncthreads and the warp reductions / shuffles -- so that the kernels'
// logic is checked against scipy where there is no GPU.
#pragma once

#ifndef LK_THREADS
#define LK_THREADS 1024
#endif
#ifndef LK_DYNAMIC_SMEM
#define LK_DYNAMIC_SMEM(name) extern __shared__ int name[]
#endif
#define LK_MLP 10                // row elements per thread in flight at a time (one batch per 10 240-leaf row)

struct LkBest { double d; int i; };

__device__ __forceinline__ LkBest lk_min(LkBest a, LkBest b) {
    // smaller distance; on equal distances the lower index (what an upward scan with a strict `<` keeps)
    if (b.d < a.d || (b.d == a.d && b.i < a.i)) return b;
    return a;
}

// grid = matrices, block = LK_THREADS; dynamic shared memory: 2 n ints (cluster sizes, the chain)
// D: [nb][n][n] float64, destroyed; Zraw: [nb][n-1][4] float64 in merge order
// linkage_methods[...] of scipy/cluster/_hierarchy_distance_update.pxi; every operation rounded on its own (the host
// builds of scipy do not contract to FMA), so the heights are bit-identical
#define LK_AVERAGE 0
#define LK_COMPLETE 1
#define LK_WEIGHTED 2
#define LK_WARD 3
#define LK_SINGLE 4
#define LK_CENTROID 5
#define LK_MEDIAN 6
template <int METHOD>
__device__ __forceinline__ double lk_new_dist(double d_xi, double d_yi, double d_xy, int nx, int ny, int ni) {
    if (METHOD == LK_AVERAGE)
        return __ddiv_rn(__dadd_rn(__dmul_rn((double)nx, d_xi), __dmul_rn((double)ny, d_yi)), (double)(nx + ny));
    if (METHOD == LK_COMPLETE) return d_xi > d_yi ? d_xi : d_yi;
    if (METHOD == LK_WEIGHTED) return __dmul_rn(0.5, __dadd_rn(d_xi, d_yi));
    if (METHOD == LK_CENTROID) {
        // sqrt((((nx d_xi d_xi) + (ny d_yi d_yi)) - (nx ny d_xy d_xy) / (nx + ny)) / (nx + ny)); nx ny is an integer product;
        // a negative radicand (distances that are not Euclidean) gives NaN, as C's sqrt does
        const double a = __dmul_rn(__dmul_rn((double)nx, d_xi), d_xi);
        const double b = __dmul_rn(__dmul_rn((double)ny, d_yi), d_yi);
        const double c = __ddiv_rn(__dmul_rn(__dmul_rn((double)(nx * ny), d_xy), d_xy), (double)(nx + ny));
        return __dsqrt_rn(__ddiv_rn(__dadd_rn(__dadd_rn(a, b), -c), (double)(nx + ny)));
    }
    if (METHOD == LK_MEDIAN)
        return __dsqrt_rn(__dadd_rn(__dmul_rn(0.5, __dadd_rn(__dmul_rn(d_xi, d_xi), __dmul_rn(d_yi, d_yi))),
                                    -__dmul_rn(__dmul_rn(0.25, d_xy), d_xy)));
    // ward: t = 1 / (nx + ny + ni); sqrt((ni + nx) t d_xi^2 + (ni + ny) t d_yi^2 - ni t d_xy^2), products left to right
    const double t = __ddiv_rn(1.0, (double)(nx + ny + ni));
    const double a = __dmul_rn(__dmul_rn(__dmul_rn((double)(ni + nx), t), d_xi), d_xi);
    const double b = __dmul_rn(__dmul_rn(__dmul_rn((double)(ni + ny), t), d_yi), d_yi);
    const double c = __dmul_rn(__dmul_rn(__dmul_rn((double)ni, t), d_xy), d_xy);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(a, b), -c));
}

// Minimum of (distance, index) over a warp in three REDUX instructions instead of fifteen shuffles: the distance as a
// sortable 64-bit key (x + 0.0 folds -0.0 into +0.0, NaN -> the largest key: never chosen, as with a strict `<`), minimum of
// the high words, then of the low words among the lanes that hold it, then of the indices among those -- "smaller distance,
// on ties the lower index", what an upward scan with a strict `<` keeps.
struct LkKey { unsigned hi, lo; int i; };
__device__ __forceinline__ LkKey lk_key(double d, int i) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d + 0.0);
    unsigned long long k = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    if (d != d) k = ~0ull;
    return LkKey{(unsigned)(k >> 32), (unsigned)k, i};
}
__device__ __forceinline__ double lk_key_value(unsigned hi, unsigned lo) {
    const unsigned long long k = ((unsigned long long)hi << 32) | lo;
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
__device__ __forceinline__ LkKey lk_warp_min(LkKey k) {
    const unsigned mh = __reduce_min_sync(0xffffffffu, k.hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, k.hi == mh ? k.lo : 0xffffffffu);
    const int mi = (int)__reduce_min_sync(0xffffffffu, (k.hi == mh && k.lo == ml) ? (unsigned)k.i : 0x7fffffffu);
    return LkKey{mh, ml, mi};
}

// THREADS = 1024: the whole SM to one matrix.  (Half-SM CTAs of 512 threads, which leave room for a CTA of the similarity
// kernels beside the clustering in the bootstrap job's dendrogram mode, were measured twice: 5.20 s against 5.19 s for 1001
// trees of 10 000 leaves with a two-part ring, 4.70 s against 4.42 s with the four-part ring.)
// A thread owns the same row positions i = tid + e THREADS in every row, so which of them are live clusters is a register
// bit mask (bit e): dead positions are neither loaded nor compared, and no shared-memory lookup sits in the scan.
// LAZY: a merge does not write the new cluster's COLUMN (n scattered 8-byte stores, one 32-byte sector each: half of the
// kernel's DRAM operations and most of a merge's time in the load/store unit).  The matrix is symmetric and a new cluster's
// ROW is complete when it is written, so an entry (x, j) of an older row is simply read from the other side, D[j][x], when it
// is next needed: per cluster born[j] (merges done when row j was written) and fresh[x] (merges done when row x was last
// brought up to date); entry (x, j) is stale iff born[j] > fresh[x].  A scan of row x loads every element from whichever
// side is current -- one load per element either way, the address is chosen from shared-memory state before any load is
// issued, so a step is still one round trip -- writes the patched values back into row x and sets fresh[x]; a merge does
// the same for the two rows it reads.  Same values, same operations, same order: the linkage matrix stays bit-identical.
// Dynamic shared memory: 2 n ints (cluster sizes, the chain) [+ 2 n ints (born, fresh)].
template <int METHOD, int THREADS, bool LAZY>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
lk_nn_chain_kernel(double* __restrict__ Dall, long long d_stride, int n, double* __restrict__ Zall) {
    LK_DYNAMIC_SMEM(lk_size);
    __shared__ LkKey s_warp[THREADS / 32];
    __shared__ int s_x, s_y, s_len, s_merge, s_first;
    __shared__ double s_min, s_dp;
    double* D = Dall + (size_t)blockIdx.x * d_stride;
    int* chain = lk_size + n;
    int* born = chain + n;                            // LAZY only
    int* fresh = born + n;                            // LAZY only
    double* Z = Zall + (size_t)blockIdx.x * (size_t)(n > 1 ? n - 1 : 0) * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += THREADS) {
        lk_size[i] = 1;
        if (LAZY) { born[i] = 0; fresh[i] = 0; }
    }
    unsigned alive = 0u;                              // n <= 32 THREADS (the shared-memory limit of the launcher is lower)
    for (int e = 0; e < 32; ++e)
        if (tid + e * THREADS < n) alive |= 1u << e;
    if (tid == 0) { s_len = 0; s_first = 0; }
    __syncthreads();
    for (int k = 0; k < n - 1; ++k) {
        if (tid == 0 && s_len == 0) {
            int f = s_first;
            while (lk_size[f] == 0) ++f;              // lowest-numbered live cluster (dead ones never come back)
            s_first = f;
            chain[0] = f;
            s_len = 1;
        }
        __syncthreads();
        // grow the chain until its two top elements are mutual nearest neighbours
        while (true) {
            const int len = s_len;
            const int x = chain[len - 1];
            const int prev = len > 1 ? chain[len - 2] : -1;
            double* row = D + (size_t)x * n;
            const int fx = LAZY ? fresh[x] : 0;
            double bd = INFINITY;
            int bi = 0x7fffffff;
            // The step is one DRAM round trip long, not one per element: a thread's loads of the row are all issued before the
            // first is looked at, LK_MLP at a time.
            for (int e0 = 0, i0 = tid; i0 < n; e0 += LK_MLP, i0 += THREADS * LK_MLP) {
                double dv[LK_MLP];
                const unsigned live = alive >> e0;
                unsigned stale = 0u;
#pragma unroll
                for (int u = 0; u < LK_MLP; ++u) {
                    const int i = i0 + u * THREADS;
                    const bool on = (live >> u) & 1u;
                    const bool st = LAZY && on && born[i] > fx;          // row i is younger than row x's last refresh
                    if (st) stale |= 1u << u;
                    dv[u] = on ? (st ? D[(size_t)i * n + x] : __ldcs(row + i)) : INFINITY;
                }
#pragma unroll
                for (int u = 0; u < LK_MLP; ++u) {
                    const int i = i0 + u * THREADS;
                    if (!((live >> u) & 1u) || i == x) continue;
                    if (LAZY && ((stale >> u) & 1u)) row[i] = dv[u];     // row x is current again
                    if (i == prev) s_dp = dv[u];                     // the incumbent's distance, for the final comparison
                    if (dv[u] < bd) { bd = dv[u]; bi = i; }          // ascending i per thread: first minimum kept
                }
            }
            LkKey best = lk_warp_min(lk_key(bd, bi));
            if (lane == 0) s_warp[warp] = best;
            __syncthreads();
            if (warp == 0) {
                best = lane < THREADS / 32 ? s_warp[lane] : LkKey{0xffffffffu, 0xffffffffu, 0x7fffffff};
                best = lk_warp_min(best);
                if (lane == 0) {
                    // the previous chain element is the incumbent: a scanned cluster replaces it only if strictly closer
                    int y = best.i;
                    double cur = lk_key_value(best.hi, best.lo);
                    const bool none = best.hi == 0xffffffffu && best.lo == 0xffffffffu;   // nothing comparable (all NaN)
                    if (none) { y = 0x7fffffff; cur = INFINITY; }
                    if (prev >= 0) {
                        const double dp = s_dp;
                        if (!(cur < dp)) { y = prev; cur = dp; }
                    } else if (y == 0x7fffffff) {
                        y = 0;                        // every distance NaN: scipy keeps its stale y; unreachable for cleaned matrices
                    }
                    if (prev >= 0 && y == prev) {
                        s_merge = 1;
                    } else {
                        s_merge = 0;
                        chain[len] = y;
                        s_len = len + 1;
                    }
                    s_x = x; s_y = y; s_min = cur;
                    if (LAZY) fresh[x] = k;           // k merges so far: every younger row has been read from its own side
                }
            }
            __syncthreads();
            if (s_merge) break;
        }
        // merge the two top elements
        int x = s_x, y = s_y;
        if (x > y) { const int t = x; x = y; y = t; }
        const int nx = lk_size[x], ny = lk_size[y];
        const double dmin = s_min;
        const int fx = LAZY ? fresh[x] : 0, fy = LAZY ? fresh[y] : 0;
        if ((x & (THREADS - 1)) == tid) alive &= ~(1u << (x / THREADS));   // x dies; its owner drops it
        __syncthreads();                              // everyone has read the sizes and s_* before they change
        if (tid == 0) {
            Z[(size_t)k * 4 + 0] = (double)x;
            Z[(size_t)k * 4 + 1] = (double)y;
            Z[(size_t)k * 4 + 2] = dmin;
            Z[(size_t)k * 4 + 3] = (double)(nx + ny);
            lk_size[x] = 0;
            lk_size[y] = nx + ny;
            s_len -= 2;
        }
        const double* rx = D + (size_t)x * n;
        double* ry = D + (size_t)y * n;
        for (int e0 = 0, i0 = tid; i0 < n; e0 += LK_MLP, i0 += THREADS * LK_MLP) {
            double ax[LK_MLP], ay[LK_MLP];
            int sz[LK_MLP];
            const unsigned live = alive >> e0;
#pragma unroll
            for (int u = 0; u < LK_MLP; ++u) {
                const int i = i0 + u * THREADS;
                const bool on = ((live >> u) & 1u) && i != y;
                const int bi2 = (LAZY && on) ? born[i] : 0;  // (i != x, y: not entries thread 0 is changing)
                ax[u] = on ? (bi2 > fx ? D[(size_t)i * n + x] : rx[i]) : 0.0;
                ay[u] = on ? (bi2 > fy ? D[(size_t)i * n + y] : ry[i]) : 0.0;
                sz[u] = on ? lk_size[i] : 0;
            }
#pragma unroll
            for (int u = 0; u < LK_MLP; ++u) {
                const int i = i0 + u * THREADS;
                if (sz[u] == 0) continue;
                const double v = lk_new_dist<METHOD>(ax[u], ay[u], dmin, nx, ny, sz[u]);
                ry[i] = v;
                if (!LAZY) D[(size_t)i * n + y] = v;
            }
        }
        if (LAZY && tid == 0) { born[y] = k + 1; fresh[y] = k + 1; }   // (read by others only after the barrier below)
        __syncthreads();
    }
}

// The chain kernel with a LIVE LIST (lazy column updates as above).  ncu on one 10 000-leaf matrix showed the kernel above
// issue-bound as much as memory-bound (issue slots 52 % busy with one CTA on the SM; 63 % of the instructions in the row
// scan, 33 % in the merge) and every scan walking all n positions, ~45 instructions each, whether the cluster there is
// alive or not -- while on average half of them are dead.  Here the live clusters are an ascending list in shared memory,
// one 32-bit entry each: (born << 16) | index, so that one LDS brings the cluster and the stamp that says which side of the
// matrix holds the current value.  Scans and merges walk the list (lane -> list slot, so the loads stay as coalesced as the
// survivors allow), a merge tombstones the entry of the cluster that dies and re-stamps the one that survives, and whenever
// an eighth of the list is tombstones the block compacts it in place (order preserving: per-thread contiguous pieces,
// counts, a two-level prefix sum).  Work per scan follows the number of live clusters: about half the element visits over a
// run, and element offsets are 32-bit (n * n < 2^32), one address computation whichever side is read.  Per thread the list
// slots ascend and the block minimum is taken on (distance, index): "first minimum of an upward scan", as before --
// bit-identical rows.  Dynamic shared memory: 3 n ints (sizes, chain, fresh) + n list entries = 16 bytes per leaf.
#define LK_TOMB 0xffffffffu
// Keep a value in its register: without this ptxas re-derives the matrix pointer (block index x stride + base, three
// constant-bank loads and a 64-bit multiply-add) and n for EVERY element under the 64-register cap of a 1024-thread CTA.
#ifdef __CUDA_ARCH__
#define LK_KEEP_PTR(p) asm volatile("" : "+l"(p))
#define LK_KEEP_U32(v) asm volatile("" : "+r"(v))
#else
#define LK_KEEP_PTR(p)
#define LK_KEEP_U32(v)
#endif
#define LK_LIST_PER 14           // list entries per thread in a compaction: n <= 14 * THREADS
static inline size_t lk_list_smem_bytes(long long n) { return (size_t)n * 16; }

// One row scan over the live list, MLP entries per thread in flight (the caller picks MLP from the list length, so that the
// instructions issued follow the number of live clusters -- predicated-off slots of a fixed batch of ten would not).
template <int MLP, int THREADS>
__device__ __forceinline__ void lk_list_scan(double* D, const unsigned* live, int len, int tid, unsigned un, unsigned ux,
                                             unsigned xrow, unsigned fx, int prev, double* dp_out, double& bd, int& bi) {
    for (int s0 = tid; s0 < len; s0 += THREADS * MLP) {
        double dv[MLP];
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            const int slot = s0 + u * THREADS;
            const unsigned e = slot < len ? live[slot] : LK_TOMB;
            const unsigned i = e & 0xffffu;
            // row i younger than row x's last refresh: the current value is on its side of the matrix
            const unsigned off = (e >> 16) > fx ? i * un + ux : xrow + i;
            dv[u] = e != LK_TOMB ? __ldcs(D + off) : INFINITY;
        }
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            const int slot = s0 + u * THREADS;
            const unsigned e = slot < len ? live[slot] : LK_TOMB;     // (re-read: entries held in registers cost more)
            const unsigned i = e & 0xffffu;                            // (a tombstone's 0xffff is no cluster: n <= 14 336)
            if (e != LK_TOMB && (e >> 16) > fx) D[xrow + i] = dv[u];   // row x is current again
            if ((int)i == prev) *dp_out = dv[u];                       // the incumbent's distance, for the final comparison
            if (dv[u] < bd) { bd = dv[u]; bi = (int)i; }               // ascending i per thread: first minimum kept
        }
    }
}

// The merge's pass over the live list: distances of every other live cluster to the union, written to the union's row.
template <int METHOD, int MLP, int THREADS>
__device__ __forceinline__ void lk_list_merge(double* D, unsigned* live, const int* sizes, int len, int tid, unsigned un,
                                              unsigned ux, unsigned uy, unsigned fx, unsigned fy, unsigned stamp, double dmin,
                                              int nx, int ny) {
    const unsigned xrow = ux * un, yrow = uy * un;
    for (int s0 = tid; s0 < len; s0 += THREADS * MLP) {
        double ax[MLP], ay[MLP];
        int sz[MLP];
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            const int slot = s0 + u * THREADS;
            const unsigned e = slot < len ? live[slot] : LK_TOMB;
            const unsigned i = e & 0xffffu;
            const bool on = e != LK_TOMB && i != ux && i != uy;
            if (e != LK_TOMB && i == ux) live[slot] = LK_TOMB;              // x dies: a tombstone
            if (e != LK_TOMB && i == uy) live[slot] = (stamp << 16) | uy;   // the union's row is written now
            const unsigned bi2 = e >> 16;
            ax[u] = on ? D[bi2 > fx ? i * un + ux : xrow + i] : 0.0;
            ay[u] = on ? D[bi2 > fy ? i * un + uy : yrow + i] : 0.0;
            sz[u] = on ? sizes[i] : 0;
        }
#pragma unroll
        for (int u = 0; u < MLP; ++u) {
            if (sz[u] == 0) continue;
            const unsigned i = live[s0 + u * THREADS] & 0xffffu;            // (this thread's own slot, re-read)
            D[yrow + i] = lk_new_dist<METHOD>(ax[u], ay[u], dmin, nx, ny, sz[u]);
        }
    }
}

template <int METHOD, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
lk_nn_chain_list_kernel(double* __restrict__ Dall, long long d_stride, int n, double* __restrict__ Zall) {
    LK_DYNAMIC_SMEM(lk_size);
    __shared__ LkKey s_warp[THREADS / 32];
    __shared__ int s_x, s_y, s_len, s_merge, s_first;
    __shared__ double s_min, s_dp;
    __shared__ int s_cnt[THREADS];
    __shared__ int s_wtot[THREADS / 32];
    double* D = Dall + (size_t)blockIdx.x * d_stride;
    int* chain = lk_size + n;
    int* fresh = chain + n;
    unsigned* live = reinterpret_cast<unsigned*>(fresh + n);
    double* Z = Zall + (size_t)blockIdx.x * (size_t)(n > 1 ? n - 1 : 0) * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned un = (unsigned)n;
    LK_KEEP_PTR(D);
    LK_KEEP_U32(un);
    for (int i = tid; i < n; i += THREADS) {
        lk_size[i] = 1;
        fresh[i] = 0;
        live[i] = (unsigned)i;                        // born 0
        D[(size_t)i * un + i] = INFINITY;             // a cluster is never its own neighbour: no `i != x` in the scan
    }
    int len = n, dead = 0;                            // list length and tombstones in it (uniform over the block)
    if (tid == 0) { s_len = 0; s_first = 0; }
    __syncthreads();
    for (int k = 0; k < n - 1; ++k) {
        if (tid == 0 && s_len == 0) {
            int f = s_first;
            while (lk_size[f] == 0) ++f;              // lowest-numbered live cluster (dead ones never come back)
            s_first = f;
            chain[0] = f;
            s_len = 1;
        }
        __syncthreads();
        // grow the chain until its two top elements are mutual nearest neighbours
        while (true) {
            const int clen = s_len;
            const int x = chain[clen - 1];
            const int prev = clen > 1 ? chain[clen - 2] : -1;
            const unsigned ux = (unsigned)x, xrow = ux * un;
            const unsigned fx = (unsigned)fresh[x];
            double bd = INFINITY;
            int bi = 0x7fffffff;
            {
                const int per = (len + THREADS - 1) / THREADS;       // uniform over the block
                if (per > 7) lk_list_scan<10, THREADS>(D, live, len, tid, un, ux, xrow, fx, prev, &s_dp, bd, bi);
                else if (per > 5) lk_list_scan<7, THREADS>(D, live, len, tid, un, ux, xrow, fx, prev, &s_dp, bd, bi);
                else if (per > 3) lk_list_scan<5, THREADS>(D, live, len, tid, un, ux, xrow, fx, prev, &s_dp, bd, bi);
                else if (per > 2) lk_list_scan<3, THREADS>(D, live, len, tid, un, ux, xrow, fx, prev, &s_dp, bd, bi);
                else if (per > 1) lk_list_scan<2, THREADS>(D, live, len, tid, un, ux, xrow, fx, prev, &s_dp, bd, bi);
                else lk_list_scan<1, THREADS>(D, live, len, tid, un, ux, xrow, fx, prev, &s_dp, bd, bi);
            }
            LkKey best = lk_warp_min(lk_key(bd, bi));
            if (lane == 0) s_warp[warp] = best;
            __syncthreads();
            if (warp == 0) {
                best = lane < THREADS / 32 ? s_warp[lane] : LkKey{0xffffffffu, 0xffffffffu, 0x7fffffff};
                best = lk_warp_min(best);
                if (lane == 0) {
                    // the previous chain element is the incumbent: a scanned cluster replaces it only if strictly closer
                    int y = best.i;
                    double cur = lk_key_value(best.hi, best.lo);
                    const bool none = best.hi == 0xffffffffu && best.lo == 0xffffffffu;   // nothing comparable (all NaN)
                    if (none) { y = 0x7fffffff; cur = INFINITY; }
                    if (prev >= 0) {
                        const double dp = s_dp;
                        if (!(cur < dp)) { y = prev; cur = dp; }
                    } else if (y == 0x7fffffff) {
                        y = 0;                        // every distance NaN: scipy keeps its stale y; unreachable for cleaned matrices
                    }
                    if (prev >= 0 && y == prev) {
                        s_merge = 1;
                    } else {
                        s_merge = 0;
                        chain[clen] = y;
                        s_len = clen + 1;
                    }
                    s_x = x; s_y = y; s_min = cur;
                    fresh[x] = k;                     // k merges so far: every younger row has been read from its own side
                }
            }
            __syncthreads();
            if (s_merge) break;
        }
        // merge the two top elements
        int x = s_x, y = s_y;
        if (x > y) { const int t = x; x = y; y = t; }
        const int nx = lk_size[x], ny = lk_size[y];
        const double dmin = s_min;
        const unsigned fx = (unsigned)fresh[x], fy = (unsigned)fresh[y];
        const unsigned ux = (unsigned)x, uy = (unsigned)y;
        __syncthreads();                              // everyone has read the sizes and s_* before they change
        if (tid == 0) {
            Z[(size_t)k * 4 + 0] = (double)x;
            Z[(size_t)k * 4 + 1] = (double)y;
            Z[(size_t)k * 4 + 2] = dmin;
            Z[(size_t)k * 4 + 3] = (double)(nx + ny);
            lk_size[x] = 0;                           // (the loop below reads the sizes of clusters other than x and y)
            lk_size[y] = nx + ny;
            fresh[y] = k + 1;                         // (read by others only after the barrier below)
            s_len -= 2;
        }
        {
            const int per = (len + THREADS - 1) / THREADS;
            const unsigned stamp = (unsigned)(k + 1);
            if (per > 3) lk_list_merge<METHOD, 5, THREADS>(D, live, lk_size, len, tid, un, ux, uy, fx, fy, stamp, dmin, nx, ny);
            else if (per > 2) lk_list_merge<METHOD, 3, THREADS>(D, live, lk_size, len, tid, un, ux, uy, fx, fy, stamp, dmin, nx, ny);
            else if (per > 1) lk_list_merge<METHOD, 2, THREADS>(D, live, lk_size, len, tid, un, ux, uy, fx, fy, stamp, dmin, nx, ny);
            else lk_list_merge<METHOD, 1, THREADS>(D, live, lk_size, len, tid, un, ux, uy, fx, fy, stamp, dmin, nx, ny);
        }
        ++dead;
        // compaction of the live list (every thread takes the same decision: len and dead are uniform)
        if (dead >= 32 && dead * 8 >= len) {
            __syncthreads();                          // this merge's tombstone is in the list
            const int per = (len + THREADS - 1) / THREADS;
            const int start = tid * per;
            unsigned e[LK_LIST_PER];
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < LK_LIST_PER; ++j) {
                e[j] = (j < per && start + j < len) ? live[start + j] : LK_TOMB;
                cnt += e[j] != LK_TOMB;
            }
            s_cnt[tid] = cnt;
            __syncthreads();                          // every piece is in registers; the counts are visible
            int off = 0;
            for (int l = 0; l < lane; ++l) off += s_cnt[(warp << 5) + l];
            if (lane == 31) s_wtot[warp] = off + cnt;
            __syncthreads();
            for (int w = 0; w < warp; ++w) off += s_wtot[w];
#pragma unroll
            for (int j = 0; j < LK_LIST_PER; ++j)
                if (e[j] != LK_TOMB) live[off++] = e[j];
            len -= dead;
            dead = 0;
        }
        __syncthreads();
    }
}

// Method "single": mst_single_linkage of scipy/cluster/_hierarchy.pyx -- Prim's minimum spanning tree from node 0; per
// step the row of the newest tree node lowers the running distances of the others, the closest (lowest index on
// ties: upward scan, strict `<`) joins.  Dynamic shared memory: n ints (merged flags) + n doubles (running distances).
__global__ void __launch_bounds__(LK_THREADS, 1)
lk_mst_single_kernel(const double* __restrict__ Dall, long long d_stride, int n, double* __restrict__ Zall) {
    LK_DYNAMIC_SMEM(lk_size);
    __shared__ LkBest s_warp[LK_THREADS / 32];
    __shared__ LkBest s_best;
    int* merged = lk_size;
    double* dmin = reinterpret_cast<double*>(lk_size + ((n + 1) & ~1));
    const double* D = Dall + (size_t)blockIdx.x * d_stride;
    double* Z = Zall + (size_t)blockIdx.x * (size_t)(n > 1 ? n - 1 : 0) * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += LK_THREADS) { merged[i] = 0; dmin[i] = INFINITY; }
    __syncthreads();
    int x = 0;
    for (int k = 0; k < n - 1; ++k) {
        if (tid == 0) merged[x] = 1;
        __syncthreads();
        const double* row = D + (size_t)x * n;
        LkBest best{INFINITY, 0x7fffffff};
        for (int i = tid; i < n; i += LK_THREADS) {
            if (merged[i]) continue;
            const double d = row[i];
            double m = dmin[i];
            if (m > d) { m = d; dmin[i] = d; }
            if (m < best.d) { best.d = m; best.i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            LkBest other;
            other.d = __shfl_xor_sync(0xffffffffu, best.d, o);
            other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
            best = lk_min(best, other);
        }
        if (lane == 0) s_warp[warp] = best;
        __syncthreads();
        if (warp == 0) {
            best = lane < LK_THREADS / 32 ? s_warp[lane] : LkBest{INFINITY, 0x7fffffff};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                LkBest other;
                other.d = __shfl_xor_sync(0xffffffffu, best.d, o);
                other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
                best = lk_min(best, other);
            }
            if (lane == 0) {
                if (best.i == 0x7fffffff) best.i = 0;     // all NaN / infinite: scipy keeps its stale y; unreachable for cleaned matrices
                s_best = best;
                Z[(size_t)k * 4 + 0] = (double)x;
                Z[(size_t)k * 4 + 1] = (double)best.i;
                Z[(size_t)k * 4 + 2] = best.d;
                Z[(size_t)k * 4 + 3] = 0.0;               // sizes come from the relabelling pass
            }
        }
        __syncthreads();
        x = s_best.i;
    }
}

// Methods "centroid" and "median": fast_linkage of scipy/cluster/_hierarchy.pyx (Muellner's generic algorithm).  Their
// heights are not monotone, so the nearest-neighbour chain does not apply; scipy keeps, per cluster x, a nearest-neighbour
// CANDIDATE among the clusters numbered above x and a lower bound of its distance in a binary min-heap (_structures.pxi):
//   initially neighbor[x], min_dist[x] = argmin / min of D[x][i], i > x (upward scan, strict `<`); heapify.
//   merge k: up to n - k times { x = heap top, y = neighbor[x]; if the bound equals D[x][y] stop; else rescan row x and
//   change x's heap value }; remove the top.  Row k = (ids of x and y in ascending order, bound, |x| + |y|), ids n + k for
//   new clusters: rows are NOT sorted or relabelled afterwards.  x dies, y becomes the union, D[z][y] updated for every
//   live z; candidates that pointed at x point at y; for z < y in ascending order: D[z][y] below z's bound replaces
//   candidate and bound (a heap change each -- their ORDER shapes the heap and with it the way later ties fall); finally
//   y's own candidate is rescanned.
// One CTA per matrix: the scans and the row / column update are spread over the block, the heap is walked by thread 0 in
// exactly scipy's operation order.  All per-cluster state lives in shared memory with 16-bit indices: 8 n (heap values)
// + 5 * 2 n (position of key, key at position, cluster size, candidate, cluster id) + n / 8 (flags) bytes, n <= 12 000.
#define LK_NONE 0xffffu
struct LkHeap {
    double* v;                  // values by heap position
    unsigned short* ibk;        // index_by_key
    unsigned short* kbi;        // key_by_index
    int size;
    __device__ __forceinline__ void swap(int i, int j) {
        const double t = v[i]; v[i] = v[j]; v[j] = t;
        const unsigned short ki = kbi[i], kj = kbi[j];
        kbi[i] = kj; kbi[j] = ki;
        ibk[ki] = (unsigned short)j; ibk[kj] = (unsigned short)i;
    }
    __device__ void sift_up(int index) {
        int parent = (index - 1) >> 1;
        while (index > 0 && v[parent] > v[index]) {
            swap(index, parent);
            index = parent;
            parent = (index - 1) >> 1;
        }
    }
    __device__ void sift_down(int index) {
        int child = (index << 1) + 1;
        while (child < size) {
            if (child + 1 < size && v[child + 1] < v[child]) ++child;
            if (v[index] > v[child]) {
                swap(index, child);
                index = child;
                child = (index << 1) + 1;
            } else break;
        }
    }
    __device__ void remove_min() {
        swap(0, size - 1);
        --size;
        sift_down(0);
    }
    __device__ void change_value(int key, double value) {
        const int index = ibk[key];
        const double old = v[index];
        v[index] = value;
        if (value < old) sift_up(index);
        else sift_down(index);
    }
};

// nothing comparable in a scan: every candidate was +inf or NaN (`dist < current_min` never held)
__device__ __forceinline__ bool lk_key_none(LkKey k) { return k.hi >= 0xfff00000u; }

static inline size_t lk_generic_smem_bytes(long long n) {
    return (size_t)n * 8 + (size_t)n * 10 + (size_t)((n + 31) / 32) * 4 + 16;
}

template <int METHOD>
__global__ void __launch_bounds__(LK_THREADS, 1)
lk_generic_kernel(double* __restrict__ Dall, long long d_stride, int n, double* __restrict__ Zall) {
    LK_DYNAMIC_SMEM(lk_size);
    __shared__ LkKey s_warp[LK_THREADS / 32];
    __shared__ int s_x, s_y, s_ok, s_hsize;
    __shared__ double s_dist;
    double* hv = reinterpret_cast<double*>(lk_size);
    unsigned short* ibk = reinterpret_cast<unsigned short*>(hv + n);
    unsigned short* kbi = ibk + n;
    unsigned short* csize = kbi + n;
    unsigned short* nbr = csize + n;
    unsigned short* cid = nbr + n;
    unsigned* flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(lk_size) + (((size_t)n * 18 + 3) & ~(size_t)3));
    double* D = Dall + (size_t)blockIdx.x * d_stride;
    double* Z = Zall + (size_t)blockIdx.x * (size_t)(n > 1 ? n - 1 : 0) * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nflag = (n + 31) / 32;
    for (int i = tid; i < n; i += LK_THREADS) { csize[i] = 1; cid[i] = (unsigned short)i; }
    for (int i = tid; i < nflag; i += LK_THREADS) flags[i] = 0u;
    // candidates of every cluster: one warp per row
    for (int x = warp; x < n - 1; x += LK_THREADS / 32) {
        const double* row = D + (size_t)x * n;
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int i = x + 1 + lane; i < n; i += 32) {
            const double d = row[i];
            if (d < bd) { bd = d; bi = i; }
        }
        const LkKey best = lk_warp_min(lk_key(bd, bi));
        if (lane == 0) {
            const bool none = lk_key_none(best);
            nbr[x] = none ? (unsigned short)LK_NONE : (unsigned short)best.i;
            hv[x] = none ? INFINITY : row[best.i];
            ibk[x] = kbi[x] = (unsigned short)x;
        }
    }
    __syncthreads();
    LkHeap heap{hv, ibk, kbi, n - 1};
    if (tid == 0) {
        for (int i = (n - 1) / 2 - 1; i >= 0; --i) heap.sift_down(i);
        s_hsize = heap.size;
    }
    // block-wide find_min_dist(x): partial minima of the warps to s_warp; the caller reduces them after a barrier
    auto scan_row = [&](int x) {
        const double* row = D + (size_t)x * n;
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int i0 = x + 1 + tid; i0 < n; i0 += LK_THREADS * 4) {
            double dv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * LK_THREADS;
                dv[u] = (i < n && csize[i] != 0) ? row[i] : INFINITY;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (dv[u] < bd) { bd = dv[u]; bi = i0 + u * LK_THREADS; }
        }
        const LkKey best = lk_warp_min(lk_key(bd, bi));
        if (lane == 0) s_warp[warp] = best;
    };
    auto scan_result = [&]() {                         // warp 0, after the barrier; valid in lane 0
        const LkKey best = lane < LK_THREADS / 32 ? s_warp[lane] : LkKey{0xffffffffu, 0xffffffffu, 0x7fffffff};
        return lk_warp_min(best);
    };
    for (int k = 0; k < n - 1; ++k) {
        // the closest pair: heap top whose bound is still the true distance to its candidate
        int it = 0;
        while (true) {
            if (tid == 0) {
                if (it == n - k) {
                    s_ok = 1;                         // scipy's bounded loop ran out: go on with the last rescan (s_x, s_y, s_dist)
                } else {
                    const int x = kbi[0];
                    const double dist = hv[0];
                    const int y = nbr[x];
                    s_x = x; s_y = y; s_dist = dist;
                    // (a cluster without a candidate -- every distance above it +inf or NaN -- makes scipy read outside its
                    // array; here it counts as "bound is stale" and the rescan settles it)
                    s_ok = (y != (int)LK_NONE && dist == D[(size_t)x * n + y]) ? 1 : 0;
                }
            }
            __syncthreads();
            if (s_ok) break;
            const int x = s_x;
            scan_row(x);
            __syncthreads();
            if (warp == 0) {
                const LkKey best = scan_result();
                if (lane == 0) {
                    const bool none = lk_key_none(best);
                    const int y = none ? (int)LK_NONE : best.i;
                    const double dist = none ? INFINITY : D[(size_t)x * n + y];
                    nbr[x] = (unsigned short)y;
                    heap.size = s_hsize;
                    heap.change_value(x, dist);
                    s_y = y; s_dist = dist;
                }
            }
            ++it;
        }
        const int x = s_x;
        int y = s_y;
        const double dist = s_dist;
        if (y == (int)LK_NONE) y = x == n - 1 ? 0 : n - 1;        // no comparable pair left (see above): keep the indices in range
        const int nx = csize[x], ny = csize[y];
        __syncthreads();                              // everyone holds x, y, dist and the sizes
        if (tid == 0) {
            heap.size = s_hsize;
            heap.remove_min();
            s_hsize = heap.size;
            int ix = cid[x], iy = cid[y];
            if (ix > iy) { const int t = ix; ix = iy; iy = t; }
            Z[(size_t)k * 4 + 0] = (double)ix;
            Z[(size_t)k * 4 + 1] = (double)iy;
            Z[(size_t)k * 4 + 2] = dist;
            Z[(size_t)k * 4 + 3] = (double)(nx + ny);
            csize[x] = 0;
            csize[y] = (unsigned short)(nx + ny);
            cid[y] = (unsigned short)(n + k);
        }
        __syncthreads();                              // the heap is at rest while the block reads bounds through it
        // distances to the union; candidates that pointed at x; bounds that the new distances undercut (flagged for
        // thread 0, which applies the heap changes in ascending z)
        {
            const double* rx = D + (size_t)x * n;
            double* ry = D + (size_t)y * n;
            for (int z = tid; z < n; z += LK_THREADS) {
                const int nz = csize[z];
                if (nz == 0 || z == y) continue;
                const double v = lk_new_dist<METHOD>(rx[z], ry[z], dist, nx, ny, nz);
                ry[z] = v;
                D[(size_t)z * n + y] = v;
                if (z < x && nbr[z] == (unsigned short)x) nbr[z] = (unsigned short)y;
                if (z < y && v < hv[ibk[z]]) {
                    nbr[z] = (unsigned short)y;
                    atomicOr(&flags[z >> 5], 1u << (z & 31));
                }
            }
        }
        __syncthreads();
        if (y < n - 1) scan_row(y);                   // y's own candidate (thread 0 takes part, then turns to the heap)
        __syncthreads();
        if (warp == 0) {
            LkKey best = LkKey{0xffffffffu, 0xffffffffu, 0x7fffffff};
            if (y < n - 1) best = scan_result();
            if (lane == 0) {
                heap.size = s_hsize;
                const double* ry = D + (size_t)y * n;
                for (int w = 0; w <= (y >> 5) && w < nflag; ++w) {
                    unsigned m = flags[w];
                    if (!m) continue;
                    flags[w] = 0u;
                    while (m) {
                        const int z = (w << 5) + __ffs((int)m) - 1;
                        m &= m - 1;
                        heap.change_value(z, ry[z]);
                    }
                }
                if (y < n - 1 && !lk_key_none(best)) {
                    nbr[y] = (unsigned short)best.i;
                    heap.change_value(y, ry[best.i]);
                }
            }
        }
        __syncthreads();
    }
}

