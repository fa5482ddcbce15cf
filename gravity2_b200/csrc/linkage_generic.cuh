// Device helpers shared by the linkage kernels and the "centroid" / "median" kernel (linkage.cu includes this file; so does
// tests/emul/lk_generic_emul.cpp, which runs the SAME kernel source on host threads -- one per CUDA thread, barriers for
// __syncthreads and the warp reductions -- so that its logic is checked against scipy where there is no GPU).
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

