// The GOM scoring kernels: genome CSR, per-GOM point sets and distance caches, pair records, weighted row sums, aggregates,
// the light and heavy score kernels.  gom_kernels.cu includes this file; so does tests/emul/gom_emul.cpp, which runs the
// SAME kernel source on host threads (cuda_block_emul.h), following gom_db_build / gom_db_score step by step, against the
// reference's dcor known answer and the CPU restatement of generate_gom_sigs where there is no GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef GOM_EMULATED            // tests/emul: cp.async as cuda_block_emul.h models it (the __syncwarp after the wait hands the data to the warp)
static inline void gom_cp_async8(void* smem, const void* gmem) { emu_cp_async(smem, gmem, 8); }
#define GOM_CP_COMMIT() emu_cp_commit()
#define GOM_CP_WAIT(n) emu_cp_wait(n)
#define GOM_DYNAMIC_SMEM(name) unsigned char* name = reinterpret_cast<unsigned char*>(emu_dynamic_smem)
#else
__device__ __forceinline__ void gom_cp_async8(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
#define GOM_CP_COMMIT() asm volatile("cp.async.commit_group;\n" ::)
#define GOM_CP_WAIT(n) asm volatile("cp.async.wait_group %0;\n" ::"n"(n) : "memory")
#define GOM_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

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
                                double* __restrict__ nrm, double* __restrict__ nrm2) {
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
    nrm2[pt] = s;
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

// wT[c][b] = (double) w[b0 + b][c] for b < nb, else 0; wT8 the same as uint8 (staged in shared memory by the
// scoring kernels; multiplicities above 255 raise *overflow).   grid = ceil(P*NBC/256)
__global__ void weights_transpose_kernel(const int* __restrict__ w, long long p, int nb,
                                         double* __restrict__ wT, unsigned char* __restrict__ wT8,
                                         int* __restrict__ overflow) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * NBC) return;
    const long long c = i / NBC;
    const int b = (int)(i % NBC);
    const int v = (w == nullptr) ? (b < nb ? 1 : 0) : (b < nb ? w[(long long)b * p + c] : 0);
    if (v < 0 || v > 255) atomicExch(overflow, 1);
    wT[i] = (double)v;
    wT8[i] = (unsigned char)v;
}

// per-GOM gathered weights: wG[pt][b] = wT[pt_col[pt]][b]   grid = ceil(npt*NBC/256)
__global__ void gom_gather_weights_kernel(const int* __restrict__ pt_col, long long npt,
                                          const double* __restrict__ wT, double* __restrict__ wG) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npt * NBC) return;
    wG[i] = wT[(long long)pt_col[i / NBC] * NBC + (i % NBC)];
}

// GA: arT[pt][b] = sum_l w_l A[k][l]  (per GOM a [ng x ng] x [ng x 32 NH] product; the squared-distance sums the centred
// variance also needs do not go through this O(ng^2) loop any more: see gom_centroid_kernel).
// grid = (ceil(max_ng/128), G), block = (32 b, 8): the block owns 128 rows k, a thread 16 of them for its replicate lane --
// of NH = 2 chunks of 32 replicates at once: every distance read from shared memory (warp-uniform, two columns per LDS.128)
// then feeds two FMAs, which turns the kernel from shared-memory-load bound (16 + 2 loads per 32 DFMA, measured: l1tex 72 %,
// fp64 pipe 30 %) into fp64-pipe bound (16 + 4 loads per 64 DFMA), and the 5 GB distance cache is read once per 64 replicates.
#define GA_ROWS 128
#define GA_LC 32
#define GA_LD (GA_LC + 2)            // even row stride: 16-byte aligned column pairs
static size_t ga_smem_bytes(int nh) { return (size_t)GA_ROWS * GA_LD * sizeof(double) + (size_t)nh * GA_LC * NBC * sizeof(double); }
template <int NH>
__global__ void __launch_bounds__(256, 2)
gom_rowsums_kernel(const int* __restrict__ pt_ptr, const long long* __restrict__ a_off,
                   const double* __restrict__ A, const double* __restrict__ wG0, const double* __restrict__ wG1,
                   double* __restrict__ arT0, double* __restrict__ arT1) {
    GOM_DYNAMIC_SMEM(ga_smem);
    double (*sA)[GA_LD] = reinterpret_cast<double (*)[GA_LD]>(ga_smem);
    double (*sW)[GA_LC][NBC] = reinterpret_cast<double (*)[GA_LC][NBC]>(ga_smem + (size_t)GA_ROWS * GA_LD * sizeof(double));
    const int g = blockIdx.y;
    const int p0 = pt_ptr[g], ng = pt_ptr[g + 1] - p0;
    const int kb = blockIdx.x * GA_ROWS;
    if (kb >= ng) return;
    const int b = threadIdx.x, ty = threadIdx.y;
    const double* a = A + a_off[g];
    const double* w0 = wG0 + (long long)p0 * NBC;
    const double* w1 = NH > 1 ? wG1 + (long long)p0 * NBC : nullptr;
    double s1[16][NH];
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int h = 0; h < NH; ++h) s1[r][h] = 0.0;
    for (int l0 = 0; l0 < ng; l0 += GA_LC) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int row = ty * 16 + r, k = min(kb + row, ng - 1), l = l0 + b;
            sA[row][b] = l < ng ? a[(long long)k * ng + l] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < GA_LC / 8; ++q) {
            const int l = ty + 8 * q;
            sW[0][l][b] = l0 + l < ng ? w0[(long long)(l0 + l) * NBC + b] : 0.0;
            if (NH > 1) sW[1][l][b] = l0 + l < ng ? w1[(long long)(l0 + l) * NBC + b] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int l = 0; l < GA_LC; l += 2) {
            double wa[NH], wb2[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) { wa[h] = sW[h][l][b]; wb2[h] = sW[h][l + 1][b]; }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const double2 d = *reinterpret_cast<const double2*>(&sA[ty * 16 + r][l]);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    s1[r][h] = fma(wa[h], d.x, s1[r][h]);         // columns in ascending order
                    s1[r][h] = fma(wb2[h], d.y, s1[r][h]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int k = kb + ty * 16 + r;
        if (k < ng) {
            arT0[(long long)(p0 + k) * NBC + b] = s1[r][0];
            if (NH > 1) arT1[(long long)(p0 + k) * NBC + b] = s1[r][NH - 1];
        }
    }
}

// GB: per-GOM aggregates, lane = replicate.  gagg[g][q][b], q in 0..GAGG-1.
//   gom_centroid_kernel   SAA = sum_kl w_k w_l |X_k - X_l|^2 = 2 (W sum_k w_k |X_k|^2 - |sum_k w_k X_k|^2) needs the weighted
//            centroid C_j = sum_k w_k X_kj: O(n_g m_g) instead of the O(n_g^2) pass over the distance cache.  One block per
//            (GOM, tile of eight member coordinates): the warps split the points, a point costs one multiplicity load and
//            eight broadcast coordinate loads for eight FMAs; partial |C|^2 per tile go to cpart[g][tile][b]
//            (grid = (G, tiles): enough blocks to hide the load latency that a one-block-per-GOM loop exposed);
//   gom_aggregates_kernel the sums over the GOM's points (n_g, N1, N2, AR1, AR2, ARN), points dealt to the warps, reduced in
//            shared memory, plus the tiles' |C|^2 in tile order (deterministic);
//   XDEG = 1 when all drawn points coincide (the distances from the first drawn point sum to exactly 0): the reference's
//            centred matrix is then exactly 0 (dcor.py:54-60), which the moment formula does not reproduce bit for bit.
#define GAGG 8   // n_g, N1, N2, AR1, AR2, ARN, SAA, XDEG
__global__ void __launch_bounds__(256)
gom_centroid_kernel(const int* __restrict__ pt_ptr, const int* __restrict__ mem_ptr, const long long* __restrict__ x_off,
                    const double* __restrict__ X, const double* __restrict__ wG, int ntile, double* __restrict__ cpart) {
    __shared__ double cred[8][8][NBC];
    const int g = blockIdx.x, tile = blockIdx.y, b = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int p0 = pt_ptr[g], ng = pt_ptr[g + 1] - p0;
    const int mg = mem_ptr[g + 1] - mem_ptr[g];
    const int j0 = tile * 8;
    double* o = cpart + ((long long)g * ntile + tile) * NBC + b;
    if (j0 >= mg) { if (wi == 0) *o = 0.0; return; }
    const double* x = X + x_off[g] + j0;
    const int jn = min(8, mg - j0);
    double c[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q] = 0.0;
#pragma unroll 4
    for (int k = wi; k < ng; k += 8) {
        const double w = wG[(long long)(p0 + k) * NBC + b];
        const double* xr = x + (long long)k * mg;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < jn) c[q] = fma(w, __ldg(xr + q), c[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) cred[wi][q][b] = c[q];
    __syncthreads();
    if (wi == 0) {
        double cs = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double t = 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) t += cred[u][q][b];
            cs = fma(t, t, cs);
        }
        *o = cs;
    }
}

__global__ void __launch_bounds__(256)
gom_aggregates_kernel(const int* __restrict__ pt_ptr, const double* __restrict__ nrm, const double* __restrict__ nrm2,
                      const double* __restrict__ wG, const double* __restrict__ arT, const double* __restrict__ cpart, int ntile,
                      double* __restrict__ gagg) {
    __shared__ double red[8][6][NBC];
    const int g = blockIdx.x, b = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int p0 = pt_ptr[g], ng = pt_ptr[g + 1] - p0;
    double n = 0, N1 = 0, N2 = 0, AR1 = 0, AR2 = 0, ARN = 0;
#pragma unroll 4
    for (int k = wi; k < ng; k += 8) {
        const double w = wG[(long long)(p0 + k) * NBC + b];
        const double nr = nrm[p0 + k], ar = arT[(long long)(p0 + k) * NBC + b];
        n += w;
        N1 = fma(w, nr, N1);
        N2 = fma(w, nrm2[p0 + k], N2);
        AR1 = fma(w, ar, AR1);
        AR2 = fma(w * ar, ar, AR2);
        ARN = fma(w * ar, nr, ARN);
    }
    red[wi][0][b] = n; red[wi][1][b] = N1; red[wi][2][b] = N2; red[wi][3][b] = AR1; red[wi][4][b] = AR2; red[wi][5][b] = ARN;
    __syncthreads();
    if (wi == 0) {
        double t[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            t[q] = 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) t[q] += red[u][q][b];
        }
        double cs = 0.0;
        for (int tile = 0; tile < ntile; ++tile) cs += cpart[((long long)g * ntile + tile) * NBC + b];
        // first drawn point of the replicate and its distance row sum
        double xdeg = 1.0;
        for (int k = 0; k < ng; ++k)
            if (wG[(long long)(p0 + k) * NBC + b] > 0.0) { xdeg = arT[(long long)(p0 + k) * NBC + b] == 0.0 ? 1.0 : 0.0; break; }
        double* o = gagg + (long long)g * GAGG * NBC + b;
        o[0 * NBC] = t[0]; o[1 * NBC] = t[1]; o[2 * NBC] = t[2]; o[3 * NBC] = t[3];
        o[4 * NBC] = t[4]; o[5 * NBC] = t[5];
        o[6 * NBC] = xdeg != 0.0 ? 0.0 : 2.0 * (t[0] * t[2] - cs);
        o[7 * NBC] = xdeg;
    }
}

#define VAGG 8   // t, Y1, Y2, Ys, RR1, RR2, RY, YC (1 = all drawn hits share one location value): per-genome aggregates of the weighted row sums r_c = sum_d w_d |y_c - y_d|
// Order of a genome's hits by location value (ties by hit index), by rank counting: ord[o + rank] = hit index.
// Built once per database.  grid = n, block = 128; the values are staged in shared memory (dynamic: t * 8 bytes).
__global__ void genome_order_kernel(const long long* __restrict__ ptr, const double* __restrict__ val, int* __restrict__ ord) {
    GOM_DYNAMIC_SMEM(vo_smem);
    double* sy = reinterpret_cast<double*>(vo_smem);
    const long long o = ptr[blockIdx.x];
    const int t = (int)(ptr[blockIdx.x + 1] - o);
    for (int d = threadIdx.x; d < t; d += blockDim.x) sy[d] = val[o + d];
    __syncthreads();
    for (int c = threadIdx.x; c < t; c += blockDim.x) {
        const double yc = sy[c];
        const bool nc = yc != yc;                  // NaN sorts last (a total order keeps the permutation valid)
        int rank = 0;
        for (int d = 0; d < t; ++d) {
            const double yd = sy[d];
            const bool nd = yd != yd;
            rank += (!nd && nc) || (nd == nc && (yd < yc || ((yd == yc || nc) && d < c)));
        }
        ord[o + rank] = c;
    }
}

// VA by prefix sums: with the hits in location order, r_c = sum_d w_d |y_c - y_d| = y_c (W_<c - W_>c) - (S_<c - S_>c),
// W / S the running sums of w and w y -- t steps per genome and replicate instead of t^2.  One warp per genome,
// lane = replicate.  grid = ceil(n / 4), block = 128
__global__ void __launch_bounds__(128)
genome_rowsums_sorted_kernel(long long n, const long long* __restrict__ ptr, const int* __restrict__ col,
                             const double* __restrict__ val, const int* __restrict__ ord,
                             const unsigned char* __restrict__ wT8, double* __restrict__ rT, double* __restrict__ vagg) {
    const long long v = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int b = threadIdx.x & 31;
    if (v >= n) return;
    const long long o = ptr[v];
    const int t = (int)(ptr[v + 1] - o);
    double Wtot = 0.0, Stot = 0.0;
    for (int c = 0; c < t; ++c) {
        const double w = (double)wT8[(long long)col[o + c] * NBC + b];
        Wtot += w;
        Stot = fma(w, val[o + c], Stot);
    }
    double Wlt = 0.0, Slt = 0.0;
    double T = 0, Y1 = 0, Y2 = 0, Ys = 0, RR1 = 0, RR2 = 0, RY = 0;
    double ylo = 0.0, yhi = 0.0;
    bool seen = false, yconst = true;
    for (int k = 0; k < t; ++k) {
        const int c = ord[o + k];
        const double yc = val[o + c], ay = fabs(yc);
        const double w = (double)wT8[(long long)col[o + c] * NBC + b];
        if (w > 0.0) {                              // hits come in location order: first and last drawn value bracket the rest
            if (!seen) { ylo = yc; seen = true; }
            yhi = yc;
            yconst = ylo == yhi;
        }
        // hits tied with y_c contribute |y_c - y_d| = 0 on either side, so "before / after in the order" is enough
        const double r = yc * (2.0 * Wlt + w - Wtot) - (2.0 * Slt + w * yc - Stot);
        rT[(o + c) * NBC + b] = r;
        T += w;
        Y1 = fma(w, ay, Y1);
        Y2 = fma(w * yc, yc, Y2);
        Ys = fma(w, yc, Ys);
        RR1 = fma(w, r, RR1);
        RR2 = fma(w * r, r, RR2);
        RY = fma(w * r, ay, RY);
        Wlt += w;
        Slt = fma(w, yc, Slt);
    }
    double* q = vagg + v * VAGG * NBC + b;
    q[0 * NBC] = T; q[1 * NBC] = Y1; q[2 * NBC] = Y2; q[3 * NBC] = Ys; q[4 * NBC] = RR1; q[5 * NBC] = RR2; q[6 * NBC] = RY;
    q[7 * NBC] = yconst ? 1.0 : 0.0;
}

// VG: score (genome, GOM) pairs; lane = replicate.  out[b][v][g].
// Everything about a pair that does not depend on the replicate is computed ONCE per database and stored as the pair's
// RECORD (genome-major: pair id = v * G + g), so that the per-chunk scoring kernels stream contiguous data instead of
// gathering 8-byte elements from the distance cache:
//   entries  the intersection list I = R_v & R_g, 32 bytes each: GOM point index k, genome hit index c, PPHMM column,
//            |X_k| and the genome's (signed) location value y;
//   M        the lower triangle (a > e, row-major) of the replicate-independent pair matrix
//            M_ae = 2 [ d_ae (|y_a - y_e| - |y_a| - |y_e|) - (|X_a| + |X_e|) |y_a - y_e| ],  d_ae = |X_a - X_e|:
//            the three I x I double sums of the expansion enter S_ab only as Q1 - 2 Q2 - 2 Q3 = sum_{a>e} w_a w_e M_ae.
//   light kernel: one block per genome, its warps walk the genome's GOMs (the per-genome aggregates stay in registers, the
//                 genome's weighted row sums stay in L1/L2); pairs with up to VG_LIGHT_CAP entries (a genome against a
//                 foreign taxon);
//   heavy kernel: one block per listed pair (a genome against its own taxon: |I| ~ all of its hits); the four warps split
//                 the rows of M, each row is one coalesced read.
#define VG_ENTRY_BYTES 56
#define VG_LIGHT_CAP 32
#define VG_LIGHT_BYTES (VG_LIGHT_CAP * 32 + VG_LIGHT_CAP * (VG_LIGHT_CAP - 1) / 2 * 8 + VG_LIGHT_CAP * NBC)   // entries + M triangle + weights
#define VG_NSUM 12

struct VgEntry { int k, c, col, pad; double nr, y; };      // 32 bytes
struct VgSums { double v[VG_NSUM]; };   // SW S_ar S_nrm S_r S_y C_arr C_ary C_nr C_ny Q1 Q2 Q3

__device__ __forceinline__ double vg_finish(const VgSums& S, const double* __restrict__ ga, const double* __restrict__ va) {
    const double SW = S.v[0], S_ar = S.v[1], S_nrm = S.v[2], S_r = S.v[3], S_y = S.v[4], C_arr = S.v[5], C_ary = S.v[6],
                 C_nr = S.v[7], C_ny = S.v[8], Q1 = S.v[9], Q2 = S.v[10], Q3 = S.v[11];
    const double n_g = ga[0], N1 = ga[1], N2 = ga[2], AR1 = ga[3], AR2 = ga[4], ARN = ga[5], SAA = ga[6];
    const double T = va[0], Y1 = va[1], Y2 = va[2], Ys = va[3], RR1 = va[4], RR2 = va[5], RY = va[6];
    const double e = T - SW, z = n_g - SW, nn = n_g + e;
    double res = 0.0;
    // Exactly degenerate Y: no zero-valued point (z == 0; the counts are small integers, exact) and every drawn hit at one
    // location value.  The reference's centred matrix is then exactly 0 and it returns 0.0 (dcor.py:54-60); the un-centred
    // expansion below would leave rounding noise in sBB instead.
    const bool y_degenerate = z == 0.0 && va[7] != 0.0;
    // Exactly degenerate X: no point outside the GOM's column set (e == 0) and all drawn GOM points coincide.
    const bool x_degenerate = e == 0.0 && ga[7] != 0.0;
    if (nn > 0.0 && !y_degenerate && !x_degenerate) {
        const double Sab = 2.0 * (C_ary - Q2) + 2.0 * (N1 - S_nrm) * (Y1 - S_y) + Q1 + 2.0 * (C_nr - Q3);
        const double cross = ((AR1 - S_ar) + e * (N1 - S_nrm)) * Y1 +
                             (C_arr + z * C_ary + e * C_nr + e * z * C_ny) +
                             N1 * ((RR1 - S_r) + z * (Y1 - S_y));
        const double a_tot = AR1 + 2.0 * e * N1, b_tot = RR1 + 2.0 * z * Y1;
        const double rowA = AR2 + 2.0 * e * ARN + e * e * N2 + e * N1 * N1;
        const double rowB = RR2 + 2.0 * z * RY + z * z * Y2 + z * Y1 * Y1;
        const double Saa = SAA + 2.0 * e * N2;
        const double Sbb = 2.0 * T * Y2 - 2.0 * Ys * Ys + 2.0 * z * Y2;
        // one reciprocal instead of nine divisions: centred sums = un-centred sum - 2/n (row products) + (total products)/n^2
        const double inn = 1.0 / nn, inn2 = inn * inn;
        const double sAB = Sab - 2.0 * inn * cross + a_tot * b_tot * inn2;
        const double sAA = Saa - 2.0 * inn * rowA + a_tot * a_tot * inn2;
        const double sBB = Sbb - 2.0 * inn * rowB + b_tot * b_tot * inn2;
        // dcor.py:10,17,54-60: dcov = sqrt(sAB) / n, dvar = sqrt(s..) / n, dcor = dcov / sqrt(dvarA dvarB) -- the n cancel:
        // dcor = sqrt(sAB / sqrt(sAA sBB)); NaN if sAB rounds below zero, as numpy; 0 unless both variances are > 0
        if (sAA > 0.0 && sBB > 0.0) res = sqrt(sAB / sqrt(sAA * sBB));
    }
    return res;
}

// Pair records, pass 1: |I| of every pair.  One warp per pair, genome-major.  grid = ceil(n*G/8), block = 256
__global__ void __launch_bounds__(256)
gom_pair_count_kernel(long long n, int G, long long p, const long long* __restrict__ ptr, const int* __restrict__ col,
                      const int* __restrict__ pos, int cap, int* __restrict__ ni_out, long long* __restrict__ bytes_out) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wid >= n * G) return;
    const long long v = wid / G;
    const int g = (int)(wid % G);
    const long long o = ptr[v];
    const int t = (int)(ptr[v + 1] - o);
    const int* posg = pos + (long long)g * p;
    int ni = 0;
    for (int c0 = 0; c0 < t; c0 += 32) {
        const int c = c0 + lane;
        const int k = c < t ? posg[col[o + c]] : -1;
        ni += __popc(__ballot_sync(0xffffffffu, k >= 0));
    }
    if (lane == 0) {
        ni_out[wid] = ni;
        // records start 32-byte aligned: the entries are read and written with 16-byte accesses.  Light records (|I| <= cap)
        // are laid out first, genome-major, so that a genome's light records are ONE contiguous range; heavy ones follow
        const long long sz = (long long)ni * 32 + (((long long)ni * (ni - 1) / 2 * 8 + 31) & ~31LL);
        bytes_out[wid] = ni <= cap ? sz : 0;
        bytes_out[n * G + wid] = ni <= cap ? 0 : sz;
    }
}

// pass 2: entries (column order preserved) and the M triangle of every pair
__global__ void __launch_bounds__(256)
gom_pair_fill_kernel(long long n, int G, long long p, const long long* __restrict__ ptr, const int* __restrict__ col,
                     const double* __restrict__ val, const int* __restrict__ pos, const int* __restrict__ pt_ptr,
                     const double* __restrict__ nrm, const long long* __restrict__ a_off, const double* __restrict__ A,
                     const int* __restrict__ pr_ni, const long long* __restrict__ pr_ptr, int cap, unsigned char* __restrict__ rec) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wid >= n * G) return;
    const int ni = pr_ni[wid];
    if (ni == 0) return;
    const long long v = wid / G;
    const int g = (int)(wid % G);
    const long long o = ptr[v];
    const int t = (int)(ptr[v + 1] - o);
    const int* posg = pos + (long long)g * p;
    const int p0 = pt_ptr[g], ng = pt_ptr[g + 1] - p0;
    VgEntry* ent = reinterpret_cast<VgEntry*>(rec + pr_ptr[ni <= cap ? wid : n * G + wid]);
    int at = 0;
    for (int c0 = 0; c0 < t; c0 += 32) {
        const int c = c0 + lane;
        int k = -1, cc = 0;
        if (c < t) { cc = col[o + c]; k = posg[cc]; }
        const unsigned m = __ballot_sync(0xffffffffu, k >= 0);
        if (k >= 0) {
            VgEntry e;
            e.k = k; e.c = c; e.col = cc; e.pad = 0; e.nr = nrm[p0 + k]; e.y = val[o + c];
            ent[at + __popc(m & ((1u << lane) - 1u))] = e;
        }
        at += __popc(m);
    }
    __syncwarp();
    double* M = reinterpret_cast<double*>(ent + ni);
    const double* Ag = A + a_off[g];
    const long long npair = (long long)ni * (ni - 1) / 2;
    for (long long q = lane; q < npair; q += 32) {
        int a = (int)((1.0 + sqrt(1.0 + 8.0 * (double)q)) * 0.5);
        while ((long long)a * (a - 1) / 2 > q) --a;
        while ((long long)(a + 1) * a / 2 <= q) ++a;
        const int e = (int)(q - (long long)a * (a - 1) / 2);
        const VgEntry ea = ent[a], ee = ent[e];
        const double dist = Ag[(long long)ea.k * ng + ee.k];
        const double dy = fabs(ea.y - ee.y);
        M[q] = 2.0 * (dist * (dy - fabs(ea.y) - fabs(ee.y)) - (ea.nr + ee.nr) * dy);
    }
}

// work list of the heavy kernel: pairs whose list does not fit the light kernel's shared-memory slots
__global__ void gom_heavy_list_kernel(long long total, const int* __restrict__ pr_ni, int cap,
                                      long long* __restrict__ heavy_list, int* __restrict__ heavy_count) {
    const long long wid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (wid >= total) return;
    if (pr_ni[wid] > cap) heavy_list[atomicAdd(heavy_count, 1)] = wid;
}

// sum over pairs of |I| (|I| - 1) / 2 and of |I|: the pair-sum terms one replicate costs (roofline bookkeeping)
__global__ void gom_mentries_kernel(long long total, const int* __restrict__ pr_ni, unsigned long long* __restrict__ out) {
    unsigned long long s = 0, e = 0;
    for (long long wid = (long long)blockIdx.x * blockDim.x + threadIdx.x; wid < total; wid += (long long)gridDim.x * blockDim.x) {
        const unsigned long long ni = (unsigned long long)pr_ni[wid];
        s += ni * (ni - (ni ? 1 : 0)) / 2;
        e += ni;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
    if ((threadIdx.x & 31) == 0) { if (s) atomicAdd(out, s); if (e) atomicAdd(out + 1, e); }
}

// grid = n (one block per genome), block = 128 (4 warps over the genome's GOMs).
// A genome's light records are one contiguous range: the block copies them to shared memory in large coalesced pieces
// (LIGHT_REC_BYTES at a time) together with the pairs' offsets, so a pair's replicate-independent data costs no global
// round trip of its own; per pair the replicate-dependent gathers (multiplicity byte, GOM-side row sum, genome-side row sum
// of each entry) are issued four entries at a time before they are consumed, and the per-GOM aggregates are requested up
// front -- one memory latency per pair instead of five dependent ones.
#define LIGHT_GCHUNK 128                 // pairs whose offsets are staged at a time: one per thread
#define LIGHT_REC_BYTES (24 * 1024)
#define LIGHT_WSTRIDE (VG_LIGHT_CAP + 4)   // bytes per lane of the multiplicity staging: 9 words, odd -> no bank conflicts
__global__ void __launch_bounds__(128, 5)
gom_score_light_kernel(long long n, int G, const long long* __restrict__ ptr, const int* __restrict__ pt_ptr,
                       const int* __restrict__ pr_ni, const long long* __restrict__ pr_ptr, const unsigned char* __restrict__ rec,
                       const unsigned char* __restrict__ wT8, const double* __restrict__ arT, const double* __restrict__ rT,
                       const double* __restrict__ gagg, const double* __restrict__ vagg, int nb, double* __restrict__ out) {
    __shared__ __align__(16) unsigned char srec[LIGHT_REC_BYTES];
    __shared__ long long soff[LIGHT_GCHUNK + 1];
    __shared__ int sni[LIGHT_GCHUNK];
    __shared__ int spt[LIGHT_GCHUNK];
    __shared__ int s_next;
    __shared__ __align__(4) unsigned char sw_all[4][NBC * LIGHT_WSTRIDE];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const long long v = blockIdx.x;
    const long long o = ptr[v];
    unsigned char* sw = sw_all[wib];
    const int b = lane;
    double va[VAGG];
#pragma unroll
    for (int q = 0; q < VAGG; ++q) va[q] = vagg[(v * VAGG + q) * NBC + b];
    for (int gc0 = 0; gc0 < G; gc0 += LIGHT_GCHUNK) {
        const int gcn = min(LIGHT_GCHUNK, G - gc0);
        __syncthreads();
        if (tid <= gcn) soff[tid] = pr_ptr[v * G + gc0 + tid];
        if (tid == 0 && gcn == LIGHT_GCHUNK) soff[LIGHT_GCHUNK] = pr_ptr[v * G + gc0 + LIGHT_GCHUNK];
        if (tid < gcn) { sni[tid] = pr_ni[v * G + gc0 + tid]; spt[tid] = pt_ptr[gc0 + tid]; }
        __syncthreads();
        int g0 = 0;
        while (g0 < gcn) {
            // pairs [g0, g1) of the chunk whose records fit the staging buffer (offsets are non-decreasing, a light record is at
            // most 5 KB: at least one pair fits; heavy pairs take no room here)
            const long long base_off = soff[g0];
            const int g1 = g0 + __syncthreads_count(tid >= g0 && tid < gcn && soff[tid + 1] - base_off <= LIGHT_REC_BYTES);
            const int bytes = (int)(soff[g1] - base_off);
            {
                const uint4* src = reinterpret_cast<const uint4*>(rec + base_off);
                uint4* dst = reinterpret_cast<uint4*>(srec);
                for (int i = tid; i < bytes / 16; i += 128) dst[i] = __ldg(src + i);
            }
            if (tid == 0) s_next = g0;
            __syncthreads();
            // the warps draw the staged pairs from a shared counter: their costs differ (|I|^2), a fixed split left warps
            // waiting at the barrier below (ncu: 1.7 barrier-stalled warps per issue)
            while (true) {
                int gi = 0;
                if (lane == 0) gi = atomicAdd(&s_next, 1);
                gi = __shfl_sync(0xffffffffu, gi, 0);
                if (gi >= g1) break;
                const int ni = sni[gi];
                if (ni > VG_LIGHT_CAP) continue;                        // on the heavy kernel's work list
                const int g = gc0 + gi;
                const int p0 = spt[gi];
                double ga[GAGG];
#pragma unroll
                for (int q = 0; q < GAGG; ++q) ga[q] = gagg[((long long)g * GAGG + q) * NBC + b];
                const VgEntry* se = reinterpret_cast<const VgEntry*>(srec + (soff[gi] - base_off));
                const double* sd = reinterpret_cast<const double*>(se + ni);    // [a (a - 1) / 2 + e], e < a
                VgSums S;
#pragma unroll
                for (int i = 0; i < VG_NSUM; ++i) S.v[i] = 0.0;
                __syncwarp();                                           // the previous pair's weights have been consumed
                for (int a0 = 0; a0 < ni; a0 += 4) {
                    // four entries' gathers in flight before the first is consumed
                    double ar4[4], r4[4], nr4[4], ay4[4];
                    unsigned char w4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int a = min(a0 + u, ni - 1);
                        const VgEntry ea = se[a];
                        w4[u] = wT8[(long long)ea.col * NBC + b];
                        ar4[u] = arT[(long long)(p0 + ea.k) * NBC + b];
                        r4[u] = rT[(o + ea.c) * NBC + b];
                        nr4[u] = ea.nr;
                        ay4[u] = fabs(ea.y);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (a0 + u < ni) {
                            const double wc = (double)w4[u], ar = ar4[u], r = r4[u], nr = nr4[u], ayc = ay4[u];
                            sw[b * LIGHT_WSTRIDE + a0 + u] = w4[u];
                            S.v[0] += wc;
                            S.v[1] = fma(wc, ar, S.v[1]);
                            S.v[2] = fma(wc, nr, S.v[2]);
                            S.v[3] = fma(wc, r, S.v[3]);
                            S.v[4] = fma(wc, ayc, S.v[4]);
                            S.v[5] = fma(wc * ar, r, S.v[5]);
                            S.v[6] = fma(wc * ar, ayc, S.v[6]);
                            S.v[7] = fma(wc * nr, r, S.v[7]);
                            S.v[8] = fma(wc * nr, ayc, S.v[8]);
                        }
                    }
                }
                // the pair sums: sum_{a > e} w_a w_e M_ae.  A lane keeps its own multiplicity bytes contiguous (row stride 36 bytes:
                // conflict-free word loads), so one 32-bit load feeds four FMAs
                double qs = 0.0;
                const unsigned char* wb = sw + b * LIGHT_WSTRIDE;
                for (int a = 1; a < ni; ++a) {
                    const double* mrow = sd + a * (a - 1) / 2;
                    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
                    int e = 0;
                    for (; e + 3 < a; e += 4) {
                        const unsigned w4e = *reinterpret_cast<const unsigned*>(wb + e);
                        q0 = fma((double)(w4e & 0xffu), mrow[e], q0);
                        q1 = fma((double)((w4e >> 8) & 0xffu), mrow[e + 1], q1);
                        q2 = fma((double)((w4e >> 16) & 0xffu), mrow[e + 2], q2);
                        q3 = fma((double)(w4e >> 24), mrow[e + 3], q3);
                    }
                    for (; e < a; ++e) q0 = fma((double)wb[e], mrow[e], q0);
                    qs = fma((double)wb[a], (q0 + q1) + (q2 + q3), qs);
                }
                S.v[9] = qs;                                            // enters vg_finish as Q1 (Q2 = Q3 = 0)
                if (lane < nb) out[((long long)lane * n + v) * G + g] = vg_finish(S, ga, va);
            }
            __syncthreads();
            g0 = g1;
        }
    }
}

// persistent blocks over the heavy list; dynamic shared memory: cap * 56 bytes + reduction scratch + 4 rows of M
__global__ void __launch_bounds__(128)
gom_score_heavy_kernel(long long n, int G, int max_t, const long long* __restrict__ ptr, const int* __restrict__ pt_ptr,
                       const int* __restrict__ pr_ni, const long long* __restrict__ pr_ptr, const unsigned char* __restrict__ rec,
                       const unsigned char* __restrict__ wT8, const double* __restrict__ arT, const double* __restrict__ rT,
                       const double* __restrict__ gagg, const double* __restrict__ vagg, int nb, double* __restrict__ out,
                       const long long* __restrict__ heavy_list, int total) {
    GOM_DYNAMIC_SMEM(vg_smem);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int cap = (max_t + 3) & ~3;                        // multiple of 4: 16-byte aligned rows of M
    const int wstride = cap + ((cap >> 2) & 1 ? 0 : 4);      // bytes per replicate lane; wstride / 4 odd
    double* sy = reinterpret_cast<double*>(vg_smem);
    double* sn = sy + cap;
    int* sk = reinterpret_cast<int*>(sn + cap);
    int* sc = sk + cap;
    unsigned char* sw = reinterpret_cast<unsigned char*>(sc + cap);
    double* red = reinterpret_cast<double*>(vg_smem + (size_t)cap * VG_ENTRY_BYTES + 128);   // [3][VG_NSUM][32]  (sw: 32 * wstride <= 32 cap + 128)
    double* mrows = red + 3 * VG_NSUM * 32;                                             // [4 warps][2 rows][cap]
    for (int it = blockIdx.x; it < total; it += gridDim.x) {
        const long long wid = heavy_list[it];
        const long long v = wid / G;
        const int g = (int)(wid % G);
        const long long o = ptr[v];
        const int p0 = pt_ptr[g];
        __syncthreads();                                                // previous pair fully consumed
        const int ni = pr_ni[wid];
        const VgEntry* ent = reinterpret_cast<const VgEntry*>(rec + pr_ptr[n * G + wid]);     // heavy records follow the light ones
        const double* Mrec = reinterpret_cast<const double*>(ent + ni);
        for (int a = threadIdx.x; a < ni; a += 128) {
            const VgEntry e = ent[a];
            sy[a] = e.y; sn[a] = e.nr; sk[a] = e.k; sc[a] = e.c;
            // (the 32 multiplicity bytes of the entry's column, stored per replicate lane: sw[b][a], so that a lane reads four
            //  consecutive entries' multiplicities with one 32-bit load; wstride / 4 is odd: conflict-free)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned w4 = reinterpret_cast<const unsigned*>(wT8 + (long long)e.col * NBC)[q];
                sw[(4 * q + 0) * wstride + a] = (unsigned char)(w4 & 0xffu);
                sw[(4 * q + 1) * wstride + a] = (unsigned char)((w4 >> 8) & 0xffu);
                sw[(4 * q + 2) * wstride + a] = (unsigned char)((w4 >> 16) & 0xffu);
                sw[(4 * q + 3) * wstride + a] = (unsigned char)(w4 >> 24);
            }
        }
        __syncthreads();
        VgSums S;
#pragma unroll
        for (int i = 0; i < VG_NSUM; ++i) S.v[i] = 0.0;
        // The four warps split the rows a of M (cost ~ a): a = wib, wib + 4, ...  Row a is one coalesced asynchronous copy of a
        // doubles into one of the warp's two shared-memory rows: the next row is in flight while lanes-as-replicates take
        // one FMA per entry of the current one.
        {
            double* mrow2 = mrows + wib * 2 * cap;
            const int b = lane;
            const unsigned char* wb = sw + b * wstride;
            auto fetch_row = [&](int a, int buf) {
                if (a < ni) {
                    const double* mr = Mrec + (long long)a * (a - 1) / 2;
                    double* dst = mrow2 + buf * cap;
                    for (int e = lane; e < a; e += 32) gom_cp_async8(dst + e, mr + e);
                }
                GOM_CP_COMMIT();
            };
            fetch_row(wib, 0);
            // the two per-replicate row sums of entry a + 4 are requested while entry a is consumed
            double ar_n = 0.0, r_n = 0.0;
            if (wib < ni) { ar_n = arT[(long long)(p0 + sk[wib]) * NBC + b]; r_n = rT[(o + sc[wib]) * NBC + b]; }
            int it = 0;
            for (int a = wib; a < ni; a += 4, ++it) {
                fetch_row(a + 4, (it + 1) & 1);
                const double yc = sy[a], ayc = fabs(yc), nr = sn[a];
                const double wc = (double)wb[a];
                const double ar = ar_n, r = r_n;
                if (a + 4 < ni) { ar_n = arT[(long long)(p0 + sk[a + 4]) * NBC + b]; r_n = rT[(o + sc[a + 4]) * NBC + b]; }
                GOM_CP_WAIT(1);
                __syncwarp();
                const double* mrow = mrow2 + (it & 1) * cap;
                S.v[0] += wc;
                S.v[1] = fma(wc, ar, S.v[1]);
                S.v[2] = fma(wc, nr, S.v[2]);
                S.v[3] = fma(wc, r, S.v[3]);
                S.v[4] = fma(wc, ayc, S.v[4]);
                S.v[5] = fma(wc * ar, r, S.v[5]);
                S.v[6] = fma(wc * ar, ayc, S.v[6]);
                S.v[7] = fma(wc * nr, r, S.v[7]);
                S.v[8] = fma(wc * nr, ayc, S.v[8]);
                double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
                int e = 0;
                for (; e + 3 < a; e += 4) {
                    const unsigned w4e = *reinterpret_cast<const unsigned*>(wb + e);
                    const double2 m01 = *reinterpret_cast<const double2*>(mrow + e), m23 = *reinterpret_cast<const double2*>(mrow + e + 2);
                    q0 = fma((double)(w4e & 0xffu), m01.x, q0);
                    q1 = fma((double)((w4e >> 8) & 0xffu), m01.y, q1);
                    q2 = fma((double)((w4e >> 16) & 0xffu), m23.x, q2);
                    q3 = fma((double)(w4e >> 24), m23.y, q3);
                }
                for (; e < a; ++e) q0 = fma((double)wb[e], mrow[e], q0);
                S.v[9] = fma(wc, (q0 + q1) + (q2 + q3), S.v[9]);       // enters vg_finish as Q1 (Q2 = Q3 = 0)
                __syncwarp();                                           // the row is overwritten by the fetch after next
            }
            GOM_CP_WAIT(0);
        }
        if (wib > 0) {
#pragma unroll
            for (int i = 0; i < VG_NSUM; ++i) red[((wib - 1) * VG_NSUM + i) * 32 + lane] = S.v[i];
        }
        __syncthreads();
        if (wib == 0) {
#pragma unroll
            for (int i = 0; i < VG_NSUM; ++i)
                S.v[i] += red[(0 * VG_NSUM + i) * 32 + lane] + red[(1 * VG_NSUM + i) * 32 + lane] + red[(2 * VG_NSUM + i) * 32 + lane];
            double ga[GAGG], va[VAGG];
#pragma unroll
            for (int q = 0; q < GAGG; ++q) ga[q] = gagg[((long long)g * GAGG + q) * NBC + lane];
#pragma unroll
            for (int q = 0; q < VAGG; ++q) va[q] = vagg[(v * VAGG + q) * NBC + lane];
            if (lane < nb) out[((long long)lane * n + v) * G + g] = vg_finish(S, ga, va);
        }
    }
}

