// The neighbourhood-weight kernel (nbhd_kernels.cu includes this file; so does tests/emul/gj_emul.cpp, which runs the SAME
// source on host threads against the golden outputs of the unmodified reference where there is no GPU).
#pragma once
#include <math.h>
#include <stdint.h>

#define NB_T 16        // pair tile edge: 16 x 16 pairs per CTA, one pair per thread
#define NB_KC 32       // columns per shared-memory stage

__device__ __forceinline__ double nbhd_value(double x, unsigned cls, double thr, double f1, double fk) {
    // cls: bit 0 = on the negative list, bit 1 = on the positive list (f1 / fk chosen by the caller)
    if (cls == 0) return (x < thr) ? 0.0 : x;
    const double x1 = x * f1, xk = x * fk;
    return (x1 < thr || xk < thr) ? 0.0 : xk;
}

// grid = (ceil(n/16) j-tiles, ceil(n/16) i-tiles); tiles below the diagonal exit.
// out [n][n]: cleaned GJ (NaN -> 0, negative -> 0, :201-205), times mul[i][j] when mul != NULL, mirrored.
__global__ void __launch_bounds__(NB_T* NB_T)
gj_nbhd_kernel(const double* __restrict__ tab, const uint8_t* __restrict__ cls, long long n, long long p, long long ld,
               double w, double thr, const double* __restrict__ mul, double* __restrict__ out) {
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bj < bi) return;
    __shared__ double sI[NB_T][NB_KC + 1], sJ[NB_T][NB_KC + 1];
    __shared__ uint8_t cI[NB_T][NB_KC], cJ[NB_T][NB_KC];
    const int tid = threadIdx.x, ti = tid / NB_T, tj = tid % NB_T;
    const long long i = (long long)bi * NB_T + ti, j = (long long)bj * NB_T + tj;
    // per-application factors by class, first application and the pair's application counts
    const double neg = 1.0 - w, pos = 1.0 + w;
    const double f1[4] = {1.0, neg, pos, neg * pos};
    // row i carries j + 2 applications, row j carries i + 1 (pairs below the diagonal of a diagonal tile are
    // computed with the roles swapped, so that (i, j) and (j, i) agree bit for bit)
    const long long lo = i < j ? i : j, hi = i < j ? j : i;
    const double k_lo = (double)(hi + 2), k_hi = (double)(lo + 1);
    double fa[4], fb[4];      // fa: factors of the row with the smaller index, fb: of the larger
    fa[0] = fb[0] = 1.0;
    for (int c = 1; c < 4; ++c) {
        fa[c] = pow(f1[c], k_lo);
        fb[c] = pow(f1[c], lo == hi ? k_lo : k_hi);
    }
    const bool swap = i > j;   // thread's "I" row is the larger index
    double smin = 0.0, smax = 0.0;
    for (long long c0 = 0; c0 < p; c0 += NB_KC) {
        for (int q = tid; q < 2 * NB_T * NB_KC; q += NB_T * NB_T) {
            const int which = q / (NB_T * NB_KC), r = (q / NB_KC) % NB_T, c = q % NB_KC;
            const long long row = (long long)(which ? bj : bi) * NB_T + r, col = c0 + c;
            double v = 0.0;
            uint8_t k = 0;
            if (row < n && col < p) { v = tab[row * ld + col]; k = cls[row * p + col] & 3; }
            if (which) { sJ[r][c] = v; cJ[r][c] = k; } else { sI[r][c] = v; cI[r][c] = k; }
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < NB_KC; ++c) {
            const unsigned ka = cI[ti][c], kb = cJ[tj][c];
            const double a = nbhd_value(sI[ti][c], ka, thr, f1[ka], swap ? fb[ka] : fa[ka]);
            const double b = nbhd_value(sJ[tj][c], kb, thr, f1[kb], swap ? fa[kb] : fb[kb]);
            // np.minimum / np.maximum propagate NaN
            smin += (a != a || b != b) ? (a + b) : fmin(a, b);
            smax += (a != a || b != b) ? (a + b) : fmax(a, b);
        }
        __syncthreads();
    }
    if (i >= n || j >= n) return;
    // `... / sum(max) if not sum(max) == 0 else 0` (:167-168); NaN and negatives are cleaned to 0 (:201-205)
    double g = (smax == 0.0) ? 0.0 : smin / smax;
    if (g != g || g < 0.0) g = 0.0;
    out[i * n + j] = mul ? g * mul[i * n + j] : g;
    if (bi != bj) out[j * n + i] = mul ? g * mul[j * n + i] : g;     // mirror (:169); diagonal tiles hold both threads
}

