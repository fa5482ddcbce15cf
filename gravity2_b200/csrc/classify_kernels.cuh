// Kernels of the classification epilogue (classify_kernels.cu includes this file; so does tests/emul/small_emul.cpp, which
// runs the SAME source on host threads so that np.max / np.argmax semantics are checked against numpy where there is no GPU).
#pragma once

// np.max / np.argmax semantics: a NaN wins (and the first NaN is the arg-maximum), else the first largest value
__device__ __forceinline__ bool cl_better(double v, long long i, double bv, long long bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (vn) return i < bi;
    return v > bv || (v == bv && i < bi);
}

// grid = rows r1 - r0, block = 256
__global__ void __launch_bounds__(256)
rowmax_block_kernel(const double* __restrict__ mat, long long ld, long long r0, long long c0, long long c1,
                    double* __restrict__ out_max, long long* __restrict__ out_arg) {
    const double* row = mat + (r0 + blockIdx.x) * ld;
    double bv = -INFINITY;
    long long bi = 0x7fffffffffffffffLL;
    for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
        const double v = row[c];
        if (cl_better(v, c, bv, bi)) { bv = v; bi = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (cl_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    __shared__ double sv[8];
    __shared__ long long si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (cl_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        out_max[blockIdx.x] = bv;
        out_arg[blockIdx.x] = bi - c0;               // index within the column block, as np.argmax of the slice
    }
}

__global__ void gather_cells_kernel(const double* __restrict__ mat, long long ld, const long long* __restrict__ ii,
                                    const long long* __restrict__ jj, long long count, double* __restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) out[k] = mat[ii[k] * ld + jj[k]];
}

