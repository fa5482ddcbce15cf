// Philox4x32-10 and the resample-index kernel (philox.cu includes this file; so does tests/emul/small_emul.cpp, which runs
// the SAME source on host threads against the Random123 known answers where there is no GPU).
#pragma once
#include <stdint.h>

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// grid = (ceil(ceil(p/4)/256), nb)
__global__ void philox_indices_kernel(uint64_t seed, long long b0, long long p, long long* __restrict__ idx,
                                      int* __restrict__ weights) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // counter = 4 draws
    const long long b = blockIdx.y;
    if (q * 4 >= p) return;
    uint32_t r[4];
    philox4x32_10((uint32_t)q, (uint32_t)(b0 + b), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int i = 0; i < 4; ++i) {
        const long long k = q * 4 + i;
        if (k >= p) break;
        const long long v = (long long)(((uint64_t)r[i] * (uint64_t)p) >> 32);
        if (idx) idx[b * p + k] = v;
        if (weights) atomicAdd(&weights[b * p + v], 1);
    }
}

