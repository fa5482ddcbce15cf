// Runs lk_generic_kernel (methods centroid / median) of gravity2_b200/csrc/linkage_generic.cuh -- the SAME source the GPU
// compiles -- on host threads (cuda_block_emul.h) with a block of 64 threads.
//   lk_generic_emul <method: 5 centroid | 6 median> <n> < matrix (n*n float64, binary) > linkage rows ((n-1)*4 float64)
#define LK_THREADS 64
#include "cuda_block_emul.h"
#define LK_DYNAMIC_SMEM(name) int* name = reinterpret_cast<int*>(emu_dynamic_smem)
#include <stdio.h>
#include <stdlib.h>

#include "../../gravity2_b200/csrc/linkage_generic.cuh"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const int method = atoi(argv[1]), n = atoi(argv[2]);
    std::vector<double> D((size_t)n * n), Z((size_t)(n > 1 ? n - 1 : 0) * 4);
    if (fread(D.data(), sizeof(double), D.size(), stdin) != D.size()) return 3;
    if (lk_generic_smem_bytes(n) > sizeof emu_dynamic_smem) return 4;
    if (n > 1)
        emu_launch_block(LK_THREADS, [&] {
            if (method == LK_CENTROID) lk_generic_kernel<LK_CENTROID>(D.data(), (long long)n * n, n, Z.data());
            else lk_generic_kernel<LK_MEDIAN>(D.data(), (long long)n * n, n, Z.data());
        });
    fwrite(Z.data(), sizeof(double), Z.size(), stdout);
    return 0;
}
