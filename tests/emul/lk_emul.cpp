// Runs the linkage kernels of gravity2_b200/csrc/linkage_kernels.cuh -- the SAME source the GPU build compiles -- on host
// threads (cuda_block_emul.h) with a block of 64 threads.
//   lk_emul <method: 0 average | 1 complete | 2 weighted | 3 ward | 4 single | 5 centroid | 6 median> <n> <columns: 0 eager | 1 lazy | 2 lazy + live list>
//       < matrix (n*n float64, binary)  > rows as the kernel writes them ((n-1)*4 float64, merge order)
#define LK_THREADS 64
#include "cuda_block_emul.h"
#define LK_DYNAMIC_SMEM(name) int* name = reinterpret_cast<int*>(emu_dynamic_smem)
#include <stdio.h>
#include <stdlib.h>

#include "../../gravity2_b200/csrc/linkage_kernels.cuh"

template <int METHOD>
static void chain(int lazy, double* D, int n, double* Z) {
    if (lazy == 2) lk_nn_chain_list_kernel<METHOD, LK_THREADS>(D, (long long)n * n, n, Z);
    else if (lazy) lk_nn_chain_kernel<METHOD, LK_THREADS, true>(D, (long long)n * n, n, Z);
    else lk_nn_chain_kernel<METHOD, LK_THREADS, false>(D, (long long)n * n, n, Z);
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int method = atoi(argv[1]), n = atoi(argv[2]);
    const int lazy = atoi(argv[3]);
    std::vector<double> D((size_t)n * n), Z((size_t)(n > 1 ? n - 1 : 0) * 4);
    if (fread(D.data(), sizeof(double), D.size(), stdin) != D.size()) return 3;
    if (lk_generic_smem_bytes(n) > sizeof emu_dynamic_smem || n > 32 * LK_THREADS || (lazy == 2 && n > LK_LIST_PER * LK_THREADS)) return 4;
    if (n > 1)
        emu_launch_block(LK_THREADS, [&] {
            switch (method) {
                case LK_AVERAGE: chain<LK_AVERAGE>(lazy, D.data(), n, Z.data()); break;
                case LK_COMPLETE: chain<LK_COMPLETE>(lazy, D.data(), n, Z.data()); break;
                case LK_WEIGHTED: chain<LK_WEIGHTED>(lazy, D.data(), n, Z.data()); break;
                case LK_WARD: chain<LK_WARD>(lazy, D.data(), n, Z.data()); break;
                case LK_SINGLE: lk_mst_single_kernel(D.data(), (long long)n * n, n, Z.data()); break;
                case LK_CENTROID: lk_generic_kernel<LK_CENTROID>(D.data(), (long long)n * n, n, Z.data()); break;
                default: lk_generic_kernel<LK_MEDIAN>(D.data(), (long long)n * n, n, Z.data()); break;
            }
        });
    fwrite(Z.data(), sizeof(double), Z.size(), stdout);
    return 0;
}
