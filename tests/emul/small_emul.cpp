// Runs the small kernels of the classification epilogue and of the device-RNG mode -- the SAME source the GPU build compiles
// (gravity2_b200/csrc/classify_kernels.cuh, philox_kernels.cuh) -- on host threads, with the launch shapes of their .cu files.
//   small_emul rowmax <rows> <ld> <r0> <r1> <c0> <c1>  < matrix (rows*ld float64)   > max (r1-r0 float64), arg (r1-r0 int64)
//   small_emul gather <rows> <ld> <count>              < matrix, ii (int64), jj (int64)  > cells (count float64)
//   small_emul philox <seed> <b0> <nb> <p>             > idx (nb*p int64), weights (nb*p int32)
#include "cuda_block_emul.h"
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "../../gravity2_b200/csrc/classify_kernels.cuh"
#include "../../gravity2_b200/csrc/philox_kernels.cuh"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string what = argv[1];
    if (what == "rowmax" && argc == 8) {
        const long long rows = atoll(argv[2]), ld = atoll(argv[3]), r0 = atoll(argv[4]), r1 = atoll(argv[5]), c0 = atoll(argv[6]),
                        c1 = atoll(argv[7]);
        std::vector<double> mat((size_t)rows * ld), mx((size_t)(r1 - r0));
        std::vector<long long> arg((size_t)(r1 - r0));
        if (fread(mat.data(), 8, mat.size(), stdin) != mat.size()) return 3;
        emu_launch_grid((int)(r1 - r0), EmuDim{256, 1, 1}, [&] { rowmax_block_kernel(mat.data(), ld, r0, c0, c1, mx.data(), arg.data()); });
        fwrite(mx.data(), 8, mx.size(), stdout);
        fwrite(arg.data(), 8, arg.size(), stdout);
        return 0;
    }
    if (what == "gather" && argc == 5) {
        const long long rows = atoll(argv[2]), ld = atoll(argv[3]), count = atoll(argv[4]);
        std::vector<double> mat((size_t)rows * ld), out((size_t)count);
        std::vector<long long> ii((size_t)count), jj((size_t)count);
        if (fread(mat.data(), 8, mat.size(), stdin) != mat.size()) return 3;
        if (fread(ii.data(), 8, ii.size(), stdin) != ii.size() || fread(jj.data(), 8, jj.size(), stdin) != jj.size()) return 3;
        emu_launch_grid((int)((count + 255) / 256), EmuDim{256, 1, 1},
                        [&] { gather_cells_kernel(mat.data(), ld, ii.data(), jj.data(), count, out.data()); });
        fwrite(out.data(), 8, out.size(), stdout);
        return 0;
    }
    if (what == "philox" && argc == 6) {
        const unsigned long long seed = strtoull(argv[2], nullptr, 0);
        const long long b0 = atoll(argv[3]), nb = atoll(argv[4]), p = atoll(argv[5]);
        std::vector<long long> idx((size_t)nb * p, -1);
        std::vector<int> w((size_t)nb * p, 0);
        emu_launch_grid(EmuDim{(int)(((p + 3) / 4 + 255) / 256), (int)nb, 1}, EmuDim{256, 1, 1},
                        [&] { philox_indices_kernel(seed, b0, p, idx.data(), w.data()); });
        fwrite(idx.data(), 8, idx.size(), stdout);
        fwrite(w.data(), 4, w.size(), stdout);
        return 0;
    }
    return 2;
}
