// Runs the presence / absence kernels of gravity2_b200/csrc/popc_kernels.cuh -- the SAME source the GPU build compiles -- on
// host threads (cuda_block_emul.h), with the launch shapes of popc_kernels.cu (pc_launch, gv2_dev_pack_presence,
// gv2_dev_unpack_tiles_i32).
//   popc_emul <n> <p> <mode: 0 direct | 1 packed tiles + unpack>  < table (n*p float64)  > counts (n*n int32), nnz (n int32)
#define PC_EMULATED
#include "cuda_block_emul.h"
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "../../gravity2_b200/csrc/popc_kernels.cuh"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const long long n = atoll(argv[1]), p = atoll(argv[2]);
    const int packed = atoi(argv[3]);
    std::vector<double> tab((size_t)n * p);
    if (fread(tab.data(), sizeof(double), tab.size(), stdin) != tab.size()) return 3;
    const long long n_pad = (n + GV2_TILE - 1) / GV2_TILE * GV2_TILE;
    const long long words = ((p + 63) / 64 + PC_BK / 2 - 1) / (PC_BK / 2) * (PC_BK / 2);   // 64-bit words per row; 2*words % PC_BK == 0
    std::vector<unsigned> bits((size_t)n_pad * 2 * words, 0xdeadbeefu);
    std::vector<int> nnz((size_t)n_pad, -1), inter((size_t)n * n, -7);
    emu_launch_grid((int)((n_pad + 7) / 8), EmuDim{256, 1, 1},
                    [&] { pack_presence_kernel(tab.data(), n, p, p, bits.data(), 2 * words, nnz.data()); });
    const size_t smem = std::max((size_t)2 * PC_STAGES * GV2_TILE * PC_LDS * sizeof(unsigned), (size_t)GV2_TILE * 129 * sizeof(int));
    if (smem > sizeof emu_dynamic_smem) return 4;
    const int nt = (int)(n_pad / GV2_TILE), ld = (int)(2 * words), kb = (int)(2 * words / PC_BK);
    const long long ntl = (long long)nt * (nt + 1) / 2;
    if (!packed) {
        emu_launch_grid((int)ntl, EmuDim{PC_THREADS, 1, 1},
                        [&] { shared_count_tile_kernel<false>(bits.data(), ld, kb, nt, 0, n, inter.data()); });
    } else {
        std::vector<int> tiles((size_t)ntl * GV2_TILE * GV2_TILE, -9);
        // two launches over two tile ranges, as a rank of the multi-GPU path would see them
        const long long half = ntl / 2;
        if (half > 0)
            emu_launch_grid((int)half, EmuDim{PC_THREADS, 1, 1},
                            [&] { shared_count_tile_kernel<true>(bits.data(), ld, kb, nt, 0, n, tiles.data()); });
        emu_launch_grid((int)(ntl - half), EmuDim{PC_THREADS, 1, 1}, [&] {
            shared_count_tile_kernel<true>(bits.data(), ld, kb, nt, half, n, tiles.data() + half * GV2_TILE * GV2_TILE);
        });
        emu_launch_grid((int)(ntl * 16), EmuDim{32, 8, 1}, [&] {
            unpack_tiles_i32_kernel(n, (int)((n + GV2_TILE - 1) / GV2_TILE), 0, tiles.data(), inter.data());
        });
    }
    fwrite(inter.data(), sizeof(int), inter.size(), stdout);
    fwrite(nnz.data(), sizeof(int), (size_t)n, stdout);
    return 0;
}
