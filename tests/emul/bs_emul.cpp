// Runs the default (pair-list) path of the sparse replicate pair sums -- the SAME kernel source the GPU build compiles
// (gravity2_b200/csrc/boot_sparse_kernels.cuh) -- on host threads, following boot_sparse_build / boot_sparse_build_lists /
// boot_sparse_launch / boot_sparse_rowsums of boot_sparse.cu step by step (the device-wide prefix sums between the passes are
// plain loops here; on the GPU they are cub::DeviceScan).
//   bs_emul <n> <k> <nb>  < table (n*k float64), multiplicities (nb*k int32)
//        > quantum (1 float64), row sums (n float64), workspace (nb * tiles * 128*128 float64; NaN where nothing is written)
// nb <= 12 runs boot_sparse_pairlane_kernel, more replicates boot_sparse_stream_kernel (as boot_sparse_launch chooses).
#define BS_EMULATED
#include "cuda_block_emul.h"
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "../../gravity2_b200/csrc/boot_sparse_kernels.cuh"

static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
static int bss_rw(long long nb) { return nb <= 128 ? 1 : nb <= 256 ? 2 : nb <= 512 ? 4 : 8; }

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const long long n = atoll(argv[1]), k = atoll(argv[2]), nb = atoll(argv[3]);
    std::vector<double> tab((size_t)n * k);
    std::vector<int> w((size_t)nb * k);
    if (fread(tab.data(), 8, tab.size(), stdin) != tab.size() || fread(w.data(), 4, w.size(), stdin) != w.size()) return 3;
    const long long n_pad = round_up(n, GV2_TILE), k_pad = round_up(k, 128), nt = n_pad / GV2_TILE, ntl = nt * (nt + 1) / 2;
    double max_value = 0.0;
    for (double v : tab) if (v > max_value) max_value = v;
    // ---- boot_sparse_build
    const double scale = max_value > 0.0 ? 281474976710656.0 * (1.0 - 9.313225746154785e-10) / max_value : 0.0;
    const double quantum = scale > 0.0 ? 1.0 / scale : 0.0;
    std::vector<long long> cnt((size_t)n_pad + 1, 0), ptr((size_t)n_pad + 1, 0);
    emu_launch_grid((int)n_pad, EmuDim{256, 1, 1}, [&] { bs_count_kernel(tab.data(), n, k, k, scale, cnt.data()); });
    for (long long r = 0; r < n_pad; ++r) ptr[r + 1] = ptr[r] + cnt[r];
    const long long nnz = ptr[n_pad];
    std::vector<int> col((size_t)std::max<long long>(1, nnz));
    std::vector<uint32_t> hi(col.size());
    std::vector<uint16_t> lo(col.size());
    emu_launch_grid((int)((n + 7) / 8), EmuDim{256, 1, 1},
                    [&] { bs_fill_kernel(tab.data(), n, k, k, scale, ptr.data(), col.data(), hi.data(), lo.data()); });
    BsCsr csr{ptr.data(), col.data(), hi.data(), lo.data()};
    // ---- boot_sparse_build_lists
    const long long nrows = ntl * 64 * 16;
    std::vector<long long> rowtot((size_t)nrows + 1, 0), rowptr((size_t)nrows + 1, 0);
    std::vector<uint32_t> pend((size_t)nrows * 16, 0xdeadbeefu);
    if (BS_LIST_SMEM(k_pad) > sizeof emu_dynamic_smem || k_pad > 65536) return 4;
    emu_launch_grid(EmuDim{64, (int)ntl, 1}, EmuDim{BS_THREADS, 1, 1},
                    [&] { bs_lists_build_kernel<false>(csr, (int)nt, (int)k_pad, rowtot.data(), nullptr, pend.data(), nullptr); });
    for (long long r = 0; r < nrows; ++r) rowptr[r + 1] = rowptr[r] + rowtot[r];
    const long long nent = rowptr[nrows];
    std::vector<uint8_t> ent((size_t)std::max<long long>(4, nent) * 8, 0);
    emu_launch_grid(EmuDim{64, (int)ntl, 1}, EmuDim{BS_THREADS, 1, 1},
                    [&] { bs_lists_build_kernel<true>(csr, (int)nt, (int)k_pad, nullptr, rowptr.data(), pend.data(), ent.data()); });
    // ---- boot_sparse_launch
    const bool pairlane = nb <= 12;
    const int rwp = bss_rw(nb);
    const long long w_ld = pairlane ? (nb <= 4 ? 4 : nb <= 8 ? 8 : 16) : round_up(nb, 128 * rwp);
    std::vector<uint8_t> wT((size_t)k_pad * w_ld, 0xee);
    int overflow = 0;
    emu_launch_grid(EmuDim{(int)((k_pad + 31) / 32), (int)((w_ld + 31) / 32), 1}, EmuDim{32, 8, 1},
                    [&] { bs_weights_transpose_kernel(w.data(), nb, k, k_pad, w_ld, wT.data(), &overflow); });
    emu_launch_grid((int)nb, EmuDim{256, 1, 1}, [&] { bs_wsum_check_kernel(w.data(), k, &overflow); });
    if (overflow) return 5;
    const long long rep_stride = ntl * GV2_TILE * GV2_TILE;
    std::vector<double> ws((size_t)nb * rep_stride, nan(""));
    if (pairlane) {
        emu_launch_grid(EmuDim{64, (int)ntl, 1}, EmuDim{BS_THREADS, 1, 1}, [&] {
            if (nb <= 4) boot_sparse_pairlane_kernel<1>(rowptr.data(), pend.data(), ent.data(), (int)nt, 0, wT.data(), w_ld, (int)nb, rep_stride, quantum, ws.data());
            else if (nb <= 8) boot_sparse_pairlane_kernel<2>(rowptr.data(), pend.data(), ent.data(), (int)nt, 0, wT.data(), w_ld, (int)nb, rep_stride, quantum, ws.data());
            else boot_sparse_pairlane_kernel<3>(rowptr.data(), pend.data(), ent.data(), (int)nt, 0, wT.data(), w_ld, (int)nb, rep_stride, quantum, ws.data());
        });
    } else {
        if ((size_t)8 * BSS_WARP_BYTES > sizeof emu_dynamic_smem) return 4;
        emu_launch_grid(EmuDim{64, (int)ntl, (int)std::max<long long>(1, w_ld / (128 * rwp))}, EmuDim{BS_THREADS, 1, 1}, [&] {
            if (rwp == 1) boot_sparse_stream_kernel<1>(rowptr.data(), pend.data(), ent.data(), (int)nt, 0, wT.data(), w_ld, (int)nb, rep_stride, quantum, ws.data());
            else if (rwp == 2) boot_sparse_stream_kernel<2>(rowptr.data(), pend.data(), ent.data(), (int)nt, 0, wT.data(), w_ld, (int)nb, rep_stride, quantum, ws.data());
            else if (rwp == 4) boot_sparse_stream_kernel<4>(rowptr.data(), pend.data(), ent.data(), (int)nt, 0, wT.data(), w_ld, (int)nb, rep_stride, quantum, ws.data());
            else boot_sparse_stream_kernel<8>(rowptr.data(), pend.data(), ent.data(), (int)nt, 0, wT.data(), w_ld, (int)nb, rep_stride, quantum, ws.data());
        });
    }
    // ---- boot_sparse_rowsums
    std::vector<double> rowsum((size_t)n, -1.0);
    emu_launch_grid((int)((n + 7) / 8), EmuDim{256, 1, 1}, [&] { bs_rowsums_kernel(csr, n, quantum, rowsum.data()); });
    fwrite(&quantum, 8, 1, stdout);
    fwrite(rowsum.data(), 8, rowsum.size(), stdout);
    fwrite(ws.data(), 8, ws.size(), stdout);
    return 0;
}
