// Runs the GOM scoring kernels of gravity2_b200/csrc/gom_score_kernels.cuh -- the SAME source the GPU build compiles -- on
// host threads, following gom_db_build / gom_db_score of gom_kernels.cu step by step as gv2_gom_table_host drives them (the
// device-wide prefix sums between the passes are plain loops here; on the GPU they are cub::DeviceScan).
//   gom_emul <n> <p> <G> <nb> <flags> <has_weights>
//        < loc (n*p f64), member offsets (G+1 int32), member rows (M*p f64), multiplicities (nb*p int32) [if has_weights]
//        > GOM signature tables (nb * n * G f64)
#define GOM_EMULATED
#include "cuda_block_emul.h"
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "../../include/gravity2_b200.h"
#include "../../gravity2_b200/csrc/gom_score_kernels.cuh"

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const long long n = atoll(argv[1]), p = atoll(argv[2]), G = atoll(argv[3]), nb = atoll(argv[4]);
    const int flags = atoi(argv[5]), has_w = atoi(argv[6]);
    std::vector<double> loc((size_t)n * p);
    std::vector<int> mem_ptr((size_t)G + 1);
    if (fread(loc.data(), 8, loc.size(), stdin) != loc.size() || fread(mem_ptr.data(), 4, mem_ptr.size(), stdin) != mem_ptr.size()) return 3;
    const long long M = mem_ptr[G];
    std::vector<double> rowtab((size_t)M * p);
    std::vector<int> weights(has_w ? (size_t)nb * p : 0);
    if (!rowtab.empty() && fread(rowtab.data(), 8, rowtab.size(), stdin) != rowtab.size()) return 3;
    if (!weights.empty() && fread(weights.data(), 4, weights.size(), stdin) != weights.size()) return 3;
    std::vector<long long> mem_row((size_t)M);
    for (long long r = 0; r < M; ++r) mem_row[r] = r;
    std::vector<double> out((size_t)nb * n * G, -7.0);
    if (n == 0 || G == 0) { fwrite(out.data(), 8, out.size(), stdout); return 0; }
    // ------------------------------------------------------------------ gom_db_build
    std::vector<int> cnt((size_t)n);
    std::vector<long long> g_ptr((size_t)n + 1, 0);
    emu_launch_grid((int)n, EmuDim{256, 1, 1}, [&] { csr_count_kernel(loc.data(), n, p, p, cnt.data()); });
    for (long long v = 0; v < n; ++v) g_ptr[v + 1] = g_ptr[v] + cnt[v];
    const long long nnz = g_ptr[n];
    std::vector<int> g_col((size_t)std::max<long long>(1, nnz));
    std::vector<double> g_val(g_col.size());
    emu_launch_grid((int)((n + 7) / 8), EmuDim{256, 1, 1}, [&] { csr_fill_kernel(loc.data(), n, p, p, g_ptr.data(), g_col.data(), g_val.data()); });
    std::vector<int> flag((size_t)(G * p)), scan((size_t)(G * p) + 1, 0), pos((size_t)(G * p));
    std::vector<int> pt_ptr((size_t)G + 1, 0);
    emu_launch_grid(EmuDim{(int)((p + 255) / 256), (int)G, 1}, EmuDim{256, 1, 1}, [&] {
        gom_colany_kernel(rowtab.data(), p, p, mem_row.data(), mem_ptr.data(), (flags & GV2_GOM_ALL_COLUMNS) ? 1 : 0, flag.data());
    });
    for (long long i = 0; i < G * p; ++i) scan[i + 1] = scan[i] + flag[i];
    for (long long g = 0; g <= G; ++g) pt_ptr[g] = scan[g * p];
    const int npt = pt_ptr[G];
    std::vector<long long> x_off((size_t)G + 1, 0), a_off((size_t)G + 1, 0);
    int max_ng = 0, max_mg = 0;
    for (long long g = 0; g < G; ++g) {
        const long long ng = pt_ptr[g + 1] - pt_ptr[g], mg = mem_ptr[g + 1] - mem_ptr[g];
        x_off[g + 1] = x_off[g] + ng * mg;
        a_off[g + 1] = a_off[g] + ng * ng;
        max_ng = std::max<int>(max_ng, (int)ng);
        max_mg = std::max<int>(max_mg, (int)mg);
    }
    std::vector<int> pt_col((size_t)std::max(1, npt));
    std::vector<double> nrm(pt_col.size()), nrm2(pt_col.size()), X((size_t)std::max<long long>(1, x_off[G])), A((size_t)std::max<long long>(1, a_off[G]));
    emu_launch_grid(EmuDim{(int)((p + 255) / 256), (int)G, 1}, EmuDim{256, 1, 1}, [&] {
        gom_fill_kernel(rowtab.data(), p, p, mem_row.data(), mem_ptr.data(), flag.data(), scan.data(), x_off.data(), pos.data(), pt_col.data(),
                        X.data(), nrm.data(), nrm2.data());
    });
    if (max_ng)
        emu_launch_grid(EmuDim{(max_ng + 15) / 16, (max_ng + 15) / 16, (int)G}, EmuDim{16, 16, 1},
                        [&] { gom_dist_cache_kernel(pt_ptr.data(), mem_ptr.data(), x_off.data(), a_off.data(), X.data(), A.data()); });
    std::vector<double> wT((size_t)p * NBC), arT((size_t)2 * std::max(1, npt) * NBC), rT((size_t)std::max<long long>(1, nnz) * NBC),
        gagg((size_t)G * GAGG * NBC), vagg((size_t)n * VAGG * NBC), wG((size_t)2 * std::max(1, npt) * NBC);
    const int n_ctile = (max_mg + 7) / 8;
    std::vector<double> cpart((size_t)G * std::max(1, n_ctile) * NBC);
    int max_t = 0;
    for (long long v = 0; v < n; ++v) max_t = std::max<int>(max_t, (int)(g_ptr[v + 1] - g_ptr[v]));
    std::vector<int> g_ord((size_t)std::max<long long>(1, nnz));
    if (max_t) emu_launch_grid((int)n, EmuDim{128, 1, 1}, [&] { genome_order_kernel(g_ptr.data(), g_val.data(), g_ord.data()); });
    std::vector<unsigned char> wT8((size_t)2 * p * NBC);
    int dflag[2] = {0, 0};
    const long long total_w = n * G;
    std::vector<long long> pbytes((size_t)(2 * total_w + 1), 0), pr_ptr((size_t)(2 * total_w + 1), 0);
    std::vector<int> pr_ni((size_t)total_w + 1, 0);
    emu_launch_grid((int)((total_w + 7) / 8), EmuDim{256, 1, 1}, [&] {
        gom_pair_count_kernel(n, (int)G, p, g_ptr.data(), g_col.data(), pos.data(), VG_LIGHT_CAP, pr_ni.data(), pbytes.data());
    });
    for (long long i = 0; i < 2 * total_w; ++i) pr_ptr[i + 1] = pr_ptr[i] + pbytes[i];
    const long long rec_bytes = pr_ptr[2 * total_w];
    std::vector<unsigned char> rec_store((size_t)std::max<long long>(16, rec_bytes) + 64);
    unsigned char* rec = rec_store.data();
    while (reinterpret_cast<uintptr_t>(rec) & 31) ++rec;                       // records are read with 16-byte accesses
    std::vector<long long> heavy_list((size_t)std::max<long long>(1, total_w));
    emu_launch_grid((int)((total_w + 7) / 8), EmuDim{256, 1, 1}, [&] {
        gom_pair_fill_kernel(n, (int)G, p, g_ptr.data(), g_col.data(), g_val.data(), pos.data(), pt_ptr.data(), nrm.data(), a_off.data(), A.data(),
                             pr_ni.data(), pr_ptr.data(), VG_LIGHT_CAP, rec);
    });
    emu_launch_grid((int)((total_w + 255) / 256), EmuDim{256, 1, 1},
                    [&] { gom_heavy_list_kernel(total_w, pr_ni.data(), VG_LIGHT_CAP, heavy_list.data(), &dflag[1]); });
    const int n_heavy = dflag[1];
    std::sort(heavy_list.begin(), heavy_list.begin() + n_heavy);               // (any order is valid: every pair is scored independently)
    const size_t vg_smem = (size_t)((max_t + 3) & ~3) * (VG_ENTRY_BYTES + 8 * sizeof(double)) + 128 + 3 * VG_NSUM * 32 * sizeof(double);
    if (vg_smem > sizeof emu_dynamic_smem || ga_smem_bytes(2) > sizeof emu_dynamic_smem) return 4;
    // ------------------------------------------------------------------ gom_db_score
    for (long long b0 = 0; b0 < nb; b0 += 2 * NBC) {
        const int nh = nb - b0 > NBC ? 2 : 1;
        for (int h = 0; h < nh; ++h) {
            const int cb = (int)std::min<long long>(NBC, nb - b0 - h * NBC);
            emu_launch_grid((int)((p * NBC + 255) / 256), EmuDim{256, 1, 1}, [&] {
                weights_transpose_kernel(has_w ? weights.data() + (b0 + h * NBC) * p : nullptr, p, cb, wT.data(), wT8.data() + (size_t)h * p * NBC, &dflag[0]);
            });
            if (max_ng)
                emu_launch_grid((int)(((long long)npt * NBC + 255) / 256), EmuDim{256, 1, 1},
                                [&] { gom_gather_weights_kernel(pt_col.data(), npt, wT.data(), wG.data() + (size_t)h * npt * NBC); });
        }
        double* wG1 = wG.data() + (size_t)npt * NBC;
        double* arT1 = arT.data() + (size_t)npt * NBC;
        if (max_ng)
            emu_launch_grid(EmuDim{(max_ng + GA_ROWS - 1) / GA_ROWS, (int)G, 1}, EmuDim{32, 8, 1}, [&] {
                if (nh == 2) gom_rowsums_kernel<2>(pt_ptr.data(), a_off.data(), A.data(), wG.data(), wG1, arT.data(), arT1);
                else gom_rowsums_kernel<1>(pt_ptr.data(), a_off.data(), A.data(), wG.data(), nullptr, arT.data(), nullptr);
            });
        for (int h = 0; h < nh; ++h) {
            const int cb = (int)std::min<long long>(NBC, nb - b0 - h * NBC);
            const double* wGh = wG.data() + (size_t)h * npt * NBC;
            const double* arTh = arT.data() + (size_t)h * npt * NBC;
            const unsigned char* wT8h = wT8.data() + (size_t)h * p * NBC;
            double* outh = out.data() + (b0 + h * NBC) * n * G;
            if (n_ctile)
                emu_launch_grid(EmuDim{(int)G, n_ctile, 1}, EmuDim{256, 1, 1},
                                [&] { gom_centroid_kernel(pt_ptr.data(), mem_ptr.data(), x_off.data(), X.data(), wGh, n_ctile, cpart.data()); });
            emu_launch_grid((int)G, EmuDim{256, 1, 1},
                            [&] { gom_aggregates_kernel(pt_ptr.data(), nrm.data(), nrm2.data(), wGh, arTh, cpart.data(), n_ctile, gagg.data()); });
            emu_launch_grid((int)((n + 3) / 4), EmuDim{128, 1, 1}, [&] {
                genome_rowsums_sorted_kernel(n, g_ptr.data(), g_col.data(), g_val.data(), g_ord.data(), wT8h, rT.data(), vagg.data());
            });
            emu_launch_grid((int)n, EmuDim{128, 1, 1}, [&] {
                gom_score_light_kernel(n, (int)G, g_ptr.data(), pt_ptr.data(), pr_ni.data(), pr_ptr.data(), rec, wT8h, arTh, rT.data(), gagg.data(),
                                       vagg.data(), cb, outh);
            });
            if (n_heavy)
                emu_launch_grid(std::min(n_heavy, 16), EmuDim{128, 1, 1}, [&] {
                    gom_score_heavy_kernel(n, (int)G, max_t, g_ptr.data(), pt_ptr.data(), pr_ni.data(), pr_ptr.data(), rec, wT8h, arTh, rT.data(),
                                           gagg.data(), vagg.data(), cb, outh, heavy_list.data(), n_heavy);
                });
        }
    }
    if (dflag[0]) return 5;
    fwrite(out.data(), 8, out.size(), stdout);
    return 0;
}
