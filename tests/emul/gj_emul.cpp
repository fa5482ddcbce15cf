// Runs the generalised-Jaccard kernels of gravity2_b200/csrc/gj_kernels.cuh -- the SAME source the GPU build compiles -- on
// host threads (cuda_block_emul.h), with the launch shapes of gj_kernels.cu / api.cu (gv2_similarity_host, gj_fused_launch,
// gom_quantize_launch).
//   gj_emul similarity <n> <p> <g> <scheme> <power> <out_kind> <thr> <apply_thr> <ksplit>
//        < sig (n*p f64) [if p > 0], gom (n*g f64) [if g > 0], aux (n*n f64) [schemes R, RG, PR, L, PL]   > out (n*n f64)
//   gj_emul fused <n> <p> <g> <fixed> <power> <out_kind>      (p == 0: scheme G, else PG)
//        < sig (n*p f64) [if p > 0], gom (n*g f64)                                                       > out (n*n f64)
//   gj_emul job_tail <n> <g> <nb> <p_scale> <power> <out_kind>   (the fused tail as the bootstrap job runs it: api.cu replicates_impl)
//        < PPHMM pair sums (nb * tiles * 128*128 f64, gj_ws_index layout), their row sums (nb*n f64), GOM tables (nb*n*g f64)
//        > out (nb * n*n f64)
//   gj_emul nbhd <n> <p> <weight> <thr> <with_mul>            (gv2_similarity_nbhd_host)
//        < sig (n*p f64), classes (n*p u8), mul (n*n f64) [if with_mul]                                    > out (n*n f64)
#define GJ_EMULATED
#include "cuda_block_emul.h"
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>

#include "../../include/gravity2_b200.h"
#include "../../gravity2_b200/csrc/gj_kernels.cuh"
#include "../../gravity2_b200/csrc/nbhd_kernels.cuh"

static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
static bool read_all(std::vector<double>& v) { return v.empty() || fread(v.data(), 8, v.size(), stdin) == v.size(); }

struct Comp {                       // one component as gv2_similarity_host prepares it
    std::vector<float> tab;
    std::vector<double> rowsum, ws;
    std::vector<unsigned char> nan;
    long long k_pad = 0;
    GjComp comp{};
};

static void prepare(const std::vector<double>& src, long long n, long long k, double thr, int apply_thr, long long n_pad, Comp& c) {
    c.k_pad = round_up(std::max<long long>(k, 1), GV2_BK);
    c.tab.assign((size_t)n_pad * c.k_pad, -1.f);
    c.rowsum.assign((size_t)n, -1.0);
    c.nan.assign((size_t)n, 7);
    emu_launch_grid((int)n_pad, EmuDim{256, 1, 1}, [&] {
        prepare_table_kernel(src.data(), n, k, k, thr, apply_thr, c.tab.data(), c.k_pad, c.rowsum.data(), c.nan.data());
    });
}

static void min_sums(Comp& c, long long n_pad, long long ntl, int ksplit) {
    c.ws.assign((size_t)ksplit * ntl * GV2_TILE * GV2_TILE, -3.0);
    const int nt = (int)(n_pad / GV2_TILE), kblocks = (int)(c.k_pad / GV2_BK);
    emu_launch_grid(EmuDim{(int)ntl, ksplit, 1}, EmuDim{GJ_THREADS, 1, 1}, [&] {
        gj_min_tile_kernel<false>(c.tab.data(), 0, (int)c.k_pad, kblocks, nt, 0, nullptr, 0, c.ws.data());
    });
    c.comp.ws = c.ws.data(); c.comp.rowsum = c.rowsum.data(); c.comp.nan = c.nan.data(); c.comp.ksplit = ksplit;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string what = argv[1];
    if (what == "similarity" && argc == 11) {
        const long long n = atoll(argv[2]), p = atoll(argv[3]), g = atoll(argv[4]);
        const int scheme = atoi(argv[5]);
        const double power = atof(argv[6]);
        const int out_kind = atoi(argv[7]);
        const double thr = atof(argv[8]);
        const int apply_thr = atoi(argv[9]), ksplit = atoi(argv[10]);
        const bool need_p = scheme == GV2_SCHEME_P || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_PR || scheme == GV2_SCHEME_PL;
        const bool need_g = scheme == GV2_SCHEME_G || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_RG;
        const bool need_r = scheme == GV2_SCHEME_R || scheme == GV2_SCHEME_RG || scheme == GV2_SCHEME_PR || scheme == GV2_SCHEME_L ||
                            scheme == GV2_SCHEME_PL;
        std::vector<double> sig((size_t)n * p), gom((size_t)n * g), aux(need_r ? (size_t)n * n : 0), out((size_t)n * n, -5.0);
        if (!read_all(sig) || !read_all(gom) || !read_all(aux)) return 3;
        const long long n_pad = round_up(n, GV2_TILE), nt = n_pad / GV2_TILE, ntl = nt * (nt + 1) / 2;
        Comp P, G;
        if (need_p) { prepare(sig, n, p, thr, apply_thr, n_pad, P); min_sums(P, n_pad, ntl, ksplit); }
        if (need_g) { prepare(gom, n, g, 0.0, 0, n_pad, G); min_sums(G, n_pad, ntl, 1); }
        emu_launch_grid(EmuDim{(int)(ntl * 16), 1, 1}, EmuDim{32, 8, 1}, [&] {
            gj_epilogue_kernel(n, (int)((n + GV2_TILE - 1) / GV2_TILE), 0, ntl, P.comp, G.comp, need_r ? aux.data() : nullptr, scheme, power,
                               out_kind, 0, out.data(), 0);
        });
        fwrite(out.data(), 8, out.size(), stdout);
        return 0;
    }
    if (what == "fused" && argc == 8) {
        const long long n = atoll(argv[2]), p = atoll(argv[3]), g = atoll(argv[4]);
        const int fixed = atoi(argv[5]);
        const double power = atof(argv[6]);
        const int out_kind = atoi(argv[7]);
        std::vector<double> sig((size_t)n * p), gom((size_t)n * g), out((size_t)n * n, -5.0);
        if (!read_all(sig) || !read_all(gom)) return 3;
        const long long n_pad = round_up(n, GV2_TILE), nt = n_pad / GV2_TILE, ntl = nt * (nt + 1) / 2;
        Comp P, G;
        double p_scale = 1.0;
        if (p > 0) {
            prepare(sig, n, p, 0.0, 1, n_pad, P);
            min_sums(P, n_pad, ntl, 1);
            if (fixed) {                              // a power of two that brings every PPHMM sum to <= 1 (api.cu: job->p_scale)
                double mx = 1.0;
                for (double v : P.rowsum) mx = std::max(mx, v);
                int e = 0;
                while (ldexp(1.0, e) < mx) ++e;
                p_scale = ldexp(1.0, -e);
            }
        }
        const long long g_pad = round_up(std::max<long long>(g, 1), GV2_BK);
        std::vector<unsigned> qtab;
        if (fixed) {                                  // gom_quantize_launch
            unsigned long long maxbits = 0;
            qtab.assign((size_t)n_pad * g_pad, 0xffffffffu);
            G.rowsum.assign((size_t)n, -1.0);
            G.nan.assign((size_t)n, 7);
            emu_launch_grid(EmuDim{(int)n, 1, 1}, EmuDim{128, 1, 1}, [&] { gom_rowmax_kernel(gom.data(), n, g, &maxbits); });
            emu_launch_grid(EmuDim{(int)n_pad, 1, 1}, EmuDim{128, 1, 1}, [&] {
                gom_quantize_kernel(gom.data(), n, g, &maxbits, qtab.data(), n_pad, g_pad, G.rowsum.data(), G.nan.data());
            });
        } else {
            prepare(gom, n, g, 0.0, 0, n_pad, G);
        }
        GjFuse fz{};
        if (p > 0) fz.P = P.comp;
        fz.g_rowsum = G.rowsum.data(); fz.g_nan = G.nan.data();
        fz.n = n; fz.scheme = p > 0 ? GV2_SCHEME_PG : GV2_SCHEME_G; fz.out_kind = out_kind; fz.pw = power; fz.out = out.data();
        fz.p_scale = p_scale;
        const float* tab = fixed ? reinterpret_cast<const float*>(qtab.data()) : G.tab.data();
        const int k_groups = (int)((g + 3) / 4), nkb = (int)((g + GV2_BK - 1) / GV2_BK);
        const int variant = (fixed ? 4 : 0) + (p > 0 ? 2 : 0) + (power != 1.0 ? 1 : 0);
        emu_launch_grid(EmuDim{(int)(2 * ntl), 1, 1}, EmuDim{GJ_THREADS, 1, 1}, [&] {
            switch (variant) {
                case 0: gj_fused_kernel<false, false, false>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
                case 1: gj_fused_kernel<false, false, true>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
                case 2: gj_fused_kernel<false, true, false>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
                case 3: gj_fused_kernel<false, true, true>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
                case 4: gj_fused_kernel<true, false, false>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
                case 5: gj_fused_kernel<true, false, true>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
                case 6: gj_fused_kernel<true, true, false>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
                default: gj_fused_kernel<true, true, true>(tab, 0, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u); break;
            }
        });
        fwrite(out.data(), 8, out.size(), stdout);
        return 0;
    }
    if (what == "job_tail" && argc == 8) {
        const long long n = atoll(argv[2]), g = atoll(argv[3]), nb = atoll(argv[4]);
        const double p_scale = atof(argv[5]), power = atof(argv[6]);
        const int out_kind = atoi(argv[7]);
        const long long n_pad = round_up(n, GV2_TILE), nt = n_pad / GV2_TILE, ntl = nt * (nt + 1) / 2;
        const long long g_pad = round_up(std::max<long long>(g, 1), GV2_BK);
        std::vector<double> ws((size_t)nb * ntl * GV2_TILE * GV2_TILE), prs((size_t)nb * n), gom((size_t)nb * n * g), out((size_t)nb * n * n, -5.0);
        if (!read_all(ws) || !read_all(prs) || !read_all(gom)) return 3;
        // job_prepare_gom -> gom_quantize_launch: one scale for the whole launch
        unsigned long long maxbits = 0;
        std::vector<unsigned> qtab((size_t)nb * n_pad * g_pad, 0xffffffffu);
        std::vector<double> grs((size_t)nb * n, -1.0);
        std::vector<unsigned char> gnan((size_t)nb * n, 7);
        emu_launch_grid(EmuDim{(int)n, (int)nb, 1}, EmuDim{128, 1, 1}, [&] { gom_rowmax_kernel(gom.data(), n, g, &maxbits); });
        emu_launch_grid(EmuDim{(int)n_pad, (int)nb, 1}, EmuDim{128, 1, 1}, [&] {
            gom_quantize_kernel(gom.data(), n, g, &maxbits, qtab.data(), n_pad, g_pad, grs.data(), gnan.data());
        });
        GjFuse fz{};
        fz.P.ws = ws.data(); fz.P.rowsum = prs.data(); fz.P.nan = nullptr; fz.P.ksplit = 1;
        fz.P.b_stride_ws = ntl * GV2_TILE * GV2_TILE; fz.P.b_stride_row = n;
        fz.g_rowsum = grs.data(); fz.g_nan = gnan.data(); fz.g_stride_row = n; fz.g_stride_nan = n;
        fz.n = n; fz.scheme = GV2_SCHEME_PG; fz.out_kind = out_kind; fz.pw = power; fz.out = out.data(); fz.out_b_stride = n * n;
        fz.p_scale = p_scale;
        const int k_groups = (int)((g + 3) / 4), nkb = (int)((g + GV2_BK - 1) / GV2_BK);
        const float* tab = reinterpret_cast<const float*>(qtab.data());
        emu_launch_grid(EmuDim{(int)(2 * ntl), 1, (int)nb}, EmuDim{GJ_THREADS, 1, 1}, [&] {
            if (power != 1.0) gj_fused_kernel<true, true, true>(tab, n_pad * g_pad, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u);
            else gj_fused_kernel<true, true, false>(tab, n_pad * g_pad, (int)g_pad, nkb, k_groups, (int)nt, fz, 1u);
        });
        fwrite(out.data(), 8, out.size(), stdout);
        return 0;
    }
    if (what == "nbhd" && argc == 7) {
        const long long n = atoll(argv[2]), p = atoll(argv[3]);
        const double w = atof(argv[4]), thr = atof(argv[5]);
        const int with_mul = atoi(argv[6]);
        std::vector<double> sig((size_t)n * p), mul(with_mul ? (size_t)n * n : 0), out((size_t)n * n, -5.0);
        std::vector<uint8_t> cls((size_t)n * p);
        if (!read_all(sig)) return 3;
        if (!cls.empty() && fread(cls.data(), 1, cls.size(), stdin) != cls.size()) return 3;
        if (!read_all(mul)) return 3;
        const int nt = (int)((n + NB_T - 1) / NB_T);
        emu_launch_grid(EmuDim{nt, nt, 1}, EmuDim{NB_T * NB_T, 1, 1}, [&] {
            gj_nbhd_kernel(sig.data(), cls.data(), n, p, p, w, thr, with_mul ? mul.data() : nullptr, out.data());
        });
        fwrite(out.data(), 8, out.size(), stdout);
        return 0;
    }
    return 2;
}
