// The generalised-Jaccard kernels: table preparation, K1 pair-sum tiles, the fused tail, K3 epilogue, tile unpacking, the
// fixed-point image of the GOM tables.  gj_kernels.cu includes this file; so does tests/emul/gj_emul.cpp, which runs the
// SAME kernel source on host threads (cuda_block_emul.h) so that the kernels' logic -- tile coordinates, ragged edges, the
// stage rings, slab promotion, the epilogue's cleaning rules, the mirror through shared memory -- is checked against the
// golden outputs of the unmodified reference where there is no GPU.
#pragma once
#include <math.h>

#include <type_traits>

#ifndef GV2_TILE
#define GV2_TILE 128
#define GV2_BK 32
#endif

#define GJ_PRAGMA_(x) _Pragma(#x)
#define GJ_UNROLL(n) GJ_PRAGMA_(unroll n)
#define GJ_THREADS 256
#define GJ_STAGES 3
#define GJ_LDS (GV2_BK + 4)          // padded row stride (floats): conflict-free LDS.128
#ifndef GJ_SLAB
#define GJ_SLAB 128                  // columns accumulated in fp32 before fp64 promotion
#endif
#ifndef GJ_KG_UNROLL
#define GJ_KG_UNROLL 2                // 4-column groups unrolled per loop trip (whole stage = 8 overflows the I-cache)
#endif
// Workspace tile layout: element (r, c) of a 128x128 pair tile lives at gj_ws_index(r, c), i.e. as
// [r/16][c/16][r%16][c%16], so that the 256 threads of the tile kernel (thread (ty, tx) owns rows
// ty+16m, columns tx+16n) read-modify-write 256 consecutive doubles per (m, n).
__host__ __device__ inline int gj_ws_index(int r, int c) { return ((((r >> 4) << 3) + (c >> 4)) << 8) + ((r & 15) << 4) + (c & 15); }

#ifdef GJ_EMULATED               // tests/emul: cp.async as cuda_block_emul.h models it; approximate / PTX instructions restated
static inline void cp_async16(void* smem, const void* gmem) { emu_cp_async(smem, gmem, 16); }
static inline void cp_async_commit() { emu_cp_commit(); }
template <int N>
static inline void cp_async_wait() { emu_cp_wait(N); }
#define GJ_DYNAMIC_SMEM(name) float* name = reinterpret_cast<float*>(emu_dynamic_smem)
#define GJ_PREFETCH_L2(ptr) (void)(ptr)
#else
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
#define GJ_DYNAMIC_SMEM(name) extern __shared__ __align__(16) float name[]
#define GJ_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#endif

// linear upper-triangular tile index -> (bi <= bj), nt = tiles per side
__host__ __device__ inline void gv2_tile_coords(long long t, int nt, int* bi, int* bj) {
    // row bi holds (nt - bi) tiles; offset(bi) = bi*nt - bi*(bi-1)/2
    double disc = (2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)t;
    int r = (int)floor(((2.0 * nt + 1.0) - sqrt(disc)) * 0.5);
    if (r < 0) r = 0;
    while (r > 0 && (long long)r * nt - (long long)r * (r - 1) / 2 > t) --r;
    while ((long long)(r + 1) * nt - (long long)(r + 1) * r / 2 <= t) ++r;
    *bi = r;
    *bj = r + (int)(t - ((long long)r * nt - (long long)r * (r - 1) / 2));
}

struct GjComp {
    const double* ws;        // [nb][ksplit][ntl][128*128] or null
    const double* rowsum;    // [nb][n]
    const unsigned char* nan;  // [nb][n] or null
    int ksplit;
    long long b_stride_ws;   // elements between replicates in ws (0 = shared by all replicates)
    long long b_stride_row;  // elements between replicates in rowsum (0 = shared)
    long long b_stride_nan;  // elements between replicates in nan (0 = shared)
};

// Fused tail of the replicate path: the tile kernel that forms the GOM-signature pair sums also does the epilogue
// (clean / combine with the PPHMM component / clamp / power / 1 - S / mirror), so the GOM pair sums never travel
// through HBM and the separate epilogue pass disappears.
struct GjFuse {
    GjComp P;                // PPHMM component (pair sums from the replicate kernels), ws == null for scheme G
    const double* g_rowsum;  // [nb][n] row sums of this kernel's own table
    const unsigned char* g_nan;
    long long g_stride_row, g_stride_nan;
    long long n;
    int scheme, out_kind;
    double pw;
    double p_scale;          // power of two that brings the PPHMM sums to <= 1 (FIXED epilogue: fp32-safe reciprocal square root seed)
    double* out;             // [nb][n][n]
    long long out_b_stride;
};
#define GJ_STAGE_LD 129      // doubles per staged column of the mirror transpose (64 x 129 x 8 B = 66 KB of the idle ring)

// 1/sqrt(x) and 1/x to full double precision without the branchy library routines: fp32 seed (22 good bits) and two
// Newton steps in fp64 (44, then 88 bits).  x > 0; values outside the fp32 range are first scaled by an even power
// of two (selects, no branch), so the whole thing is straight-line code.
__device__ __forceinline__ double gj_rsqrt(double x) {
    const bool tiny = x < 1e-30, huge = x > 1e30;
    x *= tiny ? 0x1p+200 : (huge ? 0x1p-200 : 1.0);
    double r = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    r = r * fma(-hx * r, r, 1.5);
    r = r * fma(-hx * r, r, 1.5);
    return r * (tiny ? 0x1p+100 : (huge ? 0x1p-100 : 1.0));
}
__device__ __forceinline__ double gj_rcp(double x) {
    const bool tiny = x < 1e-30, huge = x > 1e30;
    const double sc = tiny ? 0x1p+200 : (huge ? 0x1p-200 : 1.0);
    x *= sc;
    double r = (double)__frcp_rn((float)x);
    r = r * fma(-x, r, 2.0);
    r = r * fma(-x, r, 2.0);
    return r * sc;
}

// one MUFU.RSQ: the operand is a normal fp32 number by construction (no denormal handling needed)
#ifdef GJ_EMULATED
static inline float gj_rsqrt_approx(float x) { return 1.0f / sqrtf(x); }      // (the MUFU seed is good to ~22 bits; this is exact)
static inline void gj_mad1(unsigned& acc, unsigned m, unsigned one) { acc += m * one; }
#else
__device__ __forceinline__ float gj_rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void gj_mad1(unsigned& acc, unsigned m, unsigned one) {
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(m), "r"(one));
}
#endif

__device__ __forceinline__ double gj_ratio(double smin, double si, double sj, bool diag, bool isnan_row) {
    if (diag) smin = si;                           // diagonal: sum(min) == sum(a) exactly
    const double smax = si + sj - smin;
    double v = (smax == 0.0) ? 0.0 : smin / smax;  // `if not sum(max) == 0 else 0`, similarity_matrix_constructor.py:168
    if (isnan_row) v = nan("");
    if (isnan(v) || v < 0.0) v = 0.0;              // :204-205
    return v;
}

// ---------------------------------------------------------------------------------------
// K1: per-tile sum_c w_c * min(T[i,c], T[j,c])   (w == 1 when !WEIGHTED)
// grid = (tiles in [tile_begin, tile_end), ksplit, nb);  ws[((b*ksplit+ks)*ntl + t)][128*128]
template <bool WEIGHTED>
__global__ void __launch_bounds__(GJ_THREADS, 2)
gj_min_tile_kernel(const float* __restrict__ tab, long long tab_stride, int ld, int kblocks_total,
                   int nt, long long tile_begin, const float* __restrict__ wts, long long w_stride,
                   double* __restrict__ ws) {
    GJ_DYNAMIC_SMEM(smem);
    float* sA = smem;                                         // [STAGES][128][GJ_LDS]
    float* sB = sA + GJ_STAGES * GV2_TILE * GJ_LDS;           // [STAGES][128][GJ_LDS]
    float* sW = sB + GJ_STAGES * GV2_TILE * GJ_LDS;           // [STAGES][32]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int ksplit = gridDim.y, ks = blockIdx.y, b = blockIdx.z;
    const long long ntl = gridDim.x;
    int bi, bj;
    gv2_tile_coords(tile_begin + blockIdx.x, nt, &bi, &bj);

    // this CTA's K range, in 32-column blocks
    const int kb_per = (kblocks_total + ksplit - 1) / ksplit;
    const int kb0 = ks * kb_per;
    const int kb1 = min(kblocks_total, kb0 + kb_per);
    const int nkb = max(0, kb1 - kb0);

    tab += (size_t)b * tab_stride;                            // one table per replicate (GOM), or 0
    const float* gA = tab + (size_t)bi * GV2_TILE * ld;
    const float* gB = tab + (size_t)bj * GV2_TILE * ld;
    const float* gW = WEIGHTED ? wts + (size_t)b * w_stride : nullptr;

    auto load_stage = [&](int stage, int kb) {
        const int k = kb * GV2_BK;
        float* a = sA + stage * GV2_TILE * GJ_LDS;
        float* bb = sB + stage * GV2_TILE * GJ_LDS;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int q = tid + s * GJ_THREADS;
            int row = q >> 3, cc = (q & 7) * 4;
            cp_async16(a + row * GJ_LDS + cc, gA + (size_t)row * ld + k + cc);
            cp_async16(bb + row * GJ_LDS + cc, gB + (size_t)row * ld + k + cc);
        }
        if (WEIGHTED && tid < 8) cp_async16(sW + stage * GV2_BK + tid * 4, gW + k + tid * 4);
    };

    float acc[8][8];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[m][n] = 0.f;

    double* wtile = ws + (((size_t)b * ksplit + ks) * ntl + blockIdx.x) * (size_t)(GV2_TILE * GV2_TILE);

    // promote the fp32 slab totals into the fp64 workspace (first slab stores, later ones add)
    auto flush = [&](bool first) {
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                size_t o = (size_t)((m * 8 + n) * 256 + tid);      // == gj_ws_index(ty + 16m, tx + 16n)
                // first slab: plain store (no zero-init needed); later slabs: fire-and-forget RED.ADD.F64.
                // Each address is only ever touched by this one thread, in program order, so the
                // fp64 sum is formed in slab order and is deterministic.
                if (first) wtile[o] = (double)acc[m][n];
                else atomicAdd(&wtile[o], (double)acc[m][n]);
                acc[m][n] = 0.f;
            }
    };

#pragma unroll
    for (int s = 0; s < GJ_STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, kb0 + s);
        cp_async_commit();
    }

    bool first = true;
    int in_slab = 0;
    for (int it = 0; it < nkb; ++it) {
        cp_async_wait<GJ_STAGES - 2>();
        __syncthreads();
        {   // prefetch stage it + STAGES-1 (its buffer was consumed at iteration it-1)
            int nxt = it + GJ_STAGES - 1;
            if (nxt < nkb) load_stage(nxt % GJ_STAGES, kb0 + nxt);
            cp_async_commit();
        }
        const int stage = it % GJ_STAGES;
        const float* a = sA + stage * GV2_TILE * GJ_LDS + ty * GJ_LDS;
        const float* bb = sB + stage * GV2_TILE * GJ_LDS + tx * GJ_LDS;
        const float* w = sW + stage * GV2_BK;
        GJ_UNROLL(GJ_KG_UNROLL)
        for (int kg = 0; kg < GV2_BK / 4; ++kg) {
            float4 bf[8];
#pragma unroll
            for (int n = 0; n < 8; ++n)
                bf[n] = *reinterpret_cast<const float4*>(bb + n * 16 * GJ_LDS + kg * 4);
            float4 wv = make_float4(1.f, 1.f, 1.f, 1.f);
            if (WEIGHTED) wv = *reinterpret_cast<const float4*>(w + kg * 4);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                float4 af = *reinterpret_cast<const float4*>(a + m * 16 * GJ_LDS + kg * 4);
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    // 4 columns are summed as a tree before they touch the (large) accumulator: same
                    // instruction count as a chain, a quarter of the accumulator roundings
                    if (WEIGHTED) {
#ifdef GJ_W_TREE
                        float s0 = fmaf(wv.x, fminf(af.x, bf[n].x), wv.y * fminf(af.y, bf[n].y));
                        float s1 = fmaf(wv.z, fminf(af.z, bf[n].z), wv.w * fminf(af.w, bf[n].w));
                        acc[m][n] += s0 + s1;
#else
                        acc[m][n] = fmaf(wv.x, fminf(af.x, bf[n].x), acc[m][n]);
                        acc[m][n] = fmaf(wv.y, fminf(af.y, bf[n].y), acc[m][n]);
                        acc[m][n] = fmaf(wv.z, fminf(af.z, bf[n].z), acc[m][n]);
                        acc[m][n] = fmaf(wv.w, fminf(af.w, bf[n].w), acc[m][n]);
#endif
                    } else {
                        float s0 = fminf(af.x, bf[n].x) + fminf(af.y, bf[n].y);
                        float s1 = fminf(af.z, bf[n].z) + fminf(af.w, bf[n].w);
                        acc[m][n] += s0 + s1;
                    }
                }
            }
        }
        in_slab += GV2_BK;
        if (in_slab >= GJ_SLAB) {
            flush(first);
            first = false;
            in_slab = 0;
        }
    }
    if (in_slab > 0 || first) flush(first);
}

// ---------------------------------------------------------------------------------------
// Fused tail of the replicate path: GOM-signature pair sums (K1 arithmetic on a 128 x 64 pair tile) + the whole
// epilogue (clean / combine with the PPHMM component / clamp / power / 1 - S / mirror) in one kernel, so the GOM
// pair sums never travel through HBM and the separate epilogue pass over the PPHMM workspace disappears.
//   grid = (2 * tiles, 1, replicates): CTA (t, half) owns rows [0,128) x columns [64 half, 64 half + 64) of tile t;
//   thread (ty, tx) owns rows ty + 16 m (m < 8), columns 64 half + tx + 16 n (n < 4): 32 fp32 slab accumulators
//   promoted every GF_SLAB columns into 32 fp64 running sums that stay in registers.
#define GF_STAGES 3
#define GF_SLAB 64
#define GF_STAGE_FLOATS ((GV2_TILE + 64) * GJ_LDS)
static size_t gf_smem_bytes() { return (size_t)GF_STAGES * GF_STAGE_FLOATS * sizeof(float); }

// FIXED == false: fp32 table, fp32 slab sums promoted to fp64 (any table).  FIXED == true: the table holds unsigned fixed-point
// integers whose row sums fit 32 bits (the job's GOM signature tables, values in [0, 1] scaled by 2^s): integer min + integer
// add are exact, so there is no slab promotion, no fp64 running sum -- 64 registers less per thread, which pays for loading
// all 32 PPHMM pair sums of the thread in one batch -- and duplicate genomes give exactly S = 1.  The add is written as a
// multiply-add by `one` (a kernel argument equal to 1) so that it issues on the FMA pipe (IMAD) while the min issues on the ALU
// pipe: one instruction per pipe and cell, as in the fp32 variant (FMNMX + FADD).
// HAS_P / POWERED are compile-time so that the common case (scheme PG, p == 1) carries neither the pow() call nor its
// divergence bookkeeping: the epilogue is ~40 straight-line instructions per cell (it was ~200, half the main loop's cost).
template <bool FIXED, bool HAS_P, bool POWERED>
__global__ void __launch_bounds__(GJ_THREADS, 2)
gj_fused_kernel(const float* __restrict__ tab, long long tab_stride, int ld, int nkb, int k_groups, int nt, const GjFuse fz, unsigned one) {
    GJ_DYNAMIC_SMEM(smem);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long t = blockIdx.x >> 1;
    const int half = blockIdx.x & 1;
    const long long b = blockIdx.z;
    int bi, bj;
    gv2_tile_coords(t, nt, &bi, &bj);
    tab += (size_t)b * tab_stride;                            // one GOM table per replicate
    const float* gA = tab + (size_t)bi * GV2_TILE * ld;
    const float* gB = tab + ((size_t)bj * GV2_TILE + half * 64) * ld;

    auto load_stage = [&](int stage, int kb) {
        const int k = kb * GV2_BK;
        float* a = smem + stage * GF_STAGE_FLOATS;
        float* bb = a + GV2_TILE * GJ_LDS;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int q = tid + s * GJ_THREADS, row = q >> 3, cc = (q & 7) * 4;
            cp_async16(a + row * GJ_LDS + cc, gA + (size_t)row * ld + k + cc);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int q = tid + s * GJ_THREADS, row = q >> 3, cc = (q & 7) * 4;
            cp_async16(bb + row * GJ_LDS + cc, gB + (size_t)row * ld + k + cc);
        }
    };

    float acc[8][4];
    unsigned iacc[8][4];
    double sum[8][4];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) { acc[m][n] = 0.f; iacc[m][n] = 0u; sum[m][n] = 0.0; }

#pragma unroll
    for (int s = 0; s < GF_STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }
    if (HAS_P) {
        // the 64 KB of PPHMM pair sums this CTA reads in its epilogue: pull them into L2 now, under the main loop
        // (element (r, c) of the tile is at [r/16][c/16][r%16][c%16]: per r/16 the CTA's half is 8 KB contiguous)
        const double* pw = fz.P.ws + b * fz.P.b_stride_ws + t * (long long)(GV2_TILE * GV2_TILE);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int line = tid + q * GJ_THREADS;                    // 512 lines of 128 bytes
            const double* a = pw + (size_t)(line >> 6) * 2048 + (size_t)half * 1024 + (size_t)(line & 63) * 16;
            GJ_PREFETCH_L2(a);
        }
    }
    int in_slab = 0;
    for (int it = 0; it < nkb; ++it) {
        cp_async_wait<GF_STAGES - 2>();
        __syncthreads();
        {
            const int nxt = it + GF_STAGES - 1;
            if (nxt < nkb) load_stage(nxt % GF_STAGES, nxt);
            cp_async_commit();
        }
        const float* a = smem + (it % GF_STAGES) * GF_STAGE_FLOATS + ty * GJ_LDS;
        const float* bb = smem + (it % GF_STAGES) * GF_STAGE_FLOATS + (GV2_TILE + tx) * GJ_LDS;
        // the zero padding beyond the last real column (G = 200 is padded to 224) contributes nothing: skip its 4-column groups
        const int nkg = min(GV2_BK / 4, k_groups - it * (GV2_BK / 4));
#pragma unroll 2
        for (int kg = 0; kg < nkg; ++kg) {
            float4 bf[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) bf[n] = *reinterpret_cast<const float4*>(bb + n * 16 * GJ_LDS + kg * 4);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float4 af = *reinterpret_cast<const float4*>(a + m * 16 * GJ_LDS + kg * 4);
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    if (FIXED) {
                        // one min (ALU pipe) + one multiply-add by 1 (FMA pipe) per cell; written as PTX so that the four
                        // adds are not re-associated into 3-input integer adds, which issue on the ALU pipe as well
                        unsigned v = iacc[m][n];
                        gj_mad1(v, min(__float_as_uint(af.x), __float_as_uint(bf[n].x)), one);
                        gj_mad1(v, min(__float_as_uint(af.y), __float_as_uint(bf[n].y)), one);
                        gj_mad1(v, min(__float_as_uint(af.z), __float_as_uint(bf[n].z)), one);
                        gj_mad1(v, min(__float_as_uint(af.w), __float_as_uint(bf[n].w)), one);
                        iacc[m][n] = v;
                    } else {
                        const float s0 = fminf(af.x, bf[n].x) + fminf(af.y, bf[n].y);
                        const float s1 = fminf(af.z, bf[n].z) + fminf(af.w, bf[n].w);
                        acc[m][n] += s0 + s1;
                    }
                }
            }
        }
        in_slab += GV2_BK;
        if (!FIXED && (in_slab >= GF_SLAB || it + 1 == nkb)) {            // fp32 slab totals -> fp64 running sums (registers)
#pragma unroll
            for (int m = 0; m < 8; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) { sum[m][n] += (double)acc[m][n]; acc[m][n] = 0.f; }
            in_slab = 0;
        }
    }
    cp_async_wait<0>();

    // ------------------------------------------------------------------ epilogue
    const long long n = fz.n;
    const double qnan = nan("");
    const double* g_rs = fz.g_rowsum + b * fz.g_stride_row;
    const unsigned char* g_nf = fz.g_nan ? fz.g_nan + b * fz.g_stride_nan : nullptr;
    const double* p_ws = HAS_P ? fz.P.ws + b * fz.P.b_stride_ws + t * (long long)(GV2_TILE * GV2_TILE) : nullptr;
    const double* p_rs = HAS_P ? fz.P.rowsum + b * fz.P.b_stride_row : nullptr;
    const unsigned char* p_nf = (HAS_P && fz.P.nan) ? fz.P.nan + b * fz.P.b_stride_nan : nullptr;
    double* o = fz.out + b * fz.out_b_stride;
    double* stage = reinterpret_cast<double*>(smem);                 // [64 columns][GJ_STAGE_LD]: the mirror transpose
    const bool as_distance = fz.out_kind == GV2_OUT_DISTANCE;
    const double psc = fz.p_scale;
    // row sums of the partner genomes; a genome with a NaN entry gets NaN sums, which the `> 0` tests below turn into 0
    // (the PPHMM sums are scaled by a power of two: exact, and the ratio does not see it)
    double gsj[4], psj[4];
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
        const long long gj = (long long)bj * GV2_TILE + half * 64 + tx + 16 * nn;
        const bool ok = gj < n;
        gsj[nn] = ok ? ((g_nf && g_nf[gj]) ? qnan : g_rs[gj]) : 0.0;
        psj[nn] = (ok && HAS_P) ? ((p_nf && p_nf[gj]) ? qnan : p_rs[gj] * psc) : 0.0;
    }
    __syncthreads();                                                 // the ring is dead: reuse it as the transpose stage
    // Two instances of the same loop body: interior off-diagonal tiles (all but ~2 n/128 of them) need no bounds tests,
    // no diagonal special case and no index swap.
    auto rows = [&](auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
        double* orow = o + ((long long)bi * GV2_TILE + ty) * n + (long long)bj * GV2_TILE + half * 64 + tx;
        const long long row_step = 16 * n;
        // FIXED frees the registers of the fp64 running sums: the pair sums of the thread are requested 16 at a time (one
        // memory latency per batch instead of one per row group)
        double pall[(FIXED && FAST && HAS_P) ? 4 : 1][4];
#pragma unroll
        for (int m = 0; m < 8; ++m, orow += row_step) {
            if (FIXED && FAST && HAS_P && (m & 3) == 0) {
#pragma unroll
                for (int mm = 0; mm < 4; ++mm)
#pragma unroll
                    for (int nn = 0; nn < 4; ++nn) pall[mm][nn] = p_ws[(size_t)(((m + mm) * 8 + half * 4 + nn) * 256 + tid)];
            }
            const int r = ty + 16 * m;
            const long long gi = (long long)bi * GV2_TILE + r;
            const bool oki = FAST || gi < n;
            const double gsi = oki ? ((g_nf && g_nf[gi]) ? qnan : g_rs[gi]) : 0.0;
            const double psi = (oki && HAS_P) ? ((p_nf && p_nf[gi]) ? qnan : p_rs[gi] * psc) : 0.0;
            double pv[4];
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
                if (!HAS_P) {
                    pv[nn] = 0.0;
                } else if (FIXED && FAST) {
                    pv[nn] = pall[m & 3][nn] * psc;
                } else if (FAST) {
                    pv[nn] = p_ws[(size_t)((m * 8 + half * 4 + nn) * 256 + tid)] * psc;   // == gj_ws_index(r, c)
                } else {
                    const int c = half * 64 + tx + 16 * nn;
                    // diagonal tiles: the replicate kernels leave the entries below the diagonal unwritten
                    const int rr = (bi == bj && r > c) ? c : r, cc = (bi == bj && r > c) ? r : c;
                    pv[nn] = p_ws[gj_ws_index(rr, cc)] * psc;
                }
            }
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
                const int cl = tx + 16 * nn;
                const long long gj = (long long)bj * GV2_TILE + half * 64 + cl;
                // GJ = sum(min) / sum(max), 0 when sum(max) == 0 (:168), NaN and negatives cleaned to 0 (:204-205), then
                // (P * G) ** 0.5 (:214) as ONE reciprocal square root: sqrt(Np Ng / (Dp Dg)) = Np Ng rsqrt(Np Ng Dp Dg).
                // num > 0 and den > 0 say: both sum(max) != 0, both ratios positive, nothing NaN (all sums are >= 0).
                const bool diag = !FAST && gi == gj;
                const double ng = diag ? gsi : (FIXED ? (double)iacc[m][nn] : sum[m][nn]);   // diagonal: sum(min) == sum(a) exactly
                const double dg = gsi + gsj[nn] - ng;
                double sv;
                if (HAS_P) {
                    const double np_ = diag ? psi : pv[nn];
                    const double dp = psi + psj[nn] - np_;
                    const double num = np_ * ng, den = dp * dg, x = num * den;
                    double rq;
                    if (FIXED) {
                        // Ng, Dg are integers in [1, 2^32], Np, Dp lie in [2^-71, 1] after the scaling: both products are fp32
                        // normal numbers, so the seed needs no range handling; one Newton step in fp64 takes its 21 good bits
                        // to 41 (3e-13 relative, against the 1e-6 the similarity is held to)
                        rq = (double)(gj_rsqrt_approx((float)num) * gj_rsqrt_approx((float)den));
                        rq = rq * fma(-0.5 * x * rq, rq, 1.5);
                    } else {
                        rq = gj_rsqrt(x);
                    }
                    sv = (num == den) ? 1.0 : num * rq;              // == : diagonal, duplicate genomes
                    if (!(num > 0.0 && den > 0.0)) sv = 0.0;
                } else {
                    sv = (ng == dg) ? 1.0 : ng * gj_rcp(fabs(dg));
                    if (!(ng > 0.0 && dg > 0.0)) sv = 0.0;
                }
                if (POWERED) sv = pow(sv, fz.pw);                    // :226 (sv >= 0 already, :224)
                if (as_distance) sv = fmax(1.0 - sv, 0.0);           // pl1_graphs.py:52-53
                if (FAST || (oki && gj < n)) orow[16 * nn] = sv;
                stage[cl * GJ_STAGE_LD + r] = sv;
            }
        }
    };
    const bool interior = bi != bj && ((long long)bi + 1) * GV2_TILE <= n && (long long)bj * GV2_TILE + half * 64 + 64 <= n;
    if (interior) rows(std::true_type{});
    else rows(std::false_type{});
    if (bi == bj) return;                                            // diagonal tiles hold both triangles already
    __syncthreads();
    // mirror: rows bj*128 + 64 half + [0, 64), columns bi*128 + [0, 128): 128 consecutive doubles per staged row
    for (int q = tid; q < 64 * GV2_TILE; q += GJ_THREADS) {
        const int cl = q >> 7, r = q & 127;
        const long long gj = (long long)bj * GV2_TILE + half * 64 + cl, gi = (long long)bi * GV2_TILE + r;
        if (gi < n && gj < n) o[gj * n + gi] = stage[cl * GJ_STAGE_LD + r];
    }
}

// ---------------------------------------------------------------------------------------
// table preparation: float64 (n x k, ld) -> zero-padded float32 (n_pad x k_pad), score
// threshold `x < thr -> 0` (similarity_matrix_constructor.py:139-140 with weight == 0),
// per-row fp64 sum of the fp32 values and NaN flag.  One block per destination row.
__global__ void prepare_table_kernel(const double* __restrict__ src, long long n, long long k,
                                     long long ld_src, double thr, int apply_thr,
                                     float* __restrict__ dst, long long k_pad,
                                     double* __restrict__ rowsum, unsigned char* __restrict__ nanflag) {
    const long long r = blockIdx.x;
    float* d = dst + r * k_pad;
    double s = 0.0;
    int has_nan = 0;
    if (r < n) {
        const double* p = src + r * ld_src;
        for (long long c = threadIdx.x; c < k_pad; c += blockDim.x) {
            float f = 0.f;
            if (c < k) {
                double v = p[c];
                if (apply_thr && v < thr) v = 0.0;
                // a NaN entry is carried by the row flag (per replicate where the table is resampled); the image holds 0
                if (isnan(v)) { has_nan = 1; v = 0.0; }
                f = (float)v;
                s += (double)f;
            }
            d[c] = f;
        }
    } else {
        for (long long c = threadIdx.x; c < k_pad; c += blockDim.x) d[c] = 0.f;
    }
    __shared__ double red[32];
    __shared__ int rnan;
    if (threadIdx.x == 0) rnan = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    if (has_nan) rnan = 1;
    __syncthreads();
    if (threadIdx.x == 0 && r < n) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        if (rowsum) rowsum[r] = t;
        if (nanflag) nanflag[r] = (unsigned char)rnan;
    }
}

// weighted row sums: out[b][i] = sum_c w[b][c] * T[i][c]  (fp64).  grid = (n, nb)
__global__ void weighted_rowsum_kernel(const float* __restrict__ tab, long long k_pad,
                                       const float* __restrict__ w, long long w_stride,
                                       long long n, double* __restrict__ out) {
    const long long r = blockIdx.x, b = blockIdx.y;
    const float* t = tab + r * k_pad;
    const float* ww = w + b * w_stride;
    double s = 0.0;
    for (long long c = threadIdx.x * 4; c < k_pad; c += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(t + c);
        float4 q = *reinterpret_cast<const float4*>(ww + c);
        s += (double)v.x * q.x + (double)v.y * q.y + (double)v.z * q.z + (double)v.w * q.w;
    }
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tt = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tt += red[i];
        out[b * n + r] = tt;
    }
}

// int32 column multiplicities -> fp32 weights, zero-padded to k_pad.  grid = (ceil(k_pad/256), nb)
__global__ void weights_to_float_kernel(const int* __restrict__ w, long long k, long long k_pad,
                                        float* __restrict__ out) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long b = blockIdx.y;
    if (c < k_pad) out[b * k_pad + c] = c < k ? (float)w[b * k + c] : 0.f;
}

// ---------------------------------------------------------------------------------------
// K3 epilogue: NaN/negative clean, scheme combine, clamp, **p, optional 1-S, mirror.
// similarity_matrix_constructor.py:201-226, pl1_graphs.py:52-53.
// One block per 32x32 sub-block of a 128x128 tile: grid = (tiles*16, nb), block = (32, 8).
__device__ __forceinline__ double gj_value(const GjComp& c, long long b, long long ntl,
                                           long long t, int r, int cc, long long gi, long long gj) {
    // diagonal tiles: always read the upper-triangle entry (the tensor-core path leaves sub-tiles that lie
    // wholly below the diagonal unwritten; the value is symmetric)
    if (gi > gj && gi - r == gj - cc) { int tmp = r; r = cc; cc = tmp; }
    const double* w = c.ws + b * c.b_stride_ws + t * (long long)(GV2_TILE * GV2_TILE) + gj_ws_index(r, cc);
    double smin = 0.0;
    for (int ks = 0; ks < c.ksplit; ++ks) smin += w[(long long)ks * ntl * (GV2_TILE * GV2_TILE)];
    const double* rs = c.rowsum + b * c.b_stride_row;
    double si = rs[gi], sj = rs[gj];
    if (gi == gj) smin = si;                       // diagonal: sum(min) == sum(a) exactly
    double smax = si + sj - smin;
    double v = (smax == 0.0) ? 0.0 : smin / smax;  // `if not sum(max) == 0 else 0`, :168
    if (c.nan) {
        const unsigned char* nf = c.nan + b * c.b_stride_nan;
        if (nf[gi] | nf[gj]) v = nan("");
    }
    if (isnan(v) || v < 0.0) v = 0.0;              // :204-205
    return v;
}

__global__ void __launch_bounds__(256)
gj_epilogue_kernel(long long n, int nt, long long tile_begin, long long ntl, GjComp P, GjComp G,
                   const double* __restrict__ spr, int scheme, double pw, int out_kind,
                   int packed, double* __restrict__ out, long long out_b_stride) {
    __shared__ double tr[32][33];
    const long long t = blockIdx.x >> 4;
    const int sub = blockIdx.x & 15, sr = sub >> 2, sc = sub & 3;
    const long long b = blockIdx.y;
    int bi, bj;
    gv2_tile_coords(tile_begin + t, nt, &bi, &bj);
    double* o = out + b * out_b_stride;
    const int lx = threadIdx.x;
    const bool mirror = !packed && (bi != bj || sr != sc);
    if (!packed && bi == bj && sr > sc) return;     // lower sub-blocks of a diagonal tile
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const int ly = threadIdx.y + yy * 8;
        const int r = sr * 32 + ly, c = sc * 32 + lx;
        const long long gi = (long long)bi * GV2_TILE + r, gj = (long long)bj * GV2_TILE + c;
        double s = 0.0;
        if (gi < n && gj < n) {
            double vp = 0.0, vg = 0.0, vr = 0.0;
            if (scheme == GV2_SCHEME_P || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_PR)
                vp = gj_value(P, b, ntl, t, r, c, gi, gj);
            if (scheme == GV2_SCHEME_G || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_RG)
                vg = gj_value(G, b, ntl, t, r, c, gi, gj);
            if (scheme == GV2_SCHEME_PL) vp = gj_value(P, b, ntl, t, r, c, gi, gj);
            if (scheme == GV2_SCHEME_R || scheme == GV2_SCHEME_RG || scheme == GV2_SCHEME_PR)
                vr = spr[gi * n + gj];
            if (scheme == GV2_SCHEME_L || scheme == GV2_SCHEME_PL) {   // aux = location dCor matrix
                vr = spr[gi * n + gj];
                if (isnan(vr) || vr < 0.0) vr = 0.0;                   // :201-205 cleans "L" too
            }
            switch (scheme) {
                case GV2_SCHEME_P: s = vp; break;
                case GV2_SCHEME_G: s = vg; break;
                case GV2_SCHEME_PG: s = sqrt(vp * vg); break;     // (P*G)**0.5, :214
                case GV2_SCHEME_R: s = vr; break;
                case GV2_SCHEME_RG: s = sqrt(vr * vg); break;     // :220
                case GV2_SCHEME_PR: s = sqrt(vr * vp); break;     // :222
                case GV2_SCHEME_L: s = vr; break;                 // :212
                case GV2_SCHEME_PL: s = vp * vr; break;           // :216
            }
            if (s < 0.0) s = 0.0;                                 // :224
            if (pw != 1.0) s = pow(s, pw);                        // :226
            if (out_kind == GV2_OUT_DISTANCE) {                   // pl1_graphs.py:52-53
                s = 1.0 - s;
                if (s < 0.0) s = 0.0;
            }
            if (packed) o[t * (long long)(GV2_TILE * GV2_TILE) + r * GV2_TILE + c] = s;
            else o[gi * n + gj] = s;
        } else if (packed) {
            o[t * (long long)(GV2_TILE * GV2_TILE) + r * GV2_TILE + c] = 0.0;
        }
        tr[ly][lx] = s;
    }
    if (!mirror) return;
    __syncthreads();
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const int ly = threadIdx.y + yy * 8;
        // element (row = sc*32+ly of block bj, col = sr*32+lx of block bi) = value[lx][ly]
        const long long gj = (long long)bj * GV2_TILE + sc * 32 + ly;
        const long long gi = (long long)bi * GV2_TILE + sr * 32 + lx;
        if (gi < n && gj < n) o[gj * n + gi] = tr[lx][ly];
    }
}

// packed tiles [ntl][128*128] -> symmetric n x n (after an all-gather of tile blocks).
__global__ void __launch_bounds__(256)
unpack_tiles_kernel(long long n, int nt, long long tile_begin, const double* __restrict__ packed,
                    double* __restrict__ out) {
    __shared__ double tr[32][33];
    const long long t = blockIdx.x >> 4;
    const int sub = blockIdx.x & 15, sr = sub >> 2, sc = sub & 3;
    int bi, bj;
    gv2_tile_coords(tile_begin + t, nt, &bi, &bj);
    if (bi == bj && sr > sc) return;
    const int lx = threadIdx.x;
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const int ly = threadIdx.y + yy * 8;
        const int r = sr * 32 + ly, c = sc * 32 + lx;
        const long long gi = (long long)bi * GV2_TILE + r, gj = (long long)bj * GV2_TILE + c;
        double s = packed[t * (long long)(GV2_TILE * GV2_TILE) + r * GV2_TILE + c];
        if (gi < n && gj < n) out[gi * n + gj] = s;
        tr[ly][lx] = s;
    }
    if (bi == bj && sr == sc) return;
    __syncthreads();
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const int ly = threadIdx.y + yy * 8;
        const long long gj = (long long)bj * GV2_TILE + sc * 32 + ly;
        const long long gi = (long long)bi * GV2_TILE + sr * 32 + lx;
        if (gi < n && gj < n) out[gj * n + gi] = tr[lx][ly];
    }
}

// ---------------------------------------------------------------------------------------
// Fixed-point image of a job's GOM signature tables for the FIXED fused tail (see gom_quantize_launch in gj_kernels.cu)
// pass 1: grid = (n, nb), block = 128: largest fp64 row sum of the launch (bits of a non-negative double order like integers)
__global__ void gom_rowmax_kernel(const double* __restrict__ src, long long n, long long g, unsigned long long* __restrict__ maxbits) {
    const double* p = src + ((long long)blockIdx.y * n + blockIdx.x) * g;
    double s = 0.0;
    for (long long c = threadIdx.x; c < g; c += blockDim.x) {
        const double v = p[c];
        if (v > 0.0) s += v;                         // NaN and negatives contribute nothing
    }
    __shared__ double red[4];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double t = red[0] + red[1] + red[2] + red[3];
        if (t > 0.0 && t < 1e300) atomicMax(maxbits, (unsigned long long)__double_as_longlong(t));
    }
}
// pass 2: grid = (n_pad, nb), block = 128
__global__ void gom_quantize_kernel(const double* __restrict__ src, long long n, long long g, const unsigned long long* __restrict__ maxbits,
                                    unsigned* __restrict__ dst, long long n_pad, long long g_pad, double* __restrict__ rowsum,
                                    unsigned char* __restrict__ nanflag) {
    const long long r = blockIdx.x, b = blockIdx.y;
    const double mx = __longlong_as_double((long long)*maxbits) * (1.0 + 1e-9);      // fp64 row sums are rounded themselves
    int k = 30;
    while (k > 0 && ldexp(mx, k) + 0.5 * (double)g_pad >= 4294967295.0) --k;
    const double scale = ldexp(1.0, k);
    unsigned* d = dst + (b * n_pad + r) * g_pad;
    unsigned s = 0;
    int has_nan = 0;
    if (r < n) {
        const double* p = src + (b * n + r) * g;
        for (long long c = threadIdx.x; c < g_pad; c += blockDim.x) {
            unsigned q = 0;
            if (c < g) {
                const double v = p[c];
                if (isnan(v)) has_nan = 1;
                else if (v > 0.0) q = (unsigned)fmin(rint(v * scale), 4294967295.0);
            }
            d[c] = q;
            s += q;
        }
    } else {
        for (long long c = threadIdx.x; c < g_pad; c += blockDim.x) d[c] = 0u;
    }
    __shared__ unsigned red[4];
    __shared__ int rnan;
    if (threadIdx.x == 0) rnan = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    if (has_nan) rnan = 1;
    __syncthreads();
    if (threadIdx.x == 0 && r < n) {
        rowsum[b * n + r] = (double)(red[0] + red[1] + red[2] + red[3]);
        nanflag[b * n + r] = (unsigned char)rnan;
    }
}

