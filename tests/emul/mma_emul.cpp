// Runs the tensor-core replicate pair sums -- the SAME kernel source the GPU build compiles
// (gravity2_b200/csrc/boot_mma_kernels.cuh) -- on host threads, following boot_table_stats / boot_mma_prepare_table /
// boot_mma_launch of boot_mma.cu step by step.  The PTX the kernel issues through its wrappers is replaced by a FUNCTIONAL
// MODEL of the sm_100a units, written from the PTX ISA's description of them (TEST INFRASTRUCTURE, not a timing model):
//   mbarrier       pending-arrival count + transaction bytes + phase bit; try_wait.parity(p) succeeds once the phase with
//                  parity p has completed (so a wait on parity 1 of a fresh barrier passes at once, as the kernel relies on);
//   TMA            cp.async.bulk.tensor.2d of a uint8 map with SWIZZLE_128B: box rows of 128 bytes, 16-byte chunk index
//                  XOR (row % 8), out-of-bounds elements zero, then complete_tx on the mbarrier;
//   tcgen05.mma    kind::i8, cta_group::1: operands read from shared memory through the 64-bit matrix descriptors the kernel
//                  builds (start address, stride byte offset, SWIZZLE_128B applied to ABSOLUTE shared-address bits [4,7) ^=
//                  bits [7,10)), M / N from the instruction descriptor, K = 32 bytes, D[m][n] (+)= sum_k A[m][k] * B[n][k] in
//                  int32 into TMEM lane m, column n;
//   tcgen05.commit executes the MMAs issued since the last commit (operands read at the last legal moment), then arrives on
//                  the mbarrier;
//   tcgen05.ld     32x32b.x16: a warp may only touch the TMEM lane quarter 32 * (warp % 4) -- asserted;
//   cp.async       the emulator's two-extremes model (cuda_block_emul.h); bar.sync 1, 512: a second block barrier.
// What this checks without a GPU: the warp-role protocol (no deadlock, every stage handed over in order), the swizzled A-tile
// addresses the producers write against the addresses the descriptors make the tensor core read, the digit-plane
// recombination and the workspace layout.  The compiled kernel itself is checked on the B200
// (tests/test_gpu_parity.py::test_tensor_core_path_is_exact_integer_arithmetic, ::test_replicate_paths_agree).
//   mma_emul <n> <k> <nb>  < table (n*k float64), multiplicities (nb*k int32)
//        > quantum, overflow flag (2 float64), weighted row sums (nb*n float64), workspace (nb * tiles * 128*128 float64; NaN
//          where nothing is written)
#define BM_EMULATED
#include "cuda_block_emul.h"
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <condition_variable>
#include <mutex>

#define GV2_TILE 128
#define __grid_constant__
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// what cuTensorMapEncodeTiled is given in boot_mma_launch: uint8, rank 2, dims {k_pad, nb_pad}, row stride k_pad bytes,
// box {128, nblk}, SWIZZLE_128B, out-of-bounds fill zero
struct CUtensorMap { const uint8_t* base; long long dim0, dim1, stride1; int box0, box1; };

// ---- shared window: address 16 is the start of the dynamic array, so that the kernel's own 1024-byte alignment code has work to do
static const uint32_t EMU_SMEM_ORIGIN = 16;
static inline uint32_t smem_u32(const void* p) { return EMU_SMEM_ORIGIN + (uint32_t)((const uint8_t*)p - (const uint8_t*)emu_dynamic_smem); }
static inline uint8_t* emu_sptr(uint32_t a) { return (uint8_t*)emu_dynamic_smem + (a - EMU_SMEM_ORIGIN); }
static inline uint32_t emu_swizzle128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }

// ---- mbarrier
struct EmuMbar { uint32_t expected; int32_t pending; int32_t tx; uint32_t phase; };
static_assert(sizeof(EmuMbar) == 16, "two 64-bit slots");      // the kernel's barriers are 8 bytes apart: state is kept aside
static std::mutex emu_mb_m;
static std::condition_variable emu_mb_cv;
static EmuMbar emu_mb_state[16];
static uint64_t* emu_mb_first = nullptr;                        // the kernel's `bars` array: index = distance from it
static EmuMbar& emu_mb(uint64_t* bar) {
    if (!emu_mb_first || bar < emu_mb_first) emu_mb_first = bar;
    const long i = bar - emu_mb_first;
    assert(i >= 0 && i < 16);
    return emu_mb_state[i];
}
static void emu_mb_complete(EmuMbar& b) {
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = (int32_t)b.expected; emu_mb_cv.notify_all(); }
}
static inline void mbar_init(uint64_t* bar, uint32_t count) {
    std::lock_guard<std::mutex> g(emu_mb_m);
    emu_mb(bar) = EmuMbar{count, (int32_t)count, 0, 0};
}
static inline void mbar_arrive(uint64_t* bar) {
    std::lock_guard<std::mutex> g(emu_mb_m);
    EmuMbar& b = emu_mb(bar);
    assert(b.pending > 0);
    --b.pending;
    emu_mb_complete(b);
}
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    std::lock_guard<std::mutex> g(emu_mb_m);
    EmuMbar& b = emu_mb(bar);
    assert(b.pending > 0);
    b.tx += (int32_t)bytes;
    --b.pending;
    emu_mb_complete(b);
}
static inline void emu_mb_complete_tx(uint64_t* bar, uint32_t bytes) {
    std::lock_guard<std::mutex> g(emu_mb_m);
    EmuMbar& b = emu_mb(bar);
    b.tx -= (int32_t)bytes;
    emu_mb_complete(b);
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {      // bounded, as on the device: a protocol bug aborts
    std::unique_lock<std::mutex> lk(emu_mb_m);
    EmuMbar& b = emu_mb(bar);
    if (!emu_mb_cv.wait_for(lk, std::chrono::seconds(120), [&] { return (b.phase & 1u) != parity; })) {
        fprintf(stderr, "mbarrier wait timed out (protocol bug)\n");
        abort();
    }
}
static inline void fence_barrier_init() {}
static inline void fence_proxy_async() {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}

// ---- TMA
static inline void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    const uint32_t d0 = smem_u32(dst);
    assert((d0 & 1023u) == 0 && map->box0 == 128);               // SWIZZLE_128B: 128-byte inner box, 1024-byte aligned destination
    for (int r = 0; r < map->box1; ++r)
        for (int c = 0; c < map->box0; ++c) {
            const long long gr = (long long)c1 + r, gc = (long long)c0 + c;
            const uint8_t v = (gr < map->dim1 && gc < map->dim0) ? map->base[gr * map->stride1 + gc] : 0;
            *emu_sptr(emu_swizzle128(d0 + (uint32_t)r * 128u + (uint32_t)c)) = v;
        }
    emu_mb_complete_tx(bar, (uint32_t)(map->box0 * map->box1));
}

// ---- tcgen05
static int32_t emu_tmem[128][512];
static inline void tmem_alloc(uint32_t* slot) { if ((emu_linear_tid & 31) == 0) *slot = 0; }      // lane 0, column 0: all 512 columns
static inline void tmem_dealloc(uint32_t) {}
static inline uint64_t umma_desc(uint32_t saddr);              // the kernel's own encoders (boot_mma_kernels.cuh)
static inline uint32_t umma_idesc(int m, int n);
static inline uint8_t emu_operand_byte(uint64_t desc, int row, int kbyte) {
    const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4, sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
    assert((desc >> 61) == 2 && ((desc >> 46) & 3) == 1);        // SWIZZLE_128B, descriptor version 1
    const uint32_t logical = start + (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 128u + (uint32_t)kbyte;
#ifdef MMA_EMUL_LINEAR_OPERANDS    // mutant for the test suite: a tensor core that ignores the descriptor's swizzle mode
    return *emu_sptr(logical);
#else
    return *emu_sptr(emu_swizzle128(logical));
#endif
}
// The tensor core may read a stage's operands at any time between the issue of an MMA and the commit that covers it: the model
// reads them at the LAST legal moment (tcgen05.commit executes the queued MMAs, then arrives), so a stage overwritten before
// its commit -- a missing wait on the `empty` barrier -- corrupts the sums instead of going unnoticed.
struct EmuMma { uint32_t tmem_d; uint64_t da, db; uint32_t idesc, accumulate; };
static thread_local std::vector<EmuMma> emu_mma_queue;
static inline void umma_u8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    emu_mma_queue.push_back(EmuMma{tmem_d, da, db, idesc, accumulate});
}
static inline void emu_mma_execute(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    const int n = (int)((idesc >> 17) & 0x3F) << 3, m = (int)((idesc >> 24) & 0x1F) << 4;
    assert(((idesc >> 4) & 3) == 2 && ((idesc >> 7) & 7) == 0 && ((idesc >> 10) & 7) == 0 && !((idesc >> 15) & 3));   // S32 <- U8 x U8, K-major
    assert(m == 128 && n >= 16 && n <= 256 && n % 16 == 0 && (tmem_d >> 16) == 0);
    const uint32_t col0 = tmem_d & 0xFFFF;
    assert(col0 + (uint32_t)n <= 512);
    static thread_local uint8_t A[128][32], B[256][32];
    for (int r = 0; r < m; ++r) for (int k = 0; k < 32; ++k) A[r][k] = emu_operand_byte(da, r, k);
    for (int r = 0; r < n; ++r) for (int k = 0; k < 32; ++k) B[r][k] = emu_operand_byte(db, r, k);
    for (int r = 0; r < m; ++r)
        for (int c = 0; c < n; ++c) {
            int32_t s = accumulate ? emu_tmem[r][col0 + c] : 0;
            for (int k = 0; k < 32; ++k) s += (int32_t)A[r][k] * (int32_t)B[c][k];
            emu_tmem[r][col0 + c] = s;
        }
}
static inline void umma_commit(uint64_t* bar) {
    for (const EmuMma& q : emu_mma_queue) emu_mma_execute(q.tmem_d, q.da, q.db, q.idesc, q.accumulate);
    emu_mma_queue.clear();
    mbar_arrive(bar);
}
static inline void tmem_ld16(uint32_t taddr, uint32_t* v) {
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xFFFF;
    assert(lane0 == 32u * ((uint32_t)(emu_linear_tid >> 5) & 3u) && col + 16 <= 512);
    for (int e = 0; e < 16; ++e) v[e] = (uint32_t)emu_tmem[lane0 + (emu_linear_tid & 31)][col + e];
}
static inline void tmem_ld_wait() {}

// ---- cp.async ring, named barrier of the producer warps
static inline void cp16(void* smem, const void* gmem) { emu_cp_async(smem, gmem, 16); }
static inline void cp_commit() { emu_cp_commit(); }
static EmuBarrier emu_producer_bar;
static inline void cp_wait_ring();
static inline void producers_sync();

#include "../../gravity2_b200/csrc/boot_mma_kernels.cuh"

#ifdef MMA_EMUL_WEAK_WAIT          // mutant for the test suite: the producers wait for one cp.async group too few
static inline void cp_wait_ring() { emu_cp_wait(BM_RING - 1); }
#else
static inline void cp_wait_ring() { emu_cp_wait(BM_RING - 2); }
#endif
static inline void producers_sync() { emu_producer_bar.wait(); }

static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const long long n = atoll(argv[1]), k = atoll(argv[2]), nb = atoll(argv[3]);
    std::vector<double> tab((size_t)n * k);
    std::vector<int> w((size_t)nb * k);
    if (fread(tab.data(), 8, tab.size(), stdin) != tab.size() || fread(w.data(), 4, w.size(), stdin) != w.size()) return 3;
    const long long n_pad = round_up(n, GV2_TILE), k_pad = round_up(k, BM_KC), nt = n_pad / GV2_TILE, ntl = nt * (nt + 1) / 2;
    static_assert(BM_SMEM + 1024 <= sizeof emu_dynamic_smem, "dynamic shared memory of the kernel fits the emulated block");
    emu_producer_bar.reset(BM_PRODUCERS);
    // ---- boot_table_stats
    unsigned long long stats[2] = {0, 0};
    emu_launch_grid((int)n, EmuDim{256, 1, 1}, [&] { table_stats_kernel(tab.data(), n, k, k, stats); });
    double max_value;
    memcpy(&max_value, &stats[0], 8);
    if (stats[1]) return 5;                                        // negative entries: this path is not taken (api.cu)
    // ---- boot_mma_prepare_table
    const double scale = max_value > 0.0 ? 4294967295.0 / max_value : 0.0;
    const double quantum = max_value > 0.0 ? max_value / 4294967295.0 : 0.0;
    std::vector<uint32_t> fixed((size_t)n_pad * k_pad, 0xdeadbeefu);
    emu_launch_grid((int)n_pad, EmuDim{256, 1, 1}, [&] { table_to_fixed_kernel(tab.data(), n, k, k, scale, fixed.data(), k_pad); });
    // ---- boot_mma_launch
    const long long nblocks = (nb + BM_NMAX - 1) / BM_NMAX, nblk = ((nb + nblocks - 1) / nblocks + 15) / 16 * 16, nb_pad = nblocks * nblk;
    std::vector<uint8_t> wb((size_t)nb_pad * k_pad, 0xee);
    int overflow = 0;
    emu_launch_grid(EmuDim{(int)((k_pad + 255) / 256), (int)nb_pad, 1}, EmuDim{256, 1, 1},
                    [&] { weights_to_u8_kernel(w.data(), nb, k, k_pad, wb.data(), &overflow); });
    const CUtensorMap map{wb.data(), k_pad, nb_pad, k_pad, BM_KC, (int)nblk};
    const long long ws_rep_stride = ntl * GV2_TILE * GV2_TILE;
    std::vector<double> ws((size_t)nb * ws_rep_stride, NAN);
    const EmuDim grid{128, (int)ntl, (int)nblocks};
    emu_launch_grid(grid, EmuDim{BM_THREADS, 1, 1}, [&] {
        boot_mma_kernel<false>(fixed.data(), (int)k_pad, (int)(k_pad / BM_KC), (int)nt, 0, map, (int)nb, (int)nblk, ws_rep_stride, quantum, ws.data());
    });
    emu_launch_grid(grid, EmuDim{BM_THREADS, 1, 1}, [&] {
        boot_mma_kernel<true>(fixed.data(), (int)k_pad, (int)(k_pad / BM_KC), (int)nt, 0, map, (int)nb, (int)nblk, ws_rep_stride, quantum, ws.data());
    });
    // ---- boot_mma_rowsums: the diagonal pairs
    std::vector<double> rows((size_t)nb * n, NAN);
    emu_launch_grid(EmuDim{(int)((n + 255) / 256), (int)nb, 1}, EmuDim{256, 1, 1},
                    [&] { diag_rowsums_kernel(ws.data(), ws_rep_stride, (int)nt, n, rows.data()); });
    const double head[2] = {quantum, (double)overflow};
    fwrite(head, 8, 2, stdout);
    fwrite(rows.data(), 8, rows.size(), stdout);
    fwrite(ws.data(), 8, ws.size(), stdout);
    return 0;
}
