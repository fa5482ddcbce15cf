// One CUDA thread block on host threads: enough of the CUDA surface for gravity2_b200/csrc/linkage_kernels.cuh and
// popc_kernels.cuh, gj_kernels.cuh ... to compile
// with g++ and run as written -- one OS thread per CUDA thread, a barrier for __syncthreads, per-warp barriers for the
// warp reductions.  TEST INFRASTRUCTURE: lets the CPU test suite check a kernel's LOGIC (barrier placement, shared-memory
// hand-overs, tie order) against scipy where there is no GPU; the product path never sees this file.
#pragma once
#include <limits.h>
#include <linux/futex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __host__
#define __align__(x) __attribute__((aligned(x)))

struct EmuDim { int x = 0, y = 0, z = 0; };
static thread_local EmuDim threadIdx;
static thread_local int emu_linear_tid;               // threadIdx.x + blockDim.x * threadIdx.y: warps are 32 consecutive
static EmuDim blockIdx;
static EmuDim blockDim{1, 1, 1}, gridDim{1, 1, 1};
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline unsigned long long __double2ull_rn(double x) { return (unsigned long long)rint(x); }
// PRMT: byte i of the result is byte (selector nibble i) of the eight bytes {y : x}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
// IDP.4A, unsigned operands
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline void __stcs(T* p, T v) { *p = v; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }

// Counter + generation word; waiters sleep on the generation with a futex.  (A mutex + condition variable here made every one
// of up to 256 woken threads queue for the mutex again: three quarters of the CPU suite's time was spent in the kernel.)
class EmuBarrier {
  public:
    explicit EmuBarrier(int count = 1) : count_(count) {}
    void reset(int count) { count_ = count; waiting_.store(0); }
    void wait() {
        const unsigned gen = gen_.load(std::memory_order_acquire);
        if (waiting_.fetch_add(1, std::memory_order_acq_rel) + 1 == count_) {
            waiting_.store(0, std::memory_order_relaxed);         // before the release below: nobody re-enters until gen_ moves
            gen_.store(gen + 1, std::memory_order_release);
            syscall(SYS_futex, reinterpret_cast<unsigned*>(&gen_), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
        } else {
            while (gen_.load(std::memory_order_acquire) == gen)
                syscall(SYS_futex, reinterpret_cast<unsigned*>(&gen_), FUTEX_WAIT_PRIVATE, gen, nullptr, nullptr, 0);
        }
    }
  private:
    std::atomic<unsigned> gen_{0};
    std::atomic<int> waiting_{0};
    int count_;
};

static EmuBarrier emu_block_barrier;
struct EmuWarp { EmuBarrier bar{32}; unsigned slot[32]; unsigned long long wide[32]; };
static std::vector<EmuWarp>* emu_warps;

static inline void __syncthreads() { emu_block_barrier.wait(); }
static int emu_block_count;
static inline int __syncthreads_count(int pred) {
    emu_block_barrier.wait();
    if (emu_linear_tid == 0) emu_block_count = 0;
    emu_block_barrier.wait();
    if (pred) __atomic_fetch_add(&emu_block_count, 1, __ATOMIC_SEQ_CST);
    emu_block_barrier.wait();
    const int r = emu_block_count;
    emu_block_barrier.wait();
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    EmuWarp& w = (*emu_warps)[emu_linear_tid >> 5];
    w.slot[emu_linear_tid & 31] = v;
    w.bar.wait();
    unsigned m = w.slot[0];
    for (int l = 1; l < 32; ++l) m = w.slot[l] < m ? w.slot[l] : m;
    w.bar.wait();
    return m;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    EmuWarp& w = (*emu_warps)[emu_linear_tid >> 5];
    memcpy(&w.wide[emu_linear_tid & 31], &v, sizeof(T));
    w.bar.wait();
    T r;
    memcpy(&r, &w.wide[(emu_linear_tid & 31) ^ lane_mask], sizeof(T));
    w.bar.wait();
    return r;
}
static inline void __syncwarp() { (*emu_warps)[emu_linear_tid >> 5].bar.wait(); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    EmuWarp& w = (*emu_warps)[emu_linear_tid >> 5];
    memcpy(&w.wide[emu_linear_tid & 31], &v, sizeof(T));
    w.bar.wait();
    T r;
    memcpy(&r, &w.wide[src_lane & 31], sizeof(T));
    w.bar.wait();
    return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int delta, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    EmuWarp& w = (*emu_warps)[emu_linear_tid >> 5];
    const int lane = emu_linear_tid & 31;
    memcpy(&w.wide[lane], &v, sizeof(T));
    w.bar.wait();
    T r = v;
    if ((lane % width) >= delta) memcpy(&r, &w.wide[lane - delta], sizeof(T));
    w.bar.wait();
    return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    EmuWarp& w = (*emu_warps)[emu_linear_tid >> 5];
    w.slot[emu_linear_tid & 31] = v;
    w.bar.wait();
    unsigned s = 0;
    for (int l = 0; l < 32; ++l) s += w.slot[l];
    w.bar.wait();
    return s;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    EmuWarp& w = (*emu_warps)[emu_linear_tid >> 5];
    w.slot[emu_linear_tid & 31] = pred ? 1u : 0u;
    w.bar.wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= w.slot[l] << l;
    w.bar.wait();
    return m;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T>
static inline T __ldcs(const T* p) { return *p; }
static inline long long __double_as_longlong(double d) { long long b; memcpy(&b, &d, 8); return b; }
static inline double __longlong_as_double(long long b) { double d; memcpy(&d, &b, 8); return d; }
// separately rounded operations (this file is compiled with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double* p, double v) {              // (one thread per address in the kernels that use it)
    static std::mutex m;
    std::lock_guard<std::mutex> g(m);
    const double old = *p;
    *p = old + v;
    return old;
}
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

// cp.async, modelled at the two legal extremes (GV2_EMUL_CP_ASYNC, default "late"):
//   late   a copy lands when the issuing thread executes the wait_group that covers it (or never, if the kernel forgets to
//          wait): reading a stage before its wait -- a read-after-write hazard of the ring -- yields stale data and a wrong result;
//   early  a copy lands at issue: prefetching into a stage other threads still read -- a write-after-read hazard -- is a data
//          race ThreadSanitizer reports (tools/emul_tsan.py runs this mode).
struct EmuCopy { void* dst; const void* src; int bytes; };
static thread_local std::vector<std::vector<EmuCopy>> emu_cp_groups;   // committed groups, oldest first
static thread_local std::vector<EmuCopy> emu_cp_open;                  // copies issued since the last commit
static inline bool emu_cp_late() {
    static const bool late = [] { const char* e = getenv("GV2_EMUL_CP_ASYNC"); return !(e && e[0] == 'e'); }();
    return late;
}
static inline void emu_cp_async(void* dst, const void* src, int bytes) {
    if (emu_cp_late()) emu_cp_open.push_back(EmuCopy{dst, src, bytes});
    else memcpy(dst, src, (size_t)bytes);
}
static inline void emu_cp_commit() {
    if (!emu_cp_late()) return;
    emu_cp_groups.push_back(std::move(emu_cp_open));
    emu_cp_open.clear();
}
static inline void emu_cp_wait(int keep) {                             // cp.async.wait_group keep
    if (!emu_cp_late()) return;
    while ((int)emu_cp_groups.size() > keep) {
        for (const EmuCopy& c : emu_cp_groups.front()) memcpy(c.dst, c.src, (size_t)c.bytes);
        emu_cp_groups.erase(emu_cp_groups.begin());
    }
}
static inline void emu_cp_reset() { emu_cp_groups.clear(); emu_cp_open.clear(); }

// dynamic shared memory of the one emulated block
alignas(16) static unsigned char emu_dynamic_smem[256 * 1024];

// Worker threads are kept between blocks (a grid of a few hundred 256-thread blocks would otherwise spend its time in
// pthread_create): workers wait at `start`, run the block body as CUDA thread t, meet at `done`.
struct EmuPool {
    int size = 0;
    std::vector<std::thread> workers;
    std::function<void()> body;
    EmuBarrier start{1}, done{1};
    bool quit = false;
    void resize(int n) {
        stop();
        size = n;
        quit = false;
        start.reset(n + 1);
        done.reset(n + 1);
        for (int t = 0; t < n; ++t)
            workers.emplace_back([this, t] {
                for (;;) {
                    start.wait();
                    if (quit) return;
                    emu_linear_tid = t;
                    threadIdx.x = t % blockDim.x;
                    threadIdx.y = (t / blockDim.x) % blockDim.y;
                    threadIdx.z = t / (blockDim.x * blockDim.y);
                    emu_cp_reset();
                    body();
                    done.wait();
                }
            });
    }
    void stop() {
        if (workers.empty()) return;
        quit = true;
        start.wait();
        for (auto& w : workers) w.join();
        workers.clear();
        size = 0;
    }
    ~EmuPool() { stop(); }
};
static EmuPool emu_pool;
static std::vector<EmuWarp> emu_warp_store;

template <class F>
static void emu_launch_block(int threads, F&& body) {
    if (blockDim.x * blockDim.y * blockDim.z != threads) blockDim = EmuDim{threads, 1, 1};
    if (emu_pool.size != threads) {
        emu_pool.resize(threads);
        emu_warp_store = std::vector<EmuWarp>((size_t)(threads + 31) / 32);
        emu_warps = &emu_warp_store;
        emu_block_barrier.reset(threads);
    }
    emu_pool.body = [&] { body(); };
    emu_pool.start.wait();
    emu_pool.done.wait();
}

// A grid, one block after the other (blocks of a launch do not communicate).  A kernel may `return` from some threads before
// a later __syncthreads only if the whole block does (as on the GPU).
template <class F>
static void emu_launch_grid(EmuDim grid, EmuDim block, F&& body) {
    gridDim = EmuDim{grid.x, grid.y > 0 ? grid.y : 1, grid.z > 0 ? grid.z : 1};
    blockDim = block;
    for (int bz = 0; bz < gridDim.z; ++bz)
        for (int by = 0; by < gridDim.y; ++by)
            for (int bx = 0; bx < gridDim.x; ++bx) {
                blockIdx.x = bx;
                blockIdx.y = by;
                blockIdx.z = bz;
                emu_launch_block(block.x * block.y * block.z, body);
            }
}
template <class F>
static void emu_launch_grid(int blocks, EmuDim block, F&& body) { emu_launch_grid(EmuDim{blocks, 1, 1}, block, body); }
