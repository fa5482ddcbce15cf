// Shared declarations for libgravity_b200 (sm_100a).  Internal header: the public C ABI
// is include/gravity2_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gravity2_b200.h"

struct gv2_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    unsigned long long launches = 0;   // kernels launched through this context (bench: gpu_launches)
    // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember per context what has been raised
    unsigned smem_attr_done = 0;       // bit per launcher (GV2_ATTR_*)
    size_t lk_smem_attr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // optional CUDA-event timing of the hot kernels by group (gv2_timing_enable / gv2_timing_read_kind)
    bool timing = false;
    struct Timed { cudaEvent_t a, b; int kind; };
    std::vector<Timed> timed;
};

// kernel groups of the event timing: GV2_TIME_* in include/gravity2_b200.h

// RAII-free helper: t.begin() before the launches, t.end() after; no-ops unless timing is enabled
struct Gv2Timer {
    gv2_ctx* ctx;
    int kind;
    cudaEvent_t a = nullptr, b = nullptr;
    Gv2Timer(gv2_ctx* c, int k) : ctx(c), kind(k) {}
    void begin() {
        if (!ctx->timing) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
        cudaEventRecord(a, ctx->stream);
    }
    void end() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->timed.push_back({a, b, kind});
        a = b = nullptr;
    }
};

#define GV2_ATTR_GJ 1u
#define GV2_ATTR_FUSED 2u
#define GV2_ATTR_MMA 4u
#define GV2_ATTR_SPARSE 8u
#define GV2_ATTR_POPC 16u
#define GV2_ATTR_STREAM 32u

#define GV2_TILE 128          // genome-pair tile edge of the similarity kernels
#define GV2_BK 32             // PPHMM columns staged per pipeline stage

int gv2_fail(gv2_ctx* ctx, int code, const char* fmt, ...);

#define GV2_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return gv2_fail((ctx), GV2_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,  \
                            cudaGetErrorString(e__));                                        \
    } while (0)

#define GV2_LAUNCH_CHECK(ctx)                                                                \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return gv2_fail((ctx), GV2_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__,     \
                            cudaGetErrorString(e__));                                        \
    } while (0)

static inline int64_t gv2_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
