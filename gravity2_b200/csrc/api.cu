// C ABI of libgravity_b200: context, host-pointer entry points, device-resident job.
// Declarations and per-entry reference citations: include/gravity2_b200.h.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "gom.h"

int gj_min_sums_launch(gv2_ctx* ctx, const float* tab, int64_t tab_stride, int64_t n_pad, int64_t k_pad,
                       int64_t tile_begin, int64_t tile_end, const float* w, int64_t nb, int ksplit,
                       double* ws);
int gj_fused_launch(gv2_ctx* ctx, const float* tab, int64_t tab_stride, int64_t n, int64_t n_pad, int64_t k, int64_t k_pad, int64_t nb,
                    const gv2_gj_component* comp_p, const gv2_gj_component* comp_g, int scheme, double power, int out_kind,
                    double* out, int64_t out_replicate_stride, int fixed, double p_scale);
int gom_quantize_launch(gv2_ctx* ctx, const double* tab64, int64_t n, int64_t g, int64_t nb, unsigned* dst, int64_t n_pad, int64_t g_pad,
                        double* rowsum, uint8_t* nanflag, unsigned long long* scratch);

int boot_mma_available();
int boot_mma_launch(gv2_ctx* ctx, const uint32_t* tab, double quantum, int64_t n_pad, int64_t k_pad, int64_t tile_begin,
                    int64_t tile_end, const int32_t* w, int64_t k, int64_t nb, void* w8_scratch, int* overflow_flag,
                    double* ws);
size_t boot_mma_scratch_bytes(int64_t k_pad, int64_t nb);
int boot_table_stats(gv2_ctx* ctx, const double* src, int64_t n, int64_t k, int64_t ld_src, double* max_value, int* has_negative);
int boot_mma_prepare_table(gv2_ctx* ctx, const double* src, int64_t n, int64_t k, int64_t ld_src, uint32_t* dst,
                           int64_t n_pad, int64_t k_pad, double max_value, double* quantum);
int boot_mma_rowsums(gv2_ctx* ctx, const double* ws, int64_t ntl, int64_t n, int64_t nb, double* out);
struct BootSparse;
int boot_sparse_has_lists(const BootSparse* s);
int boot_sparse_rowsums(gv2_ctx* ctx, const BootSparse* s, int64_t n, double* out);
int boot_sparse_build(gv2_ctx* ctx, const double* tab, int64_t n, int64_t k, int64_t ld, int64_t n_pad, int64_t k_pad,
                      double max_value, BootSparse** out);
int boot_sparse_build_lists(gv2_ctx* ctx, BootSparse* s, size_t budget);
void boot_sparse_free(BootSparse* s);
size_t boot_sparse_bytes(const BootSparse* s);
long long boot_sparse_nnz(const BootSparse* s);
unsigned long long boot_sparse_cells(const BootSparse* s);
size_t boot_sparse_scratch_bytes(int64_t k_pad, int64_t nb);
int boot_sparse_launch(gv2_ctx* ctx, const BootSparse* s, int64_t n_pad, int64_t k_pad, int64_t tile_begin,
                       int64_t tile_end, const int32_t* w, int64_t k, int64_t nb, void* wT_scratch, int* overflow_flag,
                       double* ws);

int lk_launch(gv2_ctx* ctx, int method, double* dev_D, int64_t d_stride, int64_t n, int64_t nb, double* dev_Zraw);
int lk_method(const char* method);
void lk_finalize_host(const double* raw, int64_t n, int method, double* Z);
void lk_newick_host(const double* Z, int64_t n, const char* const* names, std::string& out);

static thread_local std::string g_create_error;

int gv2_fail(gv2_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

extern "C" int gv2_version(void) { return 100; }

extern "C" int gv2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int create_impl(int device, void* stream, bool own, gv2_ctx** out) {
    if (!out) return gv2_fail(nullptr, GV2_ERR_ARG, "gv2_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return gv2_fail(nullptr, GV2_ERR_CUDA, "gv2_create: no CUDA device (%s); libgravity_b200 has no CPU fallback",
                        e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return gv2_fail(nullptr, GV2_ERR_ARG, "gv2_create: device %d of %d", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return gv2_fail(nullptr, GV2_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return gv2_fail(nullptr, GV2_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return gv2_fail(nullptr, GV2_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                        device, prop.major, prop.minor);
    gv2_ctx* c = new gv2_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (own) {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return gv2_fail(nullptr, GV2_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        c->own_stream = true;
    } else {
        c->stream = (cudaStream_t)stream;
    }
    *out = c;
    return GV2_OK;
}

extern "C" int gv2_create(int device, gv2_ctx** out) { return create_impl(device, nullptr, true, out); }
extern "C" int gv2_create_on_stream(int device, void* s, gv2_ctx** out) { return create_impl(device, s, false, out); }

extern "C" void gv2_destroy(gv2_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* gv2_last_error(const gv2_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int gv2_synchronize(gv2_ctx* ctx) {
    if (!ctx) return GV2_ERR_ARG;
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GV2_OK;
}

extern "C" uint64_t gv2_launch_count(const gv2_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int gv2_timing_enable(gv2_ctx* ctx, int on) {
    if (!ctx) return GV2_ERR_ARG;
    ctx->timing = on != 0;
    return GV2_OK;
}

// Summed duration and number of timed launch groups of one kind since the last read of that kind (synchronises).
extern "C" int gv2_timing_read_kind(gv2_ctx* ctx, int kind, double* total_ms, int64_t* launches) {
    if (!ctx || !total_ms || !launches) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_timing_read_kind: bad arguments");
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double t = 0.0;
    int64_t cnt = 0;
    std::vector<gv2_ctx::Timed> keep;
    for (auto& e : ctx->timed) {
        if (e.kind != kind) { keep.push_back(e); continue; }
        float ms = 0.f;
        GV2_CUDA(ctx, cudaEventElapsedTime(&ms, e.a, e.b));
        t += ms;
        ++cnt;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    ctx->timed.swap(keep);
    *total_ms = t;
    *launches = cnt;
    return GV2_OK;
}

extern "C" int gv2_timing_read(gv2_ctx* ctx, double* total_ms, int64_t* launches) {
    return gv2_timing_read_kind(ctx, GV2_TIME_PAIRSUMS, total_ms, launches);
}

// ---------------------------------------------------------- memory helpers for hosts without a CUDA binding
extern "C" int gv2_dev_alloc(gv2_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_dev_alloc: bad arguments");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_NOMEM, "cudaMalloc(%.3f GB): %s", bytes / 1e9, cudaGetErrorString(e));
    return GV2_OK;
}
extern "C" int gv2_dev_free(gv2_ctx* ctx, void* p) {
    if (!ctx) return GV2_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(p);
    return GV2_OK;
}
extern "C" int gv2_host_alloc(gv2_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_host_alloc: bad arguments");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_NOMEM, "cudaMallocHost(%.3f GB): %s", bytes / 1e9, cudaGetErrorString(e));
    return GV2_OK;
}
extern "C" int gv2_host_free(gv2_ctx* ctx, void* p) {
    if (!ctx) return GV2_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaFreeHost(p);
    return GV2_OK;
}
extern "C" int gv2_memcpy_h2d(gv2_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return GV2_ERR_ARG;
    GV2_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GV2_OK;
}
extern "C" int gv2_memcpy_d2h(gv2_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return GV2_ERR_ARG;
    GV2_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GV2_OK;
}

// ------------------------------------------------------------------------------- helpers
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    cudaError_t alloc(size_t b) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = b ? b : 1;
        return cudaMalloc(&p, bytes);
    }
    template <typename T> T* as() { return (T*)p; }
};

#define ALLOC(ctx, buf, bytes)                                                                   \
    do {                                                                                         \
        cudaError_t e__ = (buf).alloc(bytes);                                                    \
        if (e__ != cudaSuccess)                                                                  \
            return gv2_fail((ctx), GV2_ERR_NOMEM, "%s:%d device allocation of %.3f GB failed: %s", \
                            __FILE__, __LINE__, (double)(bytes) / 1e9, cudaGetErrorString(e__));  \
    } while (0)

// host matrix (rows x cols doubles, stride ld) -> device, contiguous (ld = cols)
static int upload_matrix(gv2_ctx* ctx, const double* h, int64_t rows, int64_t cols, int64_t ld, DevBuf& d) {
    ALLOC(ctx, d, (size_t)rows * cols * sizeof(double));
    if (rows == 0 || cols == 0) return GV2_OK;
    GV2_CUDA(ctx, cudaMemcpy2DAsync(d.p, cols * sizeof(double), h, ld * sizeof(double), cols * sizeof(double), rows,
                                    cudaMemcpyHostToDevice, ctx->stream));
    return GV2_OK;
}

static int pick_ksplit(gv2_ctx* ctx, int64_t ntl, int64_t nb, int64_t k_pad) {
    // enough CTAs for >= 2 waves at 2 CTAs/SM; never split below one fp32 slab
    const int64_t want = 4LL * ctx->sm_count;
    int64_t ks = (want + ntl * nb - 1) / std::max<int64_t>(1, ntl * nb);
    const int64_t max_ks = std::max<int64_t>(1, k_pad / 512);
    return (int)std::max<int64_t>(1, std::min<int64_t>(ks, std::min<int64_t>(max_ks, 64)));
}

// ------------------------------------------------------------------------- device job
// parts of the device ring in the dendrogram mode: one computing, the others being clustered (host delivery uses two)
#define GV2_RING_PARTS 4

struct gv2_job {
    gv2_ctx* ctx = nullptr;
    int64_t n = 0, p = 0, g = 0, n_pad = 0, p_pad = 0, g_pad = 0, ntiles = 0;
    int scheme = GV2_SCHEME_PG, out_kind = GV2_OUT_DISTANCE;
    double power = 1.0;
    bool need_p = false, need_g = false;
    DevBuf sig64, loc64;              // owned uploads (empty when the caller's tables are on the device)
    const double* d_sig = nullptr;
    const double* d_loc = nullptr;
    DevBuf sigf, sig_rowsum, sig_nan;
    DevBuf nan_cells, rep_nan;         // (row, column) of every NaN entry of the signature table; per-replicate row flags
    int64_t n_nan_cells = 0;
    DevBuf sigfix;                     // 32-bit fixed-point signature table for the tensor-core path
    double quantum = 0.0;
    double p_scale = 1.0;              // power of two <= 1 / (largest possible weighted row sum): the fused tail's PPHMM sums land in (0, 1]
    bool mma_ok = false;
    BootSparse* sparse = nullptr;      // CSR of the fixed-point table when the sparse replicate path is chosen
    GomDb* gom = nullptr;
    DevBuf gom_base;                   // [n][g] float64 base GOM table (lazy)
    bool have_gom_base = false;
    // workspaces, grown on demand
    DevBuf wsP, wsG, gomtab, gomf, gom_rowsum, gom_nan, gom_qmax, wf, w_rowsum, wbf16, mma_flag;
    DevBuf ones_w, sig_rowsum_fix;     // all-ones multiplicities and fixed-point row sums: the base matrix through the sparse path
    DevBuf cond[2];                    // condensed (upper-triangle) images of the two ring halves, for host delivery in that format
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready[GV2_RING_PARTS] = {}, ev_drained[GV2_RING_PARTS] = {};
    cudaStream_t lk_stream[GV2_RING_PARTS] = {};         // device UPGMA of the ring's parts, concurrent with the main stream
    size_t device_bytes() const {
        size_t b = sig64.bytes + loc64.bytes + sigf.bytes + sig_rowsum.bytes + sig_nan.bytes + gom_base.bytes +
                   wsP.bytes + wsG.bytes + gomtab.bytes + gomf.bytes + gom_rowsum.bytes + gom_nan.bytes + wf.bytes + w_rowsum.bytes;
        if (gom) b += gom->bytes;
        b += sigfix.bytes + wbf16.bytes + boot_sparse_bytes(sparse);
        return b;
    }
};

static int ensure(gv2_ctx* ctx, DevBuf& b, size_t bytes) {
    if (b.p && b.bytes >= bytes) return GV2_OK;
    ALLOC(ctx, b, bytes);
    return GV2_OK;
}

extern "C" void gv2_job_destroy(gv2_job* job) {
    if (!job) return;
    cudaSetDevice(job->ctx->device);
    cudaStreamSynchronize(job->ctx->stream);
    if (job->copy_stream) {
        cudaStreamSynchronize(job->copy_stream);
        for (int h = 0; h < GV2_RING_PARTS; ++h) { cudaEventDestroy(job->ev_ready[h]); cudaEventDestroy(job->ev_drained[h]); }
        cudaStreamDestroy(job->copy_stream);
    }
    for (int h = 0; h < GV2_RING_PARTS; ++h)
        if (job->lk_stream[h]) { cudaStreamSynchronize(job->lk_stream[h]); cudaStreamDestroy(job->lk_stream[h]); }
    gom_db_free(job->gom);
    boot_sparse_free(job->sparse);
    delete job;
}

extern "C" size_t gv2_job_device_bytes(const gv2_job* job) { return job ? job->device_bytes() : 0; }

extern "C" int gv2_mem_info(gv2_ctx* ctx, size_t* free_bytes, size_t* total_bytes) {
    if (!ctx || !free_bytes || !total_bytes) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_mem_info: bad arguments");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    GV2_CUDA(ctx, cudaMemGetInfo(free_bytes, total_bytes));
    return GV2_OK;
}

// Which replicate pair-sum path the job's cost model chose (GV2_PATH_*), and the number of co-present
// (genome pair incl. diagonal, PPHMM column) cells of its signature table (0 when no CSR was built).
extern "C" int gv2_job_replicate_path(const gv2_job* job) {
    if (!job || !job->need_p) return GV2_PATH_NONE;
    return job->sparse ? GV2_PATH_SPARSE : job->mma_ok ? GV2_PATH_MMA : GV2_PATH_FP32;
}
extern "C" uint64_t gv2_job_copresent_cells(const gv2_job* job) { return job ? boot_sparse_cells(job->sparse) : 0; }


// ---- multiplicity validation, done ONCE per call before anything is computed or delivered: the uint8 operands of the
// replicate kernels hold 0..255 (flag 1), the 32-bit low-limb sums of the sparse path need sum_c w_c <= 65536 (flag 2).
__global__ void weights_validate_kernel(const int* __restrict__ w, long long k, int* __restrict__ flag) {
    const int* row = w + (long long)blockIdx.x * k;
    long long s = 0;
    int bad = 0;
    for (long long c = threadIdx.x; c < k; c += blockDim.x) {
        const int v = row[c];
        bad |= (v < 0 || v > 255);
        s += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __shared__ long long red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    if (bad) atomicOr(flag, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        if (t > 65536) atomicOr(flag, 2);
    }
}

// ---- NaN entries of the signature table.  In the reference a NaN at (i, c) turns row i of a replicate's matrix into
// zeros only when column c was drawn (np.minimum propagates it through the resampled row, :167-168, cleaned at :204);
// the fixed-point / fp32 images hold 0 there, and a per-replicate row flag carries the NaN.
__global__ void nan_cells_kernel(const double* __restrict__ src, long long n, long long k, long long ld, int2* __restrict__ out,
                                 int cap, int* __restrict__ count) {
    const long long r = blockIdx.x;
    for (long long c = threadIdx.x; c < k; c += blockDim.x)
        if (isnan(src[r * ld + c])) {
            const int at = atomicAdd(count, 1);
            if (at < cap) out[at] = make_int2((int)r, (int)c);
        }
}
__global__ void rep_nan_kernel(const int2* __restrict__ cells, int ncells, const int* __restrict__ w, long long k, long long n,
                               long long nb, unsigned char* __restrict__ flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)ncells * nb) return;
    const long long b = i / ncells;
    const int2 cell = cells[i % ncells];
    if (w[b * k + cell.y] > 0) flags[b * n + cell.x] = 1;
}

static int job_collect_nan_cells(gv2_job* j) {
    gv2_ctx* ctx = j->ctx;
    j->n_nan_cells = 0;
    if (j->n == 0 || j->p == 0) return GV2_OK;
    std::vector<uint8_t> rowflag((size_t)j->n);
    GV2_CUDA(ctx, cudaMemcpyAsync(rowflag.data(), j->sig_nan.p, (size_t)j->n, cudaMemcpyDeviceToHost, ctx->stream));
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!std::any_of(rowflag.begin(), rowflag.end(), [](uint8_t f) { return f != 0; })) return GV2_OK;
    DevBuf cnt;
    ALLOC(ctx, cnt, sizeof(int));
    int have = 0, cap = 0;
    for (int pass = 0; pass < 2; ++pass) {          // count, then fill
        GV2_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, sizeof(int), ctx->stream));
        nan_cells_kernel<<<(unsigned)j->n, 256, 0, ctx->stream>>>(j->d_sig, j->n, j->p, j->p, j->nan_cells.as<int2>(), cap, cnt.as<int>());
        GV2_LAUNCH_CHECK(ctx);
        GV2_CUDA(ctx, cudaMemcpyAsync(&have, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pass == 0) { cap = have; ALLOC(ctx, j->nan_cells, (size_t)std::max(1, cap) * sizeof(int2)); }
    }
    j->n_nan_cells = have;
    return GV2_OK;
}

extern "C" int gv2_job_create(gv2_ctx* ctx, const double* sig, const double* loc, int tables_are_device,
                              int64_t n, int64_t p, const int32_t* gom_of_row, int64_t n_gom_rows, int64_t g,
                              int scheme, double power, int out_kind, gv2_job** out) {
    if (!ctx || !out || n < 0 || p < 0) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_job_create: bad arguments");
    *out = nullptr;
    if (scheme != GV2_SCHEME_P && scheme != GV2_SCHEME_G && scheme != GV2_SCHEME_PG)
        return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "gv2_job_create: scheme %d (jobs cover P, G, PG)", scheme);
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    gv2_job* j = new gv2_job();
    j->ctx = ctx; j->n = n; j->p = p; j->g = g; j->scheme = scheme; j->power = power; j->out_kind = out_kind;
    j->need_p = scheme != GV2_SCHEME_G;
    j->need_g = scheme != GV2_SCHEME_P;
    j->n_pad = gv2_round_up(std::max<int64_t>(n, 1), GV2_TILE);
    j->p_pad = gv2_round_up(std::max<int64_t>(p, 1), 128);   // 128: one tensor-core K chunk (4 x the 32-column block)
    j->g_pad = gv2_round_up(std::max<int64_t>(g, 1), GV2_BK);
    j->ntiles = gv2_num_tiles(n);
    int rc = GV2_OK;
    auto bail = [&](int code) { gv2_job_destroy(j); return code; };
    if (j->need_p) {
        if (!sig) return bail(gv2_fail(ctx, GV2_ERR_ARG, "gv2_job_create: scheme needs the signature table"));
        if (tables_are_device) j->d_sig = sig;
        else { if ((rc = upload_matrix(ctx, sig, n, p, p, j->sig64))) return bail(rc); j->d_sig = j->sig64.as<double>(); }
        if ((rc = ensure(ctx, j->sigf, (size_t)j->n_pad * j->p_pad * sizeof(float)))) return bail(rc);
        if ((rc = ensure(ctx, j->sig_rowsum, (size_t)n * sizeof(double)))) return bail(rc);
        if ((rc = ensure(ctx, j->sig_nan, (size_t)n))) return bail(rc);
        // `row[row < thr] = 0` runs on every call of the reference, thr == 0 included (similarity_matrix_constructor.py:139-140):
        // negative scores count as absent in the Jaccard sums.  The location side (d_loc) stays as it is.
        if ((rc = gv2_dev_prepare_table(ctx, j->d_sig, n, p, p, 0.0, 1, j->sigf.as<float>(), j->n_pad, j->p_pad,
                                        j->sig_rowsum.as<double>(), j->sig_nan.as<uint8_t>()))) return bail(rc);
        if ((rc = job_collect_nan_cells(j))) return bail(rc);
        {
            // Replicate path.  Two exact integer paths on a fixed-point image of the table (both need a non-negative
            // table): the CSR walk (work ~ co-present columns; real tables are ~99 % zeros) and the dense tensor-core
            // contraction; else the fp32 tile kernel.  GV2_BOOT_PATH = sparse | mma | fp32 overrides the cost model.
            double mx = 0.0;
            int neg = 0;
            if ((rc = boot_table_stats(ctx, j->d_sig, n, p, p, &mx, &neg))) return bail(rc);
            (void)neg;                                   // negative entries are zeroed by the threshold: x <= 0 -> 0 in both images
            if (isfinite(mx) && mx > 0.0) {              // sum_c w_c x_c <= max(x) * sum_c w_c, multiplicities are <= 255 each
                int e = 0;
                frexp(mx * 255.0 * (double)std::max<int64_t>(1, p), &e);
                j->p_scale = ldexp(1.0, -e);
            }
            const bool usable = isfinite(mx);
            const char* path = getenv("GV2_BOOT_PATH");
            const bool want_fp32 = path && !strcmp(path, "fp32");
            const bool can_mma = boot_mma_available() != 0;
            if (usable && !want_fp32) {
                if ((rc = boot_sparse_build(ctx, j->d_sig, n, p, p, j->n_pad, j->p_pad, mx, &j->sparse))) return bail(rc);
                // cost model (measured rates): the CSR walk does `cells` multiply-adds per
                // replicate at ~1/75 of the per-cell rate of the dense tensor-core contraction
                const double dense_cells = 0.5 * (double)j->n_pad * (double)(j->n_pad + 1) * (double)j->p_pad;
                bool sparse_wins = (double)boot_sparse_cells(j->sparse) * 75.0 < dense_cells;
                if (!can_mma) sparse_wins = true;                        // both exact; never fall to fp32 needlessly
                // a resample of P columns has multiplicities summing to P; the low limbs of the sparse path sum in 32 bits
                if (p > 65536) sparse_wins = false;
                if (path && !strcmp(path, "sparse")) sparse_wins = true;
                if (path && !strcmp(path, "mma")) sparse_wins = false;
                if (sparse_wins) {
                    // the match phase is replicate independent: write the pair lists once if they fit a tenth of the memory
                    size_t free_b = 0, total_b = 0;
                    cudaMemGetInfo(&free_b, &total_b);
                    if ((rc = boot_sparse_build_lists(ctx, j->sparse, free_b / 10))) return bail(rc);
                }
                if (!sparse_wins) {
                    boot_sparse_free(j->sparse);
                    j->sparse = nullptr;
                    if (can_mma) {
                        if ((rc = ensure(ctx, j->sigfix, (size_t)j->n_pad * j->p_pad * sizeof(uint32_t)))) return bail(rc);
                        if ((rc = boot_mma_prepare_table(ctx, j->d_sig, n, p, p, j->sigfix.as<uint32_t>(), j->n_pad, j->p_pad, mx,
                                                         &j->quantum))) return bail(rc);
                        j->mma_ok = true;
                    }
                }
            }
        }
    }
    if (j->need_g) {
        if (!loc || !gom_of_row || g <= 0) return bail(gv2_fail(ctx, GV2_ERR_ARG, "gv2_job_create: scheme needs loc, gom_of_row and g > 0"));
        if (tables_are_device) j->d_loc = loc;
        else if (loc == sig && j->need_p) j->d_loc = j->d_sig;          // PL1 quirk Q3: one upload
        else { if ((rc = upload_matrix(ctx, loc, n, p, p, j->loc64))) return bail(rc); j->d_loc = j->loc64.as<double>(); }
        if (n_gom_rows > n) return bail(gv2_fail(ctx, GV2_ERR_ARG, "gv2_job_create: n_gom_rows > n"));
        std::vector<int32_t> mem_ptr(g + 1, 0);
        std::vector<int64_t> mem_row;
        for (int64_t r = 0; r < n_gom_rows; ++r) {
            if (gom_of_row[r] >= g) return bail(gv2_fail(ctx, GV2_ERR_ARG, "gom_of_row[%lld] = %d >= g", (long long)r, gom_of_row[r]));
            if (gom_of_row[r] >= 0) mem_ptr[gom_of_row[r] + 1]++;
        }
        for (int64_t k = 0; k < g; ++k) mem_ptr[k + 1] += mem_ptr[k];
        mem_row.resize(mem_ptr[g]);
        std::vector<int32_t> fill(mem_ptr.begin(), mem_ptr.end() - 1);
        for (int64_t r = 0; r < n_gom_rows; ++r)
            if (gom_of_row[r] >= 0) mem_row[fill[gom_of_row[r]]++] = r;
        if ((rc = gom_db_build(ctx, j->d_loc, n, p, p, j->d_loc, p, mem_row.data(), mem_ptr.data(), g, 0, &j->gom))) return bail(rc);
        j->gom->weights_validated = true;          // replicates_impl / gv2_job_gom_table check the multiplicities before launching
    }
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = j;
    return GV2_OK;
}

// ---- sparse (COO) input: the reference pickles its tables also as scipy coo_matrix (app/src/ref_virus_annotator.py:395-400);
// real tables are ~99 % zeros, so the triplets are what crosses PCIe (16 bytes per hit instead of 8 bytes per cell) and the
// dense float64 image the kernels convert from is rebuilt in HBM.
__global__ void coo_scatter_kernel(const int* __restrict__ rows, const int* __restrict__ cols, const double* __restrict__ vals,
                                   long long nnz, long long n, long long p, double* __restrict__ dense, int* __restrict__ bad) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const long long r = rows[k], c = cols[k];
    if (r < 0 || r >= n || c < 0 || c >= p) { atomicExch(bad, 1); return; }
    dense[r * p + c] = vals[k];
}

static int coo_to_dense(gv2_ctx* ctx, const int32_t* rows, const int32_t* cols, const double* vals, int64_t nnz, int64_t n, int64_t p,
                        DevBuf& dense) {
    ALLOC(ctx, dense, (size_t)std::max<int64_t>(1, n * p) * sizeof(double));
    GV2_CUDA(ctx, cudaMemsetAsync(dense.p, 0, dense.bytes, ctx->stream));
    if (nnz == 0) return GV2_OK;
    DevBuf r, c, v, flag;
    ALLOC(ctx, r, (size_t)nnz * sizeof(int32_t));
    ALLOC(ctx, c, (size_t)nnz * sizeof(int32_t));
    ALLOC(ctx, v, (size_t)nnz * sizeof(double));
    ALLOC(ctx, flag, sizeof(int));
    GV2_CUDA(ctx, cudaMemsetAsync(flag.p, 0, sizeof(int), ctx->stream));
    GV2_CUDA(ctx, cudaMemcpyAsync(r.p, rows, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    GV2_CUDA(ctx, cudaMemcpyAsync(c.p, cols, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    GV2_CUDA(ctx, cudaMemcpyAsync(v.p, vals, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    coo_scatter_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, ctx->stream>>>(r.as<int>(), c.as<int>(), v.as<double>(), nnz, n, p,
                                                                               dense.as<double>(), flag.as<int>());
    GV2_LAUNCH_CHECK(ctx);
    int bad = 0;
    GV2_CUDA(ctx, cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) return gv2_fail(ctx, GV2_ERR_ARG, "COO input: an entry lies outside the %lld x %lld table", (long long)n, (long long)p);
    return GV2_OK;
}

extern "C" int gv2_job_create_coo(gv2_ctx* ctx, int64_t n, int64_t p, const int32_t* sig_rows, const int32_t* sig_cols,
                                  const double* sig_vals, int64_t sig_nnz, const int32_t* loc_rows, const int32_t* loc_cols,
                                  const double* loc_vals, int64_t loc_nnz, const int32_t* gom_of_row, int64_t n_gom_rows, int64_t g,
                                  int scheme, double power, int out_kind, gv2_job** out) {
    if (!ctx || !out || n < 0 || p < 0 || sig_nnz < 0 || loc_nnz < 0 || (sig_nnz && (!sig_rows || !sig_cols || !sig_vals)) ||
        (loc_nnz && (!loc_rows || !loc_cols || !loc_vals)))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_job_create_coo: bad arguments");
    *out = nullptr;
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf sig, loc;
    int rc;
    if ((rc = coo_to_dense(ctx, sig_rows, sig_cols, sig_vals, sig_nnz, n, p, sig))) return rc;
    const bool own_loc = loc_rows != nullptr;                     // NULL triplets: the signature table doubles as location table (PL1)
    if (own_loc && (rc = coo_to_dense(ctx, loc_rows, loc_cols, loc_vals, loc_nnz, n, p, loc))) return rc;
    gv2_job* j = nullptr;
    rc = gv2_job_create(ctx, sig.as<double>(), scheme == GV2_SCHEME_P ? nullptr : (own_loc ? loc.as<double>() : sig.as<double>()), 1, n, p,
                        gom_of_row, n_gom_rows, g, scheme, power, out_kind, &j);
    if (rc) return rc;
    j->sig64.p = sig.p; j->sig64.bytes = sig.bytes; sig.p = nullptr;      // the job owns the dense images now
    if (own_loc) { j->loc64.p = loc.p; j->loc64.bytes = loc.bytes; loc.p = nullptr; }
    *out = j;
    return GV2_OK;
}

// GOM table (fp64 [nb][n][g]) -> padded tables [nb][n_pad][g_pad] + row sums + NaN flags: unsigned fixed point for the
// replicates' fused tail (fixed != 0, one launch for all nb tables), fp32 for the base matrix's tile kernel
static int job_prepare_gom(gv2_job* j, const double* tab64, int64_t nb, int fixed) {
    gv2_ctx* ctx = j->ctx;
    int rc;
    if ((rc = ensure(ctx, j->gomf, (size_t)nb * j->n_pad * j->g_pad * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, j->gom_rowsum, (size_t)nb * j->n * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, j->gom_nan, (size_t)nb * j->n))) return rc;
    if (fixed) {
        if ((rc = ensure(ctx, j->gom_qmax, sizeof(unsigned long long)))) return rc;
        return gom_quantize_launch(ctx, tab64, j->n, j->g, nb, j->gomf.as<unsigned>(), j->n_pad, j->g_pad, j->gom_rowsum.as<double>(),
                                   j->gom_nan.as<uint8_t>(), j->gom_qmax.as<unsigned long long>());
    }
    for (int64_t b = 0; b < nb; ++b)
        if ((rc = gv2_dev_prepare_table(ctx, tab64 + b * j->n * j->g, j->n, j->g, j->g, 0.0, 0,
                                        j->gomf.as<float>() + b * j->n_pad * j->g_pad, j->n_pad, j->g_pad,
                                        j->gom_rowsum.as<double>() + b * j->n, j->gom_nan.as<uint8_t>() + b * j->n)))
            return rc;
    return GV2_OK;
}

extern "C" int gv2_job_gom_table(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_out) {
    if (!job || !dev_out || nb <= 0) return gv2_fail(job ? job->ctx : nullptr, GV2_ERR_ARG, "gv2_job_gom_table: bad arguments");
    if (!job->gom) return gv2_fail(job->ctx, GV2_ERR_ARG, "gv2_job_gom_table: job has no GOM database");
    if (!dev_weights && nb != 1) return gv2_fail(job->ctx, GV2_ERR_ARG, "gv2_job_gom_table: nb must be 1 without weights");
    gv2_ctx* ctx = job->ctx;
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (dev_weights) {
        if (!job->mma_flag.p) ALLOC(ctx, job->mma_flag, sizeof(int));
        int flag = 0;
        GV2_CUDA(ctx, cudaMemsetAsync(job->mma_flag.p, 0, sizeof(int), ctx->stream));
        weights_validate_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(dev_weights, job->p, job->mma_flag.as<int>());
        GV2_LAUNCH_CHECK(ctx);
        GV2_CUDA(ctx, cudaMemcpyAsync(&flag, job->mma_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (flag & 1) return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "a column multiplicity is outside 0..255 (uint8 weights of the GOM kernels)");
    }
    return gom_db_score(ctx, job->gom, dev_weights, nb, dev_out);
}

extern "C" int gv2_job_gom_stats(const gv2_job* job, int64_t* stats) {
    if (!job || !stats) return GV2_ERR_ARG;
    for (int i = 0; i < 8; ++i) stats[i] = 0;
    if (const GomDb* db = job->gom) {
        stats[0] = db->npt; stats[1] = db->sum_ng2; stats[2] = db->nnz; stats[3] = db->n_heavy;
        stats[4] = db->n_isect; stats[5] = db->n_mentries; stats[6] = db->max_ng; stats[7] = db->max_t;
    }
    return GV2_OK;
}

static int job_base_impl(gv2_job* job, int64_t tile_begin, int64_t tile_end, int packed, double* dev_out);

extern "C" int gv2_job_base(gv2_job* job, int64_t tile_begin, int64_t tile_end, int packed, double* dev_out) {
    if (!job || !dev_out || tile_begin < 0 || tile_end < tile_begin || tile_end > job->ntiles)
        return gv2_fail(job ? job->ctx : nullptr, GV2_ERR_ARG, "gv2_job_base: bad arguments");
    GV2_CUDA(job->ctx, cudaSetDevice(job->ctx->device));
    Gv2Timer timer(job->ctx, GV2_TIME_BASE);
    timer.begin();
    const int rc = job_base_impl(job, tile_begin, tile_end, packed, dev_out);
    timer.end();
    return rc;
}

static int job_base_impl(gv2_job* job, int64_t tile_begin, int64_t tile_end, int packed, double* dev_out) {
    gv2_ctx* ctx = job->ctx;
    const int64_t ntl = tile_end - tile_begin;
    if (ntl == 0) return GV2_OK;
    const size_t tile_elems = (size_t)GV2_TILE * GV2_TILE;
    int rc;
    gv2_gj_component cp{}, cg{};
    if (job->need_p && boot_sparse_has_lists(job->sparse)) {
        // sparse tables: the un-resampled matrix is the all-ones replicate of the exact list path (work ~ co-present cells,
        // not pairs x P); row sums of the same fixed-point image, so that duplicate genomes come out at exactly 1
        if (!job->ones_w.p) {
            std::vector<int32_t> ones((size_t)std::max<int64_t>(1, job->p), 1);
            if ((rc = ensure(ctx, job->ones_w, ones.size() * sizeof(int32_t)))) return rc;
            GV2_CUDA(ctx, cudaMemcpyAsync(job->ones_w.p, ones.data(), ones.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
            GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if ((rc = ensure(ctx, job->sig_rowsum_fix, (size_t)std::max<int64_t>(1, job->n) * sizeof(double)))) return rc;
            if ((rc = boot_sparse_rowsums(ctx, job->sparse, job->n, job->sig_rowsum_fix.as<double>()))) return rc;
        }
        if ((rc = ensure(ctx, job->wsP, (size_t)ntl * tile_elems * sizeof(double)))) return rc;
        if ((rc = ensure(ctx, job->wbf16, boot_sparse_scratch_bytes(job->p_pad, 1)))) return rc;
        if (!job->mma_flag.p) ALLOC(ctx, job->mma_flag, sizeof(int));
        if ((rc = boot_sparse_launch(ctx, job->sparse, job->n_pad, job->p_pad, tile_begin, tile_end, job->ones_w.as<int32_t>(), job->p, 1,
                                     job->wbf16.p, job->mma_flag.as<int>(), job->wsP.as<double>()))) return rc;
        cp.min_sums = job->wsP.as<double>(); cp.rowsums = job->sig_rowsum_fix.as<double>(); cp.nanflags = job->sig_nan.as<uint8_t>();
        cp.ksplit = 1;
    } else if (job->need_p) {
        const int ks = pick_ksplit(ctx, ntl, 1, job->p_pad);
        if ((rc = ensure(ctx, job->wsP, (size_t)ks * ntl * tile_elems * sizeof(double)))) return rc;
        if ((rc = gj_min_sums_launch(ctx, job->sigf.as<float>(), 0, job->n_pad, job->p_pad, tile_begin, tile_end, nullptr, 1, ks,
                                     job->wsP.as<double>()))) return rc;
        cp.min_sums = job->wsP.as<double>(); cp.rowsums = job->sig_rowsum.as<double>(); cp.nanflags = job->sig_nan.as<uint8_t>();
        cp.ksplit = ks;
    }
    if (job->need_g) {
        if (!job->have_gom_base) {
            if ((rc = ensure(ctx, job->gom_base, (size_t)job->n * job->g * sizeof(double)))) return rc;
            if ((rc = gom_db_score(ctx, job->gom, nullptr, 1, job->gom_base.as<double>()))) return rc;
            job->have_gom_base = true;
        }
        if ((rc = job_prepare_gom(job, job->gom_base.as<double>(), 1, 0))) return rc;
        if ((rc = ensure(ctx, job->wsG, (size_t)ntl * tile_elems * sizeof(double)))) return rc;
        if ((rc = gj_min_sums_launch(ctx, job->gomf.as<float>(), 0, job->n_pad, job->g_pad, tile_begin, tile_end, nullptr, 1, 1,
                                     job->wsG.as<double>()))) return rc;
        cg.min_sums = job->wsG.as<double>(); cg.rowsums = job->gom_rowsum.as<double>(); cg.nanflags = job->gom_nan.as<uint8_t>();
        cg.ksplit = 1;
    }
    return gv2_dev_gj_epilogue(ctx, job->n, tile_begin, tile_end, 1, &cp, &cg, nullptr, job->scheme, job->power,
                               job->out_kind, packed, dev_out, 0);
}

// Replicates [0, nb): replicate b is written to ring slot (b % ring_len) of dev_ring [ring_len][n][n] and, when
// host_out != NULL, copied to host_out[b] ([nb][n][n], pinned or pageable) on a second stream while the next
// slots are being computed.  ring_len == nb keeps every replicate resident (gv2_job_replicates).
//
// Two nested batches: a SUPER-CHUNK of replicates shares one pass of the expensive per-table work (tensor-core
// pair sums for up to 128 replicates per CTA, GOM scoring 32 replicates per warp); inside it, sub-chunks of at
// most ring_len / 2 replicates run the cheap per-replicate tail (GOM Jaccard sums, epilogue) into one half of
// the ring while the other half drains to the host.
// Device-side consumer of a sub-chunk of replicate matrices: UPGMA on the ring slots, linkage rows to the host,
// Newick strings to the caller (gv2_job_replicates_newick).  Two buffers: the host finishes sub-chunk k - 1 (sort,
// relabel, print) while the GPU works on sub-chunk k.
struct TreeSink {
    gv2_ctx* ctx = nullptr;
    int64_t n = 0, cap = 0;
    int method = 0;
    const char* const* names = nullptr;
    gv2_newick_cb cb = nullptr;
    void* user = nullptr;
    double* zdev[GV2_RING_PARTS] = {};
    double* zhost[GV2_RING_PARTS] = {};
    cudaEvent_t ev[GV2_RING_PARTS] = {};
    int64_t pend_b0[GV2_RING_PARTS] = {}, pend_cnt[GV2_RING_PARTS] = {};
    int cur = 0, nparts = 2;
    std::vector<std::string> texts;
    ~TreeSink() {
        for (int h = 0; h < GV2_RING_PARTS; ++h) { cudaFree(zdev[h]); cudaFreeHost(zhost[h]); if (ev[h]) cudaEventDestroy(ev[h]); }
    }
    int init(gv2_ctx* c, int64_t n_, int64_t cap_, int nparts_) {
        ctx = c; n = n_; cap = cap_; nparts = nparts_;
        const size_t zb = (size_t)cap * std::max<int64_t>(1, n - 1) * 4 * sizeof(double);
        for (int h = 0; h < nparts; ++h) {
            GV2_CUDA(ctx, cudaMalloc((void**)&zdev[h], zb));
            GV2_CUDA(ctx, cudaMallocHost((void**)&zhost[h], zb));
            GV2_CUDA(ctx, cudaEventCreateWithFlags(&ev[h], cudaEventDisableTiming));
        }
        return GV2_OK;
    }
    // matrices of replicates [b0, b0 + cnt) are in dev[0 .. cnt) (stride n*n): cluster them, fetch the rows
    int enqueue(double* dev, int64_t b0, int64_t cnt) {
        int rc = lk_launch(ctx, method, dev, n * n, n, cnt, zdev[cur]);
        if (rc) return rc;
        if (n > 1)
            GV2_CUDA(ctx, cudaMemcpyAsync(zhost[cur], zdev[cur], (size_t)cnt * (n - 1) * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        GV2_CUDA(ctx, cudaEventRecord(ev[cur], ctx->stream));
        pend_b0[cur] = b0; pend_cnt[cur] = cnt;
        cur = (cur + 1) % nparts;                    // now the OLDEST pending part: the one the caller drains next
        return GV2_OK;
    }
    int drain(int h) {
        if (pend_cnt[h] == 0) return GV2_OK;
        cudaError_t e = cudaEventSynchronize(ev[h]);
        if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "linkage rows: %s", cudaGetErrorString(e));
        // sort + relabel + print the sub-chunk's trees on a few host threads (~5 ms of work per 10 000-leaf tree), then
        // hand them over in replicate order on the calling thread
        const int64_t cnt = pend_cnt[h];
        texts.resize((size_t)cnt);
        const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(cnt, 8), (int64_t)std::thread::hardware_concurrency()));
        auto work = [&](int t) {
            std::vector<double> Zl((size_t)std::max<int64_t>(1, n - 1) * 4);
            for (int64_t k = t; k < cnt; k += nthreads) {
                if (n > 1) lk_finalize_host(zhost[h] + (size_t)k * (n - 1) * 4, n, method, Zl.data());
                lk_newick_host(Zl.data(), n, names, texts[(size_t)k]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        pend_cnt[h] = 0;
        for (int64_t k = 0; k < cnt; ++k)
            if (cb(pend_b0[h] + k, texts[(size_t)k].c_str(), texts[(size_t)k].size(), user) != 0)
                return gv2_fail(ctx, GV2_ERR_ARG, "tree callback asked to stop at replicate %lld", (long long)(pend_b0[h] + k));
        return GV2_OK;
    }
};

// scipy's condensed form of a symmetric matrix (squareform order: row i contributes D[i][i+1 .. n-1]), what
// dist_mat_to_tree.py:8-9 hands to linkage(): half the bytes of the square matrix on the way to the host.
// grid = (n - 1 rows, matrices)
__global__ void condense_kernel(const double* __restrict__ ring, long long n, double* __restrict__ out) {
    const long long i = blockIdx.x, m = n * (n - 1) / 2;
    const double* src = ring + (size_t)blockIdx.y * n * n + i * n + i + 1;
    double* dst = out + (size_t)blockIdx.y * m + i * n - i * (i + 1) / 2;
    for (long long j = threadIdx.x; j < n - 1 - i; j += blockDim.x) dst[j] = src[j];
}

static int replicates_impl(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring, int64_t ring_len,
                           double* host_out, gv2_replicate_cb cb_fn, void* cb_user, TreeSink* sink, gv2_replicate_cb dev_cb = nullptr,
                           bool condensed = false) {
    if (!job || !dev_weights || !dev_ring || nb < 0 || ring_len <= 0 || (cb_fn && !host_out))
        return gv2_fail(job ? job->ctx : nullptr, GV2_ERR_ARG, "gv2_job_replicates_stream: bad arguments");
    // host delivery ping-pongs two halves of the ring (one computing, one draining): a single slot cannot do that
    if (host_out && ring_len < 2 && nb > 1)
        return gv2_fail(job->ctx, GV2_ERR_ARG, "gv2_job_replicates_stream: host delivery of more than one replicate needs ring_len >= 2");
    gv2_ctx* ctx = job->ctx;
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nb == 0 || job->n == 0) return GV2_OK;
    {   // multiplicities are checked before the first launch, so nothing wrong is ever delivered; the flag starts clean per call
        if (!job->mma_flag.p) ALLOC(ctx, job->mma_flag, sizeof(int));
        int flag = 0;
        GV2_CUDA(ctx, cudaMemsetAsync(job->mma_flag.p, 0, sizeof(int), ctx->stream));
        weights_validate_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(dev_weights, job->p, job->mma_flag.as<int>());
        GV2_LAUNCH_CHECK(ctx);
        GV2_CUDA(ctx, cudaMemcpyAsync(&flag, job->mma_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (flag & 1)
            return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "a column multiplicity is outside 0..255 (uint8 operands of the replicate kernels)");
        if ((flag & 2) && job->sparse)
            return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "a replicate's column multiplicities sum to more than 65536 (limit of the exact sparse path); set GV2_BOOT_PATH=fp32");
    }
    const int64_t n = job->n, p = job->p, g = job->g, ntl = job->ntiles;
    const size_t tile_elems = (size_t)GV2_TILE * GV2_TILE;
    const int64_t hstride = condensed ? n * (n - 1) / 2 : n * n;      // doubles per delivered replicate
    ring_len = std::min(ring_len, nb);
    // sub-chunk: half the ring when results stream to the host (ping-pong), else the whole ring.  Dendrogram mode: the ring is
    // cut in (up to) four parts, each clustered on its own stream while the main stream computes into the free one -- the
    // clustering is bound by DRAM operations and takes ~3x as long as the similarity kernels of the same replicates, so two or
    // three clusterings are in flight at any time and the similarity kernels (ALU bound) run beside them
    const bool halves = ring_len >= 2 && ((!sink && host_out) || (sink && !dev_cb));
    const int nparts = sink ? sink->nparts : 2;
    const int64_t sub = halves ? ring_len / nparts : ring_len;
    // super-chunk: bounded by the fp64 pair-sum workspace of the signature table (+ the GOM tables)
    const size_t per_rep = (job->need_p ? ntl * tile_elems * sizeof(double) + (size_t)job->p_pad : 0) +
                           (job->need_g ? (size_t)job->n_pad * job->g_pad * 4 + (size_t)n * g * 8 : 0) + (size_t)n * 8 * 2;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t have = job->wsP.bytes + job->gomf.bytes + job->gomtab.bytes + job->wbf16.bytes;
    const size_t sub_bytes = 0;     // the fused tail keeps the GOM pair sums in registers: no per-sub-chunk workspace
    const size_t avail = (size_t)((free_b + have) * 0.85);
    int64_t super = (int64_t)((avail > sub_bytes ? avail - sub_bytes : 0) / std::max<size_t>(1, per_rep));
    super = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(super, nb), 1024));
    if (super < nb) {                                       // several super-chunks: whole tensor-core / GOM lane blocks
        if (super >= 256) super = super / 256 * 256;
        else if (super >= 128) super = super / 128 * 128;
        else if (super >= 32) super = super / 32 * 32;
    }
    // dendrogram mode: whole ring parts per super-chunk -- a remainder of a few matrices would hold a part of the ring (and a
    // clustering launch of a few CTAs, as long as a full one) for itself
    if (sink && halves && super < nb && super > sub) super = super / sub * sub;
    if (sink && !job->lk_stream[0])
        for (int h = 0; h < GV2_RING_PARTS; ++h) GV2_CUDA(ctx, cudaStreamCreateWithFlags(&job->lk_stream[h], cudaStreamNonBlocking));
    bool lk_pending[GV2_RING_PARTS] = {};
    if (!job->copy_stream) {
        GV2_CUDA(ctx, cudaStreamCreateWithFlags(&job->copy_stream, cudaStreamNonBlocking));
        for (int h = 0; h < GV2_RING_PARTS; ++h) {
            GV2_CUDA(ctx, cudaEventCreateWithFlags(&job->ev_ready[h], cudaEventDisableTiming));
            GV2_CUDA(ctx, cudaEventCreateWithFlags(&job->ev_drained[h], cudaEventDisableTiming));
        }
    }
    bool drained_pending[2] = {false, false};
    int64_t pend_b0[2] = {0, 0}, pend_cb[2] = {0, 0}, pend_slot[2] = {0, 0};   // sub-chunk whose copy is in flight, per half
    int rc;
    bool used_mma = false;
    int64_t slot = 0;     // next ring slot
    int half = 0;
    // callback mode: once the copy of a sub-chunk has landed in the host ring, hand its replicates to the consumer.
    // Called AFTER the next sub-chunk has been enqueued, so the GPU computes while the consumer works.
    auto deliver = [&](int h) -> int {
        if (!cb_fn || !drained_pending[h]) return GV2_OK;
        cudaError_t e = cudaEventSynchronize(job->ev_drained[h]);
        if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "replicate copy: %s", cudaGetErrorString(e));
        for (int64_t k = 0; k < pend_cb[h]; ++k)
            if (cb_fn(pend_b0[h] + k, host_out + (pend_slot[h] + k) * hstride, cb_user) != 0)
                return gv2_fail(ctx, GV2_ERR_ARG, "replicate callback asked to stop at replicate %lld", (long long)(pend_b0[h] + k));
        drained_pending[h] = false;
        return GV2_OK;
    };
    for (int64_t s0 = 0; s0 < nb; s0 += super) {
        const int64_t sb = std::min<int64_t>(super, nb - s0);
        const int32_t* w = dev_weights + s0 * p;
        gv2_gj_component cp{}, cg{};
        if (job->need_p) {
            if ((rc = ensure(ctx, job->w_rowsum, (size_t)sb * n * sizeof(double)))) return rc;
            if ((rc = ensure(ctx, job->wsP, (size_t)sb * ntl * tile_elems * sizeof(double)))) return rc;
            if (job->sparse || (job->mma_ok && sb >= 8)) {
                // exact integer paths on the fixed-point table: CSR walk (sparse tables) or tensor-core contraction
                const size_t scratch = job->sparse ? boot_sparse_scratch_bytes(job->p_pad, sb) : boot_mma_scratch_bytes(job->p_pad, sb);
                if ((rc = ensure(ctx, job->wbf16, scratch))) return rc;
                if (job->sparse)
                    rc = boot_sparse_launch(ctx, job->sparse, job->n_pad, job->p_pad, 0, ntl, w, p, sb, job->wbf16.p,
                                            job->mma_flag.as<int>(), job->wsP.as<double>());
                else
                    rc = boot_mma_launch(ctx, job->sigfix.as<uint32_t>(), job->quantum, job->n_pad, job->p_pad, 0, ntl, w, p, sb,
                                         job->wbf16.p, job->mma_flag.as<int>(), job->wsP.as<double>());
                if (rc) return rc;
                if ((rc = boot_mma_rowsums(ctx, job->wsP.as<double>(), ntl, n, sb, job->w_rowsum.as<double>()))) return rc;
                used_mma = true;
            } else {
                if ((rc = ensure(ctx, job->wf, (size_t)sb * job->p_pad * sizeof(float)))) return rc;
                if ((rc = gv2_dev_weights_to_float(ctx, w, sb, p, job->p_pad, job->wf.as<float>()))) return rc;
                if ((rc = gv2_dev_weighted_rowsums(ctx, job->sigf.as<float>(), n, job->p_pad, job->wf.as<float>(), sb,
                                                   job->w_rowsum.as<double>()))) return rc;
                if ((rc = gj_min_sums_launch(ctx, job->sigf.as<float>(), 0, job->n_pad, job->p_pad, 0, ntl, job->wf.as<float>(), sb, 1,
                                             job->wsP.as<double>()))) return rc;
            }
        }
        if (job->need_p && job->n_nan_cells > 0) {      // NaN entries zero a row only in the replicates that drew their column
            if ((rc = ensure(ctx, job->rep_nan, (size_t)sb * n))) return rc;
            GV2_CUDA(ctx, cudaMemsetAsync(job->rep_nan.p, 0, (size_t)sb * n, ctx->stream));
            const long long work = (long long)job->n_nan_cells * sb;
            rep_nan_kernel<<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>(job->nan_cells.as<int2>(), (int)job->n_nan_cells, w, p, n, sb,
                                                                                   job->rep_nan.as<unsigned char>());
            GV2_LAUNCH_CHECK(ctx);
        }
        if (job->need_g) {
            if ((rc = ensure(ctx, job->gomtab, (size_t)sb * n * g * sizeof(double)))) return rc;
            if ((rc = gom_db_score(ctx, job->gom, w, sb, job->gomtab.as<double>()))) return rc;
            if ((rc = job_prepare_gom(job, job->gomtab.as<double>(), sb, 1))) return rc;
        }
        for (int64_t c0 = 0; c0 < sb;) {
            int64_t cb = std::min<int64_t>(sub, sb - c0);
            if (!host_out && !sink) {                                      // device-only: replicate b lives in slot b % ring_len
                slot = (s0 + c0) % ring_len;
                cb = std::min<int64_t>(cb, ring_len - slot);
            } else {
                if (slot + cb > ring_len) slot = 0;                        // sub-chunks never wrap inside the ring
                if (halves) { slot = half * sub; }
            }
            if (sink && lk_pending[half]) {                                // the clustering that last read this half is done
                GV2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, job->ev_drained[half], 0));
                lk_pending[half] = false;
            }
            if (cb_fn) {
                if ((rc = deliver(half))) return rc;                       // consumer is done with these host slots too
            } else if (drained_pending[half]) {                            // the copy that last read these slots is done
                GV2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, job->ev_drained[half], 0));
                drained_pending[half] = false;
            }
            if (job->need_p) {
                cp.min_sums = job->wsP.as<double>() + (size_t)c0 * ntl * tile_elems;
                cp.rowsums = job->w_rowsum.as<double>() + c0 * n;
                cp.nanflags = job->n_nan_cells > 0 ? job->rep_nan.as<uint8_t>() + c0 * n : nullptr;
                cp.replicate_stride_nan = job->n_nan_cells > 0 ? n : 0;
                cp.ksplit = 1; cp.replicate_stride_sums = (int64_t)(ntl * tile_elems); cp.replicate_stride_rows = n;
            }
            if (job->need_g) {
                // GOM Jaccard sums and the epilogue in one kernel: the GOM pair sums never go through HBM
                cg.rowsums = job->gom_rowsum.as<double>() + c0 * n; cg.nanflags = job->gom_nan.as<uint8_t>() + c0 * n;
                cg.ksplit = 1; cg.replicate_stride_rows = n; cg.replicate_stride_nan = n;
                if ((rc = gj_fused_launch(ctx, job->gomf.as<float>() + (size_t)c0 * job->n_pad * job->g_pad, job->n_pad * job->g_pad, n,
                                          job->n_pad, g, job->g_pad, cb, job->need_p ? &cp : nullptr, &cg, job->scheme, job->power,
                                          job->out_kind, dev_ring + slot * n * n, n * n, 1, job->p_scale))) return rc;
            } else {
                if ((rc = gv2_dev_gj_epilogue(ctx, n, 0, ntl, cb, &cp, &cg, nullptr, job->scheme, job->power, job->out_kind, 0,
                                              dev_ring + slot * n * n, n * n))) return rc;
            }
            if (dev_cb) {   // device consumer: the sub-chunk's matrices are complete in the ring; hand over the device pointers
                GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                for (int64_t k = 0; k < cb; ++k)
                    if (dev_cb(s0 + c0 + k, dev_ring + (slot + k) * n * n, cb_user) != 0)
                        return gv2_fail(ctx, GV2_ERR_ARG, "replicate callback asked to stop at replicate %lld", (long long)(s0 + c0 + k));
            }
            if (sink) {     // cluster this sub-chunk on the device, on the half's own stream: the main stream goes on with the
                            // next sub-chunk (and the next super-chunk's pair sums / GOM scoring); meanwhile the host finishes
                            // the previous sub-chunk's trees
                GV2_CUDA(ctx, cudaEventRecord(job->ev_ready[half], ctx->stream));
                GV2_CUDA(ctx, cudaStreamWaitEvent(job->lk_stream[half], job->ev_ready[half], 0));
                cudaStream_t main_stream = ctx->stream;
                ctx->stream = job->lk_stream[half];
                rc = sink->enqueue(dev_ring + slot * n * n, s0 + c0, cb);
                ctx->stream = main_stream;
                if (rc) return rc;
                GV2_CUDA(ctx, cudaEventRecord(job->ev_drained[half], job->lk_stream[half]));
                lk_pending[half] = true;
                if (halves) half = (half + 1) % nparts;
                if ((rc = sink->drain(sink->cur))) return rc;
            }
            if (host_out) {
                const double* dsrc = dev_ring + slot * n * n;
                if (condensed && n > 1) {            // upper triangles of the sub-chunk, packed (the half's own staging buffer)
                    if ((rc = ensure(ctx, job->cond[half], (size_t)sub * hstride * sizeof(double)))) return rc;
                    condense_kernel<<<dim3((unsigned)(n - 1), (unsigned)cb), 256, 0, ctx->stream>>>(dsrc, n, job->cond[half].as<double>());
                    GV2_LAUNCH_CHECK(ctx);
                    dsrc = job->cond[half].as<double>();
                }
                GV2_CUDA(ctx, cudaEventRecord(job->ev_ready[half], ctx->stream));
                GV2_CUDA(ctx, cudaStreamWaitEvent(job->copy_stream, job->ev_ready[half], 0));
                double* hdst = cb_fn ? host_out + slot * hstride : host_out + (s0 + c0) * hstride;   // host ring vs linear output
                if (hstride > 0)
                    GV2_CUDA(ctx, cudaMemcpyAsync(hdst, dsrc, (size_t)cb * hstride * sizeof(double), cudaMemcpyDeviceToHost, job->copy_stream));
                GV2_CUDA(ctx, cudaEventRecord(job->ev_drained[half], job->copy_stream));
                drained_pending[half] = true;
                pend_b0[half] = s0 + c0; pend_cb[half] = cb; pend_slot[half] = slot;
                half ^= 1;
                if (cb_fn && (rc = deliver(half))) return rc;              // previous sub-chunk, while this one runs
            }
            slot += cb;
            c0 += cb;
        }
    }
    if (cb_fn) {
        if ((rc = deliver(half))) return rc;
        if ((rc = deliver(half ^ 1))) return rc;
    }
    if (sink)
        for (int k = 0; k < sink->nparts; ++k)                             // oldest first: trees leave in replicate order
            if ((rc = sink->drain((sink->cur + k) % sink->nparts))) return rc;
    if (host_out) GV2_CUDA(ctx, cudaStreamSynchronize(job->copy_stream));
    (void)used_mma;
    return GV2_OK;
}

extern "C" int gv2_job_replicates_stream(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring,
                                         int64_t ring_len, double* host_out, gv2_replicate_cb cb_fn, void* cb_user) {
    return replicates_impl(job, dev_weights, nb, dev_ring, ring_len, host_out, cb_fn, cb_user, nullptr);
}

// Same, with every replicate delivered in scipy's condensed form (n (n - 1) / 2 doubles, squareform order): host_out is
// [nb][n(n-1)/2] (no callback) or a host ring [ring_len][n(n-1)/2] (callback).  Half the device-to-host bytes.
extern "C" int gv2_job_replicates_condensed(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring,
                                            int64_t ring_len, double* host_out, gv2_replicate_cb cb_fn, void* cb_user) {
    if (!host_out) return gv2_fail(job ? job->ctx : nullptr, GV2_ERR_ARG, "gv2_job_replicates_condensed: host_out is NULL");
    return replicates_impl(job, dev_weights, nb, dev_ring, ring_len, host_out, cb_fn, cb_user, nullptr, nullptr, true);
}

// The bootstrap loops keep only the dendrogram of a replicate (app/src/pl1_graphs.py:108-117): replicate distance
// matrices are clustered where they are (UPGMA, device) and only their Newick strings reach the caller, in replicate
// order.  dev_ring [ring_len][n][n] is scratch: ring_len matrices are clustered concurrently (one CTA each).
extern "C" int gv2_job_replicates_newick(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring, int64_t ring_len,
                                         const char* method, const char* const* leaf_names, gv2_newick_cb cb, void* user) {
    if (!job || !leaf_names || !cb) return gv2_fail(job ? job->ctx : nullptr, GV2_ERR_ARG, "gv2_job_replicates_newick: bad arguments");
    const int mth = lk_method(method);
    if (mth < 0)
        return gv2_fail(job->ctx, GV2_ERR_UNSUPPORTED, "gv2_job_replicates_newick: linkage method '%s' (built: average, complete, weighted, ward, single, centroid, median)", method ? method : "");
    if (job->out_kind != GV2_OUT_DISTANCE) return gv2_fail(job->ctx, GV2_ERR_ARG, "gv2_job_replicates_newick: the job must produce distances");
    if (nb <= 0 || job->n == 0) return GV2_OK;
    GV2_CUDA(job->ctx, cudaSetDevice(job->ctx->device));
    TreeSink sink;
    int nparts = std::min<int64_t>(ring_len, nb) >= 8 ? GV2_RING_PARTS : 2;
    if (const char* env = getenv("GV2_RING_PARTS_USED")) nparts = std::max(2, std::min(GV2_RING_PARTS, atoi(env)));
    if (std::min<int64_t>(ring_len, nb) < nparts) nparts = 2;
    int rc = sink.init(job->ctx, job->n, std::min<int64_t>(ring_len, nb), nparts);
    if (rc) return rc;
    sink.names = leaf_names; sink.cb = cb; sink.user = user; sink.method = mth;
    return replicates_impl(job, dev_weights, nb, dev_ring, ring_len, nullptr, nullptr, nullptr, &sink);
}

// Device consumer: callback(b, DEVICE pointer of the replicate's matrix, user) on the calling thread once the matrix is
// complete in the ring (replicate order); the pointer is valid until the callback returns.
extern "C" int gv2_job_replicates_device(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_ring, int64_t ring_len,
                                         gv2_replicate_cb callback, void* user) {
    if (!callback) return gv2_fail(job ? job->ctx : nullptr, GV2_ERR_ARG, "gv2_job_replicates_device: callback is NULL");
    return replicates_impl(job, dev_weights, nb, dev_ring, ring_len, nullptr, nullptr, user, nullptr, callback);
}

extern "C" int gv2_job_replicates(gv2_job* job, const int32_t* dev_weights, int64_t nb, double* dev_out) {
    if (nb < 0) return gv2_fail(job ? job->ctx : nullptr, GV2_ERR_ARG, "gv2_job_replicates: bad arguments");
    if (nb == 0) return GV2_OK;
    return gv2_job_replicates_stream(job, dev_weights, nb, dev_out, nb, nullptr, nullptr, nullptr);
}

// ------------------------------------------------------------------ host-pointer wrappers
extern "C" int gv2_similarity_host(gv2_ctx* ctx, const double* sig, int64_t n, int64_t p, int64_t ld_sig,
                                   const double* gom, int64_t g, int64_t ld_gom, const double* spr, int scheme,
                                   double power, double threshold, int apply_threshold, int out_kind, double* out) {
    if (!ctx || !out || n < 0) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_similarity_host: bad arguments");
    if (scheme < GV2_SCHEME_P || scheme > GV2_SCHEME_PL) return gv2_fail(ctx, GV2_ERR_ARG, "unknown scheme %d", scheme);
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return GV2_OK;
    const bool need_p = scheme == GV2_SCHEME_P || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_PR || scheme == GV2_SCHEME_PL;
    const bool need_g = scheme == GV2_SCHEME_G || scheme == GV2_SCHEME_PG || scheme == GV2_SCHEME_RG;
    const bool need_r = scheme == GV2_SCHEME_R || scheme == GV2_SCHEME_RG || scheme == GV2_SCHEME_PR ||
                        scheme == GV2_SCHEME_L || scheme == GV2_SCHEME_PL;
    if ((need_p && !sig) || (need_g && !gom) || (need_r && !spr))
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_similarity_host: scheme %d is missing an input table", scheme);
    const int64_t n_pad = gv2_round_up(n, GV2_TILE), ntl = gv2_num_tiles(n);
    const size_t tile_elems = (size_t)GV2_TILE * GV2_TILE;
    int rc;
    DevBuf t64, tf, rs_p, nan_p, ws_p, g64, gf, rs_g, nan_g, ws_g, d_spr, d_out;
    gv2_gj_component cp{}, cg{};
    if (need_p) {
        const int64_t p_pad = gv2_round_up(std::max<int64_t>(p, 1), GV2_BK);
        if ((rc = upload_matrix(ctx, sig, n, p, ld_sig, t64))) return rc;
        ALLOC(ctx, tf, (size_t)n_pad * p_pad * sizeof(float));
        ALLOC(ctx, rs_p, (size_t)n * sizeof(double));
        ALLOC(ctx, nan_p, (size_t)n);
        if ((rc = gv2_dev_prepare_table(ctx, t64.as<double>(), n, p, p, threshold, apply_threshold, tf.as<float>(), n_pad, p_pad,
                                        rs_p.as<double>(), nan_p.as<uint8_t>()))) return rc;
        const int ks = pick_ksplit(ctx, ntl, 1, p_pad);
        ALLOC(ctx, ws_p, (size_t)ks * ntl * tile_elems * sizeof(double));
        if ((rc = gj_min_sums_launch(ctx, tf.as<float>(), 0, n_pad, p_pad, 0, ntl, nullptr, 1, ks, ws_p.as<double>()))) return rc;
        cp.min_sums = ws_p.as<double>(); cp.rowsums = rs_p.as<double>(); cp.nanflags = nan_p.as<uint8_t>(); cp.ksplit = ks;
    }
    if (need_g) {
        const int64_t g_pad = gv2_round_up(std::max<int64_t>(g, 1), GV2_BK);
        if ((rc = upload_matrix(ctx, gom, n, g, ld_gom, g64))) return rc;
        ALLOC(ctx, gf, (size_t)n_pad * g_pad * sizeof(float));
        ALLOC(ctx, rs_g, (size_t)n * sizeof(double));
        ALLOC(ctx, nan_g, (size_t)n);
        if ((rc = gv2_dev_prepare_table(ctx, g64.as<double>(), n, g, g, 0.0, 0, gf.as<float>(), n_pad, g_pad, rs_g.as<double>(),
                                        nan_g.as<uint8_t>()))) return rc;
        ALLOC(ctx, ws_g, (size_t)ntl * tile_elems * sizeof(double));
        if ((rc = gj_min_sums_launch(ctx, gf.as<float>(), 0, n_pad, g_pad, 0, ntl, nullptr, 1, 1, ws_g.as<double>()))) return rc;
        cg.min_sums = ws_g.as<double>(); cg.rowsums = rs_g.as<double>(); cg.nanflags = nan_g.as<uint8_t>(); cg.ksplit = 1;
    }
    if (need_r && (rc = upload_matrix(ctx, spr, n, n, n, d_spr))) return rc;
    ALLOC(ctx, d_out, (size_t)n * n * sizeof(double));
    if ((rc = gv2_dev_gj_epilogue(ctx, n, 0, ntl, 1, &cp, &cg, d_spr.as<double>(), scheme, power, out_kind, 0, d_out.as<double>(), 0))) return rc;
    GV2_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GV2_OK;
}

extern "C" int gv2_gom_table_host(gv2_ctx* ctx, const double* loc, int64_t n, int64_t p, int64_t ld_loc,
                                  const double* gom_rows, int64_t ld_gom_rows, const int64_t* gom_offsets, int64_t g,
                                  const int32_t* weights, int64_t nb, int flags, double* out) {
    if (!ctx || !loc || !gom_offsets || !out || n < 0 || p < 0 || g < 0 || nb <= 0)
        return gv2_fail(ctx, GV2_ERR_ARG, "gv2_gom_table_host: bad arguments");
    if (!weights && nb != 1) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_gom_table_host: nb must be 1 without weights");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0 || g == 0) return GV2_OK;
    const int64_t M = gom_offsets[g];
    if (M > 0 && !gom_rows) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_gom_table_host: gom_rows is NULL");
    int rc;
    DevBuf d_loc, d_rows, d_w, d_out;
    if ((rc = upload_matrix(ctx, loc, n, p, ld_loc, d_loc))) return rc;
    if ((rc = upload_matrix(ctx, gom_rows, M, p, ld_gom_rows, d_rows))) return rc;
    std::vector<int32_t> mem_ptr(g + 1);
    std::vector<int64_t> mem_row(M);
    for (int64_t k = 0; k <= g; ++k) mem_ptr[k] = (int32_t)gom_offsets[k];
    for (int64_t r = 0; r < M; ++r) mem_row[r] = r;
    GomDb* db = nullptr;
    if ((rc = gom_db_build(ctx, d_loc.as<double>(), n, p, p, d_rows.as<double>(), p, mem_row.data(), mem_ptr.data(), g, flags, &db))) return rc;
    if (weights) {
        ALLOC(ctx, d_w, (size_t)nb * p * sizeof(int32_t));
        cudaError_t e = cudaMemcpyAsync(d_w.p, weights, (size_t)nb * p * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { gom_db_free(db); return gv2_fail(ctx, GV2_ERR_CUDA, "weights upload: %s", cudaGetErrorString(e)); }
    }
    cudaError_t e = d_out.alloc((size_t)nb * n * g * sizeof(double));
    if (e != cudaSuccess) { gom_db_free(db); return gv2_fail(ctx, GV2_ERR_NOMEM, "gom table output: %s", cudaGetErrorString(e)); }
    rc = gom_db_score(ctx, db, weights ? d_w.as<int32_t>() : nullptr, nb, d_out.as<double>());
    if (rc == GV2_OK) {
        e = cudaMemcpyAsync(out, d_out.p, (size_t)nb * n * g * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = gv2_fail(ctx, GV2_ERR_CUDA, "gom table download: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(ctx->stream);
    gom_db_free(db);
    return rc;
}

extern "C" int gv2_shared_counts_host(gv2_ctx* ctx, const double* tab, int64_t n, int64_t p, int64_t ld,
                                      int32_t* inter, int32_t* nnz) {
    if (!ctx || !tab || !inter || n < 0 || p < 0) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_shared_counts_host: bad arguments");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return GV2_OK;
    const int64_t n_pad = gv2_round_up(n, GV2_TILE);
    const int64_t words = gv2_round_up(std::max<int64_t>(1, (p + 63) / 64), 16);
    int rc;
    DevBuf d_tab, d_bits, d_nnz, d_inter;
    ALLOC(ctx, d_bits, (size_t)n_pad * words * 8);
    ALLOC(ctx, d_nnz, (size_t)n * sizeof(int32_t));
    ALLOC(ctx, d_inter, (size_t)n * n * sizeof(int32_t));
    {   // the float64 table only passes through: row chunks of ~1 GB are uploaded and bit-packed one after the other
        // (the packed table is 1/64 of it), so a 50 000 x 30 000 table never needs its 12 GB resident
        const int64_t rows = std::max<int64_t>(GV2_TILE, ((int64_t)(1LL << 30) / std::max<int64_t>(8, p * 8)) / GV2_TILE * GV2_TILE);
        ALLOC(ctx, d_tab, (size_t)std::min<int64_t>(rows, n) * p * sizeof(double));
        for (int64_t r0 = 0; r0 < n; r0 += rows) {
            const int64_t rc_rows = std::min<int64_t>(rows, n - r0);
            GV2_CUDA(ctx, cudaMemcpy2DAsync(d_tab.p, p * sizeof(double), tab + r0 * ld, ld * sizeof(double), p * sizeof(double), rc_rows,
                                            cudaMemcpyHostToDevice, ctx->stream));
            const int64_t pad_rows = r0 + rc_rows == n ? n_pad - r0 : rc_rows;
            if ((rc = gv2_dev_pack_presence(ctx, d_tab.as<double>(), rc_rows, p, p, d_bits.as<uint64_t>() + (size_t)r0 * words, pad_rows, words,
                                            d_nnz.as<int32_t>() + r0))) return rc;
        }
    }
    if ((rc = gv2_dev_shared_counts(ctx, d_bits.as<uint64_t>(), n, n_pad, words, 0, gv2_num_tiles(n), d_inter.as<int32_t>()))) return rc;
    GV2_CUDA(ctx, cudaMemcpyAsync(inter, d_inter.p, (size_t)n * n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (nnz) GV2_CUDA(ctx, cudaMemcpyAsync(nnz, d_nnz.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GV2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GV2_OK;
}

extern "C" int gv2_bootstrap_host(gv2_ctx* ctx, const double* sig, const double* loc, int64_t n, int64_t p,
                                  const int32_t* gom_of_row, int64_t n_gom_rows, int64_t g, const int32_t* weights,
                                  int64_t nb, int scheme, double power, int out_kind, double* out) {
    if (!ctx || !weights || !out || nb < 0) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_bootstrap_host: bad arguments");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nb == 0 || n == 0) return GV2_OK;
    gv2_job* job = nullptr;
    int rc = gv2_job_create(ctx, sig, loc, 0, n, p, gom_of_row, n_gom_rows, g, scheme, power, out_kind, &job);
    if (rc) return rc;
    DevBuf d_w, d_out;
    cudaError_t e = d_w.alloc((size_t)nb * p * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w.p, weights, (size_t)nb * p * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { gv2_job_destroy(job); return gv2_fail(ctx, GV2_ERR_CUDA, "weights upload: %s", cudaGetErrorString(e)); }
    // replicates stream through a device ring of at most ~8 GB (two halves: one computing, one draining)
    const int64_t ring_len = std::max<int64_t>(2, std::min<int64_t>(nb, (int64_t)((8ULL << 30) / std::max<size_t>(1, (size_t)n * n * 8))));
    e = d_out.alloc((size_t)std::min<int64_t>(ring_len, nb) * n * n * sizeof(double));
    if (e != cudaSuccess) { gv2_job_destroy(job); return gv2_fail(ctx, GV2_ERR_NOMEM, "bootstrap output ring: %s", cudaGetErrorString(e)); }
    rc = gv2_job_replicates_stream(job, d_w.as<int32_t>(), nb, d_out.as<double>(), ring_len, out, nullptr, nullptr);
    if (rc == GV2_OK) {
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = gv2_fail(ctx, GV2_ERR_CUDA, "bootstrap: %s", cudaGetErrorString(e));
    }
    gv2_job_destroy(job);
    return rc;
}
