// Multi-GPU inside the library (SURVEY.md section 8b "Threading / processes", 8e): ONE caller process -- the reference's CLI or
// FastAPI worker thread -- drives every B200 of the node.  One gv2_ctx + one host thread per GPU; the tables are uploaded
// once and broadcast over NVLink; bootstrap replicates are dealt to the GPUs by index (replicate b -> GPU b mod N, no
// exchange); the un-resampled matrix is computed as upper-triangular tile ranges, one range per GPU, and the packed tile
// blocks are all-gathered with NCCL (ncclCommInitAll in this process; the library is dlopen'ed so that libgravity_b200 has
// no link-time dependency on it -- without NCCL the blocks travel by cudaMemcpyPeerAsync).
#include "common.cuh"

#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

typedef int gvncclResult_t;
typedef void* gvncclComm_t;
struct NcclApi {
    void* lib = nullptr;
    gvncclResult_t (*CommInitAll)(gvncclComm_t*, int, const int*) = nullptr;
    gvncclResult_t (*CommDestroy)(gvncclComm_t) = nullptr;
    gvncclResult_t (*AllGather)(const void*, void*, size_t, int, gvncclComm_t, cudaStream_t) = nullptr;
    gvncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, gvncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(gvncclResult_t) = nullptr;
    bool ok() const { return CommInitAll && CommDestroy && AllGather && Broadcast; }
};
#define GV_NCCL_UINT8 1
#define GV_NCCL_FLOAT64 8

static NcclApi load_nccl() {
    NcclApi a;
    const char* env = getenv("GV2_DISABLE_NCCL");
    if (env && env[0] == '1') return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
    a.Broadcast = (decltype(a.Broadcast))dlsym(a.lib, "ncclBroadcast");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    return a;
}

struct gv2_multi {
    std::vector<int> dev;
    std::vector<gv2_ctx*> ctx;
    NcclApi nccl;
    std::vector<gvncclComm_t> comm;     // empty: no NCCL, peer copies instead
    std::string err;
};

struct gv2_multi_job {
    gv2_multi* m = nullptr;
    int64_t n = 0, p = 0, g = 0;
    bool loc_is_sig = false;
    std::vector<double*> d_sig, d_loc;  // per device (owned)
    std::vector<gv2_job*> job;
};

static thread_local std::string g_multi_create_error;

static int mfail(gv2_multi* m, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m) m->err = buf;
    else g_multi_create_error = buf;
    return code;
}

extern "C" const char* gv2_multi_last_error(const gv2_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
extern "C" int gv2_multi_device_count(const gv2_multi* m) { return m ? (int)m->dev.size() : 0; }
extern "C" int gv2_multi_uses_nccl(const gv2_multi* m) { return m && !m->comm.empty() ? 1 : 0; }
extern "C" gv2_ctx* gv2_multi_context(gv2_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

extern "C" void gv2_multi_destroy(gv2_multi* m) {
    if (!m) return;
    for (size_t i = 0; i < m->comm.size(); ++i)
        if (m->comm[i] && m->nccl.CommDestroy) { cudaSetDevice(m->dev[i]); m->nccl.CommDestroy(m->comm[i]); }
    for (gv2_ctx* c : m->ctx) gv2_destroy(c);
    delete m;
}

// devices == NULL: all visible devices
extern "C" int gv2_multi_create(const int* devices, int ndev, gv2_multi** out) {
    if (!out) return mfail(nullptr, GV2_ERR_ARG, "gv2_multi_create: out is NULL");
    *out = nullptr;
    int have = gv2_device_count();
    if (have <= 0) return mfail(nullptr, GV2_ERR_CUDA, "gv2_multi_create: no CUDA device; libgravity_b200 has no CPU fallback");
    gv2_multi* m = new gv2_multi();
    if (!devices) { ndev = have; for (int i = 0; i < have; ++i) m->dev.push_back(i); }
    else for (int i = 0; i < ndev; ++i) m->dev.push_back(devices[i]);
    if (m->dev.empty()) { delete m; return mfail(nullptr, GV2_ERR_ARG, "gv2_multi_create: no devices"); }
    for (int d : m->dev) {
        gv2_ctx* c = nullptr;
        int rc = gv2_create(d, &c);
        if (rc) { std::string e = gv2_last_error(nullptr); gv2_multi_destroy(m); return mfail(nullptr, rc, "gv2_multi_create: device %d: %s", d, e.c_str()); }
        m->ctx.push_back(c);
    }
    const int N = (int)m->dev.size();
    for (int i = 0; i < N; ++i) {            // peer access: table broadcast and the NCCL-less tile gather
        cudaSetDevice(m->dev[i]);
        for (int j = 0; j < N; ++j)
            if (i != j) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[j]);
                if (can && cudaDeviceEnablePeerAccess(m->dev[j], 0) != cudaSuccess) cudaGetLastError();   // already enabled is fine
            }
    }
    if (N > 1) {
        m->nccl = load_nccl();
        if (m->nccl.ok()) {
            m->comm.assign(N, nullptr);
            gvncclResult_t r = m->nccl.CommInitAll(m->comm.data(), N, m->dev.data());
            if (r != 0) { m->comm.clear(); cudaGetLastError(); }
        }
    }
    *out = m;
    return GV2_OK;
}

// run fn(i) on one host thread per device; first non-zero return code wins
template <typename F>
static int per_device(gv2_multi* m, F fn) {
    const int N = (int)m->dev.size();
    std::vector<int> rc(N, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < N; ++i) th.emplace_back([&, i] { cudaSetDevice(m->dev[i]); rc[i] = fn(i); });
    cudaSetDevice(m->dev[0]);
    rc[0] = fn(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < N; ++i)
        if (rc[i]) { m->err = std::string("device ") + std::to_string(m->dev[i]) + ": " + gv2_last_error(m->ctx[i]); return rc[i]; }
    return GV2_OK;
}

extern "C" void gv2_multi_job_destroy(gv2_multi_job* mj) {
    if (!mj) return;
    for (size_t i = 0; i < mj->job.size(); ++i) {
        cudaSetDevice(mj->m->dev[i]);
        if (mj->job[i]) gv2_job_destroy(mj->job[i]);
    }
    for (size_t i = 0; i < mj->d_sig.size(); ++i) {
        cudaSetDevice(mj->m->dev[i]);
        if (!mj->loc_is_sig && i < mj->d_loc.size()) cudaFree(mj->d_loc[i]);
        cudaFree(mj->d_sig[i]);
    }
    delete mj;
}

// Host tables in, one device job per GPU out: ONE upload (to the first GPU), then a broadcast over NVLink.
extern "C" int gv2_multi_job_create(gv2_multi* m, const double* sig, const double* loc, int64_t n, int64_t p, const int32_t* gom_of_row,
                                    int64_t n_gom_rows, int64_t g, int scheme, double power, int out_kind, gv2_multi_job** out) {
    if (!m || !out || !sig || n < 0 || p < 0) return mfail(m, GV2_ERR_ARG, "gv2_multi_job_create: bad arguments");
    *out = nullptr;
    const int N = (int)m->dev.size();
    gv2_multi_job* mj = new gv2_multi_job();
    mj->m = m; mj->n = n; mj->p = p; mj->g = g;
    const bool need_loc = scheme != GV2_SCHEME_P && loc != nullptr;
    mj->loc_is_sig = !need_loc || loc == sig;
    mj->d_sig.assign(N, nullptr); mj->d_loc.assign(N, nullptr); mj->job.assign(N, nullptr);
    const size_t bytes = (size_t)std::max<int64_t>(1, n * p) * sizeof(double);
    int rc = per_device(m, [&](int i) -> int {
        gv2_ctx* c = m->ctx[i];
        if (cudaMalloc((void**)&mj->d_sig[i], bytes) != cudaSuccess) return gv2_fail(c, GV2_ERR_NOMEM, "multi job: signature table (%.2f GB)", bytes / 1e9);
        if (!mj->loc_is_sig && cudaMalloc((void**)&mj->d_loc[i], bytes) != cudaSuccess) return gv2_fail(c, GV2_ERR_NOMEM, "multi job: location table");
        return GV2_OK;
    });
    if (rc) { gv2_multi_job_destroy(mj); return rc; }
    // upload to device 0
    cudaSetDevice(m->dev[0]);
    cudaError_t e = cudaMemcpyAsync(mj->d_sig[0], sig, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, m->ctx[0]->stream);
    if (e == cudaSuccess && !mj->loc_is_sig) e = cudaMemcpyAsync(mj->d_loc[0], loc, (size_t)n * p * sizeof(double), cudaMemcpyHostToDevice, m->ctx[0]->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->ctx[0]->stream);
    if (e != cudaSuccess) { gv2_multi_job_destroy(mj); return mfail(m, GV2_ERR_CUDA, "multi job: table upload: %s", cudaGetErrorString(e)); }
    // broadcast + per-device job creation (list building, GOM database) in parallel
    rc = per_device(m, [&](int i) -> int {
        gv2_ctx* c = m->ctx[i];
        if (N > 1 && n * p > 0) {
            if (!m->comm.empty()) {
                gvncclResult_t r = m->nccl.Broadcast(mj->d_sig[0 == i ? 0 : i], mj->d_sig[i], (size_t)n * p, GV_NCCL_FLOAT64, 0, m->comm[i], c->stream);
                if (r == 0 && !mj->loc_is_sig) r = m->nccl.Broadcast(mj->d_loc[i], mj->d_loc[i], (size_t)n * p, GV_NCCL_FLOAT64, 0, m->comm[i], c->stream);
                if (r != 0) return gv2_fail(c, GV2_ERR_CUDA, "ncclBroadcast: %s", m->nccl.GetErrorString ? m->nccl.GetErrorString(r) : "?");
            } else if (i > 0) {
                GV2_CUDA(c, cudaMemcpyPeerAsync(mj->d_sig[i], m->dev[i], mj->d_sig[0], m->dev[0], (size_t)n * p * sizeof(double), c->stream));
                if (!mj->loc_is_sig) GV2_CUDA(c, cudaMemcpyPeerAsync(mj->d_loc[i], m->dev[i], mj->d_loc[0], m->dev[0], (size_t)n * p * sizeof(double), c->stream));
            }
            GV2_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        const double* dl = scheme == GV2_SCHEME_P ? nullptr : (mj->loc_is_sig ? mj->d_sig[i] : mj->d_loc[i]);
        return gv2_job_create(c, mj->d_sig[i], dl, 1, n, p, gom_of_row, n_gom_rows, g, scheme, power, out_kind, &mj->job[i]);
    });
    if (rc) { gv2_multi_job_destroy(mj); return rc; }
    *out = mj;
    return GV2_OK;
}

static void tile_range(int64_t ntiles, int i, int N, int64_t* a, int64_t* b) {
    const int64_t per = (ntiles + N - 1) / N;
    *a = std::min<int64_t>(ntiles, (int64_t)i * per);
    *b = std::min<int64_t>(ntiles, (int64_t)(i + 1) * per);
}

// Un-resampled matrix: tile range i on GPU i (packed blocks), all-gather, unpack + D2H on the first GPU.
extern "C" int gv2_multi_job_base_host(gv2_multi_job* mj, double* host_out) {
    if (!mj || !host_out) return mfail(mj ? mj->m : nullptr, GV2_ERR_ARG, "gv2_multi_job_base_host: bad arguments");
    gv2_multi* m = mj->m;
    const int N = (int)m->dev.size();
    const int64_t n = mj->n;
    if (n == 0) return GV2_OK;
    const int64_t ntiles = gv2_num_tiles(n), per = (ntiles + N - 1) / N;
    const size_t tile_elems = (size_t)GV2_TILE * GV2_TILE;
    std::vector<double*> pack(N, nullptr), gath(N, nullptr);
    double* full = nullptr;
    int rc = per_device(m, [&](int i) -> int {
        gv2_ctx* c = m->ctx[i];
        GV2_CUDA(c, cudaMalloc((void**)&pack[i], (size_t)per * tile_elems * sizeof(double)));
        GV2_CUDA(c, cudaMemsetAsync(pack[i], 0, (size_t)per * tile_elems * sizeof(double), c->stream));
        if (!m->comm.empty() || i == 0) GV2_CUDA(c, cudaMalloc((void**)&gath[i], (size_t)N * per * tile_elems * sizeof(double)));
        int64_t a, b;
        tile_range(ntiles, i, N, &a, &b);
        int r = gv2_job_base(mj->job[i], a, b, 1, pack[i]);
        if (r) return r;
        if (!m->comm.empty()) {
            gvncclResult_t q = m->nccl.AllGather(pack[i], gath[i], (size_t)per * tile_elems, GV_NCCL_FLOAT64, m->comm[i], c->stream);
            if (q != 0) return gv2_fail(c, GV2_ERR_CUDA, "ncclAllGather: %s", m->nccl.GetErrorString ? m->nccl.GetErrorString(q) : "?");
        }
        GV2_CUDA(c, cudaStreamSynchronize(c->stream));
        return GV2_OK;
    });
    if (rc == GV2_OK) {
        gv2_ctx* c0 = m->ctx[0];
        cudaSetDevice(m->dev[0]);
        cudaError_t e = cudaSuccess;
        if (m->comm.empty())                      // no NCCL: gather the blocks onto the first GPU by peer copies
            for (int i = 0; i < N && e == cudaSuccess; ++i)
                e = cudaMemcpyPeerAsync(gath[0] + (size_t)i * per * tile_elems, m->dev[0], pack[i], m->dev[i], (size_t)per * tile_elems * sizeof(double), c0->stream);
        if (e == cudaSuccess) e = cudaMalloc((void**)&full, (size_t)n * n * sizeof(double));
        if (e != cudaSuccess) rc = mfail(m, GV2_ERR_CUDA, "multi base: %s", cudaGetErrorString(e));
        for (int i = 0; i < N && rc == GV2_OK; ++i) {
            int64_t a, b;
            tile_range(ntiles, i, N, &a, &b);
            if (b > a) rc = gv2_dev_unpack_tiles(c0, n, a, b, gath[0] + (size_t)i * per * tile_elems, full);
            if (rc) m->err = gv2_last_error(c0);
        }
        if (rc == GV2_OK) {
            e = cudaMemcpyAsync(host_out, full, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, c0->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c0->stream);
            if (e != cudaSuccess) rc = mfail(m, GV2_ERR_CUDA, "multi base download: %s", cudaGetErrorString(e));
        }
    }
    for (int i = 0; i < N; ++i) { cudaSetDevice(m->dev[i]); cudaFree(pack[i]); cudaFree(gath[i]); }
    cudaSetDevice(m->dev[0]);
    cudaFree(full);
    return rc;
}

// Replicate dendrograms on all GPUs: replicate b -> GPU b mod N; Newick strings are handed to the caller on the CALLING
// thread in replicate order, as soon as the prefix is complete (the GPUs keep computing meanwhile).
struct NewickCollect {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::string> text;
    std::vector<char> done;
    bool failed = false;
    int64_t stride = 1, offset = 0;
};
struct NewickSlot { NewickCollect* col; int64_t stride, offset; };
static int newick_collect_cb(int64_t b_local, const char* s, size_t len, void* user) {
    NewickSlot* sl = (NewickSlot*)user;
    const int64_t b = b_local * sl->stride + sl->offset;
    {
        std::lock_guard<std::mutex> lk(sl->col->mu);
        sl->col->text[(size_t)b].assign(s, len);
        sl->col->done[(size_t)b] = 1;
    }
    sl->col->cv.notify_all();
    return 0;
}

extern "C" int gv2_multi_job_replicates_newick(gv2_multi_job* mj, const int32_t* host_weights, int64_t nb, const char* method,
                                               const char* const* leaf_names, int64_t ring_len, gv2_newick_cb cb, void* user) {
    if (!mj || !host_weights || !leaf_names || !cb || nb < 0) return mfail(mj ? mj->m : nullptr, GV2_ERR_ARG, "gv2_multi_job_replicates_newick: bad arguments");
    if (nb == 0) return GV2_OK;
    gv2_multi* m = mj->m;
    const int N = (int)m->dev.size();
    const int64_t n = mj->n, p = mj->p;
    NewickCollect col;
    col.text.resize((size_t)nb);
    col.done.assign((size_t)nb, 0);
    std::vector<int> rcs(N, 0);
    std::vector<std::thread> th;
    auto work = [&](int i) {
        cudaSetDevice(m->dev[i]);
        gv2_ctx* c = m->ctx[i];
        const int64_t mine = (nb - i + N - 1) / N;          // replicates i, i + N, ...
        if (mine <= 0) return;
        std::vector<int32_t> w((size_t)mine * p);
        for (int64_t k = 0; k < mine; ++k) memcpy(w.data() + (size_t)k * p, host_weights + (size_t)(k * N + i) * p, (size_t)p * sizeof(int32_t));
        int32_t* d_w = nullptr;
        double* ring = nullptr;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        int64_t rl = ring_len > 0 ? ring_len : std::min<int64_t>(128, (int64_t)(0.62 * (double)free_b / std::max<double>(1.0, (double)n * n * 8)));
        rl = std::max<int64_t>(1, std::min<int64_t>(rl, mine));
        NewickSlot slot{&col, N, i};
        int rc = GV2_OK;
        if (cudaMalloc((void**)&d_w, w.size() * sizeof(int32_t)) != cudaSuccess || cudaMalloc((void**)&ring, (size_t)rl * std::max<int64_t>(1, n * n) * sizeof(double)) != cudaSuccess)
            rc = gv2_fail(c, GV2_ERR_NOMEM, "multi newick: ring of %lld matrices", (long long)rl);
        if (rc == GV2_OK && cudaMemcpy(d_w, w.data(), w.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) rc = gv2_fail(c, GV2_ERR_CUDA, "multi newick: weights upload");
        if (rc == GV2_OK) rc = gv2_job_replicates_newick(mj->job[i], d_w, mine, ring, rl, method, leaf_names, newick_collect_cb, &slot);
        if (rc == GV2_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = gv2_fail(c, GV2_ERR_CUDA, "multi newick: synchronize");
        cudaFree(d_w); cudaFree(ring);
        rcs[i] = rc;
        if (rc) { std::lock_guard<std::mutex> lk(col.mu); col.failed = true; }
        col.cv.notify_all();
    };
    for (int i = 0; i < N; ++i) th.emplace_back(work, i);
    int rc = GV2_OK;
    for (int64_t b = 0; b < nb && rc == GV2_OK; ++b) {
        std::unique_lock<std::mutex> lk(col.mu);
        col.cv.wait(lk, [&] { return col.done[(size_t)b] || col.failed; });
        if (!col.done[(size_t)b]) break;
        std::string s;
        s.swap(col.text[(size_t)b]);
        lk.unlock();
        if (cb(b, s.c_str(), s.size(), user) != 0) rc = mfail(m, GV2_ERR_ARG, "tree callback asked to stop at replicate %lld", (long long)b);
    }
    for (auto& t : th) t.join();
    for (int i = 0; i < N && rc == GV2_OK; ++i)
        if (rcs[i]) { rc = rcs[i]; m->err = std::string("device ") + std::to_string(m->dev[i]) + ": " + gv2_last_error(m->ctx[i]); }
    return rc;
}

// Replicate matrices on all GPUs, delivered in host memory on the CALLING thread in replicate order: each GPU streams its
// replicates (b = i, i + N, ...) through its own pinned ring; a GPU's worker waits inside its hand-over until the caller has
// consumed the matrix, so at most one matrix per GPU is pending and the other GPUs keep computing ahead.
struct MatHandover {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<const double*> ready;     // per device: matrix waiting to be consumed (or null)
    std::vector<int64_t> ready_b;
    bool failed = false, stop = false;
};
struct MatSlot { MatHandover* h; int dev_i; int64_t stride; };
static int mat_handover_cb(int64_t b_local, const double* matrix, void* user) {
    MatSlot* s = (MatSlot*)user;
    std::unique_lock<std::mutex> lk(s->h->mu);
    s->h->ready[(size_t)s->dev_i] = matrix;
    s->h->ready_b[(size_t)s->dev_i] = b_local * s->stride + s->dev_i;
    s->h->cv.notify_all();
    s->h->cv.wait(lk, [&] { return s->h->ready[(size_t)s->dev_i] == nullptr || s->h->stop; });
    return s->h->ready[(size_t)s->dev_i] == nullptr ? 0 : 1;      // consumed: go on; stop asked while still pending: give up
}

extern "C" int gv2_multi_job_replicates_host(gv2_multi_job* mj, const int32_t* host_weights, int64_t nb, int64_t ring_len,
                                             gv2_replicate_cb cb, void* user) {
    if (!mj || !host_weights || !cb || nb < 0) return mfail(mj ? mj->m : nullptr, GV2_ERR_ARG, "gv2_multi_job_replicates_host: bad arguments");
    if (nb == 0) return GV2_OK;
    gv2_multi* m = mj->m;
    const int N = (int)m->dev.size();
    const int64_t n = mj->n, p = mj->p;
    MatHandover h;
    h.ready.assign(N, nullptr);
    h.ready_b.assign(N, -1);
    std::vector<int> rcs(N, 0);
    std::vector<std::thread> th;
    auto work = [&](int i) {
        cudaSetDevice(m->dev[i]);
        gv2_ctx* c = m->ctx[i];
        const int64_t mine = (nb - i + N - 1) / N;
        if (mine <= 0) return;
        std::vector<int32_t> w((size_t)mine * p);
        for (int64_t k = 0; k < mine; ++k) memcpy(w.data() + (size_t)k * p, host_weights + (size_t)(k * N + i) * p, (size_t)p * sizeof(int32_t));
        int32_t* d_w = nullptr;
        double *ring = nullptr, *hring = nullptr;
        const int64_t rl = std::max<int64_t>(2, std::min<int64_t>(ring_len > 0 ? ring_len : 4, std::max<int64_t>(2, mine)));
        const size_t rb = (size_t)rl * std::max<int64_t>(1, n * n) * sizeof(double);
        MatSlot slot{&h, i, N};
        int rc = GV2_OK;
        if (cudaMalloc((void**)&d_w, w.size() * sizeof(int32_t)) != cudaSuccess || cudaMalloc((void**)&ring, rb) != cudaSuccess ||
            cudaMallocHost((void**)&hring, rb) != cudaSuccess)
            rc = gv2_fail(c, GV2_ERR_NOMEM, "multi replicates: rings of %lld matrices", (long long)rl);
        if (rc == GV2_OK && cudaMemcpy(d_w, w.data(), w.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) rc = gv2_fail(c, GV2_ERR_CUDA, "multi replicates: weights upload");
        if (rc == GV2_OK) rc = gv2_job_replicates_stream(mj->job[i], d_w, mine, ring, rl, hring, mat_handover_cb, &slot);
        cudaFree(d_w); cudaFree(ring); cudaFreeHost(hring);
        rcs[i] = rc;
        if (rc) { std::lock_guard<std::mutex> lk(h.mu); h.failed = true; }
        h.cv.notify_all();
    };
    for (int i = 0; i < N; ++i) th.emplace_back(work, i);
    int rc = GV2_OK;
    for (int64_t b = 0; b < nb; ++b) {
        const int i = (int)(b % N);
        std::unique_lock<std::mutex> lk(h.mu);
        h.cv.wait(lk, [&] { return (h.ready[(size_t)i] && h.ready_b[(size_t)i] == b) || h.failed; });
        if (!(h.ready[(size_t)i] && h.ready_b[(size_t)i] == b)) break;
        const double* mat = h.ready[(size_t)i];
        lk.unlock();
        const int r = cb(b, mat, user);
        lk.lock();
        h.ready[(size_t)i] = nullptr;
        if (r != 0) { h.stop = true; rc = mfail(m, GV2_ERR_ARG, "replicate callback asked to stop at replicate %lld", (long long)b); }
        h.cv.notify_all();
        if (r != 0) break;
    }
    { std::lock_guard<std::mutex> lk(h.mu); h.stop = true; }      // (after the last hand-over every worker has returned already)
    h.cv.notify_all();
    for (auto& t : th) t.join();
    for (int i = 0; i < N && rc == GV2_OK; ++i)
        if (rcs[i]) { rc = rcs[i]; m->err = std::string("device ") + std::to_string(m->dev[i]) + ": " + gv2_last_error(m->ctx[i]); }
    return rc;
}
