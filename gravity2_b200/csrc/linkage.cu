// Batched UPGMA dendrograms on the device + Newick strings on the host.
//
// Reference: DistMat2Tree, app/utils/dist_mat_to_tree.py:5-21 -- the distance matrix is zero-clamped, its upper
// triangle is handed to scipy.cluster.hierarchy.linkage(method) (scipy is a third-party dependency of the
// reference, pinned scipy==1.14.0 in requirements.txt:4; its published algorithm for method "average" is the
// nearest-neighbour chain of scipy/cluster/_hierarchy.pyx `nn_chain` -- also behind complete, weighted and ward -- and
// for "single" `mst_single_linkage`, for "centroid" / "median" `fast_linkage` (nearest-neighbour candidates in a binary
// heap), restated here (the test suite holds the same loops in Python, pinned bit for
// bit to the installed scipy), to_tree() turns the linkage
// matrix into nodes and GetNewick (app/utils/get_newick.py:1-13) prints them.  The bootstrap loops
// (app/src/pl1_graphs.py:108-117, app/src/virus_classification.py:316-330) do this once per replicate and keep
// only the Newick line, so with the clustering on the device a replicate's 0.8 GB distance matrix never has to
// leave HBM: 320 KB of linkage rows do.
//
// nn_chain, as scipy runs it (the order of every comparison decides how ties fall, and real distance matrices
// are full of exact ties at D = 1):
//   chain: a stack of clusters, each the nearest neighbour of the one below.  With an empty chain start at the
//   lowest-numbered live cluster.  For the top x: y = argmin_i D[x][i] over live i != x, scanning i upwards with a
//   strict `<`, starting from the previous chain element as the incumbent (so it wins ties).  If y is that previous
//   element the two are mutual nearest neighbours: merge; else push y.
//   merge (x < y after a swap): row k = (x, y, D[x][y], |x| + |y|); x dies, y becomes the union,
//   D[i][y] = (|x| D[i][x] + |y| D[i][y]) / (|x| + |y|) for every live i (three separately rounded operations).
//   Afterwards the rows are stably sorted by distance and relabelled with a union-find (host, below).
//
// One CTA per matrix (the algorithm is a chain of ~3n dependent row scans: parallelism comes from clustering
// many replicates at once); the matrix is the full symmetric n x n float64 block the fused tail wrote, so a row
// scan is one contiguous 8n-byte read; it is updated in place (row y contiguous, column y scattered).
#include "common.cuh"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "linkage_kernels.cuh"

// scipy clusters the UPPER triangle of the matrix it is given (dist_mat_to_tree.py:9): mirror it down, so that an
// asymmetric input behaves as in the reference.  grid = (ceil(n/32), ceil(n/8)), block = (32, 8)
__global__ void lk_mirror_upper_kernel(double* __restrict__ D, long long n) {
    const long long j = (long long)blockIdx.x * 32 + threadIdx.x, i = (long long)blockIdx.y * 8 + threadIdx.y;
    if (i < n && j < n && j > i) D[j * n + i] = D[i * n + j];
}

// ---------------------------------------------------------------------------------- host: sort, relabel, Newick
// Rows in merge order -> scipy's linkage matrix: stable sort by distance, then union-find labels
// (scipy/cluster/_hierarchy.pyx: nn_chain tail + label()).
static void lk_finalize(const double* raw, int64_t n, int method, double* Z) {
    const int64_t m = n - 1;
    if (method == LK_CENTROID || method == LK_MEDIAN) {       // fast_linkage returns its rows as they are
        memcpy(Z, raw, (size_t)m * 4 * sizeof(double));
        return;
    }
    std::vector<int64_t> order(m);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return raw[a * 4 + 2] < raw[b * 4 + 2]; });
    std::vector<int64_t> parent(2 * n - 1), size(2 * n - 1, 1);
    std::iota(parent.begin(), parent.end(), 0);
    int64_t next = n;
    auto find = [&](int64_t x) {
        int64_t p = x;
        while (parent[x] != x) x = parent[x];
        while (parent[p] != x) { const int64_t q = parent[p]; parent[p] = x; p = q; }
        return x;
    };
    for (int64_t i = 0; i < m; ++i) {
        const double* r = raw + order[i] * 4;
        const int64_t xr = find((int64_t)r[0]), yr = find((int64_t)r[1]);
        Z[i * 4 + 0] = (double)std::min(xr, yr);
        Z[i * 4 + 1] = (double)std::max(xr, yr);
        Z[i * 4 + 2] = r[2];
        parent[xr] = next; parent[yr] = next;
        size[next] = size[xr] + size[yr];
        Z[i * 4 + 3] = (double)size[next];
        ++next;
    }
}

// GetNewick(to_tree(Z), "", root.dist, names), app/utils/get_newick.py:1-13: children are printed right first, every
// branch length is "%f" of (parent height - node height), the root closes with ");".  Iterative (the reference recurses).
static void lk_newick(const double* Z, int64_t n, const char* const* names, std::string& out) {
    out.clear();
    if (n <= 0) return;
    char buf[64];
    if (n == 1) {                                     // to_tree of an empty linkage is the single leaf; its parent height is its own
        out += names[0];
        snprintf(buf, sizeof buf, ":%f", 0.0);
        out += buf;
        return;
    }
    struct Frame { int64_t node; double parent; int stage; };
    std::vector<Frame> st;
    const int64_t root = 2 * n - 2;
    auto height = [&](int64_t v) { return v < n ? 0.0 : Z[(v - n) * 4 + 2]; };
    st.push_back({root, height(root), 0});
    while (!st.empty()) {
        Frame& f = st.back();
        if (f.node < n) {
            out += names[f.node];
            snprintf(buf, sizeof buf, ":%f", f.parent - 0.0);
            out += buf;
            st.pop_back();
            continue;
        }
        const int64_t k = f.node - n;
        const double h = Z[k * 4 + 2];
        if (f.stage == 0) {
            out += '(';
            f.stage = 1;
            st.push_back({(int64_t)Z[k * 4 + 1], h, 0});      // right child first
        } else if (f.stage == 1) {
            out += ',';
            f.stage = 2;
            st.push_back({(int64_t)Z[k * 4 + 0], h, 0});
        } else {
            if (f.node == root) out += ");";
            else {
                snprintf(buf, sizeof buf, "):%f", f.parent - h);
                out += buf;
            }
            st.pop_back();
        }
    }
}

// ---------------------------------------------------------------------------------- entry points
int lk_method(const char* method) {
    if (!method) return -1;
    if (!strcmp(method, "average")) return LK_AVERAGE;
    if (!strcmp(method, "complete")) return LK_COMPLETE;
    if (!strcmp(method, "weighted")) return LK_WEIGHTED;
    if (!strcmp(method, "ward")) return LK_WARD;
    if (!strcmp(method, "single")) return LK_SINGLE;
    if (!strcmp(method, "centroid")) return LK_CENTROID;
    if (!strcmp(method, "median")) return LK_MEDIAN;
    return -1;
}

size_t lk_smem_bytes(int64_t n, int method) {
    return method == LK_SINGLE ? (size_t)((n + 1) & ~1) * sizeof(int) + (size_t)n * sizeof(double) : (size_t)n * 2 * sizeof(int);
}
#define LK_SMEM_MAX (220 * 1024)

template <int METHOD>
static cudaError_t lk_launch_chain(gv2_ctx* ctx, bool lazy, size_t smem, unsigned g, double* dev_D, int64_t d_stride, int64_t n,
                                   double* dev_Zraw) {
    // live-list kernel (lazy columns + work that follows the number of live clusters): 18 bytes of shared memory per leaf
    const char* env = getenv("GV2_LK_LIST");
    if (lazy && n <= (int64_t)LK_LIST_PER * LK_THREADS && lk_list_smem_bytes(n) <= LK_SMEM_MAX && !(env && atoi(env) == 0)) {
        const size_t lsm = lk_list_smem_bytes(n);
        cudaError_t e = cudaFuncSetAttribute(lk_nn_chain_list_kernel<METHOD, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(lsm, 1024));
        if (e != cudaSuccess) return e;
        lk_nn_chain_list_kernel<METHOD, 1024><<<g, LK_THREADS, lsm, ctx->stream>>>(dev_D, d_stride, (int)n, dev_Zraw);
        return cudaGetLastError();
    }
    const int sz = (int)std::max<size_t>(smem, 1024);
    cudaError_t e = lazy ? cudaFuncSetAttribute(lk_nn_chain_kernel<METHOD, 1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz)
                         : cudaFuncSetAttribute(lk_nn_chain_kernel<METHOD, 1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
    if (e != cudaSuccess) return e;
    if (lazy) lk_nn_chain_kernel<METHOD, 1024, true><<<g, LK_THREADS, smem, ctx->stream>>>(dev_D, d_stride, (int)n, dev_Zraw);
    else lk_nn_chain_kernel<METHOD, 1024, false><<<g, LK_THREADS, smem, ctx->stream>>>(dev_D, d_stride, (int)n, dev_Zraw);
    return cudaGetLastError();
}

// dev_D [nb][n][n] (stride d_stride elements; symmetric) is clustered in place (destroyed); dev_Zraw [nb][n-1][4] receives the
// merge-ordered rows.  Enqueued on the context's stream.
int lk_launch(gv2_ctx* ctx, int method, double* dev_D, int64_t d_stride, int64_t n, int64_t nb, double* dev_Zraw) {
    if (n <= 1 || nb <= 0) return GV2_OK;
    if (method < 0 || method > LK_MEDIAN) return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "device linkage: methods average, complete, weighted, ward, single, centroid, median");
    if (method == LK_CENTROID || method == LK_MEDIAN) {
        const size_t smem = lk_generic_smem_bytes(n);
        if (smem > LK_SMEM_MAX)
            return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "device linkage (centroid / median) keeps its heap in shared memory: n = %lld needs %zu bytes", (long long)n, smem);
        Gv2Timer timer(ctx, GV2_TIME_LINKAGE);
        timer.begin();
        cudaError_t e = method == LK_CENTROID
            ? cudaFuncSetAttribute(lk_generic_kernel<LK_CENTROID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024))
            : cudaFuncSetAttribute(lk_generic_kernel<LK_MEDIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
        if (e == cudaSuccess) {
            if (method == LK_CENTROID) lk_generic_kernel<LK_CENTROID><<<(unsigned)nb, LK_THREADS, smem, ctx->stream>>>(dev_D, d_stride, (int)n, dev_Zraw);
            else lk_generic_kernel<LK_MEDIAN><<<(unsigned)nb, LK_THREADS, smem, ctx->stream>>>(dev_D, d_stride, (int)n, dev_Zraw);
            e = cudaGetLastError();
        }
        ctx->launches++;
        timer.end();
        GV2_CUDA(ctx, e);
        return GV2_OK;
    }
    // lazy column updates need 16 bytes of shared memory per leaf (n <= 14 080); larger trees write the columns
    const char* env = getenv("GV2_LK_LAZY");
    const bool lazy = method != LK_SINGLE && (size_t)n * 16 <= LK_SMEM_MAX && !(env && atoi(env) == 0);
    const size_t smem = lazy ? (size_t)n * 16 : lk_smem_bytes(n, method);
    if (smem > LK_SMEM_MAX)
        return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "device linkage keeps its per-cluster state in shared memory: n = %lld needs %zu bytes", (long long)n, smem);
    Gv2Timer timer(ctx, GV2_TIME_LINKAGE);
    timer.begin();
    const unsigned g = (unsigned)nb;
    cudaError_t e = cudaSuccess;
    switch (method) {
        case LK_AVERAGE: e = lk_launch_chain<LK_AVERAGE>(ctx, lazy, smem, g, dev_D, d_stride, n, dev_Zraw); break;
        case LK_COMPLETE: e = lk_launch_chain<LK_COMPLETE>(ctx, lazy, smem, g, dev_D, d_stride, n, dev_Zraw); break;
        case LK_WEIGHTED: e = lk_launch_chain<LK_WEIGHTED>(ctx, lazy, smem, g, dev_D, d_stride, n, dev_Zraw); break;
        case LK_WARD: e = lk_launch_chain<LK_WARD>(ctx, lazy, smem, g, dev_D, d_stride, n, dev_Zraw); break;
        default:
            e = cudaFuncSetAttribute(lk_mst_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
            if (e == cudaSuccess) {
                lk_mst_single_kernel<<<g, LK_THREADS, smem, ctx->stream>>>(dev_D, d_stride, (int)n, dev_Zraw);
                e = cudaGetLastError();
            }
            break;
    }
    ctx->launches++;
    timer.end();
    GV2_CUDA(ctx, e);
    return GV2_OK;
}

void lk_finalize_host(const double* raw, int64_t n, int method, double* Z) { lk_finalize(raw, n, method, Z); }
void lk_newick_host(const double* Z, int64_t n, const char* const* names, std::string& out) { lk_newick(Z, n, names, out); }

extern "C" int gv2_newick_from_linkage(const double* Z, int64_t n, const char* const* leaf_names, char* out, size_t out_cap,
                                       size_t* out_len) {
    if (!Z && n > 1) return GV2_ERR_ARG;
    if (!leaf_names || !out_len) return GV2_ERR_ARG;
    std::string s;
    lk_newick(Z, n, leaf_names, s);
    *out_len = s.size();
    if (out && out_cap > s.size()) memcpy(out, s.c_str(), s.size() + 1);
    else if (out) return GV2_ERR_NOMEM;
    return GV2_OK;
}

extern "C" int gv2_linkage_host(gv2_ctx* ctx, const double* dist, int64_t n, const char* method, double* Z) {
    if (!ctx || !dist || !Z || n < 1) return gv2_fail(ctx, GV2_ERR_ARG, "gv2_linkage_host: bad arguments");
    const int mth = lk_method(method);
    if (mth < 0) return gv2_fail(ctx, GV2_ERR_UNSUPPORTED, "gv2_linkage_host: method '%s' (built: average, complete, weighted, ward, single, centroid, median)", method ? method : "");
    GV2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 1) return GV2_OK;
    double *d_D = nullptr, *d_Z = nullptr;
    auto cleanup = [&]() { cudaFree(d_D); cudaFree(d_Z); };
    cudaError_t e = cudaMalloc((void**)&d_D, (size_t)n * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_Z, (size_t)(n - 1) * 4 * sizeof(double));
    if (e != cudaSuccess) { cleanup(); return gv2_fail(ctx, GV2_ERR_NOMEM, "gv2_linkage_host: %s", cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(d_D, dist, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    int rc = GV2_OK;
    if (e == cudaSuccess) {
        lk_mirror_upper_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((n + 7) / 8)), dim3(32, 8), 0, ctx->stream>>>(d_D, n);
        ctx->launches++;
        rc = lk_launch(ctx, mth, d_D, n * n, n, 1, d_Z);
    }
    std::vector<double> raw((size_t)(n - 1) * 4);
    if (e == cudaSuccess && rc == GV2_OK) e = cudaMemcpyAsync(raw.data(), d_Z, raw.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == GV2_OK) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (rc) return rc;
    if (e != cudaSuccess) return gv2_fail(ctx, GV2_ERR_CUDA, "gv2_linkage_host: %s", cudaGetErrorString(e));
    lk_finalize(raw.data(), n, mth, Z);
    return GV2_OK;
}
