// extern "C" surface of libsubg_b200.so (declared in include/subg_b200.h).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace subg {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

// ---- measurement hooks
struct TimedRegion {
    cudaEvent_t a, b;
    int which;
    bool closed;
};
static std::mutex g_tm_mutex;
static std::vector<TimedRegion> g_regions;
static bool g_timing = false;
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static bool capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}
void timing_begin(int which, cudaStream_t st) {
    if (!g_timing || capturing(st)) return;
    TimedRegion r{};
    r.which = which;
    r.closed = false;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_tm_mutex);
    g_regions.push_back(r);
}
void timing_end(int which, cudaStream_t st) {
    if (!g_timing || capturing(st)) return;
    std::lock_guard<std::mutex> lk(g_tm_mutex);
    for (size_t i = g_regions.size(); i-- > 0;)
        if (g_regions[i].which == which && !g_regions[i].closed) {
            cudaEventRecord(g_regions[i].b, st);
            g_regions[i].closed = true;
            return;
        }
}

// ---- cache of large device blocks (see common.cuh)
struct BigBlock {
    void *p;
    size_t bytes;
    int device;
    cudaStream_t st;
    cudaEvent_t ev;
};
static std::mutex g_big_mutex;
static std::vector<BigBlock> g_big_free;                      // cached, oldest first
static std::vector<std::pair<void *, size_t>> g_big_live;    // handed out: (pointer, bytes)
static long long g_big_limit = -1;

static void big_release_locked(size_t i) {
    BigBlock b = g_big_free[i];
    g_big_free.erase(g_big_free.begin() + (long)i);
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != b.device) cudaSetDevice(b.device);
    cudaEventSynchronize(b.ev);
    cudaEventDestroy(b.ev);
    cudaFree(b.p);
    if (cur != b.device) cudaSetDevice(cur);
}

cudaError_t big_alloc(void **p, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_big_mutex);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < g_big_free.size(); i++) {
        const BigBlock &b = g_big_free[i];
        if (b.device == dev && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + ((size_t)64 << 20) &&
            (best == (size_t)-1 || b.bytes < g_big_free[best].bytes))
            best = i;
    }
    if (best != (size_t)-1) {
        BigBlock b = g_big_free[best];
        g_big_free.erase(g_big_free.begin() + (long)best);
        if (b.st != st) cudaStreamWaitEvent(st, b.ev, 0);   // same stream: already ordered behind the last use
        cudaEventDestroy(b.ev);
        *p = b.p;
        g_big_live.emplace_back(b.p, b.bytes);
        return cudaSuccess;
    }
    const size_t rounded = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    cudaError_t e = cudaMalloc(p, rounded);
    while (e == cudaErrorMemoryAllocation) {   // give cached blocks of this device back to the driver and retry
        cudaGetLastError();
        size_t victim = (size_t)-1;
        for (size_t i = 0; i < g_big_free.size(); i++)
            if (g_big_free[i].device == dev) { victim = i; break; }
        if (victim == (size_t)-1) break;
        big_release_locked(victim);
        e = cudaMalloc(p, rounded);
    }
    if (e == cudaSuccess) g_big_live.emplace_back(*p, rounded);
    return e;
}

bool big_free(void *p, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_big_mutex);
    size_t at = (size_t)-1;
    for (size_t i = g_big_live.size(); i-- > 0;)
        if (g_big_live[i].first == p) { at = i; break; }
    if (at == (size_t)-1) return false;
    BigBlock b{};
    b.p = p;
    b.bytes = g_big_live[at].second;
    g_big_live.erase(g_big_live.begin() + (long)at);
    cudaGetDevice(&b.device);
    b.st = st;
    if (cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(b.ev, st) != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(st);
        cudaFree(p);
        return true;
    }
    g_big_free.push_back(b);
    if (g_big_limit < 0) {   // cached bytes allowed per process: SUBG_BLOCK_CACHE_BYTES, default 45 % of the device
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const char *v = getenv("SUBG_BLOCK_CACHE_BYTES");
        g_big_limit = v ? atoll(v) : (long long)(total_b * 0.45);
    }
    size_t held = 0;
    for (const BigBlock &q : g_big_free) held += q.bytes;
    while (held > (size_t)g_big_limit && !g_big_free.empty()) {
        held -= g_big_free[0].bytes;
        big_release_locked(0);
    }
    return true;
}

// give every cached large block back to the driver (all devices); returns the bytes released
long long big_trim() {
    std::lock_guard<std::mutex> lk(g_big_mutex);
    long long freed = 0;
    while (!g_big_free.empty()) {
        freed += (long long)g_big_free[0].bytes;
        big_release_locked(0);
    }
    return freed;
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int gset_sample_impl(const Graph *g, const int32_t *seeds_hd, int64_t n_all, int64_t lo, int64_t hi, int M, int m,
                     int bucket, uint64_t seed, int rng_mode, const int32_t *walks_hd, int flags, cudaStream_t st,
                     SpG **out);
int spg_set_lp_table_impl(SpG *s, const int32_t *id_map_hd, const int16_t *enc_hd, int32_t c_new, int32_t ncol,
                          cudaStream_t st);
int spg_export_impl(const SpG *s, int32_t *nsize_hd, int32_t *remap_hd, int16_t *enc_hd, int16_t *raw_hd,
                    cudaStream_t st);
int spg_expand_rows_impl(SpG *s, int64_t num_nodes, cudaStream_t st);
int spg_from_csr_impl(const int64_t *indptr_hd, const int32_t *indices_hd, const void *data_hd, int value_kind,
                      int64_t n_rows, int64_t nnz, int device, cudaStream_t st, SpG **out);
int spg_alloc_impl(int64_t n, int64_t T, int device, cudaStream_t st, SpG **out);
int spg_seal_impl(SpG *s, cudaStream_t st);
int spjoin_plan_impl(const SpG *s, const int64_t *edge_hd, int64_t B, int arity, int64_t *edge_dev,
                     int64_t *indptr_dev, int64_t *N_out, cudaStream_t st);
int spjoin_run_impl(const SpG *s, const int64_t *edge_dev, int64_t B, int arity, const int64_t *indptr_dev,
                    const float *enc_table_dev, int k, void *out_dev, int64_t *segid_dev, cudaStream_t st);
int spjoin_fused_impl(const SpG *s, const int64_t *edge_hd, int64_t B, int arity, int64_t *edge_dev, int64_t *indptr_dev,
                      const float *enc_table_dev, int k, void *out_dev, int64_t out_capacity, int64_t *segid_dev,
                      int64_t *N_out, int *ran, cudaStream_t st);
int ppr_topk_impl(const Graph *g, const int32_t *seeds_hd, int64_t n, float alpha, float eps, int topk,
                  int normalization, const double *norm_deg_hd, cudaStream_t st, SpG **out);
int spg_encode_impl(const Graph *g, const SpG *x, int encoder, cudaStream_t st, SpG **out);
int graph_from_edges_impl(const int64_t *row_hd, const int64_t *col_hd, int64_t E, int64_t N_in, int symmetrize,
                          int drop_self_loops, int device, cudaStream_t st, Graph **out);
int graph_export_impl(const Graph *g, int64_t *rowptr_hd, int32_t *col_hd, cudaStream_t st);
struct Joiner;
int joiner_create_impl(const SpG *s, int64_t B, int arity, const float *enc_table_dev, int k, int64_t capacity_rows,
                       int want_segid, int depth, Joiner **out);
int joiner_submit_impl(Joiner *j, const int64_t *edge_hd, int edge_on_device, cudaStream_t st, void **out_dev, int64_t **indptr_dev,
                       int64_t **segid_dev, const int64_t **nrows_dev, int *slot_out);
int joiner_rows_impl(Joiner *j, int slot, int64_t *N);
void joiner_free_impl(Joiner *j);
struct Xchg;
int xchg_create_impl(int device, int rank, int world, int64_t slab_bytes, Xchg **out);
int xchg_export_impl(const Xchg *x, void *handle64);
int xchg_open_impl(Xchg *x, const void *handles);
int xchg_slab_impl(const Xchg *x, void **slab_dev, int64_t *bytes);
int xchg_pack_impl(Xchg *x, const SpG *s, int64_t n_nodes, int64_t *header, cudaStream_t st);
int xchg_assemble_impl(Xchg *x, const int64_t *headers, const void *const *srcs, int M, int ncol, cudaStream_t st, SpG **out);
int xchg_stage_impl(Xchg *x, const SpG *s, int64_t n_nodes, int64_t plane_entries, int64_t *header, cudaStream_t st);
int xchg_link_impl(Xchg *x, const int64_t *headers, const void *const *srcs, int M, int ncol, cudaStream_t st, SpG **out);
void xchg_free_impl(Xchg *x);
struct WalkSet;
int walk_sample_impl(const Graph *g, const int32_t *seeds_hd, int64_t n, int M, int m, uint64_t seed, int rng_mode,
                     int without, cudaStream_t st, WalkSet **out);
int walkset_export_impl(const WalkSet *w, int32_t *walks_hd, int64_t *off_hd, int32_t *ids_hd, int32_t *rpe_hd, cudaStream_t st);
int walkset_info_impl(const WalkSet *w, int64_t *n, int64_t *T, int32_t *M, int32_t *ncol, uint32_t *status);
int walkset_views_impl(const WalkSet *w, const int32_t **walks, const int64_t **off, const int32_t **ids, const int32_t **rpe);
void walkset_free_impl(WalkSet *w);
int batch_sample_impl(const Graph *g, const int32_t *seeds_hd, int64_t n, int M, int m, int thld, uint32_t state, int32_t *out_hd,
                      int64_t cap, int64_t *count_out, cudaStream_t st);
int walk_join_impl(const int32_t *walks_hd, int64_t n, int64_t stride, const int64_t *key_off_hd, const int32_t *key_ids_hd,
                   const int32_t *query_hd, int64_t Q, int32_t *out_hd, int32_t *xq_hd, int device, cudaStream_t st);

__global__ void widen_rowptr_kernel(const int32_t *in, long long *out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void narrow_rowptr_kernel(const long long *in, int32_t *out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)in[i];
}

// {start, degree} of every row packed in one 64-bit word (start: low 40 bits, degree: high 24 bits), so a
// walk step costs one 8-byte load whatever the rowptr width.  Degrees >= 2^24 - 1 store the escape value
// 0xFFFFFF and are read from rowptr.
template <typename P>
__global__ void rowinfo_kernel(const P *rowptr, unsigned long long *out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long start = (unsigned long long)rowptr[i];
        unsigned long long d = (unsigned long long)(rowptr[i + 1] - rowptr[i]);
        if (d > 0xFFFFFFull) d = 0xFFFFFFull;
        out[i] = (d << 40) | start;
    }
}

static int init_device(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(SUBG_ERR_CUDA, "no CUDA device: libsubg_b200 has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(SUBG_ERR_ARG, "bad device index");
    if (const char *v = getenv("SUBG_L2_FETCH")) {  // experiment knob: L2 fetch granularity hint (32 / 64 / 128 bytes)
        DeviceGuard guard(device);
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(v));
    }
    // keep freed scratch in the stream-ordered pool instead of returning it to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return SUBG_OK;
}

}  // namespace subg

using namespace subg;

extern "C" {

int subg_abi_version(void) { return SUBG_ABI_VERSION; }
const char *subg_last_error(void) { return g_last_error.c_str(); }

int subg_graph_create(const void *rowptr_hd, int rowptr_is64, const int32_t *col_hd, int64_t N, int64_t E,
                      int device, void *stream, subg_graph **out) {
    if (!rowptr_hd || !out || N < 0 || E < 0 || (E > 0 && !col_hd)) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (N >= (1ll << 31)) return fail(SUBG_ERR_ARG, "node ids must fit int32");
    if (int rc = init_device(device)) return rc;
    DeviceGuard guard(device);
    cudaStream_t st = (cudaStream_t)stream;
    Graph *g = new Graph();
    g->tag.last = st;
    g->device = device; g->N = N; g->E = E;
    g->rowptr64 = E >= (1ll << 31);
    cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, device);
    const size_t in_b = rowptr_is64 ? 8 : 4, keep_b = g->rowptr64 ? 8 : 4;
    void *tmp = nullptr;
    cudaError_t e = cudaMallocAsync(&g->rowptr, ((size_t)N + 2) * keep_b, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&g->col, ((size_t)E + 16) * 4, st);
    if (e == cudaSuccess && E > 0) e = cudaMemcpyAsync(g->col, col_hd, (size_t)E * 4, cudaMemcpyDefault, st);
    if (e == cudaSuccess) {
        if (in_b == keep_b) {
            e = cudaMemcpyAsync(g->rowptr, rowptr_hd, ((size_t)N + 1) * keep_b, cudaMemcpyDefault, st);
        } else {
            const void *src = rowptr_hd;
            if (!is_device_ptr(rowptr_hd)) {
                e = cudaMallocAsync(&tmp, ((size_t)N + 1) * in_b, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, rowptr_hd, ((size_t)N + 1) * in_b, cudaMemcpyHostToDevice, st);
                src = tmp;
            }
            if (e == cudaSuccess) {
                const unsigned blocks = (unsigned)std::min<int64_t>((N + 256) / 256, 4 * (int64_t)g->num_sms);
                if (rowptr_is64) narrow_rowptr_kernel<<<blocks, 256, 0, st>>>((const long long *)src, (int32_t *)g->rowptr, N + 1);
                else widen_rowptr_kernel<<<blocks, 256, 0, st>>>((const int32_t *)src, (long long *)g->rowptr, N + 1);
                e = cudaGetLastError();
            }
        }
    }
    if (e == cudaSuccess && E >= (1ll << 40)) e = cudaErrorInvalidValue;  // 40-bit row starts
    if (e == cudaSuccess) e = cudaMallocAsync(&g->rowinfo, ((size_t)N + 1) * 8, st);
    if (e == cudaSuccess && N > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((N + 255) / 256, 4 * (int64_t)g->num_sms);
        if (g->rowptr64) rowinfo_kernel<long long><<<blocks, 256, 0, st>>>((const long long *)g->rowptr, (unsigned long long *)g->rowinfo, N);
        else rowinfo_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t *)g->rowptr, (unsigned long long *)g->rowinfo, N);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (tmp) cudaFreeAsync(tmp, st);
    if (e != cudaSuccess) {
        if (g->rowptr) cudaFreeAsync(g->rowptr, st);
        if (g->col) cudaFreeAsync(g->col, st);
        if (g->rowinfo) cudaFreeAsync(g->rowinfo, st);
        delete g;
        return fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = reinterpret_cast<subg_graph *>(g);
    return SUBG_OK;
}

int subg_graph_from_edges(const int64_t *row_hd, const int64_t *col_hd, int64_t E, int64_t num_nodes, int symmetrize,
                          int drop_self_loops, int device, void *stream, subg_graph **out) {
    if (int rc = init_device(device)) return rc;
    return graph_from_edges_impl(row_hd, col_hd, E, num_nodes, symmetrize, drop_self_loops, device, (cudaStream_t)stream,
                                 reinterpret_cast<Graph **>(out));
}
int subg_graph_export(const subg_graph *g, int64_t *rowptr_hd, int32_t *col_hd, void *stream) {
    return graph_export_impl(reinterpret_cast<const Graph *>(g), rowptr_hd, col_hd, (cudaStream_t)stream);
}

int subg_graph_info(const subg_graph *g_, int64_t *N, int64_t *E, int *device) {
    const Graph *g = reinterpret_cast<const Graph *>(g_);
    if (!g) return fail(SUBG_ERR_ARG, "null graph");
    if (N) *N = g->N;
    if (E) *E = g->E;
    if (device) *device = g->device;
    return SUBG_OK;
}

void subg_graph_free(subg_graph *g_) {
    Graph *g = reinterpret_cast<Graph *>(g_);
    if (!g) return;
    DeviceGuard guard(g->device);
    const cudaStream_t st = g->tag.free_stream();  // behind the last kernel that read the graph
    if (g->rowptr) cudaFreeAsync(g->rowptr, st);
    if (g->col) cudaFreeAsync(g->col, st);
    if (g->rowinfo) cudaFreeAsync(g->rowinfo, st);
    if (g->col3) cudaFreeAsync(g->col3, st);
    delete g;
}

int subg_gset_sample(const subg_graph *g, const int32_t *seeds_hd, int64_t n, int num_walks, int num_steps,
                     int bucket, uint64_t seed, int rng_mode, const int32_t *walks_hd, int flags, void *stream,
                     subg_spg **out) {
    return gset_sample_impl(reinterpret_cast<const Graph *>(g), seeds_hd, n, 0, n, num_walks, num_steps, bucket, seed,
                            rng_mode, walks_hd, flags, (cudaStream_t)stream, reinterpret_cast<SpG **>(out));
}

int subg_gset_sample_shard(const subg_graph *g, const int32_t *seeds_hd, int64_t n_all, int64_t lo, int64_t hi,
                           int num_walks, int num_steps, int bucket, uint64_t seed, int rng_mode,
                           const int32_t *walks_hd, int flags, void *stream, subg_spg **out) {
    return gset_sample_impl(reinterpret_cast<const Graph *>(g), seeds_hd, n_all, lo, hi, num_walks, num_steps, bucket,
                            seed, rng_mode, walks_hd, flags, (cudaStream_t)stream, reinterpret_cast<SpG **>(out));
}

int subg_spg_set_lp_table(subg_spg *s, const int32_t *id_map_hd, const int16_t *enc_hd, int32_t c_new, int32_t ncol,
                          void *stream) {
    return spg_set_lp_table_impl(reinterpret_cast<SpG *>(s), id_map_hd, enc_hd, c_new, ncol, (cudaStream_t)stream);
}

int subg_spg_info(const subg_spg *s_, int64_t *n, int64_t *T, int32_t *c, int32_t *ncol, int32_t *max_set,
                  uint32_t *status, int32_t *value_kind) {
    const SpG *s = reinterpret_cast<const SpG *>(s_);
    if (!s) return fail(SUBG_ERR_ARG, "null SpG");
    if (n) *n = s->n;
    if (T) *T = s->T;
    if (c) *c = s->c;
    if (ncol) *ncol = s->ncol;
    if (max_set) *max_set = s->max_set;
    if (status) *status = s->status;
    if (value_kind) *value_kind = s->value_kind;
    return SUBG_OK;
}

int subg_spg_export(const subg_spg *s, int32_t *nsize_hd, int32_t *remap_hd, int16_t *enc_hd, int16_t *raw_enc_hd,
                    void *stream) {
    return spg_export_impl(reinterpret_cast<const SpG *>(s), nsize_hd, remap_hd, enc_hd, raw_enc_hd, (cudaStream_t)stream);
}

int subg_spg_rows(const subg_spg *s_, const int64_t **rowbeg, const int32_t **nsize, const int32_t **indices,
                  const void **data, int64_t *extent) {
    const SpG *s = reinterpret_cast<const SpG *>(s_);
    if (!s) return fail(SUBG_ERR_ARG, "null SpG");
    if (rowbeg) *rowbeg = s->rowbeg;
    if (nsize) *nsize = s->nsize;
    if (indices) *indices = s->indices;
    if (data) *data = s->data;
    if (extent) *extent = s->extent;
    return SUBG_OK;
}

int subg_spg_enc(const subg_spg *s_, void *stream, const int16_t **enc) {
    const SpG *s = reinterpret_cast<const SpG *>(s_);
    if (!s || !enc) return fail(SUBG_ERR_ARG, "null SpG");
    DeviceGuard guard(s->device);
    s->tag.use_on((cudaStream_t)stream);
    *enc = s->enc;
    return SUBG_OK;
}

int subg_spg_expand_rows(subg_spg *s, int64_t num_nodes, void *stream) {
    return spg_expand_rows_impl(reinterpret_cast<SpG *>(s), num_nodes, (cudaStream_t)stream);
}

int subg_spg_walks(const subg_spg *s_, void *stream, const int32_t **walks) {
    const SpG *s = reinterpret_cast<const SpG *>(s_);
    if (!s || !walks) return fail(SUBG_ERR_ARG, "null SpG");
    DeviceGuard guard(s->device);
    s->tag.use_on((cudaStream_t)stream);
    *walks = s->walks;
    return SUBG_OK;
}

int subg_spg_views(subg_spg *s_, void *stream, const int64_t **indptr, const int32_t **indices, const void **data,
                   const uint16_t **slot, const int16_t **enc, const int32_t **nsize) {
    SpG *s = reinterpret_cast<SpG *>(s_);
    if (!s) return fail(SUBG_ERR_ARG, "null SpG");
    if (int rc = spg_ensure_csr(s, (cudaStream_t)stream)) return rc;
    if (indptr) *indptr = s->indptr;
    if (indices) *indices = s->indices;
    if (data) *data = s->data;
    if (slot) *slot = s->slot;
    if (enc) *enc = s->enc;
    if (nsize) *nsize = s->nsize;
    return SUBG_OK;
}

int subg_spg_from_csr(const int64_t *indptr_hd, const int32_t *indices_hd, const void *data_hd, int value_kind,
                      int64_t n_rows, int64_t nnz, int device, void *stream, subg_spg **out) {
    if (int rc = init_device(device)) return rc;
    return spg_from_csr_impl(indptr_hd, indices_hd, data_hd, value_kind, n_rows, nnz, device, (cudaStream_t)stream,
                             reinterpret_cast<SpG **>(out));
}

int subg_spg_alloc(int64_t n, int64_t T, int device, void *stream, subg_spg **out) {
    if (int rc = init_device(device)) return rc;
    return spg_alloc_impl(n, T, device, (cudaStream_t)stream, reinterpret_cast<SpG **>(out));
}
int subg_spg_seal(subg_spg *s, void *stream) { return spg_seal_impl(reinterpret_cast<SpG *>(s), (cudaStream_t)stream); }

void subg_spg_free(subg_spg *s) { spg_free_impl(reinterpret_cast<SpG *>(s)); }

int subg_xchg_create(int device, int rank, int world, int64_t slab_bytes, subg_xchg **out) {
    if (int rc = init_device(device)) return rc;
    return xchg_create_impl(device, rank, world, slab_bytes, reinterpret_cast<Xchg **>(out));
}
int subg_xchg_export(const subg_xchg *x, void *handle64) { return xchg_export_impl(reinterpret_cast<const Xchg *>(x), handle64); }
int subg_xchg_open(subg_xchg *x, const void *handles) { return xchg_open_impl(reinterpret_cast<Xchg *>(x), handles); }
int subg_xchg_slab(const subg_xchg *x, void **slab_dev, int64_t *bytes) {
    return xchg_slab_impl(reinterpret_cast<const Xchg *>(x), slab_dev, bytes);
}
int subg_xchg_pack(subg_xchg *x, const subg_spg *shard, int64_t num_nodes, int64_t *header, void *stream) {
    return xchg_pack_impl(reinterpret_cast<Xchg *>(x), reinterpret_cast<const SpG *>(shard), num_nodes, header, (cudaStream_t)stream);
}
int subg_xchg_stage(subg_xchg *x, const subg_spg *shard, int64_t num_nodes, int64_t plane_entries, int64_t *header, void *stream) {
    return xchg_stage_impl(reinterpret_cast<Xchg *>(x), reinterpret_cast<const SpG *>(shard), num_nodes, plane_entries, header,
                           (cudaStream_t)stream);
}
int subg_xchg_link(subg_xchg *x, const int64_t *headers, const void *const *srcs, int num_walks, int ncol, void *stream,
                   subg_spg **out) {
    return xchg_link_impl(reinterpret_cast<Xchg *>(x), headers, srcs, num_walks, ncol, (cudaStream_t)stream,
                          reinterpret_cast<SpG **>(out));
}
int subg_xchg_assemble(subg_xchg *x, const int64_t *headers, const void *const *srcs, int num_walks, int ncol, void *stream,
                       subg_spg **out) {
    return xchg_assemble_impl(reinterpret_cast<Xchg *>(x), headers, srcs, num_walks, ncol, (cudaStream_t)stream,
                              reinterpret_cast<SpG **>(out));
}
void subg_xchg_free(subg_xchg *x) { xchg_free_impl(reinterpret_cast<Xchg *>(x)); }

int subg_spjoin_plan(const subg_spg *s, const int64_t *edge_hd, int64_t B, int arity, int64_t *edge_dev,
                     int64_t *indptr_dev, int64_t *N_out, void *stream) {
    return spjoin_plan_impl(reinterpret_cast<const SpG *>(s), edge_hd, B, arity, edge_dev, indptr_dev, N_out,
                            (cudaStream_t)stream);
}

int subg_spjoin_run(const subg_spg *s, const int64_t *edge_dev, int64_t B, int arity, const int64_t *indptr_dev,
                    const float *enc_table_dev, int k, void *out_dev, int64_t *segid_dev, void *stream) {
    return spjoin_run_impl(reinterpret_cast<const SpG *>(s), edge_dev, B, arity, indptr_dev, enc_table_dev, k, out_dev,
                           segid_dev, (cudaStream_t)stream);
}

int subg_spjoin(const subg_spg *s, const int64_t *edge_hd, int64_t B, int arity, int64_t *edge_dev, int64_t *indptr_dev,
                const float *enc_table_dev, int k, void *out_dev, int64_t out_capacity, int64_t *segid_dev, int64_t *N_out,
                int *ran, void *stream) {
    return spjoin_fused_impl(reinterpret_cast<const SpG *>(s), edge_hd, B, arity, edge_dev, indptr_dev, enc_table_dev, k,
                             out_dev, out_capacity, segid_dev, N_out, ran, (cudaStream_t)stream);
}

int subg_joiner_create(const subg_spg *s, int64_t B, int arity, const float *enc_table_dev, int k, int64_t capacity_rows,
                       int want_segid, int depth, subg_joiner **out) {
    return joiner_create_impl(reinterpret_cast<const SpG *>(s), B, arity, enc_table_dev, k, capacity_rows, want_segid, depth,
                              reinterpret_cast<Joiner **>(out));
}
int subg_joiner_submit(subg_joiner *j, const int64_t *edge_hd, int edge_on_device, void *stream, void **out_dev,
                       int64_t **indptr_dev, int64_t **segid_dev, const int64_t **nrows_dev, int *slot) {
    return joiner_submit_impl(reinterpret_cast<Joiner *>(j), edge_hd, edge_on_device, (cudaStream_t)stream, out_dev, indptr_dev,
                              segid_dev, nrows_dev, slot);
}
int subg_joiner_rows(subg_joiner *j, int slot, int64_t *N) { return joiner_rows_impl(reinterpret_cast<Joiner *>(j), slot, N); }
void subg_joiner_free(subg_joiner *j) { joiner_free_impl(reinterpret_cast<Joiner *>(j)); }

int subg_ppr_topk(const subg_graph *g, const int32_t *seeds_hd, int64_t n, float alpha, float eps, int topk,
                  int normalization, const double *norm_deg_hd, int encoder, void *stream, subg_spg **out) {
    if (encoder < SUBG_ENCODER_NONE || encoder > SUBG_ENCODER_SPD) return fail(SUBG_ERR_UNSUPPORTED, "unknown encoder");
    SpG *raw = nullptr;
    int rc = ppr_topk_impl(reinterpret_cast<const Graph *>(g), seeds_hd, n, alpha, eps, topk, normalization, norm_deg_hd,
                           (cudaStream_t)stream, &raw);
    if (rc != SUBG_OK) return rc;
    if (encoder == SUBG_ENCODER_NONE) {
        *out = reinterpret_cast<subg_spg *>(raw);
        return SUBG_OK;
    }
    SpG *enc = nullptr;
    rc = spg_encode_impl(reinterpret_cast<const Graph *>(g), raw, encoder, (cudaStream_t)stream, &enc);
    if (rc == SUBG_OK) {
        enc->pushes = raw->pushes;
        enc->status |= raw->status;
        *out = reinterpret_cast<subg_spg *>(enc);
    }
    spg_free_impl(raw);
    return rc;
}

int subg_spg_encode(const subg_graph *g, const subg_spg *x, int encoder, void *stream, subg_spg **out) {
    return spg_encode_impl(reinterpret_cast<const Graph *>(g), reinterpret_cast<const SpG *>(x), encoder,
                           (cudaStream_t)stream, reinterpret_cast<SpG **>(out));
}

int subg_spg_pushes(const subg_spg *s_, int64_t *pushes) {
    const SpG *s = reinterpret_cast<const SpG *>(s_);
    if (!s || !pushes) return fail(SUBG_ERR_ARG, "null SpG");
    *pushes = s->pushes;
    return SUBG_OK;
}

int subg_timing_enable(int enable) {
    g_timing = enable != 0;
    return SUBG_OK;
}
int subg_timing_read(int which, double *ms, int64_t *launches) {
    std::lock_guard<std::mutex> lk(g_tm_mutex);
    double tot = 0;
    int64_t cnt = 0;
    std::vector<TimedRegion> keep;
    for (auto &r : g_regions) {
        if (r.which != which || !r.closed) {
            keep.push_back(r);
            continue;
        }
        float t = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&t, r.a, r.b);
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
        tot += t;
        cnt++;
    }
    g_regions.swap(keep);
    if (ms) *ms = tot;
    if (launches) *launches = cnt;
    return SUBG_OK;
}
int64_t subg_launch_count(void) { return g_launches.load(); }

int subg_walk_sample(const subg_graph *g, const int32_t *seeds_hd, int64_t n, int num_walks, int num_steps,
                     uint64_t seed, int rng_mode, int replacement, void *stream, subg_walkset **out) {
    return walk_sample_impl(reinterpret_cast<const Graph *>(g), seeds_hd, n, num_walks, num_steps, seed, rng_mode,
                            replacement, (cudaStream_t)stream, reinterpret_cast<WalkSet **>(out));
}
int subg_walkset_info(const subg_walkset *w, int64_t *n, int64_t *T, int32_t *num_walks, int32_t *ncol, uint32_t *status) {
    return walkset_info_impl(reinterpret_cast<const WalkSet *>(w), n, T, num_walks, ncol, status);
}
int subg_walkset_export(const subg_walkset *w, int32_t *walks_hd, int64_t *off_hd, int32_t *ids_hd, int32_t *rpe_hd,
                        void *stream) {
    return walkset_export_impl(reinterpret_cast<const WalkSet *>(w), walks_hd, off_hd, ids_hd, rpe_hd, (cudaStream_t)stream);
}
int subg_walkset_views(const subg_walkset *w, const int32_t **walks, const int64_t **off, const int32_t **ids,
                       const int32_t **rpe) {
    return walkset_views_impl(reinterpret_cast<const WalkSet *>(w), walks, off, ids, rpe);
}
void subg_walkset_free(subg_walkset *w) { walkset_free_impl(reinterpret_cast<WalkSet *>(w)); }
int subg_walk_join(const int32_t *walks_hd, int64_t n, int64_t stride, const int64_t *key_off_hd, const int32_t *key_ids_hd,
                   const int32_t *query_hd, int64_t Q, int32_t *out_hd, int32_t *xq_hd, int device, void *stream) {
    if (int rc = init_device(device)) return rc;
    return walk_join_impl(walks_hd, n, stride, key_off_hd, key_ids_hd, query_hd, Q, out_hd, xq_hd, device, (cudaStream_t)stream);
}

int subg_batch_sample(const subg_graph *g, const int32_t *seeds_hd, int64_t n, int num_walks, int num_steps, int thld,
                      uint32_t rng_state, int32_t *out_hd, int64_t capacity, int64_t *count_out, void *stream) {
    return batch_sample_impl(reinterpret_cast<const Graph *>(g), seeds_hd, n, num_walks, num_steps, thld, rng_state, out_hd, capacity,
                             count_out, (cudaStream_t)stream);
}

int64_t subg_trim_cache(void) { return (int64_t)big_trim(); }

int subg_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(SUBG_ERR_ARG, "bad host allocation request");
    cudaError_t e = cudaHostAlloc(ptr, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(SUBG_ERR_MEM, cudaGetErrorString(e));
    return SUBG_OK;
}
void subg_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
