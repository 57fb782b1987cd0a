// Multi-GPU exchange of SpG shards over NVLink peer memory (the reference is single process).
//
// One process per GPU; the graph is replicated, every rank samples a contiguous seed range
// (subg_gset_sample_shard) and ends with the FULL SpG, joinable locally.  The exchange:
//
//   pack      the shard's rows (8 B per entry: node id + LP id) are packed to 4 / 5 / 6 / 8 bytes per entry
//             (word = node | low id bits << node bits, the remaining id bits in a byte / short / word plane) into the
//             rank's *slab*: one cudaMalloc'd buffer whose IPC handle every peer has opened, so it is mapped into all
//             the processes of the box and reachable with plain loads over NVLink.  The slab also carries the shard's
//             set sizes, row offsets and its unique LP keys with their first stream positions.
//   (barrier) the ranks all-gather an 8-word header (sizes, format) -- the only host-visible collective.
//   merge     every rank reads the peers' unique LP keys (a few thousand 16-byte records) and rebuilds the global
//             first-occurrence ids (subg_acc.c:957-978): positions are global, so inserting all keys with an atomic
//             minimum on the position and ranking the table gives the ids of the single-process scan.  The local id ->
//             global id maps stay on the device.
//   pull      ONE kernel streams every peer's packed entries over NVLink (16-byte loads from the mapped slabs, all
//             peers at once), widens them, relabels the LP ids through the maps and writes indices / data / row offsets
//             of the full SpG in place: transfer, unpack and relabel are one pass, nothing is staged.
// The full SpG is "scattered" (row u = [rowbeg[u], rowbeg[u] + nsize[u])): region r of indices / data holds rank r's
// rows at the offsets its sampler cursor gave them, which SpJoin reads in place.
//
// The same pack / merge / pull kernels also run on slabs that were brought over by an NCCL all-gather into a local
// staging buffer (`srcs` of subg_xchg_assemble): the fallback when peer mapping is unavailable, and the comparison point.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace subg {

constexpr int kMaxWorld = 16;
constexpr uint64_t kXEmptyKey = ~0ull;

enum { H_N = 0, H_T = 1, H_EXTENT = 2, H_C = 3, H_FMT = 4, H_MAXSET = 5, H_STATUS = 6, H_BYTES = 7 };
// H_FMT = extra bytes per entry (0, 1, 2, 4) | node bits << 8; negative = the pack failed on that rank

struct SlabLayout {
    int64_t key, pos, nsize, rowbeg, word, extra, end;
};
static SlabLayout slab_layout(int64_t n, int64_t extent, int64_t c, int extra_bytes) {
    auto up = [](int64_t x) { return (x + 127) & ~127ll; };
    SlabLayout L;
    int64_t o = 128;  // the first line is left for the header copy
    const int64_t e4 = (extent + 3) & ~3ll;
    L.key = o; o = up(o + 8 * c);
    L.pos = o; o = up(o + 8 * c);
    L.nsize = o; o = up(o + 4 * n);
    L.rowbeg = o; o = up(o + 8 * n);
    L.word = o; o = up(o + 4 * e4);
    L.extra = o; o = up(o + (int64_t)extra_bytes * e4);
    L.end = o;
    return L;
}
static int ceil_log2_u64(uint64_t x) {
    int b = 0;
    while ((1ull << b) < x) b++;
    return b;
}

struct Xchg {
    StreamTag tag;
    int device = 0, rank = 0, world = 1;
    int64_t slab_bytes = 0;
    unsigned char *slab = nullptr;
    bool opened = false;
    unsigned char *peer[kMaxWorld] = {};
    long long *host_words = nullptr;  // pinned: results read back at the end of assemble
};

// ------------------------------------------------------------------ pack
__device__ __forceinline__ uint32_t clamp_id(int32_t d, uint32_t c) {
    const uint32_t id = (uint32_t)(d - 1);
    return id < c ? id : 0u;  // row padding and the tail of the last vector
}

template <int E>
__global__ void xchg_pack_kernel(const int32_t *__restrict__ indices, const int32_t *__restrict__ data, int64_t e4, int nb,
                                 uint32_t c, uint32_t *__restrict__ word, void *__restrict__ extra) {
    const uint32_t nmask = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
    const int low = 32 - nb;  // id bits that ride in the word (E < 4)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (e4 >> 2); i += (int64_t)gridDim.x * blockDim.x) {
        const int4 nd = ((const int4 *)indices)[i];
        const int4 dv = ((const int4 *)data)[i];
        const uint32_t node[4] = {(uint32_t)nd.x & nmask, (uint32_t)nd.y & nmask, (uint32_t)nd.z & nmask, (uint32_t)nd.w & nmask};
        const uint32_t id[4] = {clamp_id(dv.x, c), clamp_id(dv.y, c), clamp_id(dv.z, c), clamp_id(dv.w, c)};
        uint32_t w[4], x[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (E == 4) {
                w[q] = (uint32_t)(q == 0 ? nd.x : q == 1 ? nd.y : q == 2 ? nd.z : nd.w);
                x[q] = id[q];
            } else {
                w[q] = node[q] | (low < 32 ? (id[q] << nb) : 0u);
                x[q] = low < 32 ? (id[q] >> low) : 0u;
            }
        }
        ((uint4 *)word)[i] = make_uint4(w[0], w[1], w[2], w[3]);
        if (E == 1) ((uchar4 *)extra)[i] = make_uchar4((unsigned char)x[0], (unsigned char)x[1], (unsigned char)x[2], (unsigned char)x[3]);
        if (E == 2) ((ushort4 *)extra)[i] = make_ushort4((unsigned short)x[0], (unsigned short)x[1], (unsigned short)x[2], (unsigned short)x[3]);
        if (E == 4) ((uint4 *)extra)[i] = make_uint4(x[0], x[1], x[2], x[3]);
    }
}

__global__ void xchg_rows_kernel(const long long *__restrict__ rowbeg, const int32_t *__restrict__ nsize, int64_t n,
                                 long long *__restrict__ out_rowbeg, int32_t *__restrict__ out_nsize) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        out_rowbeg[i] = rowbeg[i];
        out_nsize[i] = nsize[i];
    }
}

// ------------------------------------------------------------------ merge of the unique LP tables
struct MergeArgs {
    const unsigned char *src[kMaxWorld];
    int64_t key_off[kMaxWorld], pos_off[kMaxWorld];
    int32_t coff[kMaxWorld + 1];
    int world;
};
__device__ __forceinline__ uint32_t xlp_hash(unsigned long long key) {
    uint32_t h = (uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}
__device__ __forceinline__ int merge_owner(const MergeArgs &a, int g) {
    int r = 0;
    while (r + 1 < a.world && g >= a.coff[r + 1]) r++;
    return r;
}
// loads from mapped peer memory: read once, keep them out of the local L1
__device__ __forceinline__ unsigned long long ld_peer_u64(const void *p) {
    unsigned long long v;
    asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_peer_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld_peer_v2(const void *p) {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_peer_u32(const void *p) {
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__global__ void fill_u64x2_kernel(unsigned long long *a, unsigned long long *b, int64_t n, unsigned long long va, unsigned long long vb) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        a[i] = va;
        b[i] = vb;
    }
}

__global__ void merge_insert_kernel(const MergeArgs a, unsigned long long *tab_key, unsigned long long *tab_pos, uint32_t mask) {
    const int total = a.coff[a.world];
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        const int r = merge_owner(a, g);
        const int j = g - a.coff[r];
        const unsigned long long key = ld_peer_u64(a.src[r] + a.key_off[r] + 8ll * j);
        const unsigned long long pos = ld_peer_u64(a.src[r] + a.pos_off[r] + 8ll * j);
        uint32_t h = xlp_hash(key) & mask;
        for (;;) {  // the table has at least twice as many slots as keys
            unsigned long long cur = tab_key[h];
            if (cur == kXEmptyKey) {
                cur = atomicCAS(&tab_key[h], kXEmptyKey, key);
                if (cur == kXEmptyKey) cur = key;
            }
            if (cur == key) break;
            h = (h + 1) & mask;
        }
        atomicMin(&tab_pos[h], pos);
    }
}
// gmap[coff[r] + j] = global id + 1 of rank r's local id j
__global__ void merge_map_kernel(const MergeArgs a, const unsigned long long *tab_key, uint32_t mask, const int32_t *rank_of_slot,
                                 int32_t *gmap) {
    const int total = a.coff[a.world];
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        const int r = merge_owner(a, g);
        const unsigned long long key = ld_peer_u64(a.src[r] + a.key_off[r] + 8ll * (g - a.coff[r]));
        uint32_t h = xlp_hash(key) & mask;
        while (tab_key[h] != key) h = (h + 1) & mask;
        gmap[g] = rank_of_slot[h] + 1;
    }
}

// ------------------------------------------------------------------ pull: peers' packed entries -> the full SpG
struct PullArgs {
    const unsigned char *src[kMaxWorld];
    int64_t word_off[kMaxWorld], extra_off[kMaxWorld], nsize_off[kMaxWorld], rowbeg_off[kMaxWorld];
    int64_t e4[kMaxWorld];        // entries of the region, rounded up to 4
    int64_t dst_off[kMaxWorld];   // first entry of the region in the full arrays
    int64_t n[kMaxWorld], row_off[kMaxWorld];
    int32_t gmap_off[kMaxWorld];
    int32_t nb[kMaxWorld], extra[kMaxWorld];
    int world, rank;
    int32_t *indices;
    int32_t *data;
    long long *rowbeg;
    int32_t *nsize;
    const int32_t *gmap;
    unsigned long long *ticket;   // nullable: [world] chunk counters (windowed pull)
};

constexpr int kPullThreads = 256;
constexpr int kPullUnroll = 4;
constexpr int64_t kPullChunk = 4096;   // 16-byte words per ticket (64 KB of packed entries)

template <int E>
__device__ __forceinline__ void pull_region(const PullArgs &a, int r, int64_t first, int64_t step, int64_t limit) {
    const uint4 *word = (const uint4 *)(a.src[r] + a.word_off[r]);
    const unsigned char *extra = a.src[r] + a.extra_off[r];
    const int nb = a.nb[r];
    const uint32_t nmask = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
    const int low = 32 - nb;
    const int32_t *gmap = a.gmap + a.gmap_off[r];
    int4 *out_i = (int4 *)(a.indices + a.dst_off[r]);
    int4 *out_d = (int4 *)(a.data + a.dst_off[r]);
    const int64_t n4 = limit;
    for (int64_t i0 = first; i0 < n4; i0 += step * kPullUnroll) {
        uint4 w[kPullUnroll], x[kPullUnroll];
#pragma unroll
        for (int u = 0; u < kPullUnroll; u++) {  // all loads of the group are in flight before the first is used
            const int64_t i = i0 + u * step;
            if (i < n4) {
                w[u] = ld_peer_v4(word + i);
                if (E == 1) {
                    const uint32_t b = ld_peer_u32(extra + 4 * i);
                    x[u] = make_uint4(b & 0xffu, (b >> 8) & 0xffu, (b >> 16) & 0xffu, b >> 24);
                } else if (E == 2) {
                    const uint2 b = ld_peer_v2(extra + 8 * i);
                    x[u] = make_uint4(b.x & 0xffffu, b.x >> 16, b.y & 0xffffu, b.y >> 16);
                } else if (E == 4) {
                    x[u] = ld_peer_v4(extra + 16 * i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kPullUnroll; u++) {
            const int64_t i = i0 + u * step;
            if (i < n4) {
                const uint32_t ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
                const uint32_t xx[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                int32_t node[4], gid[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t id;
                    if (E == 4) {
                        node[q] = (int32_t)ww[q];
                        id = xx[q];
                    } else {
                        node[q] = (int32_t)(ww[q] & nmask);
                        id = low < 32 ? (ww[q] >> nb) : 0u;
                        if (E > 0) id |= xx[q] << low;
                    }
                    gid[q] = __ldg(gmap + id);
                }
                out_i[i] = make_int4(node[0], node[1], node[2], node[3]);
                out_d[i] = make_int4(gid[0], gid[1], gid[2], gid[3]);
            }
        }
    }
}

// grid = world x blocks_per_region; block b works on region (rank + b % world) % world, so every GPU reads from all of
// its peers at once and no two GPUs start on the same peer.  Two schedules inside a region:
//   grid-stride   block j of the region takes words j, j + nblk, ... (in units of a block's 4 KB): all blocks advance in
//                 lock step only as long as they run at the same speed;
//   windowed      (a.ticket) blocks take 64 KB chunks of the region in order from a counter: however far the blocks drift
//                 apart in time, the addresses in flight stay inside a window of nblk x 64 KB per region -- at the twitter
//                 size (2.5 GB per region, 20 GB per GPU) that is what keeps the peer mappings inside the TLB reach.
__global__ void __launch_bounds__(kPullThreads) xchg_pull_kernel(const PullArgs a) {
    const int r = (a.rank + (int)(blockIdx.x % a.world)) % a.world;
    const int64_t jb = blockIdx.x / a.world, nblk = gridDim.x / a.world;
    const int64_t first = jb * kPullThreads + threadIdx.x, step = nblk * kPullThreads;
    const int64_t n4 = a.e4[r] >> 2;
    if (a.ticket) {
        __shared__ unsigned long long s_chunk;
        for (;;) {
            if (threadIdx.x == 0) s_chunk = atomicAdd(a.ticket + r, 1ull);
            __syncthreads();
            const int64_t lo = (int64_t)s_chunk * kPullChunk;
            __syncthreads();
            if (lo >= n4) break;
            const int64_t hi = lo + kPullChunk < n4 ? lo + kPullChunk : n4;
            switch (a.extra[r]) {
                case 0: pull_region<0>(a, r, lo + threadIdx.x, kPullThreads, hi); break;
                case 1: pull_region<1>(a, r, lo + threadIdx.x, kPullThreads, hi); break;
                case 2: pull_region<2>(a, r, lo + threadIdx.x, kPullThreads, hi); break;
                default: pull_region<4>(a, r, lo + threadIdx.x, kPullThreads, hi); break;
            }
        }
    } else {
        switch (a.extra[r]) {
            case 0: pull_region<0>(a, r, first, step, n4); break;
            case 1: pull_region<1>(a, r, first, step, n4); break;
            case 2: pull_region<2>(a, r, first, step, n4); break;
            default: pull_region<4>(a, r, first, step, n4); break;
        }
    }
    // row offsets and set sizes of the region
    const long long *rb = (const long long *)(a.src[r] + a.rowbeg_off[r]);
    const int32_t *ns = (const int32_t *)(a.src[r] + a.nsize_off[r]);
    for (int64_t i = first; i < a.n[r]; i += step) {
        a.rowbeg[a.row_off[r] + i] = (long long)ld_peer_u64(rb + i) + a.dst_off[r];
        a.nsize[a.row_off[r] + i] = (int32_t)ld_peer_u32(ns + i);
    }
}

// ------------------------------------------------------------------ host side
// blocks of the pull kernel per region: together about 8 CTAs per SM, 2 per SM for large exchanges (measured at the twitter
// size on 8 GPUs: 1184 blocks pull 203 GB/s per GPU, 296 blocks 444 GB/s; at the ppa size 1184 blocks reach 616 GB/s;
// profiles/r2p_exchange_twitter8.txt).  SUBG_XCHG_BLOCKS overrides the total.
static int64_t env_blocks_per_region(int num_sms, int world, bool large) {
    int64_t total = (large ? 2ll : 8ll) * num_sms;
    if (const char *v = getenv("SUBG_XCHG_BLOCKS")) total = std::max<int64_t>(atoll(v), 1);
    return (total + world - 1) / world;
}

int xchg_create_impl(int device, int rank, int world, int64_t slab_bytes, Xchg **out) {
    if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || slab_bytes < 0)
        return fail(SUBG_ERR_ARG, "exchange context: need 0 <= rank < world <= 16");
    DeviceGuard guard(device);
    Xchg *x = new Xchg();
    x->device = device; x->rank = rank; x->world = world;
    x->slab_bytes = std::max<int64_t>((slab_bytes + 4095) & ~4095ll, 4096);
    // cudaMalloc, not the stream-ordered pool: the allocation has to be exportable with cudaIpcGetMemHandle
    cudaError_t e = cudaMalloc((void **)&x->slab, (size_t)x->slab_bytes);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&x->host_words, 8 * sizeof(long long), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        if (x->slab) cudaFree(x->slab);
        delete x;
        return fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA, cudaGetErrorString(e));
    }
    x->peer[rank] = x->slab;
    *out = x;
    return SUBG_OK;
}

int xchg_export_impl(const Xchg *x, void *handle64) {
    if (!x || !handle64) return fail(SUBG_ERR_ARG, "null exchange context");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard(x->device);
    cudaIpcMemHandle_t h;
    SUBG_CUDA(cudaIpcGetMemHandle(&h, x->slab));
    memcpy(handle64, &h, sizeof(h));
    return SUBG_OK;
}

int xchg_open_impl(Xchg *x, const void *handles) {
    if (!x || !handles) return fail(SUBG_ERR_ARG, "null exchange context");
    DeviceGuard guard(x->device);
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)handles + 64 * r, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; q++)
                if (q != x->rank && x->peer[q]) { cudaIpcCloseMemHandle(x->peer[q]); x->peer[q] = nullptr; }
            return fail(SUBG_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        }
        x->peer[r] = (unsigned char *)p;
    }
    x->opened = true;
    return SUBG_OK;
}

int xchg_slab_impl(const Xchg *x, void **slab_dev, int64_t *bytes) {
    if (!x) return fail(SUBG_ERR_ARG, "null exchange context");
    if (slab_dev) *slab_dev = x->slab;
    if (bytes) *bytes = x->slab_bytes;
    return SUBG_OK;
}

// shard -> slab.  header[8] (host) describes what was written; a shard that does not fit reports H_FMT < 0 (the ranks
// see it after the header all-gather and fail together instead of hanging in a collective).
int xchg_pack_impl(Xchg *x, const SpG *s, int64_t n_nodes, int64_t *header, cudaStream_t st) {
    if (!x || !s || !header || n_nodes < 1) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (s->value_kind != 0 || s->device != x->device) return fail(SUBG_ERR_ARG, "exchange takes a sampler-built LP shard on the context's device");
    if (s->c > 0 && (!s->lp_key || !s->lp_pos)) return fail(SUBG_ERR_ARG, "shard has no LP keys (not built by subg_gset_sample_shard)");
    DeviceGuard guard(x->device);
    x->tag.use_on(st);
    s->tag.use_on(st);
    const int nb = std::max(1, ceil_log2_u64((uint64_t)n_nodes));
    const int cb = ceil_log2_u64((uint64_t)std::max(s->c, 1));
    int E = 4;
    if (nb + cb <= 32) E = 0;
    else if (nb + cb <= 40) E = 1;
    else if (nb + cb <= 48) E = 2;
    if (const char *v = getenv("SUBG_XCHG_EXTRA")) {  // experiment knob: force a wider format
        const int f = atoi(v);
        if ((f == 1 || f == 2 || f == 4) && f > E) E = f;
    }
    const SlabLayout L = slab_layout(s->n, s->extent, s->c, E);
    header[H_N] = s->n; header[H_T] = s->T; header[H_EXTENT] = s->extent; header[H_C] = s->c;
    header[H_FMT] = E | (nb << 8); header[H_MAXSET] = s->max_set; header[H_STATUS] = s->status; header[H_BYTES] = L.end;
    if (L.end > x->slab_bytes || s->extent + 4 > s->cap) {
        header[H_FMT] = -1;
        return SUBG_OK;
    }
    const int64_t e4 = (s->extent + 3) & ~3ll;
    if (e4 > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((e4 / 4 + 255) / 256, 16 * 148);
        uint32_t *word = (uint32_t *)(x->slab + L.word);
        void *extra = x->slab + L.extra;
        const int32_t *ind = s->indices, *dat = (const int32_t *)s->data;
        switch (E) {
            case 0: xchg_pack_kernel<0><<<blocks, 256, 0, st>>>(ind, dat, e4, nb, (uint32_t)s->c, word, extra); break;
            case 1: xchg_pack_kernel<1><<<blocks, 256, 0, st>>>(ind, dat, e4, nb, (uint32_t)s->c, word, extra); break;
            case 2: xchg_pack_kernel<2><<<blocks, 256, 0, st>>>(ind, dat, e4, nb, (uint32_t)s->c, word, extra); break;
            default: xchg_pack_kernel<4><<<blocks, 256, 0, st>>>(ind, dat, e4, nb, (uint32_t)s->c, word, extra); break;
        }
        count_launch(1);
    }
    if (s->n > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((s->n + 255) / 256, 4 * 148);
        xchg_rows_kernel<<<blocks, 256, 0, st>>>((const long long *)s->rowbeg, s->nsize, s->n, (long long *)(x->slab + L.rowbeg),
                                                 (int32_t *)(x->slab + L.nsize));
        count_launch(1);
    }
    if (s->c > 0) {
        SUBG_CUDA(cudaMemcpyAsync(x->slab + L.key, s->lp_key, (size_t)s->c * 8, cudaMemcpyDeviceToDevice, st));
        SUBG_CUDA(cudaMemcpyAsync(x->slab + L.pos, s->lp_pos, (size_t)s->c * 8, cudaMemcpyDeviceToDevice, st));
    }
    SUBG_CUDA(cudaGetLastError());
    return SUBG_OK;
}

// headers: int64[world, 8] (host) as written by every rank's pack.  srcs: host array of `world` device pointers to the
// slabs as this GPU reaches them (NULL = the peers opened with subg_xchg_open).  Synchronises the stream at the end
// (the number of unique LP rows comes back from the device).
int xchg_assemble_impl(Xchg *x, const int64_t *headers, const void *const *srcs, int M, int ncol, cudaStream_t st, SpG **out) {
    if (!x || !headers || !out || ncol < 2) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (!srcs && !x->opened && x->world > 1) return fail(SUBG_ERR_ARG, "peer slabs are not mapped (subg_xchg_open) and no sources were given");
    DeviceGuard guard(x->device);
    x->tag.use_on(st);
    const int W = x->world;
    const int m = ncol - 1;
    int shift = 0;
    while ((M >> shift) != 0) shift++;
    MergeArgs ma{};
    PullArgs pa{};
    ma.world = W; pa.world = W; pa.rank = x->rank;
    int64_t n_tot = 0, T_tot = 0, ext_tot = 0, max_bytes = 0;
    int32_t c_sum = 0, max_set = 0;
    uint32_t status = 0;
    for (int r = 0; r < W; r++) {
        const int64_t *h = headers + 8 * r;
        if (h[H_FMT] < 0) return fail(SUBG_ERR_MEM, "a shard did not fit its exchange slab");
        const int E = (int)(h[H_FMT] & 0xff), nb = (int)(h[H_FMT] >> 8);
        if (h[H_N] < 0 || h[H_T] < 0 || h[H_EXTENT] < h[H_T] || h[H_C] < 0 || (E != 0 && E != 1 && E != 2 && E != 4) || nb < 1 || nb > 32)
            return fail(SUBG_ERR_ARG, "malformed exchange header");
        const SlabLayout L = slab_layout(h[H_N], h[H_EXTENT], h[H_C], E);
        const unsigned char *src = srcs ? (const unsigned char *)srcs[r] : x->peer[r];
        if (!src) return fail(SUBG_ERR_ARG, "missing slab source");
        ma.src[r] = src; ma.key_off[r] = L.key; ma.pos_off[r] = L.pos; ma.coff[r] = c_sum;
        pa.src[r] = src; pa.word_off[r] = L.word; pa.extra_off[r] = L.extra; pa.nsize_off[r] = L.nsize; pa.rowbeg_off[r] = L.rowbeg;
        pa.e4[r] = (h[H_EXTENT] + 3) & ~3ll; pa.dst_off[r] = ext_tot; pa.n[r] = h[H_N]; pa.row_off[r] = n_tot;
        pa.gmap_off[r] = c_sum; pa.nb[r] = nb; pa.extra[r] = E;
        n_tot += h[H_N]; T_tot += h[H_T]; ext_tot += pa.e4[r];
        if ((int64_t)c_sum + h[H_C] > INT32_MAX) return fail(SUBG_ERR_MEM, "too many unique LP rows");
        c_sum += (int32_t)h[H_C];
        max_set = std::max<int32_t>(max_set, (int32_t)h[H_MAXSET]);
        status |= (uint32_t)h[H_STATUS];
        max_bytes = std::max(max_bytes, h[H_BYTES]);
    }
    ma.coff[W] = c_sum;

    HostProf prof;
    auto pmark = [&](const char *what) {  // SUBG_PROFILE_HOST: per-phase wall time (adds a stream sync per phase)
        if (prof.on) { cudaStreamSynchronize(st); prof.mark(what); }
    };
    SpG *s = new SpG();
    s->tag.last = st;
    s->device = x->device; s->n = n_tot; s->T = T_tot; s->ncol = ncol; s->M = M; s->shift = shift; s->value_kind = 0;
    s->max_set = max_set; s->status = status;
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, x->device);
    unsigned long long *tab_key = nullptr, *tab_pos = nullptr;
    int32_t *rank_of_slot = nullptr, *gmap = nullptr;
    uint32_t *d_cnt = nullptr;
    unsigned long long *d_ticket = nullptr;
    int rc = SUBG_OK;
#define CKX(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            rc = fail(_e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,              \
                      std::string(#call) + ": " + cudaGetErrorString(_e));                         \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)
    {
        CKX(dmalloc(&s->indices, (size_t)ext_tot + 16, st));
        CKX(dmalloc((int32_t **)&s->data, (size_t)ext_tot + 16, st));
        CKX(dmalloc(&s->rowbeg, (size_t)std::max<int64_t>(n_tot, 1), st));
        CKX(dmalloc(&s->nsize, (size_t)std::max<int64_t>(n_tot, 1), st));
        CKX(dmalloc(&s->enc, (size_t)std::max(c_sum, 1) * ncol, st));
        CKX(dmalloc(&s->lp_key, (size_t)std::max(c_sum, 1), st));
        CKX(dmalloc(&s->lp_pos, (size_t)std::max(c_sum, 1), st));
        CKX(dmalloc(&gmap, (size_t)std::max(c_sum, 1), st));
        CKX(dmalloc(&d_cnt, 2, st));
        CKX(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(uint32_t), st));
        s->extent = ext_tot; s->cap = ext_tot + 16;
        pmark("xchg:alloc");
        if (c_sum > 0) {
            uint32_t cap = 1024;
            while (cap < 4u * (uint32_t)c_sum) cap <<= 1;
            CKX(dmalloc(&tab_key, (size_t)cap, st));
            CKX(dmalloc(&tab_pos, (size_t)cap, st));
            CKX(dmalloc(&rank_of_slot, (size_t)cap, st));
            const unsigned fb = (unsigned)std::min<int64_t>((cap + 255) / 256, 4 * s->num_sms);
            fill_u64x2_kernel<<<fb, 256, 0, st>>>(tab_key, tab_pos, cap, kXEmptyKey, ~0ull);
            const unsigned mb = (unsigned)std::min<int64_t>((c_sum + 255) / 256, 4 * s->num_sms);
            merge_insert_kernel<<<mb, 256, 0, st>>>(ma, tab_key, tab_pos, cap - 1);
            if (int urc = rank_unique_keys(tab_key, tab_pos, cap, (uint32_t)c_sum, false, M, m, shift, rank_of_slot, s->enc, s->lp_key,
                                           s->lp_pos, d_cnt, s->num_sms, st)) { rc = urc; goto done; }
            merge_map_kernel<<<mb, 256, 0, st>>>(ma, tab_key, cap - 1, rank_of_slot, gmap);
            CKX(cudaGetLastError());
            count_launch(3);
        }
        pmark("xchg:merge");
        if (ext_tot > 0 || n_tot > 0) {
            pa.indices = s->indices; pa.data = (int32_t *)s->data; pa.rowbeg = (long long *)s->rowbeg; pa.nsize = s->nsize; pa.gmap = gmap;
            // windowed schedule and fewer blocks for large exchanges (SUBG_XCHG_TICKET = 0 / 1 overrides; default: more
            // than 4 GB pulled)
            int64_t pulled = 0;
            for (int r = 0; r < W; r++) pulled += 4 * pa.e4[r];
            int per_region = std::max(1, (int)env_blocks_per_region(s->num_sms, W, pulled > (4ll << 30)));
            const char *tkv = getenv("SUBG_XCHG_TICKET");
            const int64_t tk = tkv ? atoll(tkv) : -1;
            if (tk > 0 || (tk < 0 && pulled > (4ll << 30))) {
                CKX(dmalloc(&d_ticket, (size_t)kMaxWorld, st));
                CKX(cudaMemsetAsync(d_ticket, 0, kMaxWorld * sizeof(unsigned long long), st));
                pa.ticket = d_ticket;
            }
            timing_begin(SUBG_TIMING_EXCHANGE, st);
            xchg_pull_kernel<<<(unsigned)(W * per_region), kPullThreads, 0, st>>>(pa);
            timing_end(SUBG_TIMING_EXCHANGE, st);
            CKX(cudaGetLastError());
            count_launch(1);
        }
        pmark("xchg:pull");
        CKX(cudaMemcpyAsync(x->host_words, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CKX(cudaStreamSynchronize(st));
        s->c = (int32_t)((uint32_t *)x->host_words)[0];
    }
done:
#undef CKX
    dfree(tab_key, st); dfree(tab_pos, st); dfree(rank_of_slot, st); dfree(gmap, st); dfree(d_cnt, st); dfree(d_ticket, st);
    if (rc != SUBG_OK) {
        spg_free_impl(s);
        return rc;
    }
    *out = s;
    return SUBG_OK;
}

// ------------------------------------------------------------------ linked SpG: shards stay where they were sampled
// The alternative to replicating the SpG: every rank STAGES its shard unpacked in its slab -- a plane
// of node ids and a plane of LP ids at the same two offsets in every slab -- and the ranks LINK them: the LP tables are
// merged as in assemble, every rank relabels its OWN id plane to the global ids in place, and only the row metadata
// (12 bytes per seed) crosses NVLink.  The result is an ordinary scattered SpG whose indices / data pointers are the
// first slab's planes and whose rowbeg[u] is row u's offset from there -- a 64-bit element offset that reaches into the
// peers' slabs, because all of them are mapped into this process (one address space).  SpJoin reads such an SpG unchanged:
// its TMA bulk copies and plain loads take `indices + rowbeg[u]` wherever that lies, so the rows of a query's endpoints
// come over NVLink at join time.  The pass costs the sampling of 1/N of the seeds and no bulk transfer; the joins pay
// for it (7/8 of the row bytes are remote at 8 GPUs).  The caller has to barrier after link (all id planes relabelled)
// and keep the exchange contexts alive for as long as the SpG is used.
static int64_t plane_round(int64_t e) { return (e + 63) & ~63ll; }
static int64_t plane_offset(int64_t slab_bytes, int64_t P) { return (slab_bytes - 8 * P) & ~255ll; }
constexpr int kStagedFormat = 8;   // H_FMT low byte of a staged shard; H_BYTES then holds the plane size in entries

__global__ void xchg_relabel_plane_kernel(int32_t *plane, int64_t e4, const int32_t *__restrict__ gmap, uint32_t c) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (e4 >> 2); i += (int64_t)gridDim.x * blockDim.x) {
        int4 v = ((int4 *)plane)[i];
        v.x = __ldg(gmap + clamp_id(v.x, c)); v.y = __ldg(gmap + clamp_id(v.y, c));
        v.z = __ldg(gmap + clamp_id(v.z, c)); v.w = __ldg(gmap + clamp_id(v.w, c));
        ((int4 *)plane)[i] = v;
    }
}
struct LinkArgs {
    const unsigned char *src[kMaxWorld];
    int64_t nsize_off[kMaxWorld], rowbeg_off[kMaxWorld], n[kMaxWorld], row_off[kMaxWorld];
    long long delta[kMaxWorld];   // elements from the first slab's plane to this slab's plane
    int world;
    long long *rowbeg;
    int32_t *nsize;
};
__global__ void xchg_link_rows_kernel(const LinkArgs a) {
    for (int r = 0; r < a.world; r++) {
        const long long *rb = (const long long *)(a.src[r] + a.rowbeg_off[r]);
        const int32_t *ns = (const int32_t *)(a.src[r] + a.nsize_off[r]);
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n[r]; i += (int64_t)gridDim.x * blockDim.x) {
            a.rowbeg[a.row_off[r] + i] = (long long)ld_peer_u64(rb + i) + a.delta[r];
            a.nsize[a.row_off[r] + i] = (int32_t)ld_peer_u32(ns + i);
        }
    }
}

// plane_entries: capacity of a plane in entries, the SAME on every rank (e.g. the largest seed range x the row capacity)
int xchg_stage_impl(Xchg *x, const SpG *s, int64_t n_nodes, int64_t plane_entries, int64_t *header, cudaStream_t st) {
    if (!x || !s || !header || n_nodes < 1 || plane_entries < 0) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (s->value_kind != 0 || s->device != x->device || s->borrowed) return fail(SUBG_ERR_ARG, "exchange takes a sampler-built LP shard on the context's device");
    if (s->c > 0 && (!s->lp_key || !s->lp_pos)) return fail(SUBG_ERR_ARG, "shard has no LP keys (not built by subg_gset_sample_shard)");
    DeviceGuard guard(x->device);
    x->tag.use_on(st);
    s->tag.use_on(st);
    const int nb = std::max(1, ceil_log2_u64((uint64_t)n_nodes));
    const int64_t P = plane_round(plane_entries + 4);
    const SlabLayout L = slab_layout(s->n, 0, s->c, 0);
    const int64_t PI = plane_offset(x->slab_bytes, P);
    header[H_N] = s->n; header[H_T] = s->T; header[H_EXTENT] = s->extent; header[H_C] = s->c;
    header[H_FMT] = kStagedFormat | (nb << 8); header[H_MAXSET] = s->max_set; header[H_STATUS] = s->status; header[H_BYTES] = P;
    const int64_t e4 = (s->extent + 3) & ~3ll;
    if (PI < L.end || e4 > P || s->extent + 4 > s->cap) {
        header[H_FMT] = -1;
        return SUBG_OK;
    }
    if (e4 > 0) {
        SUBG_CUDA(cudaMemcpyAsync(x->slab + PI, s->indices, (size_t)e4 * 4, cudaMemcpyDeviceToDevice, st));
        SUBG_CUDA(cudaMemcpyAsync(x->slab + PI + 4 * P, s->data, (size_t)e4 * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (s->n > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((s->n + 255) / 256, 4 * 148);
        xchg_rows_kernel<<<blocks, 256, 0, st>>>((const long long *)s->rowbeg, s->nsize, s->n, (long long *)(x->slab + L.rowbeg),
                                                 (int32_t *)(x->slab + L.nsize));
        count_launch(1);
    }
    if (s->c > 0) {
        SUBG_CUDA(cudaMemcpyAsync(x->slab + L.key, s->lp_key, (size_t)s->c * 8, cudaMemcpyDeviceToDevice, st));
        SUBG_CUDA(cudaMemcpyAsync(x->slab + L.pos, s->lp_pos, (size_t)s->c * 8, cudaMemcpyDeviceToDevice, st));
    }
    SUBG_CUDA(cudaGetLastError());
    return SUBG_OK;
}

int xchg_link_impl(Xchg *x, const int64_t *headers, const void *const *srcs, int M, int ncol, cudaStream_t st, SpG **out) {
    if (!x || !headers || !out || ncol < 2) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (!srcs && !x->opened && x->world > 1) return fail(SUBG_ERR_ARG, "peer slabs are not mapped (subg_xchg_open) and no sources were given");
    DeviceGuard guard(x->device);
    x->tag.use_on(st);
    const int W = x->world;
    const int m = ncol - 1;
    int shift = 0;
    while ((M >> shift) != 0) shift++;
    MergeArgs ma{};
    LinkArgs la{};
    ma.world = W; la.world = W;
    int64_t n_tot = 0, T_tot = 0, P = -1;
    int32_t c_sum = 0, max_set = 0;
    uint32_t status = 0;
    for (int r = 0; r < W; r++) {
        const int64_t *h = headers + 8 * r;
        if (h[H_FMT] < 0) return fail(SUBG_ERR_MEM, "a shard did not fit its exchange slab");
        if ((h[H_FMT] & 0xff) != kStagedFormat || h[H_N] < 0 || h[H_T] < 0 || h[H_EXTENT] < h[H_T] || h[H_C] < 0)
            return fail(SUBG_ERR_ARG, "malformed exchange header (link takes staged shards)");
        if (P < 0) P = h[H_BYTES];
        if (h[H_BYTES] != P) return fail(SUBG_ERR_ARG, "the ranks staged with different plane sizes");
        const SlabLayout L = slab_layout(h[H_N], 0, h[H_C], 0);
        const unsigned char *src = srcs ? (const unsigned char *)srcs[r] : x->peer[r];
        if (!src) return fail(SUBG_ERR_ARG, "missing slab source");
        ma.src[r] = src; ma.key_off[r] = L.key; ma.pos_off[r] = L.pos; ma.coff[r] = c_sum;
        la.src[r] = src; la.nsize_off[r] = L.nsize; la.rowbeg_off[r] = L.rowbeg; la.n[r] = h[H_N]; la.row_off[r] = n_tot;
        la.delta[r] = (long long)(((intptr_t)src - (intptr_t)(srcs ? (const unsigned char *)srcs[0] : x->peer[0])) / 4);
        n_tot += h[H_N]; T_tot += h[H_T];
        if ((int64_t)c_sum + h[H_C] > INT32_MAX) return fail(SUBG_ERR_MEM, "too many unique LP rows");
        c_sum += (int32_t)h[H_C];
        max_set = std::max<int32_t>(max_set, (int32_t)h[H_MAXSET]);
        status |= (uint32_t)h[H_STATUS];
    }
    ma.coff[W] = c_sum;
    const unsigned char *base = srcs ? (const unsigned char *)srcs[0] : x->peer[0];
    const int64_t PI = plane_offset(x->slab_bytes, P);
    SpG *s = new SpG();
    s->tag.last = st;
    s->device = x->device; s->n = n_tot; s->T = T_tot; s->ncol = ncol; s->M = M; s->shift = shift; s->value_kind = 0;
    s->max_set = max_set; s->status = status;
    s->borrowed = true;
    s->indices = (int32_t *)(base + PI);
    s->data = (void *)(base + PI + 4 * P);
    s->extent = T_tot; s->cap = 0;
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, x->device);
    unsigned long long *tab_key = nullptr, *tab_pos = nullptr;
    int32_t *rank_of_slot = nullptr, *gmap = nullptr;
    uint32_t *d_cnt = nullptr;
    int rc = SUBG_OK;
#define CKX(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            rc = fail(_e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,              \
                      std::string(#call) + ": " + cudaGetErrorString(_e));                         \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)
    {
        CKX(dmalloc(&s->rowbeg, (size_t)std::max<int64_t>(n_tot, 1), st));
        CKX(dmalloc(&s->nsize, (size_t)std::max<int64_t>(n_tot, 1), st));
        CKX(dmalloc(&s->enc, (size_t)std::max(c_sum, 1) * ncol, st));
        CKX(dmalloc(&s->lp_key, (size_t)std::max(c_sum, 1), st));
        CKX(dmalloc(&s->lp_pos, (size_t)std::max(c_sum, 1), st));
        CKX(dmalloc(&gmap, (size_t)std::max(c_sum, 1), st));
        CKX(dmalloc(&d_cnt, 2, st));
        CKX(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(uint32_t), st));
        if (c_sum > 0) {
            uint32_t cap = 1024;
            while (cap < 4u * (uint32_t)c_sum) cap <<= 1;
            CKX(dmalloc(&tab_key, (size_t)cap, st));
            CKX(dmalloc(&tab_pos, (size_t)cap, st));
            CKX(dmalloc(&rank_of_slot, (size_t)cap, st));
            const unsigned fb = (unsigned)std::min<int64_t>((cap + 255) / 256, 4 * s->num_sms);
            fill_u64x2_kernel<<<fb, 256, 0, st>>>(tab_key, tab_pos, cap, kXEmptyKey, ~0ull);
            const unsigned mb = (unsigned)std::min<int64_t>((c_sum + 255) / 256, 4 * s->num_sms);
            merge_insert_kernel<<<mb, 256, 0, st>>>(ma, tab_key, tab_pos, cap - 1);
            if (int urc = rank_unique_keys(tab_key, tab_pos, cap, (uint32_t)c_sum, false, M, m, shift, rank_of_slot, s->enc, s->lp_key,
                                           s->lp_pos, d_cnt, s->num_sms, st)) { rc = urc; goto done; }
            merge_map_kernel<<<mb, 256, 0, st>>>(ma, tab_key, cap - 1, rank_of_slot, gmap);
            CKX(cudaGetLastError());
            count_launch(3);
            // this rank's own id plane -> global ids, in place (the peers do the same with theirs)
            const int64_t *hme = headers + 8 * x->rank;
            const int64_t e4 = (hme[H_EXTENT] + 3) & ~3ll;
            if (e4 > 0 && hme[H_C] > 0) {
                const unsigned rb = (unsigned)std::min<int64_t>((e4 / 4 + 255) / 256, 16 * (int64_t)s->num_sms);
                xchg_relabel_plane_kernel<<<rb, 256, 0, st>>>((int32_t *)(x->slab + PI + 4 * P), e4, gmap + ma.coff[x->rank], (uint32_t)hme[H_C]);
                CKX(cudaGetLastError());
                count_launch(1);
            }
        }
        if (n_tot > 0) {
            la.rowbeg = (long long *)s->rowbeg; la.nsize = s->nsize;
            xchg_link_rows_kernel<<<4 * s->num_sms, 256, 0, st>>>(la);
            CKX(cudaGetLastError());
            count_launch(1);
        }
        CKX(cudaMemcpyAsync(x->host_words, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CKX(cudaStreamSynchronize(st));
        s->c = (int32_t)((uint32_t *)x->host_words)[0];
    }
done:
#undef CKX
    dfree(tab_key, st); dfree(tab_pos, st); dfree(rank_of_slot, st); dfree(gmap, st); dfree(d_cnt, st);
    if (rc != SUBG_OK) {
        spg_free_impl(s);
        return rc;
    }
    *out = s;
    return SUBG_OK;
}

void xchg_free_impl(Xchg *x) {
    if (!x) return;
    DeviceGuard guard(x->device);
    const cudaStream_t st = x->tag.free_stream();
    cudaStreamSynchronize(st);  // cudaFree is not stream ordered
    for (int r = 0; r < x->world; r++)
        if (r != x->rank && x->opened && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->slab) cudaFree(x->slab);
    if (x->host_words) cudaFreeHost(x->host_words);
    cudaGetLastError();
    delete x;
}

}  // namespace subg
