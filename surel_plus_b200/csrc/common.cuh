// Shared device/host helpers of the SubGAcc CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <string>

#include "../../include/subg_b200.h"

#ifndef __CUDA_ARCH__
#define SUBG_HOST_ONLY
#endif

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "surel_plus_b200 targets sm_100a (B200) only"
#endif

namespace subg {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------ error plumbing
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define SUBG_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess)                                                               \
            return ::subg::fail(_e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA, \
                                std::string(#call) + ": " + cudaGetErrorString(_e));         \
    } while (0)

// RAII device guard
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

bool is_device_ptr(const void *p);

// host-side phase timer (SUBG_PROFILE_HOST=1): where does a sampling call spend wall-clock besides its kernels?
struct HostProf {
    bool on;
    std::chrono::steady_clock::time_point t0;
    std::string log;
    HostProf() : on(getenv("SUBG_PROFILE_HOST") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        const auto t = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof(buf), " %s=%.3f", what, std::chrono::duration<double, std::milli>(t - t0).count());
        log += buf;
        t0 = t;
    }
    ~HostProf() { if (on) fprintf(stderr, "[subg host ms]%s\n", log.c_str()); }
};


// measurement hooks (capi.cu)
void count_launch(int n = 1);
void timing_begin(int which, cudaStream_t st);
void timing_end(int which, cudaStream_t st);

// Device allocation.  Small blocks come from the driver's stream-ordered pool.  Large blocks (>= 64 MB: SpG row arrays,
// the sampler's worst-case staging) come from a per-device cache of cudaMalloc'd blocks kept by the library (capi.cu):
// a block goes back to the cache with an event recorded on the freeing stream and is handed out again to a request of
// similar size, stream-ordered behind that event.  Measured reason: once peer mappings exist in the process (the
// multi-GPU exchange), the driver pool re-creates multi-GB allocations on every pass (300 ms per 10 GB) instead of
// reusing them.
cudaError_t big_alloc(void **p, size_t bytes, cudaStream_t st);
bool big_free(void *p, cudaStream_t st);   // true if p was a cached large block
constexpr size_t kBigBlockBytes = (size_t)64 << 20;

template <typename T>
inline cudaError_t dmalloc(T **p, size_t count, cudaStream_t st) {
    const size_t bytes = (count ? count : 1) * sizeof(T);
    if (bytes >= kBigBlockBytes) return big_alloc((void **)p, bytes, st);
    return cudaMallocAsync((void **)p, bytes, st);
}
inline void dfree(void *p, cudaStream_t st) {
    if (p && !big_free(p, st)) cudaFreeAsync(p, st);
}

// ------------------------------------------------------------------ stream ordering of handles
// Every handle remembers the stream its memory was last used on.  An entry point called with a different stream first
// makes that stream wait for the work queued so far (event), and the handle's memory is released on the last stream,
// so a kernel still queued on a non-blocking side stream never sees its operands handed out again by the pool.
struct StreamTag {
    mutable cudaStream_t last = nullptr;
    void use_on(cudaStream_t st) const {
        if (st == last) return;
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
            if (cudaEventRecord(ev, last) == cudaSuccess) cudaStreamWaitEvent(st, ev, 0);
            else cudaGetLastError();  // the old stream is gone: its work has completed
            cudaEventDestroy(ev);
        }
        last = st;
    }
    // stream to free on (if it was destroyed meanwhile, everything queued on it has run: fall back to a full sync)
    cudaStream_t free_stream() const {
        if (last && cudaStreamQuery(last) == cudaErrorInvalidResourceHandle) {
            cudaGetLastError();
            cudaDeviceSynchronize();
            last = nullptr;
        } else {
            cudaGetLastError();
        }
        return last;
    }
};

// ------------------------------------------------------------------ object layouts
struct Graph {
    StreamTag tag;
    int device = 0;
    int64_t N = 0, E = 0;
    bool rowptr64 = false;
    void *rowptr = nullptr;  // int32[N+1] or int64[N+1]
    int32_t *col = nullptr;  // int32[E]
    void *rowinfo = nullptr; // uint64[N]: row start (low 40 bits) | degree (high 24 bits, 0xFFFFFF = read rowptr)
    // neighbour ids packed three to a 64-bit word (21 bits each), built on demand for graphs with N <= 2^21 whose CSR does
    // not fit the L2: the walk's random column gathers then range over 2/3 of the bytes, so more of them hit the L2
    mutable unsigned long long *col3 = nullptr;
    mutable int col3_state = -1;  // -1 undecided, 0 not used, 1 built
    int num_sms = 148;
    // lazily computed structure properties (-1 unknown): rows strictly ascending; adjacency symmetric
    mutable int sorted_state = -1, sym_state = -1;
};

struct SpG {
    StreamTag tag;
    int device = 0;
    int64_t n = 0;         // rows (sets), in seed order
    int64_t T = 0;         // total entries
    int32_t c = 0;         // unique LP rows (0 for value SpGs)
    int32_t ncol = 0;      // num_steps + 1
    int32_t M = 0;
    int32_t max_set = 0;
    uint32_t status = 0;
    int value_kind = 0;    // 0 int32 pointers, 1 float64 values
    // Row u occupies entries [rowbeg[u], rowbeg[u] + nsize[u]) of indices/data/slot.  Two layouts:
    //   compact   indptr != null, rowbeg == indptr (CSR, rows back to back in seed order)
    //   scattered indptr == null, rowbeg owned: rows sit where the sampler's cursor put them
    //             (16-byte aligned, any order); ensure_csr() turns this into the compact layout.
    int64_t *indptr = nullptr;   // [n+1]
    int64_t *rowbeg = nullptr;   // [n]
    int64_t extent = 0;          // entries in use (== T when compact)
    int64_t cap = 0;             // entries allocated
    int32_t *indices = nullptr;  // ascending per row
    void *data = nullptr;        // int32 (id+1) or float64
    // linked SpG (multi-GPU, csrc/xchg.cu): indices / data are NOT owned -- they point into the exchange slabs of this GPU
    // and of its peers (one address space: rowbeg holds each row's offset from the first slab's plane)
    bool borrowed = false;
    uint16_t *slot = nullptr;    // first-visit rank (sampler-built SpGs that asked for it)
    int16_t *enc = nullptr;      // [c, ncol]
    // sampler-built LP SpGs keep the 64-bit key of every unique LP row and the stream position of its first occurrence
    // ((global seed index << 16) | first-visit order), both in id order: what the multi-GPU merge of the shards' tables needs
    unsigned long long *lp_key = nullptr;  // [c]
    unsigned long long *lp_pos = nullptr;  // [c]
    int32_t shift = 0;           // bits per LP column in the key (32 - clz(M))
    int32_t *nsize = nullptr;    // [n]
    int32_t *seeds = nullptr;    // [n] node id of each row
    int32_t *walks = nullptr;    // [n, M, m] the walks the sampler drew (SUBG_SAMPLE_DUMP_WALKS only)
    int64_t pushes = 0;          // PPR sampler: forward pushes performed (measurement)
    int num_sms = 148;
    // per-handle SpJoin scratch, kept between batches (handles are not thread-safe)
    mutable int32_t *join_sizes = nullptr;   // [join_cap] segment sizes
    mutable int64_t join_cap = 0;
    mutable long long *join_tot = nullptr;   // device {total rows, bad-node flag}
    mutable long long *join_host = nullptr;  // pinned copy of join_tot
};

void spg_free_impl(SpG *s);
// LP-key table -> ids in first-occurrence order (spg.cu)
int rank_unique_keys(const unsigned long long *tab_key, const unsigned long long *tab_pos, uint32_t cap, uint32_t c_max,
                     bool count_exact, int M, int m, int SHIFT, int32_t *rank_of_slot, int16_t *enc,
                     unsigned long long *lp_key, unsigned long long *lp_pos, uint32_t *d_cnt, int num_sms, cudaStream_t st);
int spg_ensure_csr(SpG *s, cudaStream_t st);  // scattered -> compact (no-op when compact)

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1) -> 4 x 32 random bits.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// glibc rand_r: one call = three LCG steps producing 11+10+10 bits.
__device__ __forceinline__ uint32_t lcg(uint32_t s) { return s * 1103515245u + 12345u; }
__device__ __forceinline__ uint32_t rand_r_dev(uint32_t &state) {
    uint32_t s = lcg(state);
    uint32_t out = (s >> 16) & 2047u;
    s = lcg(s);
    out = (out << 10) ^ ((s >> 16) & 1023u);
    s = lcg(s);
    out = (out << 10) ^ ((s >> 16) & 1023u);
    state = s;
    return out;
}
// advance an LCG state by `steps` single steps (mod 2^32) in O(log steps)
__device__ __forceinline__ uint32_t lcg_jump(uint32_t state, uint32_t steps) {
    uint32_t a = 1103515245u, c = 12345u;  // current power-of-two map  x -> a x + c
    uint32_t A = 1u, Cc = 0u;              // accumulated map
    while (steps) {
        if (steps & 1u) { A = a * A; Cc = a * Cc + c; }
        c = (a + 1u) * c;
        a = a * a;
        steps >>= 1;
    }
    return A * state + Cc;
}
#endif  // __CUDACC__

}  // namespace subg
