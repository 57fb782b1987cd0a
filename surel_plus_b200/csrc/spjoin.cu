// SpJoin on the device: outer-join of the sorted sets of a query's endpoints.
//
// Replaces (file:line relative to /root/reference)
//   bgather / gather / pgather   train.py:13-45, 75-111   (pair queries)
//   hgather                      train.py:48-72           (triplet queries, 4 segments)
// Row contract (train.py:34-36, 57-68; model.py:76-83): for every segment, the rows of the
// left set in ascending node id, each row = [value in own set, value in the other set or 0];
// segments are laid out [all left | all right] (pairs) and [u|w, w|u, v|w, w|v] (triplets).
//
// B200 design: one CTA per task (a, b).  The four row slices (ids and values of both sets)
// are staged in shared memory by 1-D TMA bulk copies (cp.async.bulk, 16-byte aligned windows
// around the rows) completing on one mbarrier; every element of S_a then does a binary search
// in S_b, writes its own row, and drops its value at the match position so that the rows of
// S_b need no second search.  Output rows are written coalesced (8 B per row, or 2k floats
// per row when the LP table lookup `encode[xz]` of train.py:37 is fused in).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace subg {

constexpr int kJoinThreads = 128;

struct JoinArgs {
    const long long *rowbeg;   // first entry of every row
    const int32_t *nsize;      // row sizes, or null: compact CSR, size = rowbeg[u + 1] - rowbeg[u]
    const int32_t *indices;
    const void *data;
    int64_t n_rows;
    const long long *edge;
    int64_t B;
    int arity;
    const long long *seg_ptr;
    const float *enc;
    int k;
    void *out;
    long long *segid;
    int64_t ntask;
    int cap;  // staged elements per row slice
    // fused plan + run (subg_spjoin): the kernel runs right behind the plan, without the host having seen the row
    // count; it backs out when the rows do not fit the caller's buffer or a query node is out of range
    int64_t max_rows;          // < 0: no check
    const long long *tot;      // {total rows, bad-node flag} written by the plan kernel (fused mode)
};

__device__ __forceinline__ bool join_must_skip(const JoinArgs &p) {
    if (p.max_rows < 0) return false;
    return p.tot[0] > p.max_rows || p.tot[1] != 0;
}

__device__ __forceinline__ int row_size(const JoinArgs &p, int64_t u) {
    return p.nsize ? p.nsize[u] : (int)(p.rowbeg[u + 1] - p.rowbeg[u]);
}

__device__ __forceinline__ void task_nodes(const JoinArgs &p, int64_t t, int64_t &a, int64_t &b, int64_t &segA,
                                           int64_t &segB) {
    if (p.arity == 2) {
        a = p.edge[t];
        b = p.edge[p.B + t];
        segA = t;
        segB = p.B + t;
    } else {
        const bool second = t >= p.B;
        const int64_t q = second ? t - p.B : t;
        a = p.edge[(second ? p.B : 0) + q];
        b = p.edge[2 * p.B + q];
        segA = (second ? 2 * p.B : 0) + q;
        segB = (second ? 3 * p.B : p.B) + q;
    }
}

// ------------------------------------------------------------------ plan: segment sizes
__global__ void join_sizes_kernel(const long long *rowbeg, const int32_t *nsize, int64_t n_rows, const long long *edge,
                                  int64_t B, int arity, int32_t *sizes, uint32_t *bad) {
    const int64_t nseg = arity == 2 ? 2 * B : 4 * B;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < nseg; g += (int64_t)gridDim.x * blockDim.x) {
        int64_t node;
        if (arity == 2) node = edge[g];
        else {
            const int64_t blk = g / B, q = g - blk * B;
            const int64_t rowsel = blk == 0 ? 0 : (blk == 2 ? 1 : 2);  // u, w, v, w
            node = edge[rowsel * B + q];
        }
        if (node < 0 || node >= n_rows) {
            atomicOr(bad, 1u);
            sizes[g] = 0;
        } else {
            sizes[g] = nsize ? nsize[node] : (int32_t)(rowbeg[node + 1] - rowbeg[node]);
        }
    }
}

// Plan of a batch in two launches (segments <= kPlanSmallMax = 1024 tiles of 256): (1) sizes of every segment and the
// tile sums; the last tile to finish scans the tile sums and publishes the total and the bad-node flag; (2) every
// tile rescans its 256 sizes behind its offset -> segment pointers.
constexpr int kPlanTile = 256;
constexpr int kPlanSmallMax = 1024 * kPlanTile;
__global__ void __launch_bounds__(kPlanTile) join_plan_sizes_kernel(const long long *rowbeg, const int32_t *nsize, int64_t n_rows,
                                                                   const long long *edge, long long *edge_out, int64_t B, int arity,
                                                                   int32_t *sizes, long long *tile_off, unsigned int *done,
                                                                   long long *tot) {
    __shared__ long long ws[32];
    __shared__ bool last;
    const int nseg = (int)(arity == 2 ? 2 * B : 4 * B);
    const int g = blockIdx.x * kPlanTile + threadIdx.x;
    int32_t sz = 0;
    bool bad = false;
    if (g < nseg) {
        int64_t at = g;
        if (arity != 2) {
            const int blk = g / (int)B, q = g - blk * (int)B;
            at = (int64_t)(blk == 0 ? 0 : (blk == 2 ? 1 : 2)) * B + q;  // u, w, v, w
        }
        const long long node = edge[at];
        if (edge_out) edge_out[at] = node;
        if (node < 0 || node >= n_rows) bad = true;
        else sz = nsize ? nsize[node] : (int32_t)(rowbeg[node + 1] - rowbeg[node]);
        sizes[g] = sz;
    }
    long long total;
    block_excl_scan((long long)sz, &total, ws);
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        tile_off[blockIdx.x] = total;
        if (any_bad) atomicExch((unsigned long long *)&tot[1], 1ull);
        __threadfence();
        last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // exclusive scan of the tile sums (<= 1024) by this block: 4 consecutive tiles per thread
    const int nt = gridDim.x;
    long long v[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int t = threadIdx.x * 4 + q;
        v[q] = t < nt ? ((volatile long long *)tile_off)[t] : 0;
        sum += v[q];
    }
    long long ex = block_excl_scan(sum, &total, ws);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int t = threadIdx.x * 4 + q;
        if (t < nt) tile_off[t] = ex;
        ex += v[q];
    }
    if (threadIdx.x == 0) {
        tot[0] = total;
        *done = 0u;  // ready for the next batch
    }
}
__global__ void __launch_bounds__(kPlanTile) join_plan_offsets_kernel(const int32_t *sizes, const long long *tile_off, int nseg,
                                                                     long long *seg_ptr, const long long *tot) {
    __shared__ long long ws[32];
    const int g = blockIdx.x * kPlanTile + threadIdx.x;
    const int32_t sz = g < nseg ? sizes[g] : 0;
    const long long ex = block_excl_scan((long long)sz, nullptr, ws) + tile_off[blockIdx.x];
    if (g < nseg) seg_ptr[g] = ex;
    if (g == nseg - 1) seg_ptr[nseg] = tot[0];
}
// Plan of a small batch in ONE launch (segments <= kPlanOneMax): a single block sizes every segment and scans them.
// Phase 1 walks the segments with a stride of the block (a warp reads 256 contiguous bytes of the edge array, which may be
// pinned HOST memory read over PCIe: one request per warp and round, all rounds in flight before the first is used) and
// leaves the sizes in shared memory; phase 2 gives every thread `per` consecutive segments of that array for the scan.
constexpr int kPlanOneThreads = 1024;
constexpr int kPlanOneItems = 8;
constexpr int kPlanOneMax = kPlanOneThreads * kPlanOneItems;
// edge_out (nullable) receives the device copy the join kernel reads
__global__ void __launch_bounds__(kPlanOneThreads) join_plan_one_kernel(const long long *__restrict__ rowbeg, const int32_t *__restrict__ nsize,
                                                                       int64_t n_rows, const long long *__restrict__ edge,
                                                                       long long *__restrict__ edge_out, int64_t B, int arity, int nseg,
                                                                       long long *__restrict__ seg_ptr, long long *__restrict__ tot) {
    __shared__ long long ws[32];
    __shared__ int32_t ssz[kPlanOneMax];
    long long node[kPlanOneItems];
    int64_t at[kPlanOneItems];
    bool bad = false;
#pragma unroll
    for (int q = 0; q < kPlanOneItems; q++) {
        const int g = threadIdx.x + q * kPlanOneThreads;
        node[q] = -1;
        at[q] = g;
        if (g < nseg) {
            if (arity != 2) {
                const int blk = g / (int)B, qq = g - blk * (int)B;
                at[q] = (int64_t)(blk == 0 ? 0 : (blk == 2 ? 1 : 2)) * B + qq;  // u, w, v, w
            }
            node[q] = edge[at[q]];
        }
    }
#pragma unroll
    for (int q = 0; q < kPlanOneItems; q++) {
        const int g = threadIdx.x + q * kPlanOneThreads;
        if (g < nseg) {
            int32_t sz = 0;
            if (node[q] < 0 || node[q] >= n_rows) bad = true;
            else sz = nsize ? nsize[node[q]] : (int32_t)(rowbeg[node[q] + 1] - rowbeg[node[q]]);
            ssz[g] = sz;
            if (edge_out) edge_out[at[q]] = node[q];
        }
    }
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    const int per = (nseg + kPlanOneThreads - 1) / kPlanOneThreads;   // consecutive segments per thread, <= kPlanOneItems
    const int first = threadIdx.x * per;
    long long sum = 0;
#pragma unroll
    for (int q = 0; q < kPlanOneItems; q++)
        if (q < per && first + q < nseg) sum += ssz[first + q];
    long long total;
    long long ex = block_excl_scan(sum, &total, ws);
#pragma unroll
    for (int q = 0; q < kPlanOneItems; q++) {
        const int g = first + q;
        if (q < per && g < nseg) {
            seg_ptr[g] = ex;
            ex += ssz[g];
        }
    }
    if (threadIdx.x == 0) {
        seg_ptr[nseg] = total;
        tot[0] = total;
        tot[1] = any_bad ? 1 : 0;
    }
}

__global__ void join_tot_kernel(const long long *seg_ptr, int64_t nseg, const uint32_t *bad, long long *tot) {
    tot[0] = seg_ptr[nseg];
    tot[1] = *bad;
}

// ------------------------------------------------------------------ TMA / mbarrier PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared; both addresses 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename V>
struct ValTraits;
template <>
struct ValTraits<int32_t> {
    static constexpr int align_elems = 4;  // 16 B
};
template <>
struct ValTraits<double> {
    static constexpr int align_elems = 2;
};

struct TaskDesc {
    long long pa, pb;      // row starts
    int sa, sb;            // set sizes
    int ia, ib;            // offset of the row inside the staged id window
    int va, vb;            // same for the value window
    long long offA, offB;  // output row offsets
    long long segA, segB;
};

// ------------------------------------------------------------------ the join kernel
// one thread per (row, side): copies the KK floats of LP row `ptr` to their place in the [N,2,KK] output
template <int KK>
__device__ __forceinline__ void put_lp_row(float *out, long long row, int side, const float *enc, int ptr) {
    float *dst = out + (row * 2 + side) * KK;
    const float *src = enc + (int64_t)ptr * KK;
    if (KK > 0 && KK % 4 == 0) {
#pragma unroll
        for (int c = 0; c < KK / 4; c++) ((float4 *)dst)[c] = __ldg((const float4 *)src + c);
    } else if (KK % 2 == 0) {
#pragma unroll
        for (int c = 0; c < KK / 2; c++) ((float2 *)dst)[c] = __ldg((const float2 *)src + c);
    } else {
        float v[KK > 0 ? KK : 1];
#pragma unroll
        for (int c = 0; c < KK; c++) v[c] = __ldg(src + c);
#pragma unroll
        for (int c = 0; c < KK; c++) dst[c] = v[c];
    }
}

// MODE 0: int32 [N,2] pointers   MODE 1: float32 [N,2,k] fused table lookup   MODE 2: float32 [N,2] values
// KK: compile-time k of MODE 1 (0 = any k, element-wise copy)
template <typename V, int MODE, int KK>
__global__ void __launch_bounds__(kJoinThreads) spjoin_kernel(const JoinArgs p) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int cap = p.cap;
    int32_t *idA = (int32_t *)sm;
    int32_t *idB = idA + cap;
    V *vA = (V *)(idB + cap);
    V *vB = vA + cap;
    V *rev = vB + cap;      // value of the S_a member matched at each S_b position (0 = none)
    V *mat = rev + cap;     // MODE 1 only: value of the S_b member matched by each S_a element
    __shared__ uint64_t bar;
    __shared__ TaskDesc td;
    const V *gdata = (const V *)p.data;
    constexpr int AE = ValTraits<V>::align_elems;

    if (join_must_skip(p)) return;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;

    for (int64_t t = blockIdx.x; t < p.ntask; t += gridDim.x) {
        if (threadIdx.x == 0) {
            int64_t a, b, segA, segB;
            task_nodes(p, t, a, b, segA, segB);
            TaskDesc d;
            d.pa = p.rowbeg[a]; d.sa = row_size(p, a);
            d.pb = p.rowbeg[b]; d.sb = row_size(p, b);
            d.offA = p.seg_ptr[segA]; d.offB = p.seg_ptr[segB];
            d.segA = segA; d.segB = segB;
            const long long a0 = d.pa & ~3ll, b0 = d.pb & ~3ll;
            const long long av0 = d.pa & ~(long long)(AE - 1), bv0 = d.pb & ~(long long)(AE - 1);
            d.ia = (int)(d.pa - a0); d.ib = (int)(d.pb - b0);
            d.va = (int)(d.pa - av0); d.vb = (int)(d.pb - bv0);
            const uint32_t nia = d.sa ? (uint32_t)(((d.pa + d.sa + 3) & ~3ll) - a0) * 4u : 0u;
            const uint32_t nib = d.sb ? (uint32_t)(((d.pb + d.sb + 3) & ~3ll) - b0) * 4u : 0u;
            const uint32_t nva = d.sa ? (uint32_t)(((d.pa + d.sa + AE - 1) & ~(long long)(AE - 1)) - av0) * (uint32_t)sizeof(V) : 0u;
            const uint32_t nvb = d.sb ? (uint32_t)(((d.pb + d.sb + AE - 1) & ~(long long)(AE - 1)) - bv0) * (uint32_t)sizeof(V) : 0u;
            td = d;
            if (nia + nib) {
                mbar_expect_tx(&bar, nia + nib + nva + nvb);
                if (nia) { tma_load_1d(idA, p.indices + a0, nia, &bar); tma_load_1d(vA, gdata + av0, nva, &bar); }
                if (nib) { tma_load_1d(idB, p.indices + b0, nib, &bar); tma_load_1d(vB, gdata + bv0, nvb, &bar); }
            }
        }
        __syncthreads();
        const TaskDesc d = td;
        for (int j = threadIdx.x; j < d.sb; j += kJoinThreads) rev[j] = (V)0;
        if (d.sa + d.sb) {
            mbar_wait(&bar, phase);
            phase ^= 1u;
        }
        __syncthreads();

        const int32_t *A = idA + d.ia, *Bq = idB + d.ib;
        const V *VA = vA + d.va, *VB = vB + d.vb;
        // ---- S_a side: search every member in S_b
        for (int j = threadIdx.x; j < d.sa; j += kJoinThreads) {
            const int32_t w = A[j];
            int lo = 0, hi = d.sb;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (Bq[mid] < w) lo = mid + 1;
                else hi = mid;
            }
            const bool hit = lo < d.sb && Bq[lo] == w;
            const V own = VA[j];
            const V other = hit ? VB[lo] : (V)0;
            if (hit) rev[lo] = own;
            if (MODE == 0) {
                ((int2 *)p.out)[d.offA + j] = make_int2((int)own, (int)other);
            } else if (MODE == 2) {
                const double o = ((double)other + 1.0) - 1.0;  // train.py:33,38 rounding in float64
                ((float2 *)p.out)[d.offA + j] = make_float2((float)own, (float)o);
            } else {
                mat[j] = other;
            }
            if (p.segid) p.segid[d.offA + j] = d.segA;
        }
        __syncthreads();
        // ---- S_b side: the match (if any) was dropped at its position
        if (MODE == 1 && KK > 0) {
            float *out = (float *)p.out;
            for (int e = threadIdx.x; e < 2 * d.sa; e += kJoinThreads) {
                const int r = e >> 1, side = e & 1;
                put_lp_row<KK>(out, d.offA + r, side, p.enc, side ? (int)mat[r] : (int)VA[r]);
            }
            for (int e = threadIdx.x; e < 2 * d.sb; e += kJoinThreads) {
                const int r = e >> 1, side = e & 1;
                put_lp_row<KK>(out, d.offB + r, side, p.enc, side ? (int)rev[r] : (int)VB[r]);
            }
            if (p.segid)
                for (int j = threadIdx.x; j < d.sb; j += kJoinThreads) p.segid[d.offB + j] = d.segB;
        } else if (MODE == 1) {
            const int k = p.k, k2 = 2 * p.k;
            float *out = (float *)p.out;
            for (int e = threadIdx.x; e < d.sa * k2; e += kJoinThreads) {
                const int r = e / k2, rem = e - r * k2;
                const int side = rem >= k, c = rem - side * k;
                const int ptr = side ? (int)mat[r] : (int)VA[r];
                out[(d.offA + r) * k2 + rem] = __ldg(p.enc + (int64_t)ptr * k + c);
            }
            for (int e = threadIdx.x; e < d.sb * k2; e += kJoinThreads) {
                const int r = e / k2, rem = e - r * k2;
                const int side = rem >= k, c = rem - side * k;
                const int ptr = side ? (int)rev[r] : (int)VB[r];
                out[(d.offB + r) * k2 + rem] = __ldg(p.enc + (int64_t)ptr * k + c);
            }
            if (p.segid)
                for (int j = threadIdx.x; j < d.sb; j += kJoinThreads) p.segid[d.offB + j] = d.segB;
        } else {
            for (int j = threadIdx.x; j < d.sb; j += kJoinThreads) {
                const V own = VB[j];
                const V other = rev[j];
                if (MODE == 0) {
                    ((int2 *)p.out)[d.offB + j] = make_int2((int)own, (int)other);
                } else {
                    const double o = ((double)other + 1.0) - 1.0;
                    ((float2 *)p.out)[d.offB + j] = make_float2((float)own, (float)o);
                }
                if (p.segid) p.segid[d.offB + j] = d.segB;
            }
        }
        __syncthreads();  // smem is reused by the next task
    }
}

// ------------------------------------------------------------------ generic path (sets too large for smem)
template <typename V, int MODE>
__global__ void __launch_bounds__(kJoinThreads) spjoin_global_kernel(const JoinArgs p) {
    const V *gdata = (const V *)p.data;
    if (join_must_skip(p)) return;
    for (int64_t t = blockIdx.x; t < p.ntask; t += gridDim.x) {
        int64_t a, b, segA, segB;
        task_nodes(p, t, a, b, segA, segB);
        for (int dir = 0; dir < 2; dir++) {
            const int64_t x = dir ? b : a, y = dir ? a : b;
            const int64_t seg = dir ? segB : segA;
            const long long px = p.rowbeg[x], py = p.rowbeg[y];
            const int64_t sx = row_size(p, x), sy = row_size(p, y);
            const long long off = p.seg_ptr[seg];
            for (int64_t j = threadIdx.x; j < sx; j += kJoinThreads) {
                const int32_t w = p.indices[px + j];
                int64_t lo = 0, hi = sy;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (p.indices[py + mid] < w) lo = mid + 1;
                    else hi = mid;
                }
                const bool hit = lo < sy && p.indices[py + lo] == w;
                const V own = gdata[px + j];
                const V other = hit ? gdata[py + lo] : (V)0;
                if (MODE == 0) {
                    ((int2 *)p.out)[off + j] = make_int2((int)own, (int)other);
                } else if (MODE == 2) {
                    const double o = ((double)other + 1.0) - 1.0;
                    ((float2 *)p.out)[off + j] = make_float2((float)own, (float)o);
                } else {
                    float *out = (float *)p.out + (off + j) * 2 * p.k;
                    for (int c = 0; c < p.k; c++) out[c] = __ldg(p.enc + (int64_t)own * p.k + c);
                    for (int c = 0; c < p.k; c++) out[p.k + c] = __ldg(p.enc + (int64_t)other * p.k + c);
                }
                if (p.segid) p.segid[off + j] = seg;
            }
        }
    }
}

// ------------------------------------------------------------------ host side
int spjoin_plan_impl(const SpG *s, const int64_t *edge_hd, int64_t B, int arity, int64_t *edge_dev,
                     int64_t *indptr_dev, int64_t *N_out, cudaStream_t st) {
    if (!s || !edge_hd || !indptr_dev || !N_out || B < 0 || (arity != 2 && arity != 3))
        return fail(SUBG_ERR_ARG, "Input parsing error.");
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    const int64_t nseg = (arity == 2 ? 2 : 4) * B;
    const long long *edge = (const long long *)edge_hd;
    if (!is_device_ptr(edge_hd)) {
        if (!edge_dev) return fail(SUBG_ERR_ARG, "host edge list needs a device staging buffer");
        SUBG_CUDA(cudaMemcpyAsync(edge_dev, edge_hd, (size_t)arity * B * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        edge = (const long long *)edge_dev;
    } else if (edge_dev && edge_dev != edge_hd) {
        SUBG_CUDA(cudaMemcpyAsync(edge_dev, edge_hd, (size_t)arity * B * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    }
    int32_t *sizes = nullptr;
    long long *scratch = nullptr;
    uint32_t *bad = nullptr;
    SUBG_CUDA(dmalloc(&sizes, (size_t)nseg, st));
    SUBG_CUDA(dmalloc(&scratch, (size_t)std::max(1, scan_num_blocks(nseg)), st));
    SUBG_CUDA(dmalloc(&bad, 1, st));
    SUBG_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), st));
    if (nseg > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>((nseg + 255) / 256, 4 * (int64_t)s->num_sms);
        join_sizes_kernel<<<blocks, 256, 0, st>>>((const long long *)s->rowbeg, s->indptr ? nullptr : s->nsize, s->n, edge, B,
                                                  arity, sizes, bad);
    }
    SUBG_CUDA(exclusive_scan_i32_i64(sizes, (long long *)indptr_dev, nseg, 0, scratch, st));
    count_launch(4);
    long long N = 0;
    uint32_t hbad = 0;
    SUBG_CUDA(cudaMemcpyAsync(&N, indptr_dev + nseg, sizeof(long long), cudaMemcpyDeviceToHost, st));
    SUBG_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SUBG_CUDA(cudaStreamSynchronize(st));
    dfree(sizes, st); dfree(scratch, st); dfree(bad, st);
    if (hbad) return fail(SUBG_ERR_ARG, "query node id outside the SpG");
    *N_out = N;
    return SUBG_OK;
}

template <typename V, int MODE, int KK = 0>
static cudaError_t launch_join(const SpG *s, JoinArgs &p, cudaStream_t st) {
    if (p.ntask <= 0) return cudaSuccess;
    static thread_local int c_smem_dev = -1, dev_smem = 0;
    if (c_smem_dev != s->device) {
        cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device);
        c_smem_dev = s->device;
    }
    const int cap = ((s->max_set + 3) & ~3) + 8;
    const size_t smem = (size_t)cap * (8 + 4 * sizeof(V)) + 128;
    if ((int64_t)smem <= std::min<int64_t>(dev_smem - 1024, 96 * 1024)) {
        p.cap = cap;
        auto kern = spjoin_kernel<V, MODE, KK>;
        // attribute + occupancy query once per (kernel, device, shared-memory size): they cost more than the launch
        static thread_local int c_dev = -1, c_per_sm = 0;
        static thread_local size_t c_smem = 0;
        cudaError_t e;
        if (c_dev != s->device || c_smem != smem) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c_per_sm, kern, kJoinThreads, smem);
            if (e != cudaSuccess) return e;
            c_dev = s->device;
            c_smem = smem;
        }
        const int per_sm = c_per_sm;
        const int64_t blocks = std::min<int64_t>(p.ntask, (int64_t)s->num_sms * std::max(per_sm, 1));
        kern<<<(unsigned)blocks, kJoinThreads, smem, st>>>(p);
    } else {
        const int64_t blocks = std::min<int64_t>(p.ntask, (int64_t)s->num_sms * 16);
        spjoin_global_kernel<V, MODE><<<(unsigned)blocks, kJoinThreads, 0, st>>>(p);
    }
    return cudaGetLastError();
}

static int join_launch(const SpG *s, const int64_t *edge_dev, int64_t B, int arity, const int64_t *indptr_dev,
                       const float *enc_table_dev, int k, void *out_dev, int64_t *segid_dev, int64_t max_rows,
                       const long long *tot, cudaStream_t st) {
    if (!s || !edge_dev || !indptr_dev || B < 0 || (arity != 2 && arity != 3)) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (B > 0 && !out_dev) return fail(SUBG_ERR_ARG, "null output");
    if (enc_table_dev && (s->value_kind != 0 || k < 1)) return fail(SUBG_ERR_ARG, "table lookup needs an int SpG and k >= 1");
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    JoinArgs p{};
    p.rowbeg = (const long long *)s->rowbeg; p.nsize = s->indptr ? nullptr : s->nsize; p.indices = s->indices; p.data = s->data; p.n_rows = s->n;
    p.edge = (const long long *)edge_dev; p.B = B; p.arity = arity; p.seg_ptr = (const long long *)indptr_dev;
    p.enc = enc_table_dev; p.k = k; p.out = out_dev; p.segid = (long long *)segid_dev;
    p.ntask = arity == 2 ? B : 2 * B;
    p.max_rows = max_rows; p.tot = tot;
    cudaError_t e;
    timing_begin(SUBG_TIMING_SPJOIN, st);
    if (s->value_kind == 1) e = launch_join<double, 2>(s, p, st);
    else if (enc_table_dev) {
        // the LP table must allow vector loads of a row (torch allocations are 512-byte aligned)
        const bool al = ((uintptr_t)enc_table_dev & 15) == 0;
        switch (al ? k : 0) {
            case 2: e = launch_join<int32_t, 1, 2>(s, p, st); break;
            case 3: e = launch_join<int32_t, 1, 3>(s, p, st); break;
            case 4: e = launch_join<int32_t, 1, 4>(s, p, st); break;
            case 5: e = launch_join<int32_t, 1, 5>(s, p, st); break;
            default: e = launch_join<int32_t, 1, 0>(s, p, st);
        }
    }
    else e = launch_join<int32_t, 0>(s, p, st);
    timing_end(SUBG_TIMING_SPJOIN, st);
    count_launch(1);
    if (e != cudaSuccess) return fail(SUBG_ERR_CUDA, cudaGetErrorString(e));
    return SUBG_OK;
}

int spjoin_run_impl(const SpG *s, const int64_t *edge_dev, int64_t B, int arity, const int64_t *indptr_dev,
                    const float *enc_table_dev, int k, void *out_dev, int64_t *segid_dev, cudaStream_t st) {
    return join_launch(s, edge_dev, B, arity, indptr_dev, enc_table_dev, k, out_dev, segid_dev, -1, nullptr, st);
}

// plan + run with ONE host synchronisation: the plan kernel(s) and the join kernel are queued back to back; the join
// kernel checks on the device that the rows fit `out_capacity`.  *ran = 0 -> nothing was written (call run with a
// buffer of *N_out rows; edge_dev / indptr_dev are already filled).  Scratch lives on the SpG handle between batches.
int spjoin_fused_impl(const SpG *s, const int64_t *edge_hd, int64_t B, int arity, int64_t *edge_dev, int64_t *indptr_dev,
                      const float *enc_table_dev, int k, void *out_dev, int64_t out_capacity, int64_t *segid_dev,
                      int64_t *N_out, int *ran, cudaStream_t st) {
    if (!s || !edge_hd || !edge_dev || !indptr_dev || !N_out || !ran || B < 0 || out_capacity < 0 || (arity != 2 && arity != 3))
        return fail(SUBG_ERR_ARG, "Input parsing error.");
    DeviceGuard guard(s->device);
    s->tag.use_on(st);
    HostProf prof;
    const int64_t nseg = (arity == 2 ? 2 : 4) * B;
    if (edge_dev != edge_hd)
        SUBG_CUDA(cudaMemcpyAsync(edge_dev, edge_hd, (size_t)arity * B * sizeof(int64_t), cudaMemcpyDefault, st));
    if (s->join_cap < nseg + 1 || !s->join_tot) {
        dfree(s->join_sizes, st);
        s->join_sizes = nullptr;
        s->join_cap = 0;
        SUBG_CUDA(dmalloc(&s->join_sizes, (size_t)nseg + 1, st));
        s->join_cap = nseg + 1;
        if (!s->join_tot) {  // [0] total rows  [1] bad-node flag  [2] flag word of the large path  [3] tile counter  [8..] tile sums
            SUBG_CUDA(dmalloc(&s->join_tot, 8 + 1024, st));
            SUBG_CUDA(cudaMemsetAsync(s->join_tot, 0, (8 + 1024) * sizeof(long long), st));
        }
        if (!s->join_host) SUBG_CUDA(cudaHostAlloc((void **)&s->join_host, 4 * sizeof(long long), cudaHostAllocDefault));
    }
    prof.mark("scratch");
    SUBG_CUDA(cudaMemsetAsync(s->join_tot, 0, 2 * sizeof(long long), st));
    if (nseg <= kPlanSmallMax) {
        const int tiles = (int)((nseg + kPlanTile - 1) / kPlanTile);
        if (tiles > 0) {
            long long *tile_off = s->join_tot + 8;
            join_plan_sizes_kernel<<<tiles, kPlanTile, 0, st>>>((const long long *)s->rowbeg, s->indptr ? nullptr : s->nsize, s->n,
                                                                (const long long *)edge_dev, nullptr, B, arity, s->join_sizes, tile_off,
                                                                (unsigned int *)(s->join_tot + 3), s->join_tot);
            join_plan_offsets_kernel<<<tiles, kPlanTile, 0, st>>>(s->join_sizes, tile_off, (int)nseg, (long long *)indptr_dev,
                                                                  s->join_tot);
            count_launch(2);
        } else {
            SUBG_CUDA(cudaMemsetAsync(indptr_dev, 0, sizeof(long long), st));
        }
    } else {
        long long *scratch = nullptr;
        uint32_t *bad = (uint32_t *)(s->join_tot + 2);
        SUBG_CUDA(dmalloc(&scratch, (size_t)std::max(1, scan_num_blocks(nseg)), st));
        SUBG_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), st));
        const unsigned blocks = (unsigned)std::min<int64_t>((nseg + 255) / 256, 4 * (int64_t)s->num_sms);
        join_sizes_kernel<<<blocks, 256, 0, st>>>((const long long *)s->rowbeg, s->indptr ? nullptr : s->nsize, s->n,
                                                  (const long long *)edge_dev, B, arity, s->join_sizes, bad);
        SUBG_CUDA(exclusive_scan_i32_i64(s->join_sizes, (long long *)indptr_dev, nseg, 0, scratch, st));
        join_tot_kernel<<<1, 1, 0, st>>>((const long long *)indptr_dev, nseg, bad, s->join_tot);
        dfree(scratch, st);
        count_launch(5);
    }
    SUBG_CUDA(cudaGetLastError());
    prof.mark("plan");
    int rc = SUBG_OK;
    if (B > 0 && out_dev)
        rc = join_launch(s, edge_dev, B, arity, indptr_dev, enc_table_dev, k, out_dev, segid_dev, out_capacity, s->join_tot, st);
    prof.mark("join");
    SUBG_CUDA(cudaMemcpyAsync(s->join_host, s->join_tot, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    SUBG_CUDA(cudaStreamSynchronize(st));
    prof.mark("sync");
    if (rc != SUBG_OK) return rc;
    if (s->join_host[1]) return fail(SUBG_ERR_ARG, "query node id outside the SpG");
    const long long N = s->join_host[0];
    *N_out = N;
    *ran = ((B > 0 && out_dev && N <= out_capacity) || N == 0) ? 1 : 0;
    return SUBG_OK;
}

// ------------------------------------------------------------------ joiner: the per-batch join as a replayed CUDA graph
// The training loop of the reference joins one mini-batch per step (train.py:121-127, batch 1024; main_horder.py:33,
// 2048 triplets): at those sizes the join kernel runs for 10-20 us and everything around it -- output allocation, three
// launches, the device-to-host copy of the row count, the stream synchronisation -- costs several times that.  A joiner
// fixes (SpG, batch size, arity, LP table, output capacity) once, captures [edge upload -> plan -> join -> row count to
// pinned memory] as a CUDA graph per ring slot, and a submit is one small memcpy into pinned staging plus one
// cudaGraphLaunch: no allocation, no host synchronisation.  The row count stays on the device (and lands in pinned
// memory for whoever wants it later); rows beyond it in the slot's output buffer are not written.
struct JoinSlot {
    long long *edge_dev = nullptr, *indptr_dev = nullptr, *segid_dev = nullptr, *tot_dev = nullptr;
    int32_t *sizes_dev = nullptr;
    void *out_dev = nullptr;
    long long *edge_pin = nullptr, *tot_pin = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaEvent_t done = nullptr;
    cudaStream_t last_stream = nullptr;
};
struct Joiner {
    const SpG *s = nullptr;
    int64_t B = 0, cap_rows = 0;
    int arity = 2, k = 0, depth = 0, next = 0, launches = 0;
    size_t row_bytes = 0;
    const float *enc = nullptr;
    cudaStream_t cap_stream = nullptr;
    // Host-edge batches alternate between two internal streams, so that the plan kernel of batch k+1 (and the PCIe read of
    // its edges) runs beside the join kernel of batch k; the caller's stream is made to wait for each batch (see submit).
    cudaStream_t lane[2] = {nullptr, nullptr};
    cudaEvent_t ev_user = nullptr;   // the caller's stream at the previous submit
    bool ev_user_set = false;
    int lanes = 2;
    unsigned seq = 0;
    std::vector<JoinSlot> slot;
};

void joiner_free_impl(Joiner *j) {
    if (!j) return;
    DeviceGuard guard(j->s ? j->s->device : 0);
    for (auto &q : j->slot) {
        if (q.done) { cudaEventSynchronize(q.done); cudaEventDestroy(q.done); }
        if (q.exec) cudaGraphExecDestroy(q.exec);
        if (q.edge_dev) cudaFree(q.edge_dev);
        if (q.indptr_dev) cudaFree(q.indptr_dev);
        if (q.segid_dev) cudaFree(q.segid_dev);
        if (q.sizes_dev) cudaFree(q.sizes_dev);
        if (q.tot_dev) cudaFree(q.tot_dev);
        if (q.out_dev) cudaFree(q.out_dev);
        if (q.edge_pin) cudaFreeHost(q.edge_pin);
        if (q.tot_pin) cudaFreeHost(q.tot_pin);
    }
    for (int l = 0; l < 2; l++)
        if (j->lane[l]) { cudaStreamSynchronize(j->lane[l]); cudaStreamDestroy(j->lane[l]); }
    if (j->ev_user) cudaEventDestroy(j->ev_user);
    if (j->cap_stream) cudaStreamDestroy(j->cap_stream);
    cudaGetLastError();
    delete j;
}

// plan + join + row count to pinned memory of one slot, queued on st (captured into the slot's graph, and run once
// un-captured beforehand so that the launch attributes are cached before the capture starts)
// plan + join of one slot, queued on st (captured into the slot's graph, and run once un-captured beforehand so that the
// launch attributes are cached before the capture starts).  edge_src: where the plan kernel reads the batch's edges --
// the slot's pinned host staging (zero-copy over PCIe, 16-48 KB) or the caller's device array; the plan kernel leaves
// the device copy that the join kernel reads.  Two kernel nodes per batch, nothing else.
static int joiner_enqueue(Joiner *j, JoinSlot &sl, cudaStream_t st, const long long *edge_src) {
    const SpG *s = j->s;
    const int64_t B = j->B, nseg = (j->arity == 2 ? 2 : 4) * B;
    if (nseg <= kPlanOneMax) {
        join_plan_one_kernel<<<1, kPlanOneThreads, 0, st>>>((const long long *)s->rowbeg, s->indptr ? nullptr : s->nsize, s->n,
                                                            edge_src, sl.edge_dev, B, j->arity, (int)nseg, sl.indptr_dev, sl.tot_dev);
        j->launches = 2;
    } else {
        const int tiles = (int)((nseg + kPlanTile - 1) / kPlanTile);
        long long *tile_off = sl.tot_dev + 8;
        cudaMemsetAsync(sl.tot_dev, 0, 2 * sizeof(long long), st);
        join_plan_sizes_kernel<<<tiles, kPlanTile, 0, st>>>((const long long *)s->rowbeg, s->indptr ? nullptr : s->nsize, s->n,
                                                            edge_src, sl.edge_dev, B, j->arity, sl.sizes_dev, tile_off,
                                                            (unsigned int *)(sl.tot_dev + 3), sl.tot_dev);
        join_plan_offsets_kernel<<<tiles, kPlanTile, 0, st>>>(sl.sizes_dev, tile_off, (int)nseg, sl.indptr_dev, sl.tot_dev);
        j->launches = 3;
    }
    return join_launch(s, (const int64_t *)sl.edge_dev, B, j->arity, (const int64_t *)sl.indptr_dev, j->enc, j->k, sl.out_dev,
                       (int64_t *)sl.segid_dev, j->cap_rows, sl.tot_dev, st);
}

int joiner_create_impl(const SpG *s, int64_t B, int arity, const float *enc_table_dev, int k, int64_t capacity_rows,
                       int want_segid, int depth, Joiner **out) {
    if (!s || !out || B < 1 || (arity != 2 && arity != 3) || capacity_rows < 1 || depth < 1 || depth > 16 || s->n < 1)
        return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (enc_table_dev && (s->value_kind != 0 || k < 1)) return fail(SUBG_ERR_ARG, "table lookup needs an int SpG and k >= 1");
    const int64_t nseg = (arity == 2 ? 2 : 4) * B;
    if (nseg > kPlanSmallMax) return fail(SUBG_ERR_UNSUPPORTED, "joiner: batch too large (use subg_spjoin)");
    DeviceGuard guard(s->device);
    cudaDeviceSynchronize();  // the SpG is complete before the graphs that read it are built
    Joiner *j = new Joiner();
    j->s = s; j->B = B; j->arity = arity; j->k = k; j->enc = enc_table_dev; j->cap_rows = capacity_rows; j->depth = depth;
    j->row_bytes = s->value_kind == 1 ? 8 : (enc_table_dev ? (size_t)8 * k : 8);
    j->slot.resize(depth);
    cudaError_t e = cudaStreamCreateWithFlags(&j->cap_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) s->tag.last = j->cap_stream;  // no cross-stream event inside the capture (everything has completed)
    if (const char *v = getenv("SUBG_JOIN_LANES")) j->lanes = atoi(v) >= 2 ? 2 : 1;
    for (int l = 0; l < j->lanes && e == cudaSuccess; l++) e = cudaStreamCreateWithFlags(&j->lane[l], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&j->ev_user, cudaEventDisableTiming);
    int rc = SUBG_OK;
    for (int q = 0; q < depth && e == cudaSuccess && rc == SUBG_OK; q++) {
        JoinSlot &sl = j->slot[q];
        e = cudaMalloc((void **)&sl.edge_dev, (size_t)arity * B * 8);
        if (e == cudaSuccess) e = cudaMalloc((void **)&sl.indptr_dev, (size_t)(nseg + 1) * 8);
        if (e == cudaSuccess) e = cudaMalloc((void **)&sl.sizes_dev, (size_t)(nseg + 1) * 4);
        if (e == cudaSuccess && want_segid) e = cudaMalloc((void **)&sl.segid_dev, (size_t)capacity_rows * 8);
        if (e == cudaSuccess) e = cudaMalloc((void **)&sl.tot_dev, (8 + 1024) * 8);
        if (e == cudaSuccess) e = cudaMemset(sl.tot_dev, 0, (8 + 1024) * 8);
        if (e == cudaSuccess) e = cudaMemset(sl.edge_dev, 0, (size_t)arity * B * 8);
        if (e == cudaSuccess) e = cudaMalloc(&sl.out_dev, (size_t)capacity_rows * j->row_bytes);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&sl.edge_pin, (size_t)arity * B * 8, cudaHostAllocDefault);
        if (e == cudaSuccess) memset(sl.edge_pin, 0, (size_t)arity * B * 8);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&sl.tot_pin, 4 * 8, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
        if (e != cudaSuccess) break;
        sl.tot_pin[0] = sl.tot_pin[1] = 0;
        if (q == 0) {  // dry run (all queries = node 0): launch attributes and occupancy are cached outside the capture
            rc = joiner_enqueue(j, sl, j->cap_stream, sl.edge_pin);
            e = cudaStreamSynchronize(j->cap_stream);
            if (rc != SUBG_OK || e != cudaSuccess) break;
        }
        cudaGraph_t graph = nullptr;
        e = cudaStreamBeginCapture(j->cap_stream, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) break;
        rc = joiner_enqueue(j, sl, j->cap_stream, sl.edge_pin);
        e = cudaStreamEndCapture(j->cap_stream, &graph);
        if (e == cudaSuccess && rc == SUBG_OK) e = cudaGraphInstantiate(&sl.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
    }
    if (e != cudaSuccess || rc != SUBG_OK) {
        cudaGetLastError();
        joiner_free_impl(j);
        if (rc != SUBG_OK) return rc;
        return fail(e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA, std::string("joiner: ") + cudaGetErrorString(e));
    }
    *out = j;
    return SUBG_OK;
}

// edge_hd: int64[arity * B].  edge_on_device: 1 device memory, 0 host memory, < 0 ask the driver.  Host edges are copied
// into the slot's pinned staging (the caller's array may be reused at once) and the slot's graph -- upload, plan, join,
// row count to pinned memory -- is launched; device edges are copied on the stream and the same kernels are launched
// directly.  Returns the slot's buffers; nothing is synchronised.
int joiner_submit_impl(Joiner *j, const int64_t *edge_hd, int edge_on_device, cudaStream_t st, void **out_dev, int64_t **indptr_dev,
                       int64_t **segid_dev, const int64_t **nrows_dev, int *slot_out) {
    if (!j || !edge_hd) return fail(SUBG_ERR_ARG, "Input parsing error.");
    DeviceGuard guard(j->s->device);
    j->s->tag.use_on(st);
    const int q = j->next;
    j->next = (j->next + 1) % j->depth;
    JoinSlot &sl = j->slot[q];
    const size_t eb = (size_t)j->arity * j->B * 8;
    if (edge_on_device < 0) edge_on_device = is_device_ptr(edge_hd) ? 1 : 0;
    if (edge_on_device) {
        if (int rc = joiner_enqueue(j, sl, st, (const long long *)edge_hd)) return rc;
        SUBG_CUDA(cudaEventRecord(sl.done, st));
        sl.last_stream = st;
    } else {
        SUBG_CUDA(cudaEventSynchronize(sl.done));  // the slot's previous batch has read its staging (long ago, unless the ring is lapped)
        memcpy(sl.edge_pin, edge_hd, eb);
        if (j->lanes < 2) {
            SUBG_CUDA(cudaGraphLaunch(sl.exec, st));
            SUBG_CUDA(cudaEventRecord(sl.done, st));
            sl.last_stream = st;
        } else {
            // The batch runs on an internal stream behind the caller's stream AS OF THE PREVIOUS SUBMIT: that point covers
            // the consumer's reads of this slot's last contents (queued at least depth - 1 submits ago) but not the wait
            // for the previous batch, so two batches are in flight.  The caller's stream then waits for this batch.
            cudaStream_t is = j->lane[j->seq++ & 1];
            if (!j->ev_user_set) SUBG_CUDA(cudaEventRecord(j->ev_user, st));   // first batch: behind everything queued so far
            SUBG_CUDA(cudaStreamWaitEvent(is, j->ev_user, 0));
            SUBG_CUDA(cudaGraphLaunch(sl.exec, is));
            SUBG_CUDA(cudaEventRecord(sl.done, is));
            SUBG_CUDA(cudaEventRecord(j->ev_user, st));
            j->ev_user_set = true;
            SUBG_CUDA(cudaStreamWaitEvent(st, sl.done, 0));
            sl.last_stream = is;
        }
    }
    count_launch(j->launches);
    if (out_dev) *out_dev = sl.out_dev;
    if (indptr_dev) *indptr_dev = (int64_t *)sl.indptr_dev;
    if (segid_dev) *segid_dev = (int64_t *)sl.segid_dev;
    if (nrows_dev) *nrows_dev = (const int64_t *)sl.tot_dev;
    if (slot_out) *slot_out = q;
    return SUBG_OK;
}

// waits for the batch last submitted to `slot` and returns its row count; SUBG_ERR_MEM if the rows did not fit the
// capacity (nothing was written: re-run that batch through subg_spjoin), SUBG_ERR_ARG for a node id outside the SpG
int joiner_rows_impl(Joiner *j, int slot, int64_t *N) {
    if (!j || slot < 0 || slot >= j->depth || !N) return fail(SUBG_ERR_ARG, "Input parsing error.");
    JoinSlot &sl = j->slot[slot];
    DeviceGuard guard(j->s->device);
    SUBG_CUDA(cudaMemcpyAsync(sl.tot_pin, sl.tot_dev, 2 * sizeof(long long), cudaMemcpyDeviceToHost, sl.last_stream));
    SUBG_CUDA(cudaStreamSynchronize(sl.last_stream));
    if (sl.tot_pin[1]) return fail(SUBG_ERR_ARG, "query node id outside the SpG");
    *N = sl.tot_pin[0];
    if (sl.tot_pin[0] > j->cap_rows) return fail(SUBG_ERR_MEM, "joiner: rows of the batch exceed the slot capacity");
    return SUBG_OK;
}

}  // namespace subg
