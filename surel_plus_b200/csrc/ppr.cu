// PPR set sampler and the PPR / SPD structure encoders on the device.
//
// Replaces (file:line relative to /root/reference)
//   _calc_ppr_node             sampler/pprgo.py:9-38     ACL forward push, LIFO queue, float32 state
//   calc_ppr_topk_parallel     sampler/pprgo.py:52-62    per-seed top-k by score
//   construct_sparse + norm.   sampler/pprgo.py:65-111   CSR assembly, 'sym' / 'col' / 'row' normalisation
//   encoding(...,'PPR'|'SPD')  utils.py:29-36
//
// The push result depends on the LIFO order and on the float32 accumulation order, so the
// queue discipline of the reference is kept exactly: one warp owns one seed and replays the
// sequential pop loop; the parallelism inside a seed is across the neighbours of the popped node
// (distinct columns => independent r[v] updates), whose queue appends are ordered by CSR
// position with a ballot.  Parallelism across seeds: persistent warps pulling seeds from an
// atomic counter.  Per-warp state (hash node -> record, record arrays, queue) lives in a
// private global-memory workspace that stays L2-warm; seeds whose support outgrows the
// first-pass workspace are re-run in a second pass sized by the push bound 1/(alpha*eps).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace subg {

constexpr int kPprWarps = 8;  // warps per CTA
constexpr unsigned long long kHEmpty = ~0ull;

struct PprArgs {
    const void *rowptr;
    int rowptr64;
    const int32_t *col;
    const int32_t *seeds;
    const long long *work;  // nullable: indices into seeds[] to process
    int64_t nwork;
    float alpha, alpha_eps;
    double one_minus_alpha;
    int topk;
    int R;          // record capacity per warp
    uint32_t hmask;  // hash capacity - 1 (capacity = 2R rounded up to a power of two)
    // per-warp workspace, indexed [warp * R + i] (hash: [warp * (hmask+1) + h])
    unsigned long long *htab;
    int32_t *node, *pord, *hslot, *q;
    float *r, *p;
    uint8_t *inq;
    unsigned long long *counter;  // [0] next work item, [1] pushes, [2] failed seeds
    uint8_t *fail;                // per seed
    // staged output rows, pitch topk
    int32_t *st_node;
    float *st_val;
    int32_t *cnt;
};

__device__ __forceinline__ int64_t ld_rowptr(const void *rowptr, int is64, int64_t i) {
    return is64 ? __ldg((const long long *)rowptr + i) : (int64_t)__ldg((const int *)rowptr + i);
}
__device__ __forceinline__ uint32_t hash_node(uint32_t v, uint32_t mask) { return mix32(v) & mask; }

__global__ void __launch_bounds__(kPprWarps * 32) ppr_push_kernel(const PprArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int per_warp = 1024 + 8 * a.topk;
    uint32_t *hist = (uint32_t *)(smem_raw + (size_t)wib * per_warp);
    int32_t *sel_node = (int32_t *)(hist + 256);
    float *sel_val = (float *)(sel_node + a.topk);

    const int64_t gw = (int64_t)blockIdx.x * kPprWarps + wib;
    unsigned long long *htab = a.htab + gw * ((int64_t)a.hmask + 1);
    int32_t *node = a.node + gw * a.R, *pord = a.pord + gw * a.R, *hslot = a.hslot + gw * a.R, *q = a.q + gw * a.R;
    float *r = a.r + gw * a.R, *p = a.p + gw * a.R;
    uint8_t *inq = a.inq + gw * a.R;
    unsigned long long pushes = 0;

    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.counter, 1ull);
        t = __shfl_sync(FULL, t, 0);
        if ((int64_t)t >= a.nwork) break;
        const int64_t i = a.work ? a.work[t] : (int64_t)t;
        const int32_t s = __ldg(a.seeds + i);

        // p = {s: 0}; r = {s: alpha}; q = [s]                          pprgo.py:12-16
        int nrec = 1, np = 1, qlen = 1;
        if (lane == 0) {
            const uint32_t h = hash_node((uint32_t)s, a.hmask);
            htab[h] = ((unsigned long long)0 << 32) | (uint32_t)s;
            hslot[0] = (int32_t)h;
            node[0] = s; r[0] = a.alpha; p[0] = 0.f; pord[0] = 0; inq[0] = 1; q[0] = 0;
        }
        __syncwarp();
        bool overflow = false;
        unsigned long long seed_pushes = 0;
        while (qlen > 0) {
            const int ui = q[qlen - 1];  // q.pop()                      pprgo.py:18
            qlen--;
            const int32_t u = node[ui];
            const float res = r[ui];
            const int po = pord[ui];
            __syncwarp();
            if (lane == 0) {
                inq[ui] = 0;
                if (po < 0) { pord[ui] = np; p[ui] = res; }              // pprgo.py:21-24
                else p[ui] += res;
                r[ui] = 0.f;                                             // pprgo.py:25
            }
            if (po < 0) np++;
            const int64_t rp0 = ld_rowptr(a.rowptr, a.rowptr64, u);
            const int64_t d = ld_rowptr(a.rowptr, a.rowptr64, (int64_t)u + 1) - rp0;
            // (1 - alpha) * res / deg[u]: float64 arithmetic rounded to float32 (pprgo.py:8,27)
            const float val = (float)(a.one_minus_alpha * (double)res / (double)d);
            __syncwarp();
            for (int64_t j0 = 0; j0 < d; j0 += 32) {
                const int64_t j = j0 + lane;
                const bool act = j < d;
                int32_t v = -1;
                int idx = -1;
                uint32_t h = 0;
                if (act) {
                    v = __ldg(a.col + rp0 + j);
                    h = hash_node((uint32_t)v, a.hmask);
                    for (;;) {
                        const unsigned long long e = __ldcg(htab + h);
                        if (e == kHEmpty) break;
                        if ((int32_t)(uint32_t)e == v) { idx = (int)(e >> 32); break; }
                        h = (h + 1) & a.hmask;
                    }
                }
                const bool isnew = act && idx < 0;
                const uint32_t newm = __ballot_sync(FULL, isnew);
                const int nnew = __popc(newm);
                if (nrec + nnew > a.R) { overflow = true; break; }
                float rv;
                if (isnew) {
                    idx = nrec + __popc(newm & lt);
                    const unsigned long long ent = ((unsigned long long)(uint32_t)idx << 32) | (uint32_t)v;
                    while (atomicCAS(htab + h, kHEmpty, ent) != kHEmpty) h = (h + 1) & a.hmask;
                    hslot[idx] = (int32_t)h;
                    node[idx] = v; p[idx] = 0.f; pord[idx] = -1;
                    rv = val;                                            // pprgo.py:30-31
                } else if (act) {
                    rv = r[idx] + val;                                   // pprgo.py:28-29
                }
                nrec += nnew;
                bool push = false;
                if (act) {
                    r[idx] = rv;
                    const int64_t dv = ld_rowptr(a.rowptr, a.rowptr64, (int64_t)v + 1) - ld_rowptr(a.rowptr, a.rowptr64, v);
                    // res_vnode >= alpha_eps * deg[vnode] (float32 product widened, pprgo.py:33-34); vnode not in q
                    push = ((double)rv >= (double)a.alpha_eps * (double)dv) && (isnew || inq[idx] == 0);
                }
                const uint32_t pm = __ballot_sync(FULL, push);
                if (act) {
                    if (push) q[qlen + __popc(pm & lt)] = idx;           // pprgo.py:35-36, CSR order
                    if (push || isnew) inq[idx] = push ? 1 : 0;
                }
                qlen += __popc(pm);
                __syncwarp();
            }
            if (overflow) break;
            seed_pushes++;
        }
        if (!overflow) pushes += seed_pushes;  // a seed re-run in the second pass is counted there

        if (overflow) {
            if (lane == 0) {
                a.fail[i] = 1;
                a.cnt[i] = 0;
                atomicAdd(a.counter + 2, 1ull);
            }
        } else {
            // ---- top-k of p by (score, insertion rank): argsort(val)[-topk:] with the ties at the k-th
            // score resolved towards later insertion (stable ascending sort, pprgo.py:59)
            unsigned long long thr = 0ull;
            if (np > a.topk) {
                unsigned long long prefix = 0ull;
                int want = a.topk;
                for (int pass = 7; pass >= 0; pass--) {
                    for (int b = lane; b < 256; b += 32) hist[b] = 0u;
                    __syncwarp();
                    for (int t0 = lane; t0 < nrec; t0 += 32) {
                        const int po = pord[t0];
                        if (po >= 0) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(p[t0]) << 32) | (uint32_t)po;
                            if (pass == 7 || (key >> (8 * (pass + 1))) == prefix)
                                atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
                        }
                    }
                    __syncwarp();
                    // lane L owns bins [8L, 8L+8); walk from the top bin down
                    uint32_t mine = 0;
#pragma unroll
                    for (int b = 0; b < 8; b++) mine += hist[lane * 8 + b];
                    // suffix sum over lanes: above = sum of lanes > lane
                    uint32_t incl = mine;
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const uint32_t o = __shfl_down_sync(FULL, incl, dd);
                        if (lane + dd < 32) incl += o;
                    }
                    const uint32_t above = incl - mine;
                    const bool here = above < (uint32_t)want && incl >= (uint32_t)want;
                    const int owner = __ffs(__ballot_sync(FULL, here)) - 1;
                    int digit = 0, rem = 0, binc = 0;
                    if (lane == owner) {
                        uint32_t acc = above;
                        for (int b = 7; b >= 0; b--) {
                            const uint32_t c = hist[lane * 8 + b];
                            if (acc + c >= (uint32_t)want) { digit = lane * 8 + b; rem = want - (int)acc; binc = (int)c; break; }
                            acc += c;
                        }
                    }
                    digit = __shfl_sync(FULL, digit, owner);
                    rem = __shfl_sync(FULL, rem, owner);
                    binc = __shfl_sync(FULL, binc, owner);
                    prefix = (prefix << 8) | (unsigned long long)digit;
                    want = rem;
                    __syncwarp();
                    if (binc == want) {  // the whole bin is selected: done
                        thr = prefix << (8 * pass);
                        break;
                    }
                    thr = prefix;  // pass 0 falls through with the exact k-th key
                }
            }
            // ---- gather the selected entries, rank them by node id, stage the row
            int nsel = 0;
            for (int t0 = 0; t0 < nrec; t0 += 32) {
                const int t = t0 + lane;
                bool keep = false;
                float pv = 0.f;
                if (t < nrec) {
                    const int po = pord[t];
                    if (po >= 0) {
                        pv = p[t];
                        const unsigned long long key = ((unsigned long long)__float_as_uint(pv) << 32) | (uint32_t)po;
                        keep = key >= thr;
                    }
                }
                const uint32_t km = __ballot_sync(FULL, keep);
                if (keep) {
                    const int o = nsel + __popc(km & lt);
                    sel_node[o] = node[t];
                    sel_val[o] = pv;
                }
                nsel += __popc(km);
            }
            __syncwarp();
            const int64_t row = i * (int64_t)a.topk;
            for (int t = lane; t < nsel; t += 32) {
                const int32_t me = sel_node[t];
                int rank = 0;
                for (int o = 0; o < nsel; o++) rank += sel_node[o] < me;
                a.st_node[row + rank] = me;
                a.st_val[row + rank] = sel_val[t];
            }
            if (lane == 0) {
                a.cnt[i] = nsel;
                a.fail[i] = 0;
            }
        }
        // ---- leave the hash empty for the next seed
        __syncwarp();
        for (int t = lane; t < nrec; t += 32) htab[hslot[t]] = kHEmpty;
        __syncwarp();
    }
    if (lane == 0 && pushes) atomicAdd(a.counter + 1, pushes);
}


// ------------------------------------------------------------------ forward push, fast path
// Same algorithm and the same floating-point operations in the same order as ppr_push_kernel (pprgo.py:9-38), with the
// per-seed state arranged so that a push costs three dependent memory round trips instead of five:
//   shared memory (per warp)  the LIFO queue (node ids) and the p-list: (node, p) of the nodes popped so far, in
//                             insertion order -- a few hundred entries; the top-k selection scans THIS list instead of
//                             every touched node (3 000 touched vs 160 popped nodes on the citation2 shape) and compacts
//                             it in place; its histogram aliases the queue, which is empty by then.  4.5 KB per warp:
//                             six CTAs (48 warps) per SM.
//   global memory             one open-addressing table per warp whose 16-byte slot IS the record:
//                               x node id | y degree | z epoch << 16 | in-queue bit 15 | p-list index + 1 | w residual r
//                             one vector load answers "seen?", r, the queue flag and the degree; an update is one 4- or
//                             8-byte store into the same sector.  A new seed bumps the epoch instead of clearing slots.
// Inserts need no atomics: the lanes of a warp hold distinct neighbours, so only the target SLOT can collide, and
// __match_any_sync elects one writer per slot; the losers probe on after a __syncwarp.  The degree of every neighbour is
// fetched (row info, 8 bytes, L2 resident) together with its first probe, not after it.
// Seeds that outgrow the queue, the p-list or the table are flagged and redone by the general kernel.
#ifndef SUBG_PPR_LAZY_DEG
#define SUBG_PPR_LAZY_DEG 1   // 1: fetch a neighbour's degree only once the probe says it is new (half the row-info requests; 197 vs 201 ms on citation2), 0: with the first probe
#endif
constexpr int kFastQ = 512;    // queue entries per warp
constexpr int kFastPDefault = 320;    // popped nodes per warp (SUBG_PPR_FAST_P)

struct PprFastArgs {
    const unsigned long long *rowinfo;  // [N] row start (low 40 bits) | degree (high 24 bits, 0xFFFFFF = read rowptr)
    const void *rowptr;
    int rowptr64;
    const int32_t *col;
    const int32_t *seeds;
    int64_t nwork;
    float alpha, alpha_eps;
    double one_minus_alpha;
    int topk;
    int R;           // touched nodes per seed (<= half the slots)
    int P;           // p-list capacity per warp
    uint32_t hmask;
    uint4 *htab;                // [warp][hmask + 1]
    uint32_t *epoch;            // [warp]: current epoch (persists across launches)
    unsigned long long *counter;  // [0] next work item, [1] pushes, [2] failed seeds
    uint8_t *fail;
    int32_t *st_node;
    float *st_val;
    int32_t *cnt;
};

__device__ __forceinline__ uint4 ld_slot(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// the row info (8 B per node, tens of MB) is what the L2 should keep; the tables stream through it
__device__ __forceinline__ unsigned long long ld_rowinfo(const PprFastArgs &a, uint32_t v, uint64_t keep) {
    unsigned long long q;
    asm("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(q) : "l"(a.rowinfo + v), "l"(keep));
    return q;
}
__device__ __forceinline__ void row_of(const PprFastArgs &a, uint32_t v, uint64_t keep, int64_t &start, uint32_t &deg) {
    const unsigned long long q = ld_rowinfo(a, v, keep);
    start = (int64_t)(q & 0xffffffffffull);
    deg = (uint32_t)(q >> 40);
    if (deg == 0xFFFFFFu) deg = (uint32_t)min(ld_rowptr(a.rowptr, a.rowptr64, (int64_t)v + 1) - start, (int64_t)0xffffffffll);
}

__global__ void __launch_bounds__(kPprWarps * 32) ppr_push_fast_kernel(const PprFastArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int per_warp = 4 * kFastQ + 8 * a.P;
    unsigned char *wsm = smem_raw + (size_t)wib * per_warp;
    int32_t *q_node = (int32_t *)wsm;
    int32_t *p_node = q_node + kFastQ;
    float *p_val = (float *)(p_node + a.P);
    uint32_t *hist = (uint32_t *)wsm;               // the queue is empty when the selection runs
    int32_t *sel_node = p_node;                     // the selection compacts the p-list in place
    float *sel_val = p_val;
    uint64_t keep;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));

    const int64_t gw = (int64_t)blockIdx.x * kPprWarps + wib;
    uint4 *htab = a.htab + gw * ((int64_t)a.hmask + 1);
    uint32_t epoch = a.epoch[gw];
    unsigned long long pushes = 0;

    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.counter, 1ull);
        t = __shfl_sync(FULL, t, 0);
        if ((int64_t)t >= a.nwork) break;
        const int64_t i = (int64_t)t;
        const int32_t s = __ldg(a.seeds + i);
        // a new epoch invalidates every slot of the previous seed; on wrap-around the table is really cleared
        epoch++;
        if (epoch >= 0xffffu) {
            for (uint32_t h = lane; h <= a.hmask; h += 32) htab[h] = make_uint4(0u, 0u, 0u, 0u);
            epoch = 1;
        }
        const uint32_t etag = epoch << 16;
        __syncwarp();

        // p = {s: 0}; r = {s: alpha}; q = [s]                          pprgo.py:12-16
        int nrec = 1, np = 0, qlen = 1;
        if (lane == 0) {
            int64_t rs;
            uint32_t ds;
            row_of(a, (uint32_t)s, keep, rs, ds);
            const uint32_t h = hash_node((uint32_t)s, a.hmask);
            htab[h] = make_uint4((uint32_t)s, ds, etag | 0x8000u, __float_as_uint(a.alpha));
            q_node[0] = s;
        }
        __syncwarp();
        bool overflow = false;
        unsigned long long seed_pushes = 0;
        while (qlen > 0) {
            qlen--;
            const int32_t u = q_node[qlen];  // q.pop()                  pprgo.py:18
            const unsigned long long urow = ld_rowinfo(a, (uint32_t)u, keep);
            uint32_t uh = hash_node((uint32_t)u, a.hmask);   // u is in the table: almost always at its first slot
            uint4 us = ld_slot(htab + uh);
            while (us.x != (uint32_t)u || (us.z & 0xffff0000u) != etag) {
                uh = (uh + 1) & a.hmask;
                us = ld_slot(htab + uh);
            }
            const float res = __uint_as_float(us.w);
            const int64_t rp0 = (int64_t)(urow & 0xffffffffffull);
            const int64_t d = (int64_t)us.y;
            int pi = (int)(us.z & 0x7fffu);  // p-list index + 1
            if (pi == 0) {                   // first pop of u: p[u] = res      pprgo.py:21-24
                if (np >= a.P) { overflow = true; break; }
                pi = ++np;
                if (lane == 0) { p_node[pi - 1] = u; p_val[pi - 1] = res; }
            } else if (lane == 0) {
                p_val[pi - 1] += res;
            }
            if (lane == 0)                   // out of the queue, r[u] = 0      pprgo.py:25
                *(uint2 *)((uint32_t *)(htab + uh) + 2) = make_uint2(etag | (uint32_t)pi, 0u);
            // (1 - alpha) * res / deg[u]: float64 arithmetic rounded to float32 (pprgo.py:8,27)
            const float val = (float)(a.one_minus_alpha * (double)res / (double)d);
            __syncwarp();
            for (int64_t j0 = 0; j0 < d; j0 += 32) {
                const int64_t j = j0 + lane;
                const bool act = j < d;
                uint32_t v = 0xffffffffu, h = 0xffffff00u + (uint32_t)lane, dv = 0;
                uint4 sl = make_uint4(0u, 0u, 0u, 0u);
                if (act) {
                    v = (uint32_t)__ldg(a.col + rp0 + j);
                    h = hash_node(v, a.hmask);
                    sl = ld_slot(htab + h);
#if !SUBG_PPR_LAZY_DEG
                    int64_t rs;
                    row_of(a, v, keep, rs, dv);   // wanted for new nodes only, but asked for before the probe answers
#endif
                }
                // lookup: probe until the node or a free (stale / never used) slot; the loop ends on a warp vote
                bool look = act, isnew = false;
                while (__any_sync(FULL, look)) {
                    if (look) {
                        if ((sl.z & 0xffff0000u) != etag) { isnew = true; look = false; }     // free slot: v is new
                        else if (sl.x == v) look = false;
                        else {
                            h = (h + 1) & a.hmask;
                            sl = ld_slot(htab + h);
                        }
                    }
                }
                const uint32_t newm = __ballot_sync(FULL, isnew);
                if (nrec + __popc(newm) > a.R) { overflow = true; break; }
                nrec += __popc(newm);
                const float rv = isnew ? val : __uint_as_float(sl.w) + val;            // pprgo.py:28-31
                const uint32_t fl = isnew ? 0u : (sl.z & 0xffffu);
#if SUBG_PPR_LAZY_DEG
                if (isnew) {
                    int64_t rs;
                    row_of(a, v, keep, rs, dv);
                }
#endif
                if (!isnew) dv = sl.y;
                // res_vnode >= alpha_eps * deg[vnode] (float32 product widened, pprgo.py:33-34); vnode not in q
                const bool push = act && ((double)rv >= (double)a.alpha_eps * (double)dv) && (fl & 0x8000u) == 0u;
                const uint32_t pm = __ballot_sync(FULL, push);
                if (qlen + __popc(pm) > kFastQ) { overflow = true; break; }
                const uint32_t tagw = etag | fl | (push ? 0x8000u : 0u);
                if (act && !isnew) {
                    if (push) *(uint2 *)((uint32_t *)(htab + h) + 2) = make_uint2(tagw, __float_as_uint(rv));
                    else ((uint32_t *)(htab + h))[3] = __float_as_uint(rv);
                }
                // insert the new nodes: distinct columns, so only the slot can collide; one writer per slot and round
                bool ins = isnew;
                while (__any_sync(FULL, ins)) {
                    const uint32_t grp = __match_any_sync(FULL, ins ? h : (0xffffff00u + (uint32_t)lane));
                    const bool win = ins && (__ffs((int)grp) - 1) == lane;
                    if (win) {
                        htab[h] = make_uint4(v, dv, tagw, __float_as_uint(rv));
                        ins = false;
                    }
                    __syncwarp();
                    if (ins) {                                   // lost the slot: probe on (sees this round's writes)
                        for (;;) {
                            h = (h + 1) & a.hmask;
                            if ((ld_slot(htab + h).z & 0xffff0000u) != etag) break;
                        }
                    }
                }
                if (push) {                                              // pprgo.py:35-36, CSR order
                    const int at = qlen + __popc(pm & lt);
                    q_node[at] = (int32_t)v;
                }
                qlen += __popc(pm);
                __syncwarp();
            }
            if (overflow) break;
            seed_pushes++;
        }

        if (overflow) {
            if (lane == 0) {
                a.fail[i] = 1;
                a.cnt[i] = 0;
                atomicAdd(a.counter + 2, 1ull);
            }
            __syncwarp();
            continue;
        }
        pushes += seed_pushes;
        __syncwarp();
        // ---- top-k of the p-list by (score, insertion rank): argsort(val)[-topk:], ties at the k-th score towards
        // later insertion (pprgo.py:59) -- the same radix selection as ppr_push_kernel, over shared memory
        unsigned long long thr = 0ull;
        if (np > a.topk) {
            unsigned long long prefix = 0ull;
            int want = a.topk;
            for (int pass = 7; pass >= 0; pass--) {
                for (int b = lane; b < 256; b += 32) hist[b] = 0u;
                __syncwarp();
                for (int t0 = lane; t0 < np; t0 += 32) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(p_val[t0]) << 32) | (uint32_t)t0;
                    if (pass == 7 || (key >> (8 * (pass + 1))) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
                }
                __syncwarp();
                uint32_t mine = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) mine += hist[lane * 8 + b];
                uint32_t incl = mine;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t o = __shfl_down_sync(FULL, incl, dd);
                    if (lane + dd < 32) incl += o;
                }
                const uint32_t above = incl - mine;
                const bool here = above < (uint32_t)want && incl >= (uint32_t)want;
                const int owner = __ffs(__ballot_sync(FULL, here)) - 1;
                int digit = 0, rem = 0, binc = 0;
                if (lane == owner) {
                    uint32_t acc = above;
                    for (int b = 7; b >= 0; b--) {
                        const uint32_t c = hist[lane * 8 + b];
                        if (acc + c >= (uint32_t)want) { digit = lane * 8 + b; rem = want - (int)acc; binc = (int)c; break; }
                        acc += c;
                    }
                }
                digit = __shfl_sync(FULL, digit, owner);
                rem = __shfl_sync(FULL, rem, owner);
                binc = __shfl_sync(FULL, binc, owner);
                prefix = (prefix << 8) | (unsigned long long)digit;
                want = rem;
                __syncwarp();
                if (binc == want) {
                    thr = prefix << (8 * pass);
                    break;
                }
                thr = prefix;
            }
        }
        int nsel = 0;
        for (int t0 = 0; t0 < np; t0 += 32) {
            const int t = t0 + lane;
            bool keep = false;
            float pv = 0.f;
            int32_t pn = 0;
            if (t < np) {
                pv = p_val[t];
                pn = p_node[t];
                keep = (((unsigned long long)__float_as_uint(pv) << 32) | (uint32_t)t) >= thr;
            }
            const uint32_t km = __ballot_sync(FULL, keep);   // every read of this batch precedes its in-place writes (o <= t)
            if (keep) {
                const int o = nsel + __popc(km & lt);
                sel_node[o] = pn;
                sel_val[o] = pv;
            }
            nsel += __popc(km);
        }
        __syncwarp();
        const int64_t row = i * (int64_t)a.topk;
        for (int t = lane; t < nsel; t += 32) {
            const int32_t me = sel_node[t];
            int rank = 0;
            for (int o = 0; o < nsel; o++) rank += sel_node[o] < me;
            a.st_node[row + rank] = me;
            a.st_val[row + rank] = sel_val[t];
        }
        if (lane == 0) {
            a.cnt[i] = nsel;
            a.fail[i] = 0;
        }
        __syncwarp();
    }
    if (lane == 0) {
        a.epoch[gw] = epoch;
        if (pushes) atomicAdd(a.counter + 1, pushes);
    }
}

// ------------------------------------------------------------------ CSR assembly + normalisation
// pprgo.py:87-106: 'sym' sqrt(max(deg_u,1e-12)) * p * (1/sqrt(max(deg_w,1e-12))), 'col' deg_u * p * (1/max(deg_w,1e-12)),
// evaluated left to right in float64; 'row' keeps p.  deg = adj.sum(1) (caller-supplied, else the row length).
__device__ __forceinline__ double deg_of(const double *ndeg, const void *rowptr, int is64, int64_t v) {
    if (ndeg) return ndeg[v];
    return (double)(ld_rowptr(rowptr, is64, v + 1) - ld_rowptr(rowptr, is64, v));
}

__global__ void ppr_assemble_kernel(const int32_t *st_node, const float *st_val, int topk, const int32_t *cnt,
                                    const long long *indptr, const int32_t *seeds, int64_t n, int norm,
                                    const double *ndeg, const void *rowptr, int is64, int32_t *indices, double *data,
                                    int32_t *max_set) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int mx = 0;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int c = cnt[i];
        const int64_t src = i * (int64_t)topk, dst = indptr[i];
        const int32_t u = seeds[i];
        double du = 0.0;
        if (norm == 1) du = sqrt(fmax(deg_of(ndeg, rowptr, is64, u), 1e-12));
        else if (norm == 2) du = deg_of(ndeg, rowptr, is64, u);
        for (int j = lane; j < c; j += 32) {
            const int32_t w = st_node[src + j];
            double x = (double)st_val[src + j];
            if (norm == 1) {
                const double dw = 1.0 / sqrt(fmax(deg_of(ndeg, rowptr, is64, w), 1e-12));
                x = __dmul_rn(__dmul_rn(du, x), dw);
            } else if (norm == 2) {
                const double dw = 1.0 / fmax(deg_of(ndeg, rowptr, is64, w), 1e-12);
                x = __dmul_rn(__dmul_rn(du, x), dw);
            }
            indices[dst + j] = w;
            data[dst + j] = x;
        }
        mx = max(mx, c);
    }
    if (lane == 0 && mx > 0) atomicMax(max_set, mx);
}

// ------------------------------------------------------------------ encoder 'PPR' (utils.py:35-36)
__global__ void max_f64_kernel(const double *x, int64_t n, unsigned long long *out_bits) {
    double m = 0.0;  // data > 0 on this path; bit patterns of non-negative doubles order like the values
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmax(m, x[i]);
    for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));
}
__global__ void ppr_rescale_kernel(const double *in, double *out, int64_t n, const unsigned long long *max_bits) {
    const double den = __dadd_rn(__longlong_as_double((long long)*max_bits), 0.1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __ddiv_rn(__dadd_rn(in[i], 0.1), den);
}

// ------------------------------------------------------------------ graph property checks (cached on the Graph)
__global__ void check_sorted_kernel(const void *rowptr, int is64, const int32_t *col, int64_t N, uint32_t *bad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < N; u += nwarps) {
        const int64_t b = ld_rowptr(rowptr, is64, u), e = ld_rowptr(rowptr, is64, u + 1);
        for (int64_t j = b + 1 + lane; j < e; j += 32)
            if (col[j - 1] >= col[j]) atomicOr(bad, 1u);
    }
}
__device__ __forceinline__ bool row_contains(const int32_t *col, int64_t b, int64_t e, int32_t w) {
    while (b < e) {
        const int64_t mid = (b + e) >> 1;
        const int32_t c = __ldg(col + mid);
        if (c < w) b = mid + 1;
        else if (c > w) e = mid;
        else return true;
    }
    return false;
}
__global__ void check_symmetric_kernel(const void *rowptr, int is64, const int32_t *col, int64_t N, uint32_t *bad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < N; u += nwarps) {
        const int64_t b = ld_rowptr(rowptr, is64, u), e = ld_rowptr(rowptr, is64, u + 1);
        for (int64_t j = b + lane; j < e; j += 32) {
            const int32_t v = col[j];
            if (!row_contains(col, ld_rowptr(rowptr, is64, v), ld_rowptr(rowptr, is64, (int64_t)v + 1), (int32_t)u))
                atomicOr(bad, 1u);
        }
    }
}

// ------------------------------------------------------------------ encoder 'SPD' (utils.py:29-34)
// x0 = x > 0 (the PPR set S_u), x1 = adj > 0, x2 = x1**2 (boolean square: some v with u->v and v->w);
// x = x1 + x0.multiply(x2*0.5) + x0*0.3; setdiag(2.3).  Row u = N(u) U S_u U {u}, ascending.
// Pass 1 classifies every member of S_u (bit0: w in N(u), bit1: two-hop) and counts the row;
// pass 2 writes it, every element computing its own output position by binary searches.
struct SpdArgs {
    const void *rowptr;
    int rowptr64;
    const int32_t *col;
    int symmetric;
    int64_t N;
    const long long *s_indptr;  // input value SpG (rows = nodes 0..N-1)
    const int32_t *s_indices;
    uint8_t *code;              // [T_in]
    int32_t *rowcnt;            // [N]
    const long long *o_indptr;
    int32_t *o_indices;
    double *o_data;
    int32_t *max_set;
    int cap;                    // smem entries per warp (>= max set size of the input + 1)
};

__device__ __forceinline__ bool two_hop(const SpdArgs &a, int64_t ub, int64_t ue, int32_t w, int lane) {
    // exists v in N(u) with w in N(v).  Symmetric graphs: intersect N(u) with N(w), scanning the shorter list.
    const int64_t wb = ld_rowptr(a.rowptr, a.rowptr64, w), we = ld_rowptr(a.rowptr, a.rowptr64, (int64_t)w + 1);
    if (a.symmetric) {
        int64_t sb = ub, se = ue, lb = wb, le = we;
        if (we - wb < ue - ub) { sb = wb; se = we; lb = ub; le = ue; }
        for (int64_t j0 = sb; j0 < se; j0 += 32) {
            const int64_t j = j0 + lane;
            const bool hit = j < se && row_contains(a.col, lb, le, __ldg(a.col + j));
            if (__any_sync(FULL, hit)) return true;
        }
        return false;
    }
    for (int64_t j0 = ub; j0 < ue; j0 += 32) {
        const int64_t j = j0 + lane;
        bool hit = false;
        if (j < ue) {
            const int32_t v = __ldg(a.col + j);
            hit = row_contains(a.col, ld_rowptr(a.rowptr, a.rowptr64, v), ld_rowptr(a.rowptr, a.rowptr64, (int64_t)v + 1), w);
        }
        if (__any_sync(FULL, hit)) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128) spd_classify_kernel(const SpdArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < a.N; u += nwarps) {
        const int64_t ub = ld_rowptr(a.rowptr, a.rowptr64, u), ue = ld_rowptr(a.rowptr, a.rowptr64, u + 1);
        const int64_t sb = a.s_indptr[u], se = a.s_indptr[u + 1];
        int extra = 0;
        bool self_seen = false;
        for (int64_t t = sb; t < se; t++) {  // members of S_u one at a time, lanes share the searches
            const int32_t w = a.s_indices[t];
            bool in_n = false;
            if (lane == 0) in_n = row_contains(a.col, ub, ue, w);
            in_n = __shfl_sync(FULL, (int)in_n, 0) != 0;
            const bool th = two_hop(a, ub, ue, w, lane);
            if (lane == 0) a.code[t] = (uint8_t)((in_n ? 1 : 0) | (th ? 2 : 0));
            if (!in_n) extra++;
            if (w == (int32_t)u) self_seen = true;
        }
        if (!self_seen) {
            bool in_n = false;
            if (lane == 0) in_n = row_contains(a.col, ub, ue, (int32_t)u);
            in_n = __shfl_sync(FULL, (int)in_n, 0) != 0;
            if (!in_n) extra++;
        }
        if (lane == 0) a.rowcnt[u] = (int32_t)(ue - ub) + extra;
    }
}

__global__ void __launch_bounds__(128) spd_fill_kernel(const SpdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t *s_node = (int32_t *)(smem_raw + (size_t)wib * a.cap * 12);
    int32_t *x_node = s_node + a.cap;            // extras: S_u \ N(u) (+ u), ascending
    uint8_t *s_code = (uint8_t *)(x_node + a.cap);
    uint8_t *x_code = s_code + a.cap;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int mx = 0;
    for (int64_t u = warp; u < a.N; u += nwarps) {
        const int64_t ub = ld_rowptr(a.rowptr, a.rowptr64, u), ue = ld_rowptr(a.rowptr, a.rowptr64, u + 1);
        const int64_t sb = a.s_indptr[u];
        const int ns = (int)(a.s_indptr[u + 1] - sb);
        const int64_t ob = a.o_indptr[u];
        // stage S_u and compact the extras (bit2 of a code marks "u inserted only for the diagonal")
        int nx = 0;
        bool self_in_s = false;
        for (int t0 = 0; t0 < ns; t0 += 32) {
            const int t = t0 + lane;
            int32_t w = -1;
            uint8_t c = 0;
            if (t < ns) {
                w = a.s_indices[sb + t];
                c = a.code[sb + t];
                s_node[t] = w;
                s_code[t] = c;
            }
            const bool ex = t < ns && !(c & 1);
            const uint32_t em = __ballot_sync(FULL, ex);
            if (ex) {
                const int o = nx + __popc(em & ((1u << lane) - 1u));
                x_node[o] = w;
                x_code[o] = c;
            }
            nx += __popc(em);
            if (__any_sync(FULL, t < ns && w == (int32_t)u)) self_in_s = true;
        }
        __syncwarp();
        if (!self_in_s) {
            bool in_n = false;
            if (lane == 0) in_n = row_contains(a.col, ub, ue, (int32_t)u);
            in_n = __shfl_sync(FULL, (int)in_n, 0) != 0;
            if (!in_n) {  // insert u into the sorted extras
                int pos = 0;
                for (int o = lane; o < nx; o += 32) pos += x_node[o] < (int32_t)u;
                for (int dd = 16; dd > 0; dd >>= 1) pos += __shfl_xor_sync(FULL, pos, dd);
                __syncwarp();
                if (lane == 0) {
                    for (int o = nx; o > pos; o--) { x_node[o] = x_node[o - 1]; x_code[o] = x_code[o - 1]; }
                    x_node[pos] = (int32_t)u;
                    x_code[pos] = 4;
                }
                nx++;
                __syncwarp();
            }
        }
        // neighbours: position = own index + #extras below
        for (int64_t j = ub + lane; j < ue; j += 32) {
            const int32_t w = __ldg(a.col + j);
            int lo = 0, hi = nx;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (x_node[mid] < w) lo = mid + 1; else hi = mid; }
            const int below = lo;
            lo = 0; hi = ns;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_node[mid] < w) lo = mid + 1; else hi = mid; }
            const bool in_s = lo < ns && s_node[lo] == w;
            double x = 1.0;                                               // x1
            if (in_s && (s_code[lo] & 2)) x = __dadd_rn(x, 0.5);          // + x0.multiply(x2 * 0.5)
            if (in_s) x = __dadd_rn(x, 0.3);                              // + x0 * 0.3
            if (w == (int32_t)u) x = 2.3;                                 // setdiag(2.3)
            a.o_indices[ob + (j - ub) + below] = w;
            a.o_data[ob + (j - ub) + below] = x;
        }
        // extras: position = own index + #neighbours below
        for (int o = lane; o < nx; o += 32) {
            const int32_t w = x_node[o];
            int64_t lo = ub, hi = ue;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(a.col + mid) < w) lo = mid + 1; else hi = mid; }
            double x = 0.0;
            if (x_code[o] & 2) x = __dadd_rn(x, 0.5);
            if (!(x_code[o] & 4)) x = __dadd_rn(x, 0.3);
            if (w == (int32_t)u) x = 2.3;
            a.o_indices[ob + (lo - ub) + o] = w;
            a.o_data[ob + (lo - ub) + o] = x;
        }
        mx = max(mx, (int)(ue - ub) + nx);
        __syncwarp();
    }
    if (lane == 0 && mx > 0) atomicMax(a.max_set, mx);
}

// ------------------------------------------------------------------ host side
static int64_t env_i64(const char *name, int64_t dflt) {
    const char *v = getenv(name);
    return v ? atoll(v) : dflt;
}

#define CKG(call)                                                                              \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            rc = fail(_e == cudaErrorMemoryAllocation ? SUBG_ERR_MEM : SUBG_ERR_CUDA,          \
                      std::string(#call) + ": " + cudaGetErrorString(_e));                     \
            goto done;                                                                         \
        }                                                                                      \
    } while (0)

// 1 = every row strictly ascending (sorted, no duplicate columns), 0 = not; cached
static int graph_sorted(const Graph *g, cudaStream_t st, int *out) {
    if (g->sorted_state < 0) {
        uint32_t *bad = nullptr, h = 0;
        SUBG_CUDA(dmalloc(&bad, 1, st));
        SUBG_CUDA(cudaMemsetAsync(bad, 0, 4, st));
        check_sorted_kernel<<<8 * g->num_sms, 256, 0, st>>>(g->rowptr, g->rowptr64, g->col, g->N, bad);
        SUBG_CUDA(cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, st));
        SUBG_CUDA(cudaStreamSynchronize(st));
        dfree(bad, st);
        count_launch(1);
        g->sorted_state = h ? 0 : 1;
    }
    *out = g->sorted_state;
    return SUBG_OK;
}
static int graph_symmetric(const Graph *g, cudaStream_t st, int *out) {
    if (g->sym_state < 0) {
        uint32_t *bad = nullptr, h = 0;
        SUBG_CUDA(dmalloc(&bad, 1, st));
        SUBG_CUDA(cudaMemsetAsync(bad, 0, 4, st));
        check_symmetric_kernel<<<8 * g->num_sms, 256, 0, st>>>(g->rowptr, g->rowptr64, g->col, g->N, bad);
        SUBG_CUDA(cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, st));
        SUBG_CUDA(cudaStreamSynchronize(st));
        dfree(bad, st);
        count_launch(1);
        g->sym_state = h ? 0 : 1;
    }
    *out = g->sym_state;
    return SUBG_OK;
}

struct PushWorkspace {
    unsigned long long *htab = nullptr;
    int32_t *node = nullptr, *pord = nullptr, *hslot = nullptr, *q = nullptr;
    float *r = nullptr, *p = nullptr;
    uint8_t *inq = nullptr;
    void release(cudaStream_t st) {
        dfree(htab, st); dfree(node, st); dfree(pord, st); dfree(hslot, st); dfree(q, st);
        dfree(r, st); dfree(p, st); dfree(inq, st);
        *this = PushWorkspace();
    }
};

static __global__ void ppr_check_seeds_kernel(const int32_t *seeds, int64_t n, int64_t N, uint32_t *bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (seeds[i] < 0 || seeds[i] >= N) atomicOr(bad, 1u);
}

__global__ void fill_empty_kernel(unsigned long long *p, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = kHEmpty;
}

int ppr_topk_impl(const Graph *g, const int32_t *seeds_hd, int64_t n, float alpha, float eps, int topk,
                  int normalization, const double *norm_deg_hd, cudaStream_t st, SpG **out) {
    if (!g || !out || n < 0 || (n > 0 && !seeds_hd)) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (topk < 1 || topk > 4096) return fail(SUBG_ERR_ARG, "topk must be in [1, 4096]");
    if (!(alpha > 0.f && alpha < 1.f) || !(eps > 0.f)) return fail(SUBG_ERR_ARG, "need 0 < alpha < 1 and eps > 0");
    if (normalization < 0 || normalization > 2) return fail(SUBG_ERR_ARG, "Unknown PPR normalization");  // pprgo.py:109
    DeviceGuard guard(g->device);
    int sorted = 0;
    if (int rc0 = graph_sorted(g, st, &sorted)) return rc0;
    if (!sorted)
        return fail(SUBG_ERR_UNSUPPORTED, "PPR sampler needs CSR rows with strictly ascending columns (scipy canonical format)");

    g->tag.use_on(st);
    SpG *s = new SpG();
    s->tag.last = st;
    s->device = g->device; s->n = n; s->value_kind = 1; s->num_sms = g->num_sms; s->ncol = 1;
    int rc = SUBG_OK;
    PushWorkspace ws;
    unsigned long long *counter = nullptr;
    uint8_t *fail_d = nullptr;
    int32_t *st_node = nullptr, *cnt = nullptr, *d_max = nullptr;
    float *st_val = nullptr;
    long long *scan_scratch = nullptr, *work = nullptr;
    double *ndeg = nullptr;
    bool ndeg_owned = false;
    std::vector<uint8_t> hfail;
    std::vector<long long> hwork;
    {
        CKG(dmalloc(&s->seeds, (size_t)n, st));
        CKG(dmalloc(&s->indptr, (size_t)n + 1, st));
        CKG(dmalloc(&counter, 4, st));
        CKG(dmalloc(&fail_d, (size_t)n, st));
        CKG(dmalloc(&st_node, (size_t)n * topk, st));
        CKG(dmalloc(&st_val, (size_t)n * topk, st));
        CKG(dmalloc(&cnt, (size_t)n, st));
        CKG(dmalloc(&d_max, 2, st));
        CKG(dmalloc(&scan_scratch, (size_t)std::max(1, scan_num_blocks(n)), st));
        CKG(cudaMemsetAsync(d_max, 0, 8, st));
        if (n > 0) CKG(cudaMemcpyAsync(s->seeds, seeds_hd, (size_t)n * 4, cudaMemcpyDefault, st));
        if (n > 0) {
            uint32_t hbad = 0;
            ppr_check_seeds_kernel<<<std::min<int64_t>((n + 255) / 256, 4 * g->num_sms), 256, 0, st>>>(s->seeds, n, g->N, (uint32_t *)(d_max + 1));
            CKG(cudaMemcpyAsync(&hbad, d_max + 1, 4, cudaMemcpyDeviceToHost, st));
            CKG(cudaStreamSynchronize(st));
            if (hbad) { rc = fail(SUBG_ERR_ARG, "idx contains node ids outside [0, N)"); goto done; }
        }
        if (norm_deg_hd && normalization != 0) {
            if (is_device_ptr(norm_deg_hd)) ndeg = const_cast<double *>(norm_deg_hd);
            else {
                CKG(dmalloc(&ndeg, (size_t)g->N, st));
                ndeg_owned = true;
                CKG(cudaMemcpyAsync(ndeg, norm_deg_hd, (size_t)g->N * 8, cudaMemcpyHostToDevice, st));
            }
        }

        const float alpha_eps = alpha * eps;  // float32 product (pprgo.py:11 under numba typing)
        const int smem = kPprWarps * (1024 + 8 * topk);
        if (smem > 200 * 1024) { rc = fail(SUBG_ERR_UNSUPPORTED, "topk too large for the shared-memory selection buffers"); goto done; }
        CKG(cudaFuncSetAttribute(ppr_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        CKG(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppr_push_kernel, kPprWarps * 32, smem));
        per_sm = std::max(per_sm, 1);

        // pass 0 (fast kernel): all seeds, queue and p-list in shared memory, records capped; pass 1 (general kernel): the
        // seeds that outgrew one of those; pass 2: what is left, with workspaces sized by the push bound
        //   #records <= 1 + deg(seed) + sum_pushes deg(u) <= ~ 1/(alpha*eps) + max_deg  (capped by N)
        int64_t nwork = n;
        const bool use_fast = env_i64("SUBG_PPR_FAST", 1) != 0;
        for (int pass = use_fast ? 0 : 1; pass < 3 && nwork > 0; pass++) {
            const bool fast = pass == 0;
            int64_t R;
            if (pass < 2) R = std::min<int64_t>(env_i64("SUBG_PPR_RECORDS", 8192), g->N + 1);
            else R = g->N + 1;  // a record per node can never overflow
            R = std::max<int64_t>(R, 64);
            if (fast) R = std::min<int64_t>(R, 65535);
            uint32_t H = 64;
            while ((int64_t)H < 2 * R) H <<= 1;
            int smem_k = smem, per_sm_k = per_sm, fast_p = 0;
            if (fast) {
                fast_p = (int)std::min<int64_t>(std::max<int64_t>(env_i64("SUBG_PPR_FAST_P", kFastPDefault), 32), 2048);
                smem_k = kPprWarps * (4 * kFastQ + 8 * fast_p);
                CKG(cudaFuncSetAttribute(ppr_push_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_k));
                CKG(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_k, ppr_push_fast_kernel, kPprWarps * 32, smem_k));
                per_sm_k = std::max(per_sm_k, 1);
                const int64_t capb = env_i64("SUBG_PPR_BLOCKS", 0);
                if (capb > 0) per_sm_k = (int)std::min<int64_t>(per_sm_k, capb);
            }
            int64_t blocks = (int64_t)g->num_sms * per_sm_k;
            blocks = std::min<int64_t>(blocks, (nwork + kPprWarps - 1) / kPprWarps);
            const int64_t bytes_per_warp = fast ? (int64_t)H * 16 : (int64_t)H * 8 + R * 25;
            const int64_t budget = env_i64("SUBG_PPR_WORKSPACE_BYTES", 24ll << 30);
            blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, budget / (bytes_per_warp * kPprWarps)));
            const int64_t nw = blocks * kPprWarps;
            CKG(dmalloc(&ws.htab, (size_t)nw * H * (fast ? 2 : 1), st));     // fast: 16-byte slots
            if (!fast) {
                CKG(dmalloc(&ws.node, (size_t)nw * R, st));
                CKG(dmalloc(&ws.r, (size_t)nw * R, st));
                CKG(dmalloc(&ws.pord, (size_t)nw * R, st));
                CKG(dmalloc(&ws.hslot, (size_t)nw * R, st)); CKG(dmalloc(&ws.q, (size_t)nw * R, st));
                CKG(dmalloc(&ws.p, (size_t)nw * R, st));
                CKG(dmalloc(&ws.inq, (size_t)nw * R, st));
                fill_empty_kernel<<<8 * g->num_sms, 256, 0, st>>>(ws.htab, nw * (int64_t)H);
            } else {
                CKG(dmalloc(&ws.q, (size_t)nw, st));                        // fast: the per-warp epochs
                CKG(cudaMemsetAsync(ws.q, 0, (size_t)nw * 4, st));
                CKG(cudaMemsetAsync(ws.htab, 0, (size_t)nw * H * 16, st));  // epoch 0 = never written
            }
            CKG(cudaMemsetAsync(counter, 0, 4 * sizeof(unsigned long long), st));
            timing_begin(SUBG_TIMING_PPR, st);
            if (fast) {
                PprFastArgs a{};
                a.rowinfo = (const unsigned long long *)g->rowinfo;
                a.rowptr = g->rowptr; a.rowptr64 = g->rowptr64 ? 1 : 0; a.col = g->col; a.seeds = s->seeds; a.nwork = nwork;
                a.alpha = alpha; a.alpha_eps = alpha_eps; a.one_minus_alpha = 1.0 - (double)alpha;
                a.topk = topk; a.R = (int)R; a.P = fast_p; a.hmask = H - 1;
                a.htab = (uint4 *)ws.htab; a.epoch = (uint32_t *)ws.q;
                a.counter = counter; a.fail = fail_d; a.st_node = st_node; a.st_val = st_val; a.cnt = cnt;
                ppr_push_fast_kernel<<<(unsigned)blocks, kPprWarps * 32, smem_k, st>>>(a);
            } else {
                PprArgs a{};
                a.rowptr = g->rowptr; a.rowptr64 = g->rowptr64 ? 1 : 0; a.col = g->col; a.seeds = s->seeds;
                a.work = work; a.nwork = nwork;   // work == nullptr: every seed (no earlier pass ran)
                a.alpha = alpha; a.alpha_eps = alpha_eps; a.one_minus_alpha = 1.0 - (double)alpha;
                a.topk = topk; a.R = (int)R; a.hmask = H - 1;
                a.htab = ws.htab; a.node = ws.node; a.pord = ws.pord; a.hslot = ws.hslot; a.q = ws.q;
                a.r = ws.r; a.p = ws.p; a.inq = ws.inq; a.counter = counter; a.fail = fail_d;
                a.st_node = st_node; a.st_val = st_val; a.cnt = cnt;
                ppr_push_kernel<<<(unsigned)blocks, kPprWarps * 32, smem, st>>>(a);
            }
            timing_end(SUBG_TIMING_PPR, st);
            CKG(cudaGetLastError());
            count_launch(2);
            unsigned long long hc[3];
            CKG(cudaMemcpyAsync(hc, counter, sizeof(hc), cudaMemcpyDeviceToHost, st));
            CKG(cudaStreamSynchronize(st));
            ws.release(st);
            s->pushes += (int64_t)hc[1];
            const int64_t nfail = (int64_t)hc[2];
            if (nfail == 0) break;
            if (pass == 2) { rc = fail(SUBG_ERR_MEM, "PPR push workspace overflow in the full-size pass"); goto done; }
            hfail.resize((size_t)n);
            CKG(cudaMemcpyAsync(hfail.data(), fail_d, (size_t)n, cudaMemcpyDeviceToHost, st));
            CKG(cudaStreamSynchronize(st));
            hwork.clear();
            for (int64_t i = 0; i < n; i++)
                if (hfail[i]) hwork.push_back(i);
            nwork = (int64_t)hwork.size();
            dfree(work, st);
            work = nullptr;
            CKG(dmalloc(&work, (size_t)nwork, st));
            CKG(cudaMemcpyAsync(work, hwork.data(), (size_t)nwork * 8, cudaMemcpyHostToDevice, st));
            if (pass >= 1) s->status |= SUBG_STATUS_PPR_SECOND_PASS;
        }

        // ---- CSR assembly
        timing_begin(SUBG_TIMING_BUILD, st);
        CKG(exclusive_scan_i32_i64(cnt, (long long *)s->indptr, n, 0, scan_scratch, st));
        long long T = 0;
        CKG(cudaMemcpyAsync(&T, s->indptr + n, 8, cudaMemcpyDeviceToHost, st));
        CKG(cudaStreamSynchronize(st));
        s->T = T; s->rowbeg = s->indptr; s->extent = T; s->cap = T;
        CKG(dmalloc(&s->indices, (size_t)T + 16, st));
        CKG(dmalloc((unsigned char **)&s->data, ((size_t)T + 16) * 8, st));   // multi-GB at full size: the block cache, not the driver pool
        if (n > 0) {
            const int64_t blocks = std::min<int64_t>((n * 32 + 255) / 256, 8 * (int64_t)g->num_sms);
            ppr_assemble_kernel<<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, st>>>(
                st_node, st_val, topk, cnt, (const long long *)s->indptr, s->seeds, n, normalization, ndeg, g->rowptr,
                g->rowptr64 ? 1 : 0, s->indices, (double *)s->data, d_max);
            CKG(cudaGetLastError());
        }
        timing_end(SUBG_TIMING_BUILD, st);
        count_launch(4);
        int32_t mx = 0;
        CKG(cudaMemcpyAsync(&mx, d_max, 4, cudaMemcpyDeviceToHost, st));
        CKG(cudaStreamSynchronize(st));
        s->max_set = mx;
    }
done:
    ws.release(st);
    dfree(counter, st); dfree(fail_d, st); dfree(st_node, st); dfree(st_val, st); dfree(cnt, st); dfree(d_max, st);
    dfree(scan_scratch, st); dfree(work, st);
    if (ndeg_owned) dfree(ndeg, st);
    if (rc != SUBG_OK) {
        spg_free_impl(s);
        return rc;
    }
    *out = s;
    return SUBG_OK;
}

int spg_encode_impl(const Graph *g, const SpG *x, int encoder, cudaStream_t st, SpG **out) {
    if (!x || !out) return fail(SUBG_ERR_ARG, "Input parsing error.");
    if (x->value_kind != 1) return fail(SUBG_ERR_ARG, "structure encoders take a value SpG (PPR scores)");
    if (encoder != SUBG_ENCODER_PPR && encoder != SUBG_ENCODER_SPD) return fail(SUBG_ERR_UNSUPPORTED, "encoder must be 'PPR' or 'SPD'");  // utils.py:37-38
    DeviceGuard guard(x->device);
    int rc = SUBG_OK;
    if (g) g->tag.use_on(st);
    x->tag.use_on(st);
    SpG *s = new SpG();
    s->tag.last = st;
    s->device = x->device; s->n = x->n; s->value_kind = 1; s->num_sms = x->num_sms; s->ncol = 1;
    unsigned long long *mx_bits = nullptr;
    uint8_t *code = nullptr;
    int32_t *rowcnt = nullptr, *d_max = nullptr;
    long long *scan_scratch = nullptr;
    {
        const int64_t n = x->n, T = x->T;
        if (encoder == SUBG_ENCODER_PPR) {
            // x.data = (x.data + 0.1) / (x.data.max() + 0.1)                      utils.py:36
            CKG(dmalloc(&s->indptr, (size_t)n + 1, st));
            CKG(dmalloc(&s->indices, (size_t)T + 16, st));
            CKG(dmalloc((unsigned char **)&s->data, ((size_t)T + 16) * 8, st));   // multi-GB at full size: the block cache, not the driver pool
            CKG(dmalloc(&mx_bits, 1, st));
            CKG(cudaMemsetAsync(mx_bits, 0, 8, st));
            CKG(cudaMemcpyAsync(s->indptr, x->indptr, ((size_t)n + 1) * 8, cudaMemcpyDeviceToDevice, st));
            if (T > 0) {
                CKG(cudaMemcpyAsync(s->indices, x->indices, (size_t)T * 4, cudaMemcpyDeviceToDevice, st));
                const unsigned blocks = (unsigned)std::min<int64_t>((T + 255) / 256, 8 * (int64_t)x->num_sms);
                max_f64_kernel<<<blocks, 256, 0, st>>>((const double *)x->data, T, mx_bits);
                ppr_rescale_kernel<<<blocks, 256, 0, st>>>((const double *)x->data, (double *)s->data, T, mx_bits);
                CKG(cudaGetLastError());
                count_launch(2);
            }
            s->T = T; s->max_set = x->max_set; s->rowbeg = s->indptr; s->extent = T; s->cap = T;
            if (x->seeds) {
                CKG(dmalloc(&s->seeds, (size_t)n, st));
                if (n > 0) CKG(cudaMemcpyAsync(s->seeds, x->seeds, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            }
            CKG(cudaStreamSynchronize(st));
        } else {
            if (!g) { rc = fail(SUBG_ERR_ARG, "the SPD encoder needs the graph"); goto done; }
            if (g->device != x->device) { rc = fail(SUBG_ERR_ARG, "graph and SpG live on different devices"); goto done; }
            if (n != g->N) { rc = fail(SUBG_ERR_ARG, "SPD encoder: the SpG must have one row per graph node (idx = arange(N))"); goto done; }
            int sorted = 0, sym = 0;
            if ((rc = graph_sorted(g, st, &sorted))) goto done;
            if (!sorted) { rc = fail(SUBG_ERR_UNSUPPORTED, "SPD encoder needs CSR rows with strictly ascending columns"); goto done; }
            if ((rc = graph_symmetric(g, st, &sym))) goto done;
            CKG(dmalloc(&code, (size_t)T + 1, st));
            CKG(dmalloc(&rowcnt, (size_t)n + 1, st));
            CKG(dmalloc(&d_max, 1, st));
            CKG(cudaMemsetAsync(d_max, 0, 4, st));
            CKG(dmalloc(&scan_scratch, (size_t)std::max(1, scan_num_blocks(n)), st));
            CKG(dmalloc(&s->indptr, (size_t)n + 1, st));
            SpdArgs a{};
            a.rowptr = g->rowptr; a.rowptr64 = g->rowptr64 ? 1 : 0; a.col = g->col; a.symmetric = sym; a.N = n;
            a.s_indptr = (const long long *)x->indptr; a.s_indices = x->indices; a.code = code; a.rowcnt = rowcnt;
            a.max_set = d_max; a.cap = ((x->max_set + 1 + 3) & ~3) + 4;
            const size_t smem = (size_t)4 * a.cap * 12;
            if (smem > 200 * 1024) { rc = fail(SUBG_ERR_UNSUPPORTED, "SPD encoder: input sets too large for shared memory"); goto done; }
            CKG(cudaFuncSetAttribute(spd_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 3) / 4, 16 * (int64_t)g->num_sms));
            timing_begin(SUBG_TIMING_BUILD, st);
            spd_classify_kernel<<<blocks, 128, 0, st>>>(a);
            CKG(cudaGetLastError());
            CKG(exclusive_scan_i32_i64(rowcnt, (long long *)s->indptr, n, 0, scan_scratch, st));
            long long To = 0;
            CKG(cudaMemcpyAsync(&To, s->indptr + n, 8, cudaMemcpyDeviceToHost, st));
            CKG(cudaStreamSynchronize(st));
            CKG(dmalloc(&s->indices, (size_t)To + 16, st));
            CKG(dmalloc((unsigned char **)&s->data, ((size_t)To + 16) * 8, st));
            a.o_indptr = (const long long *)s->indptr; a.o_indices = s->indices; a.o_data = (double *)s->data;
            spd_fill_kernel<<<blocks, 128, smem, st>>>(a);
            CKG(cudaGetLastError());
            timing_end(SUBG_TIMING_BUILD, st);
            count_launch(5);
            int32_t mx = 0;
            CKG(cudaMemcpyAsync(&mx, d_max, 4, cudaMemcpyDeviceToHost, st));
            CKG(cudaStreamSynchronize(st));
            s->T = To; s->max_set = mx; s->rowbeg = s->indptr; s->extent = To; s->cap = To;
        }
    }
done:
    dfree(mx_bits, st); dfree(code, st); dfree(rowcnt, st); dfree(d_max, st); dfree(scan_scratch, st);
    if (rc != SUBG_OK) {
        spg_free_impl(s);
        return rc;
    }
    *out = s;
    return SUBG_OK;
}

}  // namespace subg
