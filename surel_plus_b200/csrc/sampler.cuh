// Walk-based set sampler + LP encoder: one warp per seed; a seed's walks, its dedup/sort and its
// landing counts never leave the SM, and the finished set is written once, straight into the SpG.
//
// Reference behaviour reproduced (file:line relative to /root/reference):
//   first hop without replacement, later hops uniform    subg_acc/subg_acc.c:763-809
//   per-seed dedup, first-visit slot order, LP counts    subg_acc/subg_acc.c:784-844
//   bucket overflow drops late nodes, walk continues      subg_acc/subg_acc.c:814-828
//   64-bit LP key, LEAD bit on the root                   subg_acc/subg_acc.c:900-955
//   first-occurrence ids of unique LP rows                subg_acc/subg_acc.c:957-978
//
// B200 design (per warp, per seed):
//   1. walk: lanes own walks; every hop issues GW x 32 independent gathers (row info as one
//      8/16-byte load with an L2 evict_last policy, neighbour column with evict_first), draws
//      come from Philox4x32-10 (one call = one hop of four walks).  Every visit becomes a key
//      (node << OB | order) in the warp's shared-memory key buffer; order = 0 for the root and
//      ((walk + 1) << LS | step) otherwise, so ascending order == the reference's first-visit order.
//   2. sort: blocked load (EPL keys per lane), register sorting network per lane, then five
//      merge-path rounds through shared memory.  Equal nodes become runs: the run head carries
//      the first visit, the run's population per step is the LP row.
//   3. encode: per-lane packed (4 x 16 bit) step counts, a warp scan stitches runs that straddle
//      lanes; members are compacted to (key, counts) records in shared memory.
//   4. emit: the row is allocated with one atomic on a global cursor (16-byte aligned rows) and
//      written coalesced in ascending node order, the order the SpG CSR-of-sets needs; the LP row
//      is interned in an L2-resident hash table that also tracks its first stream position.
#pragma once
#include <type_traits>
#include <utility>

#include "common.cuh"

namespace subg {

constexpr int kWarpsPerBlock = 4;
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint32_t kStatusTableFull = 1u << 31;  // internal status bits
constexpr uint32_t kStatusBadSeed = 1u << 30;
constexpr int kFirstHopCap = 1000000;  // NEBMAX, subg_acc.c:13,750
constexpr int kGW = 8;                 // walks advanced together per lane (template GW: 8, or 4 when num_walks <= 128)
constexpr int kCtrCursor = 16, kCtrTotal = 32, kCtrWords = 48;

struct SamplerArgs {
    const unsigned long long *rowinfo;  // [N] row start (low 40 bits) | degree (high 24 bits, 0xFFFFFF = escape)
    const void *rowptr;    // escape path only
    int rowptr64;
    const int32_t *col;
    const unsigned long long *col3;  // nullable: the same ids, three per 64-bit word (21 bits each)
    const int32_t *seeds;  // chunk-local [n_chunk]
    int64_t n_chunk;
    int64_t seed_base;     // global index of seeds[0]
    int64_t N;             // nodes of the graph
    int M, m, stride, Kt;  // stride = reference's bucket stride (cap on set size); Kt = M*m+1 keys per seed
    int OB, LS;            // bits of the order field; log2 of the step slots per walk
    int SHIFT;             // 32 - clz(M), bits per LP column in the key
    int rng_mode;
    uint32_t rng_lo, rng_hi;
    const int64_t *call_base;  // RAND_R: exclusive prefix of rand_r calls, global seed index
    const int32_t *walks;      // TRACE: chunk-local [n_chunk, M, m]
    int32_t *dump_walks;       // nullable, chunk-local [n_chunk, M, m]: the walks drawn (SUBG_SAMPLE_DUMP_WALKS)
    // output rows: row i occupies [rowbeg[i], rowbeg[i] + nsize[i]) of the three arrays
    int32_t *out_node;
    int32_t *out_prov;
    uint16_t *out_slot;        // nullable: first-visit ranks are not wanted
    long long *rowbeg;         // chunk-local
    int32_t *nsize;            // chunk-local
    unsigned long long *ctr;   // [0] seed ticket  [16] row cursor (entries, rows padded to 4)  [32] sum of set sizes
                               // (one 128-byte line each: same-address atomics serialise in their L2 slice)
    int32_t *max_set;
    int want_rank;
    int blocks_per_sm;         // 0 = as many as fit; otherwise a cap (fewer blocks leave more of the SM's 228 KB to L1)
    int stop_after;            // measurement only (SUBG_SAMPLER_STOP): 1 = walks, 2 = + sort, 3 = + counts, 4 = all but the LP-row lookups,
                               // 5 = all but the row stores; 0 = full kernel
    // LP-key intern table (global, L2 resident)
    unsigned long long *tab_key;
    unsigned long long *tab_pos;
    uint32_t tab_mask;
    uint32_t *tab_count;
    uint32_t *status;
    // shared memory carve-up (per warp)
    int nbw;         // bitmap words
    int fy_cap;      // Fisher-Yates overflow map capacity (power of two)
    int lp_off;      // byte offset of the member LP rows (region 2; region 1 at 0 = key buffer / member keys)
    int lp64;        // LP rows need 64 bits (m * SHIFT + 1 > 32)
    int bitmap_off;  // byte offset of the rank bitmap
    int smem_per_warp;
};

// ---------------------------------------------------------------- cache-policy loads
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint2 ldg_v2_hint(const void *p, uint64_t pol) {
    uint2 v;
    asm("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
struct Policies { uint64_t keep; };

// issue / decode split so that a group's row loads are all in flight before the first is consumed
__device__ __forceinline__ uint2 load_row_raw(const SamplerArgs &a, const Policies &pol, uint32_t v) {
    return ldg_v2_hint(a.rowinfo + v, pol.keep);  // the row info (8 B per node) is what the L2 should keep
}
__device__ __forceinline__ void decode_row(const SamplerArgs &a, uint32_t v, uint2 q, int64_t &start, uint32_t &deg) {
    start = (int64_t)(((uint64_t)(q.y & 0xffu) << 32) | q.x);
    deg = q.y >> 8;
    if (deg == 0xFFFFFFu) {  // hub with >= 2^24 - 1 neighbours
        const int64_t e = a.rowptr64 ? (int64_t)__ldg((const long long *)a.rowptr + v + 1)
                                     : (int64_t)__ldg((const int *)a.rowptr + v + 1);
        deg = (uint32_t)min(e - start, (int64_t)0xffffffffll);
    }
}
__device__ __forceinline__ uint32_t load_col(const SamplerArgs &a, const Policies &pol, int64_t e) {
    if (a.col3) {  // packed columns (E < 2^32): word e / 3, field e % 3
        const uint32_t e32 = (uint32_t)e;
        const uint32_t w = __umulhi(e32, 0xAAAAAAABu) >> 1;
        const unsigned long long v = __ldg(a.col3 + w);
        return (uint32_t)(v >> (21u * (e32 - 3u * w))) & 0x1fffffu;
    }
    return (uint32_t)__ldg(a.col + e);  // cache hints on these make no difference (profiles/r1_gather_micro.txt)
}

// ---------------------------------------------------------------- warp merge sort
template <typename K>
__device__ __forceinline__ void cswap(K &a, K &b) {
    const K lo = a < b ? a : b;
    const K hi = a < b ? b : a;
    a = lo;
    b = hi;
}

// Batcher odd-even merge sort network on EPL registers (any EPL: comparators that would touch
// indices >= EPL are those of the next power of two with +inf inputs, i.e. no-ops).  The comparator
// list is computed at compile time and applied through a pack expansion, so every register index is
// a constant (a loop nest with these bounds is not reliably unrolled and would spill k[] to local memory).
struct NetCE { int a, b; };
constexpr NetCE net_walk(int n, int want, int *count) {
    int c = 0;
    for (int p = 1; p < n; p <<= 1)
        for (int q = p; q >= 1; q >>= 1)
            for (int j = q % p; j + q < n; j += 2 * q)
                for (int i = 0; i < q && i + j + q < n; i++)
                    if ((i + j) / (2 * p) == (i + j + q) / (2 * p)) {
                        if (c == want) return NetCE{i + j, i + j + q};
                        c++;
                    }
    if (count) *count = c;
    return NetCE{0, 0};
}
constexpr int net_size(int n) {
    int c = 0;
    net_walk(n, -1, &c);
    return c;
}
template <typename K, int EPL, int I>
__device__ __forceinline__ void net_ce(K (&k)[EPL]) {
    constexpr NetCE c = net_walk(EPL, I, nullptr);
    cswap(k[c.a], k[c.b]);
}
template <typename K, int EPL, int... I>
__device__ __forceinline__ void net_apply(K (&k)[EPL], std::integer_sequence<int, I...>) {
    (net_ce<K, EPL, I>(k), ...);
}
template <typename K, int EPL>
__device__ __forceinline__ void lane_sort(K (&k)[EPL]) {
    net_apply<K, EPL>(k, std::make_integer_sequence<int, net_size(EPL)>{});
}

// Sorts the 32*EPL keys of a warp ascending.  In: lane L holds elements [L*EPL, (L+1)*EPL) of any
// order.  Out: the same blocked layout, globally sorted.  buf: 32*EPL + 64 keys of shared memory owned by
// the warp.  Keys are distinct except for the padding (~0 - 1); ~0 itself never occurs in the data: it is the
// sentinel that follows every run in shared memory, so the serial merge reads without bounds tests (an exhausted
// run presents ~0, which loses against every key including the padding).
#ifndef SUBG_MERGE_BIDIR
#define SUBG_MERGE_BIDIR 1
#endif
template <typename K, int EPL>
__device__ __forceinline__ void warp_merge_sort(K (&k)[EPL], K *buf, int lane) {
    constexpr K SENT = ~(K)0;
    lane_sort<K, EPL>(k);
#if SUBG_MERGE_BIDIR
    // Every lane merges its EPL outputs from both ends at once: the first half forwards from its own merge-path point,
    // the second half backwards from the next lane's point (one shuffle) -- two independent chains of dependent
    // shared-memory loads instead of one.  Run q occupies buf[q * (L + 2) + 1 ..] between a 0 slot (an exhausted run
    // presents the minimum to the backward chain) and a ~0 slot (the maximum, for the forward chain).  The only key that
    // can equal 0 is the root of node 0, which is output position 0 of the whole sort: forward territory.
    constexpr int HF = (EPL + 1) / 2;
#pragma unroll 1
    for (int r = 0; r < 5; r++) {
        const int L = EPL << r;
        const int rl = lane & ((1 << r) - 1);    // lane index inside its run
        const int own = lane * EPL + 2 * (lane >> r) + 1;
#pragma unroll
        for (int j = 0; j < EPL; j++) buf[own + j] = k[j];
        if (rl == 0) buf[own - 1] = (K)0;
        if (rl == (1 << r) - 1) buf[own + EPL] = SENT;
        __syncwarp();
        const int t = lane & ((2 << r) - 1);     // lane index inside the pair of runs
        const int a0 = (lane - t) * EPL + 2 * ((lane - t) >> r) + 1;   // first key of run A; run B starts L + 2 later
        const K *A = buf + a0;
        const K *B = A + L + 2;
        const int diag = t * EPL;
        int lo = diag > L ? diag - L : 0;
        int hi = diag < L ? diag : L;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (A[mid] <= B[diag - 1 - mid]) lo = mid + 1;
            else hi = mid;
        }
        // the next lane's split is where this lane's outputs end; the last lane of the pair ends at (L, L)
        int lo_n = __shfl_down_sync(FULL, lo, 1);
        if (t == (2 << r) - 1) lo_n = L;
        int pa = a0 + lo, pb = a0 + L + 2 + diag - lo;
        int qa = a0 + lo_n - 1, qb = a0 + L + 2 + (diag + EPL - lo_n) - 1;
        K ka = buf[pa], kb = buf[pb];
        K la = buf[qa], lb = buf[qb];
#pragma unroll
        for (int j = 0; j < HF; j++) {
            const bool ta = ka <= kb;
            k[j] = ta ? ka : kb;
            if (j + 1 < HF) {
                pa += ta ? 1 : 0;
                pb += ta ? 0 : 1;
                const K v = buf[ta ? pa : pb];
                ka = ta ? v : ka;
                kb = ta ? kb : v;
            }
            const int jb = EPL - 1 - j;
            if (jb >= HF) {
                const bool tb = la > lb;         // the larger tail goes last; ties (padding only) take B
                k[jb] = tb ? la : lb;
                if (jb - 1 >= HF) {
                    qa -= tb ? 1 : 0;
                    qb -= tb ? 0 : 1;
                    const K v = buf[tb ? qa : qb];
                    la = tb ? v : la;
                    lb = tb ? lb : v;
                }
            }
        }
        __syncwarp();
    }
#else
#pragma unroll 1
    for (int r = 0; r < 5; r++) {
        const int L = EPL << r;                  // run length; run q occupies buf[q * (L + 1) ..] + one sentinel slot
        const int own = lane * EPL + (lane >> r);
#pragma unroll
        for (int j = 0; j < EPL; j++) buf[own + j] = k[j];
        if ((lane & ((1 << r) - 1)) == (1 << r) - 1) buf[own + EPL] = SENT;  // last lane of its run
        __syncwarp();
        const int t = lane & ((2 << r) - 1);     // lane index inside the pair of runs
        const int a0 = (lane - t) * EPL + ((lane - t) >> r);   // first key of run A; run B starts L + 1 later
        const K *A = buf + a0;
        const K *B = A + L + 1;
        const int diag = t * EPL;
        int lo = diag > L ? diag - L : 0;
        int hi = diag < L ? diag : L;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (A[mid] <= B[diag - 1 - mid]) lo = mid + 1;
            else hi = mid;
        }
        // serial merge of this lane's EPL outputs; pa / pb index buf
        int pa = a0 + lo, pb = a0 + L + 1 + diag - lo;
        K ka = buf[pa];
        K kb = buf[pb];
#pragma unroll
        for (int j = 0; j < EPL; j++) {
            const bool ta = ka <= kb;
            k[j] = ta ? ka : kb;
            if (j + 1 < EPL) {
                pa += ta ? 1 : 0;
                pb += ta ? 0 : 1;
                const K v = buf[ta ? pa : pb];
                ka = ta ? v : ka;
                kb = ta ? kb : v;
            }
        }
        __syncwarp();
    }
#endif
}

// Power-of-two EPL: bitonic sort of the 32*EPL keys entirely in registers -- in-lane compare-exchanges for strides below
// EPL, one shuffle per key for the strides that cross lanes.  Written in the all-ascending form (the first step of every
// merge pairs i with i ^ (size - 1), the later steps i with i ^ stride), so no comparator needs a direction flag: the lower
// index keeps the minimum (a predicated VIMNMX).  Same blocked layout in and out as warp_merge_sort, no shared memory, no
// searches: at 8 keys per lane 36 steps of 8 two-instruction operations against five merge-path rounds (binary search +
// serial merge each).  The steps that cross lanes differ only in the lane mask, so they are ROLLED loops over one code
// block each (fully unrolled, the 16-key network alone is 24 KB of SASS and the kernel waits on instruction fetch:
// stalled_no_instruction 4.5, profiles/r3a_sampler_collab.txt); only the in-lane steps are unrolled.
template <typename K, int EPL>
__device__ __forceinline__ void bitonic_inlane_halving(K (&k)[EPL], int first_stride) {
#pragma unroll
    for (int stride = EPL / 2; stride >= 1; stride >>= 1) {
        if (stride <= first_stride) {
#pragma unroll
            for (int j = 0; j < EPL; j++) {
                const int p = j ^ stride;
                if (p > j) cswap(k[j], k[p]);
            }
        }
    }
}
template <typename K, int EPL>
__device__ __forceinline__ void warp_bitonic_sort(K (&k)[EPL], int lane) {
    static_assert((EPL & (EPL - 1)) == 0, "EPL must be a power of two");
    // merges of size 2 .. EPL stay inside the lane
#pragma unroll
    for (int size = 2; size <= EPL; size <<= 1) {
#pragma unroll
        for (int j = 0; j < EPL; j++) {
            const int p = j ^ (size - 1);
            if (p > j) cswap(k[j], k[p]);
        }
        bitonic_inlane_halving<K, EPL>(k, size / 4);
    }
    // merges of size 2 EPL .. 32 EPL: lanes per merge lm2 = 2, 4, .., 32
#pragma unroll 1
    for (int lm2 = 2; lm2 <= 32; lm2 <<= 1) {
        {   // flip step: lane ^ (lm2 - 1), register EPL - 1 - j
            const bool lower = (lane & (lm2 >> 1)) == 0;
            K o[EPL];
#pragma unroll
            for (int j = 0; j < EPL; j++) o[j] = __shfl_xor_sync(FULL, k[EPL - 1 - j], lm2 - 1);
#pragma unroll
            for (int j = 0; j < EPL; j++) {
                const K mn = k[j] < o[j] ? k[j] : o[j];
                const K mx = k[j] < o[j] ? o[j] : k[j];
                k[j] = lower ? mn : mx;
            }
        }
#pragma unroll 1
        for (int lm = lm2 >> 2; lm >= 1; lm >>= 1) {   // halving steps that cross lanes: lane ^ lm, same register
            const bool lower = (lane & lm) == 0;
#pragma unroll
            for (int j = 0; j < EPL; j++) {
                const K o = __shfl_xor_sync(FULL, k[j], lm);
                const K mn = k[j] < o ? k[j] : o;
                const K mx = k[j] < o ? o : k[j];
                k[j] = lower ? mn : mx;
            }
        }
        bitonic_inlane_halving<K, EPL>(k, EPL / 2);
    }
}

template <typename T>
__device__ __forceinline__ T warp_incl_scan_acc(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T t = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

// ---------------------------------------------------------------- LP-key interning
// An LP row is keyed as in the reference: SHIFT bits per step, step 1 in the top field, the root
// marker (LEAD) above them (subg_acc.c:936-949); the landing counts are accumulated directly in
// that packing (a field cannot overflow: counts <= M < 2^SHIFT).
// Open-addressing table in global memory (a few MB, L2 resident).  intern returns the slot of
// `key`; tab_pos[slot] keeps the smallest stream position at which the key occurs, which
// later yields the reference's first-occurrence ids (subg_acc.c:957-978).
__device__ __forceinline__ uint32_t lp_hash(unsigned long long key) {
    uint32_t h = (uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h;
}
// first probe issued by the caller (so that several lookups are in flight): cur0 / pos0 are the
// table words at h0 = lp_hash(key) & mask
__device__ __forceinline__ uint32_t intern_key(const SamplerArgs &a, unsigned long long key, unsigned long long pos,
                                               uint32_t h, unsigned long long cur, unsigned long long seen_pos) {
    bool first = true;
    for (uint32_t probes = 0;; probes++) {
        if (probes > a.tab_mask) {  // table full: the host sees tab_count > cap/2 and reruns with a larger one
            atomicOr(a.status, kStatusTableFull);
            return 0;
        }
        if (!first) cur = a.tab_key[h];
        if (cur == kEmptyKey) {
            cur = atomicCAS(&a.tab_key[h], kEmptyKey, key);
            if (cur == kEmptyKey) {
                atomicAdd(a.tab_count, 1u);
                cur = key;
            }
        }
        if (cur == key) break;
        h = (h + 1) & a.tab_mask;
        first = false;
    }
    if (!first) seen_pos = a.tab_pos[h];
    if (pos < seen_pos) atomicMin(&a.tab_pos[h], pos);
    return h;
}

template <typename K, int EPL>
constexpr int sampler_min_blocks() {
    constexpr int W = (int)sizeof(K) / 4;
    constexpr int smem_warp = ((int)sizeof(K) + 4) * 32 * EPL + 64 * (int)sizeof(K) + 256;
    constexpr int by_smem = 232448 / (kWarpsPerBlock * smem_warp);
    // registers the compiler is asked to fit: keys + 45 (19 keys per lane -> 64 registers; measured on ppa against 72: same 7
    // resident CTAs, kernel 5.96 -> 5.79 ms, profiles/r2_sampler_sweeps.txt)
    constexpr int by_regs = 65536 / (kWarpsPerBlock * 32 * (EPL * W + 45));
#ifndef SUBG_SAMPLER_EXTRA_BLOCK
#define SUBG_SAMPLER_EXTRA_BLOCK 0   // experiment: ask the compiler for one more resident CTA than the register estimate gives
#endif
    constexpr int b = (by_smem < by_regs ? by_smem : by_regs) + SUBG_SAMPLER_EXTRA_BLOCK;
    return b < 1 ? 1 : (b > 8 ? 8 : b);
}

// ---------------------------------------------------------------- the sampler kernel
// PARITY = false: Philox draws only (the fast path); true: rand_r replay and supplied traces (kept out
// of the fast kernel: its code has to stay inside the instruction cache).
// LEAN = true: additionally no first-visit ranks, no bucket cap (stride >= M*m+1) and LP rows of at most 32 bits -- the
// configuration subg_matrix runs in -- with those paths compiled out (the kernel is bound by instruction issue and
// fetch: every kilobyte of SASS that is not executed still competes for the instruction caches).
// GW: walks a lane advances together.  A lane owns ceil(M / 32) walks; with M <= 128 the slots 4..7 of a group of 8 would
// never hold a walk but would still run their share of the Philox calls and address arithmetic (dblp / twitter shapes,
// M = 100: a quarter of the kernel's instructions).  The Philox counters of the first group do not depend on GW (and GW < 8
// is only chosen when one group holds all of a lane's walks): same walks either way.  GW = 7 for 128 < M <= 224 (slot 7 idle
// at M = 200) was measured too: ppa 5.790 -> 5.777 ms, collab 0.963 -> 0.957 ms -- not worth its instantiations.
template <typename K, int EPL, bool PARITY, bool LEAN, int GW = kGW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, sampler_min_blocks<K, EPL>()) gset_sample_kernel(const SamplerArgs a) {
    using Acc = std::conditional_t<LEAN, uint32_t, unsigned long long>;   // packed landing counts of one member
    const int stop_after = LEAN ? 0 : a.stop_after;                      // the measurement hooks are not in the lean kernel
    const bool want_rank = LEAN ? false : (a.want_rank != 0);
    const bool lp64 = LEAN ? false : (a.lp64 != 0);
    uint16_t *const out_slot = LEAN ? nullptr : a.out_slot;
    constexpr K SENT = ~(K)0;
    constexpr K PAD = SENT - 1;  // fills the key slots beyond M*m+1: above every real key, below the merge sentinel
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char *wsm = smem_raw + (size_t)wib * a.smem_per_warp;
    K *keys = (K *)wsm;
    // member records alias the key buffer (the keys are in registers by then)
    K *rec_key = (K *)wsm;  // member i's head key overwrites the key buffer (i <= position of the head)
    uint32_t *rec_lp32 = (uint32_t *)(wsm + a.lp_off);   // low words of the member LP rows
    uint32_t *rec_lphi = rec_lp32 + a.Kt;                // high words (lp64 only)
    // Fisher-Yates scratch: picks and hash keys in region 1 (dead before the first key is written),
    // the permutation and the hash values in region 2 (read by the first hop)
    int32_t *fy_pick = (int32_t *)wsm;
    int32_t *fy_key = fy_pick + a.M;
    int32_t *fy_dense = (int32_t *)(wsm + a.lp_off);
    int32_t *fy_val = fy_dense + a.M;
    uint32_t *bitmap = (uint32_t *)(wsm + a.bitmap_off);
    uint32_t *bprefix = bitmap + a.nbw;
    const int M = a.M, m = a.m, OB = a.OB, LS = a.LS;
    const uint32_t ord_mask = (1u << OB) - 1u;
    const uint32_t step_mask = (1u << LS) - 1u;
    const K no_node = SENT >> OB;
    const int lp_top = a.SHIFT * (m - 1);  // step s lands in bits [SHIFT*(m-1-s), SHIFT*(m-s)) (subg_acc.c:936-943)
    Policies pol;
    pol.keep = l2_policy_evict_last();
    int mx = 0;

    // seeds are handed out by a ticket counter; the ticket of the NEXT seed is requested before this one is processed
    unsigned long long ticket = 0;
    if (lane == 0) ticket = atomicAdd(&a.ctr[0], 1ull);
    for (;;) {
        const int64_t i = (int64_t)__shfl_sync(FULL, ticket, 0);
        if (i >= a.n_chunk) break;
        if (lane == 0) ticket = atomicAdd(&a.ctr[0], 1ull);
        const int64_t gi = a.seed_base + i;
        const int32_t u = __ldg(a.seeds + i);
        if ((uint64_t)(int64_t)u >= (uint64_t)a.N) {  // the host turns this into the reference's TypeError
            if (lane == 0) {
                atomicOr(a.status, kStatusBadSeed);
                a.nsize[i] = 0;
                a.rowbeg[i] = 0;
            }
            continue;
        }

        if (want_rank)
            for (int b = lane; b < a.nbw; b += 32) bitmap[b] = 0u;

        if (PARITY && a.rng_mode == SUBG_RNG_TRACE) {
            const int32_t *wk = a.walks + i * (int64_t)M * m;
            for (int j = lane; j < M * m; j += 32) {
                const uint32_t v = (uint32_t)__ldg(wk + j);
                const int w = j / m, s = j - w * m;
                keys[1 + j] = ((K)v << OB) | (K)((((uint32_t)w + 1u) << LS) | (uint32_t)s);
            }
        } else {
            int64_t rp0;
            uint32_t dfull;
            decode_row(a, (uint32_t)u, load_row_raw(a, pol, (uint32_t)u), rp0, dfull);
            const int d = dfull > (uint32_t)kFirstHopCap ? kFirstHopCap : (int)dfull;
            const bool replay = PARITY && a.rng_mode == SUBG_RNG_RAND_R;
            const uint32_t gi_lo = (uint32_t)gi, gi_hi = (uint32_t)((uint64_t)gi >> 32);
            int64_t calls0 = 0;
            if (replay) calls0 = __ldg((const long long *)a.call_base + gi);

            // ---- first hop without replacement (subg_acc.c:763-776, 790-800)
            if (d > M) {
                if (replay) {
                    for (int k = lane; k < M; k += 32) {
                        uint32_t st = lcg_jump(a.rng_lo, 3u * (uint32_t)(calls0 + k));
                        fy_pick[k] = (int32_t)(rand_r_dev(st) % (uint32_t)(d - k) + k);
                        fy_dense[k] = k;
                    }
                } else {
                    for (int c = lane; 4 * c < M; c += 32) {
                        const uint4 r4 = philox4x32_10(make_uint4(gi_lo, gi_hi, (uint32_t)c, 0x46597331u),
                                                       make_uint2(a.rng_lo, a.rng_hi));
                        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int k = 4 * c + q;
                            if (k < M) {
                                fy_pick[k] = (int32_t)(k + __umulhi(rr[q], (uint32_t)(d - k)));
                                fy_dense[k] = k;
                            }
                        }
                    }
                }
                for (int h = lane; h < a.fy_cap; h += 32) fy_key[h] = -1;
                __syncwarp();
                if (lane == 0) {
                    const int hm = a.fy_cap - 1;
                    for (int k = 0; k < M; k++) {
                        const int s = fy_pick[k];
                        const int vk = fy_dense[k];
                        if (s < M) {
                            const int vs = fy_dense[s];
                            fy_dense[s] = vk;
                            fy_dense[k] = vs;
                        } else {
                            int p = (int)(mix32((uint32_t)s) & (uint32_t)hm);
                            while (fy_key[p] != -1 && fy_key[p] != s) p = (p + 1) & hm;
                            const int vs = (fy_key[p] == s) ? fy_val[p] : s;
                            fy_key[p] = s;
                            fy_val[p] = vk;
                            fy_dense[k] = vs;
                        }
                    }
                }
                __syncwarp();
                calls0 += M;
            }

            // Walks are advanced in groups of GW per lane: all loads of one hop of a group are
            // issued back to back (GW x 32 gathers in flight per warp).
            const int rstep = d > 0 ? 32 % d : 0;   // w % d for w = lane + 32 t, kept incrementally
            int rr0 = d > 0 ? lane % d : 0;
            for (int g = 0; g * 32 < M; g += GW) {
                uint32_t cur[GW];
#pragma unroll
                for (int tt = 0; tt < GW; tt++) {
                    const int w = lane + 32 * (g + tt);
                    cur[tt] = (uint32_t)u;
                    if (w < M && d > 0) {
                        const int off = (d <= M) ? rr0 : fy_dense[w];
                        cur[tt] = load_col(a, pol, rp0 + off);
                    }
                    if (d > 0) {
                        rr0 += rstep;
                        if (rr0 >= d) rr0 -= d;
                    }
                    if (w < M) {
                        keys[1 + w] = ((K)cur[tt] << OB) | (K)(((uint32_t)w + 1u) << LS);
                        if (a.dump_walks) a.dump_walks[(i * M + w) * m] = (int32_t)cur[tt];
                    }
                }
                // ---- later hops, uniform with replacement (subg_acc.c:802-809)
                uint32_t rst[GW];
                if (replay && m > 1) {
#pragma unroll
                    for (int tt = 0; tt < GW; tt++) {
                        const int w = lane + 32 * (g + tt);
                        rst[tt] = lcg_jump(a.rng_lo, 3u * (uint32_t)(calls0 + (int64_t)w * (m - 1)));
                    }
                }
                for (int s = 1; s < m; s++) {
                    uint2 raw[GW];
#pragma unroll
                    for (int tt = 0; tt < GW; tt++) raw[tt] = load_row_raw(a, pol, cur[tt]);
                    constexpr int NC = (GW + 3) / 4;   // Philox calls per hop and lane (GW = 7: the 8th draw is not used)
                    uint32_t draw[4 * NC];
                    if (!replay) {  // one Philox call = this hop of four walks
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            const uint4 r4 = philox4x32_10(
                                make_uint4(gi_lo, gi_hi, (uint32_t)(lane + 8 * g + 32 * c) | ((uint32_t)s << 16), 0x57414c4bu),
                                make_uint2(a.rng_lo, a.rng_hi));
                            draw[4 * c] = r4.x; draw[4 * c + 1] = r4.y; draw[4 * c + 2] = r4.z; draw[4 * c + 3] = r4.w;
                        }
                    }
#pragma unroll
                    for (int tt = 0; tt < GW; tt++) {
                        const int w = lane + 32 * (g + tt);
                        if (w < M) {
                            int64_t rp;
                            uint32_t dn;
                            decode_row(a, cur[tt], raw[tt], rp, dn);
                            if (dn > 0) {
                                uint32_t off;
                                if (replay) off = rand_r_dev(rst[tt]) % dn;
                                else off = __umulhi(draw[tt], dn);
                                cur[tt] = load_col(a, pol, rp + off);
                            } else if (replay && d > 0) {
                                atomicOr(a.status, SUBG_STATUS_DEAD_END);
                            }
                            keys[1 + s * M + w] = ((K)cur[tt] << OB) | (K)((((uint32_t)w + 1u) << LS) | (uint32_t)s);
                            if (a.dump_walks) a.dump_walks[(i * M + w) * m + s] = (int32_t)cur[tt];
                        }
                    }
                }
            }
        }
        if (lane == 0) keys[0] = (K)(uint32_t)u << OB;  // the root: order 0
        for (int j = a.Kt + lane; j < 32 * EPL; j += 32) keys[j] = PAD;
        __syncwarp();

        if (stop_after == 1) continue;
        K k[EPL];
#pragma unroll
        for (int r = 0; r < EPL; r++) k[r] = keys[lane * EPL + r];
        __syncwarp();
        if constexpr ((EPL & (EPL - 1)) == 0 && sizeof(K) == 4) warp_bitonic_sort<K, EPL>(k, lane);
        else warp_merge_sort<K, EPL>(k, keys, lane);
        if (stop_after == 2) {
            if (k[0] == 1 && lane == 33) a.nsize[i] = 0;  // keep the sort alive
            continue;
        }

        // ---- runs of equal node = one set member each; heads per lane
        const K prev_last = __shfl_up_sync(FULL, k[EPL - 1], 1);
        const K pn0 = lane ? (prev_last >> OB) : no_node;
        uint32_t nhead = 0;
#pragma unroll
        for (int r = 0; r < EPL; r++) {
            const K pn = r ? (k[r - 1] >> OB) : pn0;
            const bool head = k[r] < PAD && (k[r] >> OB) != pn;
            nhead += head ? 1u : 0u;
        }
        const uint32_t incl = warp_incl_scan(nhead);
        const int s_total = (int)__shfl_sync(FULL, incl, 31);
        const int kept = LEAN ? s_total : (s_total < a.stride ? s_total : a.stride);
        const int kept4 = (kept + 3) & ~3;
        unsigned long long base_u = 0;
        if (lane == 0) {
            base_u = atomicAdd(&a.ctr[kCtrCursor], (unsigned long long)kept4);
            atomicAdd(&a.ctr[kCtrTotal], (unsigned long long)kept);
            a.rowbeg[i] = (long long)base_u;
            a.nsize[i] = kept;
            if (kept < s_total) atomicOr(a.status, SUBG_STATUS_BUCKET_OVERFLOW);
        }
        mx = kept > mx ? kept : mx;

        // ---- landing counts per run: packed 4 x 16 bit, runs that straddle lanes are stitched by a scan
        {
            Acc acc = 0, lead = 0;
            bool have = false;
            K curk = 0;
            int idx = (int)(incl - nhead) - 1;
#pragma unroll
            for (int r = 0; r < EPL; r++) {
                const K pn = r ? (k[r - 1] >> OB) : pn0;
                const bool valid = k[r] < PAD;
                const bool head = valid && (k[r] >> OB) != pn;
                const uint32_t ord = (uint32_t)k[r] & ord_mask;
                if (head) {
                    if (have) {
                        rec_key[idx] = curk;
                        rec_lp32[idx] = (uint32_t)acc;
                        if (lp64) rec_lphi[idx] = (uint32_t)((unsigned long long)acc >> 32);
                    } else {
                        lead = acc;
                    }
                    have = true;
                    curk = k[r];
                    acc = 0;
                    idx++;
                    if (want_rank) atomicOr(&bitmap[ord >> 5], 1u << (ord & 31));
                }
                if (valid && ord) acc += (Acc)1 << (lp_top - a.SHIFT * (int)(ord & step_mask));
            }
            if (!have) lead = acc;
            const Acc S = warp_incl_scan_acc(lead);
            const uint32_t H = __ballot_sync(FULL, have);
            const uint32_t above = lane == 31 ? 0u : (H & ~((2u << lane) - 1u));
            const int nh = above ? (__ffs((int)above) - 1) : 31;
            const Acc S_nh = __shfl_sync(FULL, S, nh);
            if (have) {
                rec_key[idx] = curk;
                const Acc tot = acc + (S_nh - S);
                rec_lp32[idx] = (uint32_t)tot;
                if (lp64) rec_lphi[idx] = (uint32_t)((unsigned long long)tot >> 32);
            }
        }
        __syncwarp();

        if (stop_after == 3) continue;
        // ---- first-visit rank of every member = popcount prefix over the order bitmap
        if (want_rank) {
            uint32_t running = 0;
            for (int b0 = 0; b0 < a.nbw; b0 += 32) {
                const int b = b0 + lane;
                const uint32_t cnt = b < a.nbw ? __popc(bitmap[b]) : 0u;
                const uint32_t inc = warp_incl_scan(cnt);
                if (b < a.nbw) bprefix[b] = running + inc - cnt;
                running += __shfl_sync(FULL, inc, 31);
            }
            __syncwarp();
        }

        // ---- emit the set: ascending node id, provisional LP id, first-visit rank.  Two members per
        // lane and iteration so that two table lookups are in flight.
        const long long base = (long long)__shfl_sync(FULL, base_u, 0);
        const bool overflow = LEAN ? false : kept < s_total;
        int done = 0;
        for (int t0 = 0; t0 < s_total; t0 += 64) {
            K kk[2];
            unsigned long long lp[2], cur[2], seen[2];
            uint32_t h[2], ord[2], rank[2];
            bool keep[2];
            int o[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int t = t0 + 32 * q + lane;
                const bool act = t < s_total;
                kk[q] = 0;
                lp[q] = 0ull;
                if (act) {
                    kk[q] = rec_key[t];
                    lp[q] = rec_lp32[t];
                    if (lp64) lp[q] |= (unsigned long long)rec_lphi[t] << 32;
                }
                ord[q] = (uint32_t)kk[q] & ord_mask;
                rank[q] = 0;
                if (want_rank && act)
                    rank[q] = bprefix[ord[q] >> 5] + __popc(bitmap[ord[q] >> 5] & ((1u << (ord[q] & 31)) - 1u));
                keep[q] = act;
                o[q] = t;
                if (overflow) {
                    keep[q] = act && (int)rank[q] < a.stride;
                    const uint32_t km = __ballot_sync(FULL, keep[q]);
                    o[q] = done + __popc(km & ((1u << lane) - 1u));
                    done += __popc(km);
                }
                if (ord[q] == 0) lp[q] |= 1ull << (m * a.SHIFT);  // the root row (LEAD, subg_acc.c:944-949)
                h[q] = lp_hash(lp[q]) & a.tab_mask;
                cur[q] = kEmptyKey;
                seen[q] = 0ull;
                if (keep[q] && stop_after != 4) {
                    cur[q] = a.tab_key[h[q]];
                    seen[q] = a.tab_pos[h[q]];
                }
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (keep[q]) {
                    // measurement knobs: stop_after 4 = no LP-row lookups, 5 = lookups but no row stores (results invalid)
                    const uint32_t prov = stop_after == 4 ? h[q]
                                                            : intern_key(a, lp[q], ((unsigned long long)gi << 16) | ord[q], h[q], cur[q], seen[q]);
                    if (stop_after != 5) {
                        a.out_node[base + o[q]] = (int32_t)(kk[q] >> OB);
                        a.out_prov[base + o[q]] = (int32_t)prov;
                        if (out_slot) out_slot[base + o[q]] = (uint16_t)rank[q];
                    } else if (prov == 0xffffffffu) {
                        a.out_prov[base] = 0;  // keeps the lookup alive
                    }
                }
            }
        }
        if (lane < kept4 - kept) {  // keep the row padding defined (ids are remapped in place later)
            a.out_node[base + kept + lane] = 0x7fffffff;
            a.out_prov[base + kept + lane] = 0;
            if (out_slot) out_slot[base + kept + lane] = 0;
        }
        __syncwarp();  // the record area is the next seed's key buffer
    }
    if (lane == 0 && mx > 0) atomicMax(a.max_set, mx);
}

// rand_r calls consumed per seed in the reference's single stream:
// M for the Fisher-Yates draw if deg > M, plus M*(m-1) later hops if deg > 0.
static __global__ void rand_r_calls_kernel(const void *rowptr, int rowptr64, const int32_t *seeds, int64_t n,
                                    int64_t N, int M, int m, int32_t *calls) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = seeds[i];
        if (u < 0 || u >= N) { calls[i] = 0; continue; }  // reported by check_seeds_kernel
        int64_t d = rowptr64 ? ((const long long *)rowptr)[u + 1] - ((const long long *)rowptr)[u]
                             : (int64_t)((const int *)rowptr)[u + 1] - ((const int *)rowptr)[u];
        if (d > kFirstHopCap) d = kFirstHopCap;
        calls[i] = (d > M ? M : 0) + (d > 0 ? M * (m - 1) : 0);
    }
}

}  // namespace subg
