// Walk-based set sampler, fast path: dedup + landing counts in a per-warp shared-memory hash table while the walks
// run, then a register-resident warp sort of the DISTINCT members only.
//
// Same outputs, bit for bit, as gset_sample_kernel (sampler.cuh) for the configurations it accepts -- Philox draws, no
// first-visit ranks wanted, no bucket cap, LP row <= 32 bits, num_walks <= 256 -- and the same reference semantics
// (subg_acc/subg_acc.c:763-844, 900-978).  Why a second kernel: in sampler.cuh every visit (M*m+1 of them) goes through
// the warp sort and the landing counts are recovered from runs of equal nodes afterwards.  On the collab / dblp shapes
// only 103 of 401 / 57 of 201 visits are distinct, and even on ppa (495 of 601) the sort + count phases were 48 % of the
// kernel's instructions.  Here
//   1. walk    every visit is inserted into the warp's open-addressing table (atomicCAS on the entry
//              node << OB | order, atomicMin keeps the first visit order of a revisited node) and bumps the packed
//              landing counts of its slot (one shared-memory atomicAdd): dedup and LP counts are done when the walks are;
//   2. compact occupied slots are squeezed to the front in place (ballot / popc), entry -> sort key node << IB | index,
//              first-visit order and counts follow their member into dense arrays;
//   3. sort    the S distinct keys are sorted in registers by a bitonic network over the warp: blocked layout, the
//              sort class (2 / 4 / 8 / 16 keys per lane, or the seed's maximum) is picked per seed from S, cross-lane
//              stages are SHFL.BFLY + VIMNMX, in-lane stages are compile-time networks; nothing goes through shared
//              memory and there is no dependent shared-memory load chain as in a merge-path merge;
//   4. emit    as sampler.cuh: row allocated with one atomic on the global cursor, written coalesced in ascending node
//              order, LP rows interned in the L2-resident table with their first stream position.
#pragma once
#include "sampler.cuh"

namespace subg {

struct HashPlan {
    int cap;          // table slots per warp (multiple of 32)
    int IB;           // bits of the dense member index in the sort key
    int ord_off;      // byte offsets inside the warp's shared memory: first-visit orders (uint16 [Kt])
    int cnt_off;      // packed landing counts (uint32 [cap])
    int smem_per_warp;
};

template <typename K>
__device__ __forceinline__ K shfl_xor_key(K v, int mask) {
    return __shfl_xor_sync(FULL, v, mask);
}

// in-lane bitonic merge (EPL a power of two): a bitonic sequence in k[] -> ascending
template <typename K, int EPL, int S>
__device__ __forceinline__ void lane_bitonic_merge(K (&k)[EPL]) {
    if constexpr (S >= 1) {
#pragma unroll
        for (int j = 0; j < EPL; j++)
            if ((j & S) == 0) cswap(k[j], k[j | S]);
        lane_bitonic_merge<K, EPL, S / 2>(k);
    }
}

// Sorts the 32 * EPL keys of a warp ascending, entirely in registers.  In: any arrangement.  Out: blocked layout
// (lane L holds global positions [L * EPL, (L + 1) * EPL)).  Merge of two sorted runs A, B of 2^r lanes each:
//   flip          A[i] <-> B[L-1-i]: lane ^ (2^(r+1) - 1), register EPL-1-j; the A side keeps the minima.  Afterwards
//                 every key of A' is <= every key of B' and both are bitonic sequences;
//   half-cleaners lane distance 2^(r-1) .. 1, same register: each lane ends with a bitonic sequence of EPL keys whose
//                 values lie between those of its neighbours;
//   in-lane       power-of-two EPL: bitonic merge network (EPL/2 log2 EPL comparators); otherwise (the class sized for
//                 the largest possible set) a full sorting network, which sorts a bitonic sequence like any other.
template <typename K, int EPL>
__device__ __forceinline__ void warp_bitonic_sort(K (&k)[EPL], int lane) {
    lane_sort<K, EPL>(k);
#pragma unroll 1
    for (int r = 0; r < 5; r++) {
        {
            const int mask = (2 << r) - 1;
            const bool up = (lane >> r) & 1;
            K o[EPL];
#pragma unroll
            for (int j = 0; j < EPL; j++) o[j] = shfl_xor_key(k[EPL - 1 - j], mask);
#pragma unroll
            for (int j = 0; j < EPL; j++) k[j] = up ? (k[j] > o[j] ? k[j] : o[j]) : (k[j] < o[j] ? k[j] : o[j]);
        }
#pragma unroll 1
        for (int d = r - 1; d >= 0; d--) {
            const bool up = (lane >> d) & 1;
#pragma unroll
            for (int j = 0; j < EPL; j++) {
                const K o = shfl_xor_key(k[j], 1 << d);
                k[j] = up ? (k[j] > o ? k[j] : o) : (k[j] < o ? k[j] : o);
            }
        }
        if constexpr ((EPL & (EPL - 1)) == 0) lane_bitonic_merge<K, EPL, EPL / 2>(k);
        else lane_sort<K, EPL>(k);
    }
}

// striped load (conflict-free), sort, blocked store with one word of padding per 32 (conflict-free for every EPL used)
__device__ __forceinline__ int pad_idx(int e) { return e + (e >> 5); }

template <typename K, int EPL>
__device__ __noinline__ void sort_class(K *list, int lane) {
    K k[EPL];
#pragma unroll
    for (int j = 0; j < EPL; j++) k[j] = list[j * 32 + lane];
    __syncwarp();
    warp_bitonic_sort<K, EPL>(k, lane);
#pragma unroll
    for (int j = 0; j < EPL; j++) list[pad_idx(lane * EPL + j)] = k[j];
}

template <typename K, int TOP>
constexpr int hash_min_blocks() {
    return 4;
}

// K: entry / sort-key type.  TOP: keys per lane of the largest sort class (32 * TOP >= M*m + 1).
template <typename K, int TOP>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, hash_min_blocks<K, TOP>()) gset_hash_kernel(const SamplerArgs a, const HashPlan hp) {
    constexpr K EMPTY = ~(K)0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char *wsm = smem_raw + (size_t)wib * hp.smem_per_warp;
    K *tab = (K *)wsm;                                   // [cap] entries, later the compacted / sorted key list
    uint32_t *cnt = (uint32_t *)(wsm + hp.cnt_off);      // [cap] packed landing counts, later dense per member
    uint16_t *ordv = (uint16_t *)(wsm + hp.ord_off);     // [Kt] first-visit order per member (dense)
    // Fisher-Yates scratch overlays the table before it is initialised
    int32_t *fy_pick = (int32_t *)wsm;
    int32_t *fy_key = fy_pick + a.M;
    int32_t *fy_dense = fy_key + a.fy_cap;
    int32_t *fy_val = fy_dense + a.M;
    const int M = a.M, m = a.m, OB = a.OB, LS = a.LS, cap = hp.cap, IB = hp.IB;
    const uint32_t ord_mask = (1u << OB) - 1u;
    const int lp_top = a.SHIFT * (m - 1);
    Policies pol;
    pol.keep = l2_policy_evict_last();
    int mx = 0;

    unsigned long long ticket = 0;
    if (lane == 0) ticket = atomicAdd(&a.ctr[0], 1ull);
    for (;;) {
        const int64_t i = (int64_t)__shfl_sync(FULL, ticket, 0);
        if (i >= a.n_chunk) break;
        if (lane == 0) ticket = atomicAdd(&a.ctr[0], 1ull);
        const int64_t gi = a.seed_base + i;
        const int32_t u = __ldg(a.seeds + i);
        if ((uint64_t)(int64_t)u >= (uint64_t)a.N) {
            if (lane == 0) {
                atomicOr(a.status, kStatusBadSeed);
                a.nsize[i] = 0;
                a.rowbeg[i] = 0;
            }
            continue;
        }
        int64_t rp0;
        uint32_t dfull;
        decode_row(a, (uint32_t)u, load_row_raw(a, pol, (uint32_t)u), rp0, dfull);
        const int d = dfull > (uint32_t)kFirstHopCap ? kFirstHopCap : (int)dfull;
        const uint32_t gi_lo = (uint32_t)gi, gi_hi = (uint32_t)((uint64_t)gi >> 32);

        // ---- first hop without replacement (subg_acc.c:763-776, 790-800): offsets of this lane's walks in registers
        int off0[8];
        if (d > M) {
            for (int c = lane; 4 * c < M; c += 32) {
                const uint4 r4 = philox4x32_10(make_uint4(gi_lo, gi_hi, (uint32_t)c, 0x46597331u), make_uint2(a.rng_lo, a.rng_hi));
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
#pragma unroll
            for (int t = 0; t < 8; t++) off0[t] = (lane + 32 * t < M) ? fy_dense[lane + 32 * t] : 0;
            __syncwarp();
        } else {
#pragma unroll
            for (int t = 0; t < 8; t++) off0[t] = d > 0 ? (lane + 32 * t) % d : 0;
        }

        // ---- empty table
        {
            uint4 *t4 = (uint4 *)tab;
            const uint4 e4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            for (int q = lane; q < cap * (int)sizeof(K) / 16; q += 32) t4[q] = e4;
            uint4 *c4 = (uint4 *)cnt;
            for (int q = lane; q < cap / 4; q += 32) c4[q] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        if (lane == 0) tab[__umulhi((uint32_t)u * 0x9E3779B1u, (uint32_t)cap)] = (K)(uint32_t)u << OB;  // the root: order 0
        __syncwarp();

        // insert one visit: entry = node << OB | order; the smallest order of a node survives; its step count is bumped
        auto visit = [&](uint32_t node, uint32_t order, int step) {
            const K e = ((K)node << OB) | (K)order;
            uint32_t h = __umulhi(node * 0x9E3779B1u, (uint32_t)cap);
            for (;;) {
                const K old = atomicCAS(&tab[h], EMPTY, e);
                if (old == EMPTY) break;
                if ((uint32_t)(old >> OB) == node) {
                    if (e < old) atomicMin(&tab[h], e);
                    break;
                }
                h = h + 1 == (uint32_t)cap ? 0u : h + 1;
            }
            atomicAdd(&cnt[h], 1u << (lp_top - a.SHIFT * step));
        };

        // ---- walks: lane l owns walks l, l + 32, ... (at most 8: num_walks <= 256), all advanced together.  The visits of
        // hop s - 1 are inserted while the row-info loads of hop s are in flight.
        {
            uint32_t cur[kGW];
#pragma unroll
            for (int tt = 0; tt < kGW; tt++) {
                const int w = lane + 32 * tt;
                cur[tt] = (uint32_t)u;
                if (w < M && d > 0) cur[tt] = load_col(a, pol, rp0 + off0[tt]);
            }
            for (int s = 1; s < m; s++) {
                uint2 raw[kGW];
#pragma unroll
                for (int tt = 0; tt < kGW; tt++) raw[tt] = load_row_raw(a, pol, cur[tt]);
#pragma unroll
                for (int tt = 0; tt < kGW; tt++) {
                    const int w = lane + 32 * tt;
                    if (w < M) {
                        visit(cur[tt], (((uint32_t)w + 1u) << LS) | (uint32_t)(s - 1), s - 1);
                        if (a.dump_walks) a.dump_walks[(i * M + w) * m + s - 1] = (int32_t)cur[tt];
                    }
                }
                uint32_t draw[kGW];
#pragma unroll
                for (int c = 0; c < kGW / 4; c++) {
                    const uint4 r4 = philox4x32_10(make_uint4(gi_lo, gi_hi, (uint32_t)(lane + 32 * c) | ((uint32_t)s << 16), 0x57414c4bu),
                                                   make_uint2(a.rng_lo, a.rng_hi));
                    draw[4 * c] = r4.x; draw[4 * c + 1] = r4.y; draw[4 * c + 2] = r4.z; draw[4 * c + 3] = r4.w;
                }
#pragma unroll
                for (int tt = 0; tt < kGW; tt++) {
                    const int w = lane + 32 * tt;
                    if (w < M) {
                        int64_t rp;
                        uint32_t dn;
                        decode_row(a, cur[tt], raw[tt], rp, dn);
                        if (dn > 0) cur[tt] = load_col(a, pol, rp + __umulhi(draw[tt], dn));
                    }
                }
            }
#pragma unroll
            for (int tt = 0; tt < kGW; tt++) {
                const int w = lane + 32 * tt;
                if (w < M) {
                    visit(cur[tt], (((uint32_t)w + 1u) << LS) | (uint32_t)(m - 1), m - 1);
                    if (a.dump_walks) a.dump_walks[(i * M + w) * m + m - 1] = (int32_t)cur[tt];
                }
            }
        }
        __syncwarp();
        if (a.stop_after == 1) continue;

        // ---- compact in place: slot q = j * 32 + lane -> position p <= q (earlier iterations only wrote below their slots)
        int S = 0;
        for (int q0 = 0; q0 < cap; q0 += 32) {
            const K e = tab[q0 + lane];
            const uint32_t c = cnt[q0 + lane];
            __syncwarp();  // every slot of this group is in registers before anything is written below it
            const bool occ = e != EMPTY;
            const uint32_t bal = __ballot_sync(FULL, occ);
            const int p = S + __popc(bal & ((1u << lane) - 1u));
            if (occ) {
                tab[p] = ((e >> OB) << IB) | (K)p;
                cnt[p] = c;
                ordv[p] = (uint16_t)((uint32_t)e & ord_mask);
            }
            S += __popc(bal);
            __syncwarp();
        }
        // ---- sort class from the set size; keys beyond S are padding
        const int epl = S <= 64 ? 2 : S <= 128 ? 4 : S <= 256 ? 8 : (S <= 512 && TOP > 16) ? 16 : TOP;
        const int cls = (epl < TOP) ? epl : TOP;
        for (int q = S + lane; q < 32 * cls; q += 32) tab[q] = EMPTY;
        __syncwarp();
        if (TOP > 2 && cls == 2) sort_class<K, 2>(tab, lane);
        else if (TOP > 4 && cls == 4) sort_class<K, 4>(tab, lane);
        else if (TOP > 8 && cls == 8) sort_class<K, 8>(tab, lane);
        else if (TOP > 16 && cls == 16) sort_class<K, 16>(tab, lane);
        else sort_class<K, TOP>(tab, lane);
        __syncwarp();
        if (a.stop_after == 2) continue;

        // ---- row allocation
        const int kept4 = (S + 3) & ~3;
        unsigned long long base_u = 0;
        if (lane == 0) {
            base_u = atomicAdd(&a.ctr[kCtrCursor], (unsigned long long)kept4);
            atomicAdd(&a.ctr[kCtrTotal], (unsigned long long)S);
            a.rowbeg[i] = (long long)base_u;
            a.nsize[i] = S;
        }
        mx = S > mx ? S : mx;
        const long long base = (long long)__shfl_sync(FULL, base_u, 0);

        // ---- emit: ascending node id, provisional LP id; two members per lane and iteration (two lookups in flight)
        const K imask = ((K)1 << IB) - 1;
        for (int t0 = 0; t0 < S; t0 += 64) {
            uint32_t node[2], h[2], ord[2];
            unsigned long long lp[2], cur[2], seen[2];
            bool act[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int t = t0 + 32 * q + lane;
                act[q] = t < S;
                node[q] = 0; lp[q] = 0ull; ord[q] = 0; h[q] = 0; cur[q] = kEmptyKey; seen[q] = 0ull;
                if (act[q]) {
                    const K key = tab[pad_idx(t)];
                    const int idx = (int)(key & imask);
                    node[q] = (uint32_t)(key >> IB);
                    ord[q] = ordv[idx];
                    lp[q] = cnt[idx];
                    if (ord[q] == 0) lp[q] |= 1ull << (m * a.SHIFT);  // the root row (LEAD, subg_acc.c:944-949)
                    h[q] = lp_hash(lp[q]) & a.tab_mask;
                    cur[q] = a.tab_key[h[q]];
                    seen[q] = a.tab_pos[h[q]];
                }
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (act[q]) {
                    const uint32_t prov = intern_key(a, lp[q], ((unsigned long long)gi << 16) | ord[q], h[q], cur[q], seen[q]);
                    const int t = t0 + 32 * q + lane;
                    a.out_node[base + t] = (int32_t)node[q];
                    a.out_prov[base + t] = (int32_t)prov;
                }
            }
        }
        if (lane < kept4 - S) {  // keep the row padding defined (ids are remapped in place later)
            a.out_node[base + S + lane] = 0x7fffffff;
            a.out_prov[base + S + lane] = 0;
        }
        __syncwarp();  // the table is re-initialised by the next seed
    }
    if (lane == 0 && mx > 0) atomicMax(a.max_set, mx);
}

}  // namespace subg
