/*
 * oracle/subg_oracle.c -- CPU restatement of the SubGAcc hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is compared
 * against; it is never part of the product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the library built
 * from it (oracle/_ref/libsubg_oracle.so).
 *
 * Parity status: PINNED.  Every function below is validated in
 * tests/test_oracle_vs_reference.py against the UNMODIFIED reference compiled
 * from /root/reference (oracle/Makefile target `ref`) when that build is present,
 * and against the committed fixtures in tests/golden/ (generated from the
 * reference by tests/golden/make_golden.py) everywhere else.
 *
 * The reference algorithm each function follows is cited as
 * <file>:<lines> relative to /root/reference.  The code is a restatement with
 * its own data structures (flat open-addressing tables instead of uthash),
 * written so that the two observable orders of the reference are preserved:
 *   - within a set, slots are in first-visit order (walk-major, step-minor,
 *     root first)                                     subg_acc/subg_acc.c:784-844
 *   - LP-row ids are in first-occurrence order of the seed-major stream
 *                                                     subg_acc/subg_acc.c:957-978
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_NEB_CAP 1000000 /* first-hop neighbourhood cap, subg_acc.c:13,750 */

/* ------------------------------------------------------------------------- */
/* glibc rand_r: three rounds of a 32-bit LCG giving 11+10+10 bits.          */
/* (the reference calls libc's rand_r at subg_acc.c:771,807)                 */
/* ------------------------------------------------------------------------- */
static inline uint32_t lcg_step(uint32_t s) { return s * 1103515245u + 12345u; }

int orc_rand_r(uint32_t *state)
{
    uint32_t s = lcg_step(*state);
    uint32_t out = (s >> 16) & 2047u;
    s = lcg_step(s);
    out = (out << 10) ^ ((s >> 16) & 1023u);
    s = lcg_step(s);
    out = (out << 10) ^ ((s >> 16) & 1023u);
    *state = s;
    return (int)out;
}

/* ------------------------------------------------------------------------- */
/* Walk traces, single rand_r stream (== reference with nthread=1).          */
/* Follows subg_acc.c:745-809: first hop without replacement (round-robin if  */
/* deg<=M, partial Fisher-Yates otherwise), later hops uniform with          */
/* replacement, a walk that reaches a node without out-neighbours stays put. */
/* walks[i][w][s] = node after step s.  An isolated seed yields walks that   */
/* never leave the seed (the reference special-cases it at :753-761 to the   */
/* same LP row [M,...,M]).  calls[i] (optional) = rand_r calls spent on seed */
/* i.  Returns the final RNG state.                                          */
/* ------------------------------------------------------------------------- */
uint32_t orc_walks_rand_r(const int64_t *rowptr, const int32_t *col,
                          const int32_t *query, int64_t n, int M, int m,
                          uint32_t seed, int32_t *walks, int64_t *calls)
{
    uint32_t st = seed;
    int32_t *perm = NULL;
    int64_t perm_cap = 0;
    for (int64_t i = 0; i < n; i++) {
        const int32_t u = query[i];
        int64_t d = rowptr[u + 1] - rowptr[u];
        if (d > ORC_NEB_CAP) d = ORC_NEB_CAP;
        int32_t *out = walks + i * (int64_t)M * m;
        int64_t ncall = 0;
        if (d == 0) {
            for (int64_t t = 0; t < (int64_t)M * m; t++) out[t] = u;
            if (calls) calls[i] = 0;
            continue;
        }
        if (d > M) {
            if (d > perm_cap) {
                perm_cap = d;
                perm = (int32_t *)realloc(perm, (size_t)perm_cap * sizeof(int32_t));
            }
            for (int64_t j = 0; j < d; j++) perm[j] = (int32_t)j;
            for (int k = 0; k < M; k++) {
                int64_t pick = orc_rand_r(&st) % (d - k) + k;
                ncall++;
                int32_t tmp = perm[k];
                perm[k] = perm[pick];
                perm[pick] = tmp;
            }
        }
        for (int w = 0; w < M; w++) {
            int32_t cur = u;
            for (int s = 0; s < m; s++) {
                if (s == 0) {
                    int64_t off = (d <= M) ? (w % d) : perm[w];
                    cur = col[rowptr[cur] + off];
                } else {
                    int64_t dn = rowptr[cur + 1] - rowptr[cur];
                    if (dn > 0) {
                        cur = col[rowptr[cur] + (orc_rand_r(&st) % dn)];
                        ncall++;
                    }
                }
                out[(int64_t)w * m + s] = cur;
            }
        }
        if (calls) calls[i] = ncall;
    }
    free(perm);
    return st;
}

/* ------------------------------------------------------------------------- */
/* Per-seed node -> slot map (stands in for the uthash dict_int).            */
/* ------------------------------------------------------------------------- */
typedef struct {
    int32_t *key, *val;
    uint32_t *stamp;
    uint32_t mask, gen;
} slotmap;

static int slotmap_init(slotmap *h, int64_t want)
{
    uint32_t cap = 16;
    while ((int64_t)cap < 2 * want + 2) cap <<= 1;
    h->key = (int32_t *)malloc(cap * sizeof(int32_t));
    h->val = (int32_t *)malloc(cap * sizeof(int32_t));
    h->stamp = (uint32_t *)calloc(cap, sizeof(uint32_t));
    h->mask = cap - 1;
    h->gen = 0;
    return (h->key && h->val && h->stamp) ? 0 : -1;
}
static void slotmap_free(slotmap *h) { free(h->key); free(h->val); free(h->stamp); }
static inline uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
/* returns slot index of key or -1; *pos receives the probe position */
static inline int32_t slotmap_find(const slotmap *h, int32_t k, uint32_t *pos)
{
    uint32_t p = mix32((uint32_t)k) & h->mask;
    while (h->stamp[p] == h->gen) {
        if (h->key[p] == k) { *pos = p; return h->val[p]; }
        p = (p + 1) & h->mask;
    }
    *pos = p;
    return -1;
}

/* ------------------------------------------------------------------------- */
/* 64-bit key -> id map in first-occurrence order (stands in for dict_long). */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t *key;
    int32_t *val;
    uint64_t cap, used;
} idmap;
static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static int idmap_init(idmap *h, uint64_t cap)
{
    h->cap = cap; h->used = 0;
    h->key = (uint64_t *)malloc(cap * sizeof(uint64_t));
    h->val = (int32_t *)malloc(cap * sizeof(int32_t));
    if (!h->key || !h->val) return -1;
    for (uint64_t i = 0; i < cap; i++) h->val[i] = -1;
    return 0;
}
static void idmap_free(idmap *h) { free(h->key); free(h->val); }
static int idmap_get_or_add(idmap *h, uint64_t k, int32_t next_id, int *added);
static int idmap_grow(idmap *h)
{
    idmap big;
    if (idmap_init(&big, h->cap * 2)) return -1;
    for (uint64_t i = 0; i < h->cap; i++)
        if (h->val[i] >= 0) { int a; idmap_get_or_add(&big, h->key[i], h->val[i], &a); }
    idmap_free(h);
    *h = big;
    return 0;
}
static int idmap_get_or_add(idmap *h, uint64_t k, int32_t next_id, int *added)
{
    if (2 * (h->used + 1) > h->cap && idmap_grow(h)) return -2;
    uint64_t p = mix64(k) & (h->cap - 1);
    while (h->val[p] >= 0) {
        if (h->key[p] == k) { *added = 0; return h->val[p]; }
        p = (p + 1) & (h->cap - 1);
    }
    h->key[p] = k; h->val[p] = next_id; h->used++;
    *added = 1;
    return next_id;
}

/* ------------------------------------------------------------------------- */
/* Trace-driven set sampler + LP encoder + global LP de-dup.                 */
/*   per-seed dedup and landing counts      subg_acc.c:778-844               */
/*   dense compaction                        subg_acc.c:848-872               */
/*   64-bit LP key (cols 1..m, SHIFT bits each, LEAD bit on roots) :900-955   */
/*   first-occurrence unique ids + table     subg_acc.c:957-1000              */
/* Inputs : query[n], walks[n][M][m] (node after each step).                 */
/* Outputs: nsize[n]; nidx[T] (first-visit order); sfptr[T]; raw[T][m+1]     */
/*          (nullable); enc[c][m+1]; *T_out, *c_out; *dropped = 1 if a       */
/*          bucket overflowed (the reference only prints a warning, :835).   */
/* Capacities: nidx/sfptr >= n*stride, raw >= n*stride*(m+1), enc likewise   */
/*          unless enc_cap_rows is given.                                    */
/* Returns 0, -1 out of memory, -2 key too wide (AssertionError at :913),    */
/*         -3 enc capacity exceeded.                                         */
/* ------------------------------------------------------------------------- */
int orc_gset_from_walks(const int32_t *query, int64_t n, int M, int m, int bucket,
                        const int32_t *walks, int32_t *nsize, int32_t *nidx,
                        int32_t *sfptr, int16_t *raw, int16_t *enc,
                        int64_t enc_cap_rows, int64_t *T_out, int32_t *c_out,
                        int32_t *dropped)
{
    const int ncol = m + 1;
    const int64_t stride = bucket < 0 ? (int64_t)M * m + 1 : bucket;
    int shift = 0;
    while ((M >> shift) != 0) shift++; /* 32 - clz(M) */
    if ((int64_t)m * shift + 1 > 64) return -2;
    const uint64_t lead = (m * shift == 64) ? 0 : (1ULL << (m * shift));

    slotmap sm;
    if (slotmap_init(&sm, stride)) return -1;
    int16_t *rows = (int16_t *)malloc((size_t)stride * ncol * sizeof(int16_t));
    idmap uq;
    if (!rows || idmap_init(&uq, 1024)) return -1;

    int64_t T = 0;
    int32_t c = 0;
    *dropped = 0;
    for (int64_t i = 0; i < n; i++) {
        const int32_t u = query[i];
        const int32_t *wk = walks + i * (int64_t)M * m;
        sm.gen++;
        memset(rows, 0, (size_t)stride * ncol * sizeof(int16_t));
        /* root occupies slot 0 and carries M in column 0 */
        uint32_t pos;
        slotmap_find(&sm, u, &pos);
        sm.key[pos] = u; sm.val[pos] = 0; sm.stamp[pos] = sm.gen;
        rows[0] = (int16_t)M;
        int32_t *ids = nidx + T;
        ids[0] = u;
        int32_t cnt = 1;
        for (int w = 0; w < M; w++) {
            for (int s = 0; s < m; s++) {
                const int32_t v = wk[(int64_t)w * m + s];
                int32_t slot = slotmap_find(&sm, v, &pos);
                if (slot < 0) {
                    if (cnt >= stride) { *dropped = 1; continue; } /* walk goes on, visit not counted */
                    slot = cnt++;
                    sm.key[pos] = v; sm.val[pos] = slot; sm.stamp[pos] = sm.gen;
                    ids[slot] = v;
                }
                rows[(int64_t)slot * ncol + s + 1]++;
            }
        }
        nsize[i] = cnt;
        for (int32_t j = 0; j < cnt; j++) {
            const int16_t *r = rows + (int64_t)j * ncol;
            uint64_t key = 0;
            for (int q = 1; q < ncol; q++) key = (key << shift) | (uint64_t)(uint16_t)r[q];
            if (j == 0) key |= lead;
            int added;
            int32_t id = idmap_get_or_add(&uq, key, c, &added);
            if (id == -2) return -1;
            if (added) {
                if (enc_cap_rows >= 0 && c >= enc_cap_rows) return -3;
                memcpy(enc + (int64_t)c * ncol, r, ncol * sizeof(int16_t));
                c++;
            }
            sfptr[T + j] = id;
            if (raw) memcpy(raw + (T + j) * ncol, r, ncol * sizeof(int16_t));
        }
        T += cnt;
    }
    *T_out = T;
    *c_out = c;
    slotmap_free(&sm);
    idmap_free(&uq);
    free(rows);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* SpG build: COO (row=seed, col=node, val=sfptr+1) -> CSR with sorted cols.  */
/* sampler/random_walks.py:79-80 (scipy csr_matrix ctor; has_sorted_indices). */
/* Seeds must be distinct (scipy would sum duplicates; callers pass arange). */
/* ------------------------------------------------------------------------- */
typedef struct { int32_t node, val; } nv_pair;
static int cmp_nv(const void *a, const void *b)
{
    int32_t x = ((const nv_pair *)a)->node, y = ((const nv_pair *)b)->node;
    return (x > y) - (x < y);
}
int orc_spg_build(int64_t N, const int32_t *query, int64_t n, const int32_t *nsize,
                  const int32_t *nidx, const int32_t *sfptr, int64_t *indptr,
                  int32_t *indices, int32_t *data)
{
    memset(indptr, 0, (size_t)(N + 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) indptr[query[i] + 1] += nsize[i];
    for (int64_t r = 0; r < N; r++) indptr[r + 1] += indptr[r];
    int32_t maxs = 0;
    for (int64_t i = 0; i < n; i++) if (nsize[i] > maxs) maxs = nsize[i];
    nv_pair *buf = (nv_pair *)malloc((size_t)(maxs > 0 ? maxs : 1) * sizeof(nv_pair));
    if (!buf) return -1;
    int64_t src = 0;
    for (int64_t i = 0; i < n; i++) {
        const int32_t s = nsize[i];
        for (int32_t j = 0; j < s; j++) { buf[j].node = nidx[src + j]; buf[j].val = sfptr[src + j] + 1; }
        qsort(buf, (size_t)s, sizeof(nv_pair), cmp_nv);
        int64_t dst = indptr[query[i]];
        for (int32_t j = 0; j < s; j++) { indices[dst + j] = buf[j].node; data[dst + j] = buf[j].val; }
        src += s;
    }
    free(buf);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* SpJoin.  For a pair (a,b): every w in S_a (ascending) emits               */
/* [val_a(w), val_b(w) or 0].  train.py:75-85 (bgather) computes this as     */
/* xb.multiply(xa>0) + (xa>0) - 1 on scipy CSR row slices.                   */
/* Integer (LP pointer) flavour.                                             */
/* ------------------------------------------------------------------------- */
static int64_t join_rows_i32(const int64_t *indptr, const int32_t *indices,
                             const int32_t *data, int32_t a, int32_t b, int32_t *out)
{
    int64_t pa = indptr[a], ea = indptr[a + 1], pb = indptr[b], eb = indptr[b + 1];
    int64_t k = 0;
    for (; pa < ea; pa++, k++) {
        const int32_t w = indices[pa];
        while (pb < eb && indices[pb] < w) pb++;
        out[2 * k] = data[pa];
        out[2 * k + 1] = (pb < eb && indices[pb] == w) ? data[pb] : 0;
    }
    return k;
}
/* Floating flavour (PPR/SPD values, fp64 store).  The reference forms the    */
/* matched value as (x*1 + 1) - 1 in fp64 (train.py:33,38-41), which rounds;  */
/* the same two roundings are restated here.                                  */
static int64_t join_rows_f64(const int64_t *indptr, const int32_t *indices,
                             const double *data, int32_t a, int32_t b, double *out)
{
    int64_t pa = indptr[a], ea = indptr[a + 1], pb = indptr[b], eb = indptr[b + 1];
    int64_t k = 0;
    for (; pa < ea; pa++, k++) {
        const int32_t w = indices[pa];
        while (pb < eb && indices[pb] < w) pb++;
        out[2 * k] = data[pa];
        volatile double t = (pb < eb && indices[pb] == w) ? data[pb] + 1.0 : 1.0;
        out[2 * k + 1] = t - 1.0;
    }
    return k;
}

/* Pair join over a batch: rows of all left sets, then rows of all right sets */
/* (train.py:34-36 vstack([xl, xr])); sizes_l/sizes_r as train.py:85.         */
/* out must hold 2*(sum S_u + sum S_v) values.  Returns total rows N.         */
int64_t orc_spjoin_pair_i32(const int64_t *indptr, const int32_t *indices,
                            const int32_t *data, const int64_t *edge, int64_t B,
                            int32_t *out, int64_t *sizes_l, int64_t *sizes_r)
{
    int64_t N = 0;
    for (int64_t q = 0; q < B; q++) {
        int64_t k = join_rows_i32(indptr, indices, data, (int32_t)edge[q], (int32_t)edge[B + q], out + 2 * N);
        sizes_l[q] = k; N += k;
    }
    for (int64_t q = 0; q < B; q++) {
        int64_t k = join_rows_i32(indptr, indices, data, (int32_t)edge[B + q], (int32_t)edge[q], out + 2 * N);
        sizes_r[q] = k; N += k;
    }
    return N;
}
int64_t orc_spjoin_pair_f64(const int64_t *indptr, const int32_t *indices,
                            const double *data, const int64_t *edge, int64_t B,
                            double *out, int64_t *sizes_l, int64_t *sizes_r)
{
    int64_t N = 0;
    for (int64_t q = 0; q < B; q++) {
        int64_t k = join_rows_f64(indptr, indices, data, (int32_t)edge[q], (int32_t)edge[B + q], out + 2 * N);
        sizes_l[q] = k; N += k;
    }
    for (int64_t q = 0; q < B; q++) {
        int64_t k = join_rows_f64(indptr, indices, data, (int32_t)edge[B + q], (int32_t)edge[q], out + 2 * N);
        sizes_r[q] = k; N += k;
    }
    return N;
}

/* Triplet join (u,v,w): blocks [u|w], [w|u], [v|w], [w|v]; u and v are not    */
/* joined with each other.  train.py:48-72 (hgather).  sizes has 4*B entries  */
/* laid out cat[usize, wsize, vsize, wsize] (train.py:57).                     */
int64_t orc_spjoin_triplet_i32(const int64_t *indptr, const int32_t *indices,
                               const int32_t *data, const int64_t *hedge, int64_t B,
                               int32_t *out, int64_t *sizes)
{
    int64_t N = 0;
    static const int lhs[4] = {0, 2, 1, 2}, rhs[4] = {2, 0, 2, 1};
    for (int blk = 0; blk < 4; blk++)
        for (int64_t q = 0; q < B; q++) {
            int32_t a = (int32_t)hedge[lhs[blk] * B + q], b = (int32_t)hedge[rhs[blk] * B + q];
            int64_t k = join_rows_i32(indptr, indices, data, a, b, out + 2 * N);
            sizes[blk * B + q] = k; N += k;
        }
    return N;
}

/* ------------------------------------------------------------------------- */
/* ACL forward push for one seed.  sampler/pprgo.py:9-38 (_calc_ppr_node):    */
/* p, r are float32 maps, q is a LIFO list with an O(len) membership test;    */
/* the pushed amount is formed in double and rounded to float                 */
/* ((1 - alpha) is int - float32 -> float64 under numba's typing; `_val` is   */
/* declared float32), the threshold alpha*eps is a float32 product compared   */
/* in double against r[v].  Output: keys/vals of p in insertion order.        */
/* ------------------------------------------------------------------------- */
typedef struct {
    int32_t *key; int32_t *idx; uint32_t cap, used;
} nodemap;
static void nodemap_init(nodemap *h, uint32_t cap)
{
    h->cap = cap; h->used = 0;
    h->key = (int32_t *)malloc(cap * sizeof(int32_t));
    h->idx = (int32_t *)malloc(cap * sizeof(int32_t));
    for (uint32_t i = 0; i < cap; i++) h->idx[i] = -1;
}
static void nodemap_free(nodemap *h) { free(h->key); free(h->idx); }
static int32_t nodemap_get(nodemap *h, int32_t k, int create, int32_t next)
{
    if (create && 2 * (h->used + 1) > h->cap) {
        nodemap big; nodemap_init(&big, h->cap * 2);
        for (uint32_t i = 0; i < h->cap; i++)
            if (h->idx[i] >= 0) nodemap_get(&big, h->key[i], 1, h->idx[i]);
        nodemap_free(h); *h = big;
    }
    uint32_t p = mix32((uint32_t)k) & (h->cap - 1);
    while (h->idx[p] >= 0) {
        if (h->key[p] == k) return h->idx[p];
        p = (p + 1) & (h->cap - 1);
    }
    if (!create) return -1;
    h->key[p] = k; h->idx[p] = next; h->used++;
    return next;
}

/* returns support size; fills keys/vals (capacity cap). -1 if cap too small. */
int64_t orc_ppr_push(const int64_t *rowptr, const int32_t *col, const int64_t *deg,
                     int32_t seed_node, float alpha, float eps,
                     int32_t *keys, float *vals, int64_t cap, int64_t *n_push)
{
    const float alpha_eps = alpha * eps;
    const double one_minus_alpha = 1.0 - (double)alpha;
    /* one record per touched node; p-insertion order tracked separately */
    nodemap map; nodemap_init(&map, 64);
    int64_t rcap = 64, nrec = 0;
    float *r = (float *)malloc(rcap * sizeof(float));
    float *p = (float *)malloc(rcap * sizeof(float));
    int32_t *node = (int32_t *)malloc(rcap * sizeof(int32_t));
    int64_t *pord = (int64_t *)malloc(rcap * sizeof(int64_t)); /* p insertion rank or -1 */
    uint8_t *inq = (uint8_t *)malloc(rcap);
    int64_t qcap = 64, qlen = 0, np = 0, pushes = 0;
    int32_t *q = (int32_t *)malloc(qcap * sizeof(int32_t));
#define NEWREC(v) do { if (nrec == rcap) { rcap *= 2; \
        r = (float *)realloc(r, rcap * sizeof(float)); p = (float *)realloc(p, rcap * sizeof(float)); \
        node = (int32_t *)realloc(node, rcap * sizeof(int32_t)); pord = (int64_t *)realloc(pord, rcap * sizeof(int64_t)); \
        inq = (uint8_t *)realloc(inq, rcap); } \
        node[nrec] = (v); r[nrec] = 0.f; p[nrec] = 0.f; pord[nrec] = -1; inq[nrec] = 0; nrec++; } while (0)
    int32_t s = nodemap_get(&map, seed_node, 1, 0);
    NEWREC(seed_node);
    pord[s] = np++;        /* p = {inode: 0} */
    r[s] = alpha;          /* r[inode] = alpha */
    q[qlen++] = s; inq[s] = 1;
    while (qlen > 0) {
        const int32_t ui = q[--qlen];
        inq[ui] = 0;
        const int32_t u = node[ui];
        const float res = r[ui];
        if (pord[ui] < 0) { pord[ui] = np++; p[ui] = res; } else p[ui] += res;
        r[ui] = 0.f;
        pushes++;
        const float val = (float)(one_minus_alpha * (double)res / (double)deg[u]);
        for (int64_t e = rowptr[u]; e < rowptr[u + 1]; e++) {
            const int32_t v = col[e];
            int32_t vi = nodemap_get(&map, v, 0, 0);
            if (vi < 0) { vi = nodemap_get(&map, v, 1, (int32_t)nrec); NEWREC(v); r[vi] = val; }
            else r[vi] += val;
            if ((double)r[vi] >= (double)alpha_eps * (double)deg[v]) {
                if (!inq[vi]) {
                    if (qlen == qcap) { qcap *= 2; q = (int32_t *)realloc(q, qcap * sizeof(int32_t)); }
                    q[qlen++] = vi; inq[vi] = 1;
                }
            }
        }
    }
#undef NEWREC
    int64_t ret = np;
    if (np > cap) ret = -1;
    else
        for (int64_t i = 0; i < nrec; i++)
            if (pord[i] >= 0) { keys[pord[i]] = node[i]; vals[pord[i]] = p[i]; }
    if (n_push) *n_push = pushes;
    nodemap_free(&map);
    free(r); free(p); free(node); free(pord); free(inq); free(q);
    return ret;
}

/* OpenMP driver over seeds (pprgo.py:52-56 prange); writes each seed's full  */
/* support into a ragged buffer: off[i]..off[i]+cnt[i], row capacity `cap`.   */
int orc_ppr_push_many(const int64_t *rowptr, const int32_t *col, const int64_t *deg,
                      const int32_t *seeds, int64_t n, float alpha, float eps,
                      int32_t *keys, float *vals, int64_t cap, int64_t *cnt,
                      int64_t *pushes, int nthread)
{
    int bad = 0;
    if (nthread <= 0) nthread = 1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthread)
    for (int64_t i = 0; i < n; i++) {
        int64_t np_ = 0;
        int64_t c = orc_ppr_push(rowptr, col, deg, seeds[i], alpha, eps,
                                 keys + i * cap, vals + i * cap, cap, &np_);
        cnt[i] = c;
        if (pushes) pushes[i] = np_;
        if (c < 0) bad = 1;
    }
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------- */
/* SUREL-v1 walk_sampler: walks with the root in column 0   */
/* and the relative-position encoder.  Single rand_r stream == reference with */
/* nthread=1.                                                                 */
/*   without <= 0: every hop uniform with replacement   subg_acc.c:144-181    */
/*   without  > 0: first hop without replacement (round-robin if deg <= M,    */
/*                 partial Fisher-Yates otherwise; no neighbourhood cap in    */
/*                 this function)                       subg_acc.c:183-248    */
/* walks[i][w][0..m], walks[i][w][0] = query[i].  Returns the RNG state.      */
/* ------------------------------------------------------------------------- */
uint32_t orc_walk_sampler_walks(const int64_t *rowptr, const int32_t *col,
                                const int32_t *query, int64_t n, int M, int m,
                                uint32_t seed, int without, int32_t *walks)
{
    uint32_t st = seed;
    int32_t *perm = NULL;
    int64_t perm_cap = 0;
    for (int64_t i = 0; i < n; i++) {
        const int32_t u = query[i];
        const int64_t d = rowptr[u + 1] - rowptr[u];
        if (without > 0 && d > M) {
            if (d > perm_cap) {
                perm_cap = d;
                perm = (int32_t *)realloc(perm, (size_t)perm_cap * sizeof(int32_t));
            }
            for (int64_t j = 0; j < d; j++) perm[j] = (int32_t)j;
            for (int k = 0; k < M; k++) {
                const int64_t pick = orc_rand_r(&st) % (d - k) + k;
                const int32_t tmp = perm[k];
                perm[k] = perm[pick];
                perm[pick] = tmp;
            }
        }
        for (int w = 0; w < M; w++) {
            int32_t *out = walks + (i * (int64_t)M + w) * (m + 1);
            int32_t cur = u;
            out[0] = cur;
            for (int s = 0; s < m; s++) {
                if (without > 0 && s == 0) {
                    if (d >= 1) cur = col[rowptr[cur] + (d <= M ? (w % d) : perm[w])];
                } else {
                    const int64_t dn = rowptr[cur + 1] - rowptr[cur];
                    if (dn > 0) cur = col[rowptr[cur] + (orc_rand_r(&st) % dn)];
                }
                out[s + 1] = cur;
            }
        }
    }
    free(perm);
    return st;
}

/* ------------------------------------------------------------------------- */
/* rpe_encoder, subg_acc.c:250-314: per seed, unique nodes numbered in        */
/* first-visit order of the STEP-major, walk-minor scan (root = 0), and       */
/* rpe[node][step] = number of walks at that node after `step` hops;          */
/* rpe[root][0] = M.  Output is ragged: seed i owns ids[off[i]..off[i+1]) and */
/* the matching rows of rpe[., m+1].  Returns the total or -1 if cap is short.*/
/* ------------------------------------------------------------------------- */
int64_t orc_rpe_encode(const int32_t *walks, int64_t n, int M, int m,
                       int64_t *off, int32_t *ids, int32_t *rpe, int64_t cap)
{
    slotmap h;
    if (slotmap_init(&h, (int64_t)M * m + 1)) return -2;
    int64_t total = 0;
    const int ncol = m + 1;
    for (int64_t i = 0; i < n; i++) {
        const int32_t *wk = walks + i * (int64_t)M * ncol;
        h.gen++;
        off[i] = total;
        int32_t count = 0;
        uint32_t pos;
        /* the root first (:256-260) */
        if (total + 1 > cap) { slotmap_free(&h); return -1; }
        slotmap_find(&h, wk[0], &pos);
        h.key[pos] = wk[0]; h.val[pos] = 0; h.stamp[pos] = h.gen;
        ids[total] = wk[0];
        memset(rpe + total * ncol, 0, sizeof(int32_t) * ncol);
        rpe[total * ncol] = M;
        count = 1;
        for (int s = 1; s <= m; s++)
            for (int w = 0; w < M; w++) {
                const int32_t v = wk[(int64_t)w * ncol + s];
                int32_t slot = slotmap_find(&h, v, &pos);
                if (slot < 0) {
                    if (total + count + 1 > cap) { slotmap_free(&h); return -1; }
                    slot = count++;
                    h.key[pos] = v; h.val[pos] = slot; h.stamp[pos] = h.gen;
                    ids[total + slot] = v;
                    memset(rpe + (total + slot) * ncol, 0, sizeof(int32_t) * ncol);
                }
                rpe[(total + slot) * ncol + s]++;
            }
        total += count;
    }
    off[n] = total;
    slotmap_free(&h);
    return total;
}

/* ------------------------------------------------------------------------- */
/* walk_join (SUREL v1), subg_acc.c:509-647.  Two-level lookup as the         */
/* reference builds it: root node -> row (:566-575), and per row node ->      */
/* running index starting at 1 over the concatenated key sets (:577-588).     */
/* find_idx (:87-100): index, 0 if the node is not in that root's set, -1 if  */
/* the root is unknown.  out[2][Q*2*stride], xq[Q][2] (:618-633).  Entries    */
/* that depend on the walks of an unknown root are set to -1 (the reference   */
/* reads out of bounds there).  Duplicate roots / duplicate ids within a key  */
/* set are outside the contract (uthash keeps both and finds one of them).    */
/* ------------------------------------------------------------------------- */
int orc_walk_join(const int32_t *walks, int64_t n, int64_t stride, const int64_t *key_off,
                  const int32_t *key_ids, const int32_t *query, int64_t Q, int32_t *out, int32_t *xq)
{
    idmap roots, members;
    uint64_t cap = 64;
    while (cap < 4 * (uint64_t)(key_off[n] + n + 1)) cap <<= 1;
    if (idmap_init(&roots, cap) || idmap_init(&members, cap)) return -1;
    int added;
    for (int64_t i = 0; i < n; i++) {
        idmap_get_or_add(&roots, (uint64_t)(uint32_t)walks[i * stride], (int32_t)i, &added);
        for (int64_t j = key_off[i]; j < key_off[i + 1]; j++)
            idmap_get_or_add(&members, ((uint64_t)i << 32) | (uint32_t)key_ids[j], (int32_t)(j + 1), &added);
    }
    const int64_t half = Q * 2 * stride;
#define ORC_FIND_ROW(node, dst)                                                            \
    do {                                                                                   \
        uint64_t p_ = mix64((uint64_t)(uint32_t)(node)) & (roots.cap - 1);                 \
        (dst) = -1;                                                                        \
        while (roots.val[p_] >= 0) {                                                       \
            if (roots.key[p_] == (uint64_t)(uint32_t)(node)) { (dst) = roots.val[p_]; break; } \
            p_ = (p_ + 1) & (roots.cap - 1);                                               \
        }                                                                                  \
    } while (0)
#define ORC_FIND_IDX(row, node, dst)                                                       \
    do {                                                                                   \
        if ((row) < 0) { (dst) = -1; break; }                                              \
        const uint64_t k_ = ((uint64_t)(row) << 32) | (uint32_t)(node);                    \
        uint64_t p_ = mix64(k_) & (members.cap - 1);                                       \
        (dst) = 0;                                                                         \
        while (members.val[p_] >= 0) {                                                     \
            if (members.key[p_] == k_) { (dst) = members.val[p_]; break; }                 \
            p_ = (p_ + 1) & (members.cap - 1);                                             \
        }                                                                                  \
    } while (0)
    for (int64_t x = 0; x < Q; x++) {
        int32_t r1, r2;
        ORC_FIND_ROW(query[2 * x], r1);
        ORC_FIND_ROW(query[2 * x + 1], r2);
        xq[2 * x] = r1;
        xq[2 * x + 1] = r2;
        for (int64_t j = 0; j < stride; j++) {
            const int64_t o = 2 * x * stride + 2 * j;
            int32_t a = -1, b = -1, c = -1, d = -1;
            if (r1 >= 0) {
                const int32_t w1 = walks[(int64_t)r1 * stride + j];
                ORC_FIND_IDX(r1, w1, a);
                ORC_FIND_IDX(r2, w1, b);
            }
            if (r2 >= 0) {
                const int32_t w2 = walks[(int64_t)r2 * stride + j];
                ORC_FIND_IDX(r1, w2, c);
                ORC_FIND_IDX(r2, w2, d);
            }
            out[o] = a; out[o + 1] = b; out[half + o] = c; out[half + o + 1] = d;
        }
    }
#undef ORC_FIND_ROW
#undef ORC_FIND_IDX
    idmap_free(&roots);
    idmap_free(&members);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* batch_sampler (subg_acc.c:391-507): the serial mini-batch node sampler.    */
/* One rand_r stream (the reference seeds it with seed + getpid(), :423; the  */
/* caller passes that sum).  For every query node in order: a partial Fisher- */
/* Yates over its neighbours if it has more than num_walks of them (:430-441; */
/* num_walks draws, before any walk), the node itself joins the batch (:443), */
/* then up to num_walks walks of num_steps nodes each (first hop w % deg or   */
/* the shuffled pick, later hops uniform with replacement, a hop from a node  */
/* without neighbours is skipped without a draw, :445-470); after every walk  */
/* the seed stops as soon as the batch holds (i + 1) * thld / n distinct      */
/* nodes (:472-473, int arithmetic).  Output: the distinct nodes in insertion */
/* order (uthash iterates in insertion order, :484-490).  N = node count      */
/* (sizes the seen map).  Returns the count, or -1 if it exceeds cap.         */
/* ------------------------------------------------------------------------- */
int64_t orc_batch_sampler(const int64_t *rowptr, const int32_t *col, int64_t N,
                          const int32_t *query, int64_t n, int num_walks, int num_steps, int thld,
                          uint32_t state, int32_t *out, int64_t cap)
{
    unsigned char *seen = (unsigned char *)calloc((size_t)(N > 0 ? N : 1), 1);
    int32_t *rseq = NULL;
    int64_t rcap = 0, count = 0;
    int overflow = 0;
    if (!seen) return -2;
#define ORC_ADD(v)                                         \
    do {                                                   \
        const int32_t _v = (v);                            \
        if (!seen[_v]) {                                   \
            seen[_v] = 1;                                  \
            if (count < cap) out[count] = _v;              \
            else overflow = 1;                             \
            count++;                                       \
        }                                                  \
    } while (0)
    for (int64_t i = 0; i < n; i++) {
        const int32_t u = query[i];
        const int64_t hop1 = rowptr[u + 1] - rowptr[u];
        if (hop1 > num_walks) {
            if (hop1 > rcap) {
                free(rseq);
                rcap = hop1;
                rseq = (int32_t *)malloc((size_t)rcap * sizeof(int32_t));
                if (!rseq) { free(seen); return -2; }
            }
            for (int64_t j = 0; j < hop1; j++) rseq[j] = (int32_t)j;
            for (int k = 0; k < num_walks; k++) {
                const int64_t s = (int64_t)((uint32_t)orc_rand_r(&state) % (uint32_t)(hop1 - k)) + k;
                const int32_t t = rseq[k];
                rseq[k] = rseq[s];
                rseq[s] = t;
            }
        }
        ORC_ADD(u);
        for (int walk = 0; walk < num_walks; walk++) {
            int32_t curr = u;
            if (hop1 < 1) break;
            else if (hop1 <= num_walks) curr = col[rowptr[curr] + walk % hop1];
            else curr = col[rowptr[curr] + rseq[walk]];
            ORC_ADD(curr);
            for (int step = 1; step < num_steps; step++) {
                const int64_t nn = rowptr[curr + 1] - rowptr[curr];
                if (nn > 0) {
                    curr = col[rowptr[curr] + (int64_t)((uint32_t)orc_rand_r(&state) % (uint32_t)nn)];
                    ORC_ADD(curr);
                }
            }
            if ((int)count >= (int)((i + 1) * (int64_t)thld / n)) break;
        }
    }
#undef ORC_ADD
    free(seen);
    free(rseq);
    return overflow ? -1 : count;
}
