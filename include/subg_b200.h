/*
 * subg_b200.h -- C ABI of the B200-native SubGAcc hot path.
 *
 * One shared library (surel_plus_b200/_lib/libsubg_b200.so), plain pointers and
 * sizes, no Python.h and no torch types.  This is the drop-in boundary: the
 * reference binds this path through a CPython extension module (`subg_acc`,
 * method table at subg_acc/subg_acc.c:1036-1043) and through SciPy calls in
 * train.py; each entry point below cites the reference interface it replaces.
 * INTEGRATION.md shows the ctypes stubs a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; subg_last_error()
 *     returns the message of the last failure on the calling thread.
 *         SUBG_ERR_ARG    -> reference raises TypeError      (subg_acc.c:658)
 *         SUBG_ERR_MEM    -> reference raises MemoryError    (subg_acc.c:688-721)
 *         SUBG_ERR_ASSERT -> reference raises AssertionError (subg_acc.c:913,1003)
 *   - "hd" pointers may point to host or device memory; the library inspects
 *     the pointer (cudaPointerGetAttributes) and copies when needed.  "dev"
 *     pointers must be device memory on the object's device.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls are asynchronous with respect to the host except where a size has
 *     to be returned (documented per function).
 *   - handles are not thread-safe; distinct handles may be used concurrently.
 *   - there is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef SUBG_B200_H
#define SUBG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUBG_ABI_VERSION 4

#define SUBG_OK          0
#define SUBG_ERR_ARG    -1
#define SUBG_ERR_MEM    -2
#define SUBG_ERR_ASSERT -3
#define SUBG_ERR_CUDA   -4
#define SUBG_ERR_UNSUPPORTED -5

/* RNG modes of the walk sampler */
#define SUBG_RNG_PHILOX 0 /* counter-based Philox4x32-10 keyed by (seed, seed index, walk): fast path, statistical parity */
#define SUBG_RNG_RAND_R 1 /* replays glibc rand_r exactly as the reference with nthread=1 (bit-exact end to end) */
#define SUBG_RNG_TRACE  2 /* walks supplied by the caller: int32 [n, num_walks, num_steps] */

/* status bits reported by subg_spg_info */
#define SUBG_STATUS_BUCKET_OVERFLOW 1u /* a set hit `bucket`; reference prints a warning (subg_acc.c:835-836) */
#define SUBG_STATUS_DEAD_END        2u /* RAND_R replay met a node without out-neighbours: stream no longer matches */
#define SUBG_STATUS_PPR_SECOND_PASS 4u /* some PPR seeds outgrew the first-pass workspace and were re-run (result unaffected) */

/* flags of subg_gset_sample* */
#define SUBG_SAMPLE_NO_RANKS 1 /* skip the first-visit ranks (`slot`): the SpG (sorted CSR-of-sets + LP table, what
                                  subg_matrix builds) is identical, only subg_spg_export needs the ranks */

#define SUBG_SAMPLE_DUMP_WALKS 2 /* keep the walks the sampler drew (int32 [n, num_walks, num_steps], subg_spg_walks): feeding
                                   them back as SUBG_RNG_TRACE, or to a CPU restatement of subg_acc.c:778-844, must
                                   reproduce the SpG bit for bit -- the parity hook of the Philox fast path */

#define SUBG_SAMPLE_NO_COMPACT 4 /* keep the rows where the sampler put them even when the worst-case allocation is mostly
                                   empty (a shard that only lives until it is packed for the multi-GPU exchange) */

/* structure encoders of utils.py:20-39 (the 'DEG' branch is broken upstream and not provided) */
#define SUBG_ENCODER_NONE 0
#define SUBG_ENCODER_PPR  1 /* utils.py:35-36 */
#define SUBG_ENCODER_SPD  2 /* utils.py:29-34 */

typedef struct subg_graph subg_graph; /* CSR graph resident in HBM (int64 or int32 rowptr, int32 col) */
typedef struct subg_spg subg_spg;     /* SpG: CSR-of-sets resident in HBM */
typedef struct subg_walkset subg_walkset; /* SUREL-v1 walks + relative-position encodings resident in HBM */

int subg_abi_version(void);
const char *subg_last_error(void);

/* ---- graph ------------------------------------------------------------------
 * Replaces the (indptr, indices) numpy arguments of gset_sampler
 * (subg_acc/subg_acc.c:655-671).  rowptr_is64: 1 -> const int64_t*, 0 -> const
 * int32_t* (what the reference accepts).  Graphs with E < 2^31 are stored with a
 * 32-bit rowptr in HBM (8-byte rowptr pair per walk step), larger ones with 64-bit. */
int subg_graph_create(const void *rowptr_hd, int rowptr_is64, const int32_t *col_hd,
                      int64_t N, int64_t E, int device, void *stream, subg_graph **out);
/* Edge list -> CSR on the device.  Replaces edge2csr (subg_acc/test/test.py:15-19:
 * csr_matrix((ones, (row, col)), shape=(nmax+1, nmax+1)): duplicate edges coalesced, columns ascending)
 * and, with symmetrize != 0, the symmetrisation of dataloader.py:119-129 (every edge stored in both
 * directions).  row_hd / col_hd: int64[E], host or device.  num_nodes < 0 -> largest id + 1.
 * drop_self_loops != 0 removes (u, u) entries.  The row pointer is kept 64-bit when the coalesced
 * graph has >= 2^31 entries.  Synchronises the stream. */
int subg_graph_from_edges(const int64_t *row_hd, const int64_t *col_hd, int64_t E, int64_t num_nodes,
                          int symmetrize, int drop_self_loops, int device, void *stream, subg_graph **out);
/* CSR of a resident graph back to the caller: rowptr int64[N+1], col int32[E] (host or device; either may be NULL) */
int subg_graph_export(const subg_graph *g, int64_t *rowptr_hd, int32_t *col_hd, void *stream);
int subg_graph_info(const subg_graph *g, int64_t *N, int64_t *E, int *device);
void subg_graph_free(subg_graph *g);

/* ---- set sampler + LP encoder + SpG build -----------------------------------
 * Replaces gset_sampler / C function set_sampler (subg_acc/subg_acc.c:649-1034)
 * and the CSR construction of subg_matrix (sampler/random_walks.py:79-80).
 *   seeds_hd   int32[n]  (the reference's `query`; must be distinct for SpJoin use)
 *   num_walks  M (1..32767), num_steps m >= 1 (walk length; CLI --num_steps - 1)
 *   bucket     < 0 -> M*m+1 slots per set, else the reference's `bucket` cap
 *   rng_mode   SUBG_RNG_*; walks_hd only for SUBG_RNG_TRACE
 *   flags      SUBG_SAMPLE_* bits
 * The sampler writes every set once, at a device-side cursor, into arrays sized for the worst case,
 * so the rows of the returned SpG are 16-byte aligned but not back to back ("scattered" layout:
 * subg_spg_rows); SpJoin reads that layout in place.  subg_spg_views compacts it into the CSR.
 * Synchronises the stream (set sizes decide allocations). */
int subg_gset_sample(const subg_graph *g, const int32_t *seeds_hd, int64_t n,
                     int num_walks, int num_steps, int bucket, uint64_t seed,
                     int rng_mode, const int32_t *walks_hd, int flags, void *stream, subg_spg **out);

/* Multi-GPU shard of the same call: seeds_hd is the WHOLE query (n_all entries)
 * and only the sets of the contiguous window [lo, hi) are sampled.  Seed indices stay global
 * (Philox counters, rand_r call offsets, first-occurrence positions of the LP rows), so the shards
 * of a range-partitioned query concatenate to exactly what one subg_gset_sample call returns.
 * walks_hd (TRACE mode) holds the walks of the window only: int32[hi-lo, num_walks, num_steps]. */
int subg_gset_sample_shard(const subg_graph *g, const int32_t *seeds_hd, int64_t n_all, int64_t lo,
                           int64_t hi, int num_walks, int num_steps, int bucket, uint64_t seed,
                           int rng_mode, const int32_t *walks_hd, int flags, void *stream, subg_spg **out);

/* Re-label the LP rows of a shard after the unique tables of all shards were merged in rank
 * order (= the first-occurrence order of subg_acc.c:957-978 over the whole query):
 * data <- id_map[data-1] + 1 for the c old ids, enc <- enc_hd int16[c_new, ncol].  Also attaches
 * a table to an SpG wrapped by subg_spg_from_csr (id_map_hd NULL, ncol > 0); ncol <= 0 keeps the width. */
int subg_spg_set_lp_table(subg_spg *s, const int32_t *id_map_hd, const int16_t *enc_hd,
                          int32_t c_new, int32_t ncol, void *stream);

/* n = sets, T = total entries, c = unique LP rows, ncol = num_steps+1 */
int subg_spg_info(const subg_spg *s, int64_t *n, int64_t *T, int32_t *c, int32_t *ncol,
                  int32_t *max_set, uint32_t *status, int32_t *value_kind);

/* The reference's return list [nsize, remap, enc(, raw_enc)] (subg_acc.c:1017-1024),
 * bit-compatible: remap[0] node ids in first-visit order, remap[1] LP-row ids in
 * first-occurrence order, enc int16[c, ncol].  raw_enc_hd (int16[T, ncol]) may be NULL.
 * Destinations may be host or device. */
int subg_spg_export(const subg_spg *s, int32_t *nsize_hd, int32_t *remap_hd,
                    int16_t *enc_hd, int16_t *raw_enc_hd, void *stream);

/* Device views of the sorted CSR-of-sets (what subg_matrix returns as scipy CSR,
 * sampler/random_walks.py:79): indptr int64[n+1], indices int32[T] ascending per
 * row, data int32[T] = LP-row id + 1 (0 = absent) or float64[T] for value SpGs,
 * slot uint16[T] first-visit rank (NULL without ranks), enc int16[c, ncol], nsize int32[n]
 * (NULL for wrapped CSRs).  Any out pointer may be NULL.  A sampler-built SpG is compacted into
 * this layout on the first call (one pass over its entries on `stream`, synchronised); earlier
 * subg_spg_rows pointers become invalid.  Views stay valid until subg_spg_free. */
int subg_spg_views(subg_spg *s, void *stream, const int64_t **indptr, const int32_t **indices,
                   const void **data, const uint16_t **slot, const int16_t **enc,
                   const int32_t **nsize);

/* The row layout SpJoin reads, without compaction: row u = entries [rowbeg[u], rowbeg[u] + nsize[u])
 * of indices/data (nsize == NULL for wrapped CSRs: the size is rowbeg[u+1] - rowbeg[u]);
 * *extent = entries in use (>= T because scattered rows are padded to 16 bytes). */
int subg_spg_rows(const subg_spg *s, const int64_t **rowbeg, const int32_t **nsize,
                  const int32_t **indices, const void **data, int64_t *extent);

/* The LP table alone: enc int16[c, ncol] on the device (c, ncol from subg_spg_info), without compacting the rows as
 * subg_spg_views does.  Work queued on `stream` afterwards is ordered behind the kernels that fill the table. */
int subg_spg_enc(const subg_spg *s, void *stream, const int16_t **enc);

/* Rows in seed order -> one row per graph node: row seeds[i] = set i, every other row empty.  This is the (N, N) matrix
 * subg_matrix builds for a query that is not arange(N) (csr_matrix((data, (repeat(idx, nsize), nodes)), (N, N)),
 * sampler/random_walks.py:79), so that SpJoin can index rows by node id.  Seeds must be distinct.  The entries are not
 * moved.  Afterwards subg_spg_info reports n = num_nodes. */
int subg_spg_expand_rows(subg_spg *s, int64_t num_nodes, void *stream);

/* The walks kept by SUBG_SAMPLE_DUMP_WALKS: int32 [n, num_walks, num_steps] on the device (NULL otherwise);
 * walk w of seed i visits walks[i][w][0..num_steps) after the seed itself. */
int subg_spg_walks(const subg_spg *s, void *stream, const int32_t **walks);

/* Wrap an existing CSR (e.g. the scipy matrix produced by the reference's
 * subg_matrix / topk_ppr_matrix+encoding) as an SpG for the join.
 * value_kind 0: int32 data (LP pointers), 1: float64 data (PPR / SPD values). */
int subg_spg_from_csr(const int64_t *indptr_hd, const int32_t *indices_hd, const void *data_hd,
                      int value_kind, int64_t n_rows, int64_t nnz, int device, void *stream,
                      subg_spg **out);
/* Multi-GPU assembly: an empty LP SpG in the CSR layout whose arrays are filled in place -- the
 * all-gather writes the shards of every rank straight into nsize int32[n], indices int32[T], data int32[T]
 * (pointers from subg_spg_views; indptr is derived by subg_spg_seal, which also finds the largest set and
 * checks that the sizes add up to T).  The LP table is attached with subg_spg_set_lp_table(id_map NULL). */
int subg_spg_alloc(int64_t n, int64_t T, int device, void *stream, subg_spg **out);
int subg_spg_seal(subg_spg *s, void *stream);
void subg_spg_free(subg_spg *s);

/* ---- multi-GPU exchange over NVLink peer memory ------------------------------
 * Nothing like it exists in the reference (single process, README.md:24); this is the scaling path of
 * gset_sampler + subg_matrix: one process per GPU, the graph replicated, every rank samples a contiguous
 * seed range (subg_gset_sample_shard) and ends with the full SpG.
 *   create   allocates the rank's *slab* (cudaMalloc, IPC-exportable), slab_bytes >= the packed size of the
 *            largest shard the rank will publish (8 bytes per entry + 12 per seed + 16 per unique LP row is safe)
 *   export   the 64-byte cudaIpcMemHandle_t of the slab; the host exchanges the handles of all ranks ...
 *   open     ... and maps the peers' slabs (handles: world x 64 bytes in rank order; the own entry is ignored)
 *   pack     shard -> slab in the packed wire format (4 / 5 / 6 / 8 bytes per entry, chosen from the node-id
 *            and LP-id widths) + set sizes, row offsets, unique LP keys with their first positions;
 *            header int64[8] (host) = {n, T, extent, c, format, max_set, status, bytes used}.  A shard that
 *            does not fit reports format < 0.  Asynchronous on `stream`.
 *   assemble headers int64[world, 8] of all ranks (the host all-gathers them; that collective is also the
 *            barrier after which the peers' slabs are complete) -> the full SpG on this GPU: the LP tables are
 *            merged on the device into the global first-occurrence order (subg_acc.c:957-978) and ONE kernel
 *            pulls every peer's packed entries over NVLink, widens and relabels them in flight.  srcs: NULL =
 *            the mapped peers; else `world` device pointers where the slabs can be read (e.g. slices of a
 *            staging buffer filled by an NCCL all-gather).  num_walks / ncol = the sampling call's M and
 *            num_steps + 1.  Synchronises the stream.
 * The slab may be re-packed only after every peer has finished its assemble (the host runs a barrier).
 * Time of the pull kernel: subg_timing_read(SUBG_TIMING_EXCHANGE). */
typedef struct subg_xchg subg_xchg;
#define SUBG_XCHG_HANDLE_BYTES 64
#define SUBG_XCHG_HEADER_WORDS 8
int subg_xchg_create(int device, int rank, int world, int64_t slab_bytes, subg_xchg **out);
int subg_xchg_export(const subg_xchg *x, void *handle64);
int subg_xchg_open(subg_xchg *x, const void *handles);
int subg_xchg_slab(const subg_xchg *x, void **slab_dev, int64_t *bytes);
int subg_xchg_pack(subg_xchg *x, const subg_spg *shard, int64_t num_nodes, int64_t *header, void *stream);
int subg_xchg_assemble(subg_xchg *x, const int64_t *headers, const void *const *srcs, int num_walks, int ncol,
                       void *stream, subg_spg **out);
/* The alternative to replicating the SpG: LINKED shards.  stage copies the shard UNPACKED into the slab (a plane of node
 * ids and a plane of LP ids at the same offsets in every slab; plane_entries = capacity of a plane, the same on every
 * rank; header as for pack, *header[4] < 0 if it did not fit).  link merges the LP tables, relabels THIS rank's id plane
 * to the global ids in place, fetches the row offsets / sizes of all shards (12 bytes per seed) and returns an SpG whose
 * rows stay where they were sampled: indices / data point at the first slab's planes and rowbeg[u] is row u's offset from
 * there, reaching into the peers' mapped slabs.  SpJoin takes it unchanged and reads remote rows over NVLink at join
 * time.  The caller must barrier after link (every rank has relabelled its plane) and keep every exchange context alive
 * while the SpG is in use; the SpG does not own its rows (subg_spg_views makes a local compact copy). */
int subg_xchg_stage(subg_xchg *x, const subg_spg *shard, int64_t num_nodes, int64_t plane_entries, int64_t *header, void *stream);
int subg_xchg_link(subg_xchg *x, const int64_t *headers, const void *const *srcs, int num_walks, int ncol, void *stream,
                   subg_spg **out);
void subg_xchg_free(subg_xchg *x);

/* ---- SpJoin -------------------------------------------------------------------
 * Replaces gather / bgather / pgather / hgather (train.py:13-111).
 * arity 2: edge int64[2,B]  -> segments [all left sets | all right sets]
 * arity 3: hedge int64[3,B] -> segments [u|w, w|u, v|w, w|v] (train.py:57-68)
 * plan:  edge_hd is copied to edge_dev (int64[arity*B], may equal edge_hd when it
 *        is already device memory), indptr_dev int64[nseg+1] receives the segment
 *        offsets (nseg = 2B or 4B) and *N_out the total number of rows.
 *        Synchronises the stream (N sizes the caller's output tensor).
 * run:   out_dev is  int32[N,2]      (enc_table_dev == NULL, int SpG)   -> train.py:82-83
 *                    float32[N,2,k]  (enc_table_dev float32[c+1,k])     -> train.py:37 encode[xz]
 *                    float32[N,2]    (float64 SpG, PPR/SPD mode)        -> train.py:39-43
 *        segid_dev (nullable) int64[N] receives the segment id of every row
 *        (train.py:27-30 ptr=False, and hgather's `ind`). */
int subg_spjoin_plan(const subg_spg *s, const int64_t *edge_hd, int64_t B, int arity,
                     int64_t *edge_dev, int64_t *indptr_dev, int64_t *N_out, void *stream);
int subg_spjoin_run(const subg_spg *s, const int64_t *edge_dev, int64_t B, int arity,
                    const int64_t *indptr_dev, const float *enc_table_dev, int k,
                    void *out_dev, int64_t *segid_dev, void *stream);

/* plan + run in one call with a single host synchronisation (the per-batch call of the training loop):
 * the caller passes out_dev (and segid_dev) with room for out_capacity rows, e.g. sized from the previous
 * batch; sizes, scan and the join kernel are queued back to back and the kernel checks on the device that
 * the rows fit.  *N_out = total rows; *ran = 1 if out_dev was written, 0 if out_capacity was too small (then
 * allocate *N_out rows and call subg_spjoin_run: edge_dev and indptr_dev are already filled). */
int subg_spjoin(const subg_spg *s, const int64_t *edge_hd, int64_t B, int arity, int64_t *edge_dev,
                int64_t *indptr_dev, const float *enc_table_dev, int k, void *out_dev, int64_t out_capacity,
                int64_t *segid_dev, int64_t *N_out, int *ran, void *stream);

/* The per-batch join of a training / evaluation loop (train.py:121-127: one gather per mini-batch of 1024 queries;
 * main_horder.py:33: 2048 triplets) without allocation or host synchronisation: a joiner fixes the SpG, the batch size,
 * the arity, the LP table and the output capacity, and captures [plan -> join] as a CUDA graph (two kernel nodes) for each
 * of `depth` ring slots; the plan kernel reads the batch's edges straight from the slot's pinned staging.  submit copies the batch's edges (int64[arity*B]; edge_on_device: 1 device memory,
 * 0 host memory, < 0 ask the driver) into the next slot and launches its graph on `stream` (host edges: one memcpy into
 * pinned staging + one cudaGraphLaunch; device edges: the same two kernels launched directly on the caller's array); it
 * returns the slot's device buffers at once:
 *   out_dev     rows [0, N) of the layout of subg_spjoin_run; rows beyond N are not written
 *   indptr_dev  int64[nseg+1] segment pointers (train.py:21-30)      segid_dev  int64 rows' segment ids (want_segid)
 *   nrows_dev   int64[2] on the device: {N, bad-node flag}
 * Host-edge batches alternate between two streams owned by the joiner (the plan kernel of batch k+1 and the PCIe read of
 * its edges run beside the join kernel of batch k; SUBG_JOIN_LANES=1 keeps everything on `stream`); each runs behind
 * `stream` as of the PREVIOUS submit, and `stream` is made to wait for the batch, so consumers simply use the buffers on
 * `stream`.  A slot is reused after `depth` submits: queue the work that reads its buffers on `stream` before the
 * (depth - 1)-th following submit (depth 3: submit(k), consume(k), or one batch of prefetch).
 * rows waits for the slot's last batch and returns N; SUBG_ERR_MEM if N exceeded capacity_rows (the batch was not
 * joined: run it through subg_spjoin), SUBG_ERR_ARG if a node id was outside the SpG.
 * The SpG must outlive the joiner and must not be compacted (subg_spg_views) while it exists. */
typedef struct subg_joiner subg_joiner;
int subg_joiner_create(const subg_spg *s, int64_t B, int arity, const float *enc_table_dev, int k,
                       int64_t capacity_rows, int want_segid, int depth, subg_joiner **out);
int subg_joiner_submit(subg_joiner *j, const int64_t *edge_hd, int edge_on_device, void *stream, void **out_dev,
                       int64_t **indptr_dev, int64_t **segid_dev, const int64_t **nrows_dev, int *slot);
int subg_joiner_rows(subg_joiner *j, int slot, int64_t *N);
void subg_joiner_free(subg_joiner *j);

/* ---- PPR set sampler -----------------------------------------------------------
 * Replaces topk_ppr_matrix (sampler/pprgo.py:83-111): ACL forward push per seed
 * (_calc_ppr_node, pprgo.py:9-38: LIFO queue, float32 p and r, float64 intermediate for the
 * pushed amount), top-k by score (pprgo.py:59; ties at the k-th score, which the reference's
 * unstable argsort leaves unspecified, go to the later-inserted node), CSR assembly and
 * normalisation (0 'row', 1 'sym', 2 'col'; pprgo.py:87-106, float64).  The result is a
 * float64 value SpG with rows in seed order and ascending node ids.
 *   norm_deg_hd  float64[N] = adj_matrix.sum(1) (the weighted degree used by 'sym'/'col'),
 *                or NULL to use the row length (unweighted adjacency).
 *   encoder      SUBG_ENCODER_*: applied after the normalisation (main.py:181-183).
 * The push needs CSR rows with strictly ascending columns (scipy canonical format); the
 * degree used inside the push is the row length (np.sum(adj > 0), pprgo.py:68).
 * Synchronises the stream. */
int subg_ppr_topk(const subg_graph *g, const int32_t *seeds_hd, int64_t n, float alpha,
                  float eps, int topk, int normalization, const double *norm_deg_hd,
                  int encoder, void *stream, subg_spg **out);

/* encoding(x, adj, 'PPR' | 'SPD') (utils.py:29-36) on a value SpG; returns a new SpG.
 * 'PPR': (x + 0.1) / (max(x) + 0.1) with the global max; g may be NULL.
 * 'SPD': 1*[w in N(u)] + 0.5*[w in S_u and two-hop(u,w)] + 0.3*[w in S_u], diagonal 2.3;
 *        the set becomes N(u) U S_u U {u}; needs one SpG row per graph node (idx = arange(N)). */
int subg_spg_encode(const subg_graph *g, const subg_spg *x, int encoder, void *stream,
                    subg_spg **out);

/* forward pushes performed while building a PPR SpG (0 for other SpGs) */
int subg_spg_pushes(const subg_spg *s, int64_t *pushes);

/* ---- SUREL-v1 walk sampler ---------------------------------------------------------
 * Replaces walk_sampler (subg_acc/subg_acc.c:316-389): random_walk (:144-181), random_walk_wo
 * (:183-248) and rpe_encoder (:250-314).  The reference returns [walks, obj]: walks int32
 * [n, num_walks*(num_steps+1)] (column 0 of every walk = the seed) and an object array whose row i
 * holds the unique nodes of seed i's walks in first-visit order of the step-major scan (root first)
 * and their int32 [count, num_steps+1] landing counts (entry [0][0] = num_walks).  Here the ragged
 * part is one CSR-like triple: off int64[n+1], ids int32[T], rpe int32[T, num_steps+1].
 *   replacement > 0 selects the first hop WITHOUT replacement, as the reference's flag does
 *                   (subg_acc.c:359-367); <= 0: every hop uniform with replacement.
 *   rng_mode        SUBG_RNG_PHILOX, or SUBG_RNG_RAND_R = the reference's nthread=1 stream (bit-exact).
 * Limits: num_walks*num_steps + 1 <= 16384; num_walks <= 4096 when replacement > 0.
 * Synchronises the stream (T sizes the allocations). */
int subg_walk_sample(const subg_graph *g, const int32_t *seeds_hd, int64_t n, int num_walks, int num_steps,
                     uint64_t seed, int rng_mode, int replacement, void *stream, subg_walkset **out);
int subg_walkset_info(const subg_walkset *w, int64_t *n, int64_t *T, int32_t *num_walks, int32_t *ncol,
                      uint32_t *status);
/* copies to host or device destinations (any may be NULL); synchronises the stream */
int subg_walkset_export(const subg_walkset *w, int32_t *walks_hd, int64_t *off_hd, int32_t *ids_hd,
                        int32_t *rpe_hd, void *stream);
/* device views, valid until subg_walkset_free */
int subg_walkset_views(const subg_walkset *w, const int32_t **walks, const int64_t **off, const int32_t **ids,
                       const int32_t **rpe);
void subg_walkset_free(subg_walkset *w);

/* Replaces walk_join (subg_acc/subg_acc.c:509-647), the SUREL-v1 join over walk sets.
 *   walks   int32[n, stride]: row i holds the walks of root walks[i*stride] (walk_sampler's first output)
 *   key     the node set of every row as a CSR pair: key_off int64[n+1], key_ids int32[T] (the reference takes a
 *           sequence of n arrays; walk_sampler's obj[:,0] concatenated); ids are unique within a row
 *   query   int32[Q, 2] node ids
 *   out     int32[2, Q*2*stride]: for query x = (u, v) and position j, with w1 = walks[row(u)][j], w2 = walks[row(v)][j]:
 *             out[0][2x*stride + 2j] = idx(u, w1)   out[0][.. + 1] = idx(v, w1)
 *             out[1][2x*stride + 2j] = idx(u, w2)   out[1][.. + 1] = idx(v, w2)
 *           idx(r, w) = 1 + position of w in the concatenated key sets if w is in the set of root r, else 0
 *           (-1 if r is not a root, subg_acc.c:87-100)
 *   xq      int32[Q, 2] (nullable): row of every query node, -1 if it is not a root (the reference's `return_idx`).
 * Entries that depend on the walks of an unknown root are -1 (the reference reads out of bounds there).
 * All pointers host or device.  Synchronises the stream. */
int subg_walk_join(const int32_t *walks_hd, int64_t n, int64_t stride, const int64_t *key_off_hd,
                   const int32_t *key_ids_hd, const int32_t *query_hd, int64_t Q, int32_t *out_hd, int32_t *xq_hd,
                   int device, void *stream);

/* ---- measurement hooks (bench.py / profiles) -------------------------------------
 * When enabled, the library brackets its dominant kernels with CUDA events on the
 * launching stream.  subg_timing_read synchronises those events, returns the summed
 * device time and number of launches of kernel class `which` since the last read and
 * clears them.  which: 0 set-sampler kernel, 1 SpJoin kernel, 2 SpG build (scan,
 * compaction, unique ranking, id remap), 3 PPR push kernel, 4 multi-GPU pull kernel.  subg_launch_count: kernels launched by the
 * library since load (all classes). */
#define SUBG_TIMING_SAMPLER 0
#define SUBG_TIMING_SPJOIN  1
#define SUBG_TIMING_BUILD   2
#define SUBG_TIMING_PPR     3
#define SUBG_TIMING_EXCHANGE 4 /* pull kernel of subg_xchg_assemble */
int subg_timing_enable(int enable);
int subg_timing_read(int which, double *ms, int64_t *launches);
int64_t subg_launch_count(void);

/* Replaces batch_sampler (subg_acc/subg_acc.c:391-507), the serial walk-based mini-batch node sampler: for every query node
 * in order, up to num_walks walks of num_steps nodes (first hop without replacement, later hops uniform); a seed stops as
 * soon as the batch holds (i + 1) * thld / n distinct nodes.  One rand_r stream starting at rng_state (the reference seeds
 * it with seed + getpid(), :423): same state, same batch, bit for bit.  out_hd (host or device, `capacity` entries)
 * receives the distinct nodes in insertion order, *count_out their number; SUBG_ERR_MEM if capacity was too small (an upper
 * bound is min(N, n * (num_walks * num_steps + 1))).  One warp replays the stream; synchronises the stream. */
int subg_batch_sample(const subg_graph *g, const int32_t *seeds_hd, int64_t n, int num_walks, int num_steps, int thld,
                      uint32_t rng_state, int32_t *out_hd, int64_t capacity, int64_t *count_out, void *stream);

/* The library keeps freed device blocks of 64 MB and more (SpG row arrays, sampler staging) in a per-process cache and
 * hands them out again (SUBG_BLOCK_CACHE_BYTES caps it, default 45 % of the device memory).  subg_trim_cache returns
 * every cached block to the driver and reports the bytes released. */
int64_t subg_trim_cache(void);

/* pinned host memory helpers (so numpy arrays handed back by the Python shim can be
 * filled with asynchronous copies) */
int subg_host_alloc(void **ptr, int64_t bytes);
void subg_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* SUBG_B200_H */
