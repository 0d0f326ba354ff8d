/*
 * semcode_ivf.h -- C ABI of the B200-native IVF_FLAT engine (libsemcode_ivf.so).
 *
 * The reference has no FFI for this path: its "FFI" is pymilvus -> gRPC -> Milvus.  Each entry
 * point below cites the reference call (path:line under /root/reference) whose work it replaces;
 * INTEGRATION.md shows the ctypes binding a semcode maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 (sc_status) on failure; sc_last_error() gives the
 *     thread-local message of the last failure on this thread.
 *   - plain pointers + sizes only.  Bulk array pointers may be HOST or DEVICE (same GPU) pointers;
 *     the library detects which (cudaPointerGetAttributes).  Device arrays must be 16-byte aligned,
 *     C-contiguous, row stride = dim floats.
 *   - `stream` is a cudaStream_t (void* here so that C callers need no CUDA headers); NULL = the
 *     legacy default stream.  All work of a call is ordered on that stream.
 *   - DEVICE outputs are valid once the caller synchronises `stream` (no hidden sync);
 *     HOST outputs are valid on return (the call synchronises `stream`).
 *   - the caller owns every input/output buffer; the library owns index storage and scratch.
 *   - a handle may be used from several threads.  Searches (sc_index_search*, sc_index_probe, sc_index_assign) share the
 *     handle: two of them run in the library at a time, each on its own scratch slot, and overlap on the device when
 *     their streams differ (one host thread alternating between two streams gets the same overlap: a stream keeps the
 *     slot it used last).  Every other call runs alone: it waits for running searches and holds new ones back, on the
 *     host and -- through events -- on the device, so a search never sees a half-applied insert.  This is the
 *     concurrent-search contract of the reference's server.
 *   - NVTX: every call and every search phase is an NVTX range (zero cost without a tool attached).
 *   - there is NO CPU fallback: every call fails with SC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SEMCODE_IVF_H
#define SEMCODE_IVF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_ABI_VERSION 3

typedef enum sc_status {
    SC_OK = 0,
    SC_ERR_INVALID = -1, /* bad argument */
    SC_ERR_CUDA = -2,    /* CUDA runtime / launch failure, or no usable device */
    SC_ERR_STATE = -3,   /* e.g. search/add before centroids exist */
    SC_ERR_OOM = -4      /* device allocation failed */
} sc_status;

typedef enum sc_metric {
    SC_METRIC_IP = 0, /* inner product, larger is better  (milvus_store.py:79 "metric_type": "IP") */
    SC_METRIC_L2 = 1  /* squared L2, smaller is better    (FAISS METRIC_L2, no sqrt)               */
} sc_metric;

typedef struct sc_index sc_index_t; /* opaque */

/* Scalar filter pushed into the list scan (replaces the client-side post-filter of
 * src/semcode/frontend/app.py:100-116 / gradio_app.py:79-94; Milvus would take it as `expr`).
 * HOST arrays.  n_repos == 0 -> any repo; n_langs == 0 -> any language. A row passes when its
 * repo tag is in repo_tags[] AND its language tag is in lang_tags[]. Removed rows never pass. */
typedef struct sc_filter {
    const uint32_t *repo_tags;
    int32_t n_repos;
    const uint8_t *lang_tags;
    int32_t n_langs;
} sc_filter_t;

typedef struct sc_stats {
    int32_t dim, dim_padded, metric, nlist, device, trained;
    int64_t ntotal;        /* live rows (added - removed) */
    int64_t nremoved;      /* tombstoned rows still occupying list slots */
    int64_t npages;        /* 32-row list pages handed out so far (including the free ones) */
    int64_t nfree_pages;   /* pages returned by sc_index_compact and not yet reused */
    int64_t bytes_lists;   /* device bytes held by list slabs (vectors + ids + tags) */
    int64_t bytes_scratch; /* device bytes held by scratch */
    int32_t max_list_len, min_list_len;
    /* device-side timing of the last sc_index_search call (valid after the stream has been
       synchronised and sc_index_last_search_times() has been called) */
} sc_stats_t;

/* per-phase device times (ms) of the most recent sc_index_search on this handle, measured with
 * CUDA events on the caller's stream when profiling is enabled with sc_index_set_profiling(). */
typedef struct sc_search_times {
    float coarse_ms, probe_select_ms, plan_ms, scan_ms, topk_ms, total_ms;
    int64_t scanned_rows;  /* rows whose distance was evaluated (sum over queries) */
    int64_t unique_rows;   /* list-major scan only: rows of the DISTINCT lists probed by the batch, i.e. the
                              compulsory rows; 0 when the query-major scan ran */
    int32_t scan_launches; /* launches of the list-scan kernel */
    int32_t total_launches;
} sc_search_times_t;

const char *sc_last_error(void);
int sc_abi_version(void);
/* number of visible CUDA devices with compute capability 10.x (0 => nothing will work) */
int sc_device_count(void);

/* -- lifecycle.  Replaces Collection(...)+create_index(IVF_FLAT, metric, nlist) at
 *    src/semcode/storage/milvus_store.py:75-84 ------------------------------------------------- */
int sc_index_create(int32_t dim, int32_t metric, int32_t nlist, int32_t device, sc_index_t **out);
int sc_index_destroy(sc_index_t *idx);
/* drop all rows (keeps centroids) */
int sc_index_reset(sc_index_t *idx);

/* -- coarse quantizer.  Replaces the server-side k-means that Milvus runs when a segment is
 *    sealed after Collection.upsert (milvus_store.py:128-130) [FAISS IndexIVFFlat::train] ------- */
/* Lloyd k-means on x[n,dim]; niter iterations; init rows / subsample drawn from `seed` exactly as
 * oracle/ivf_numpy.py does when init_rows/sub_rows are given by the host wrapper:
 *   init_rows [nlist] : rows (of the *subsampled* training set) that seed the centroids
 *   objective_out     : HOST array [niter] or NULL, objective entering each iteration */
int sc_index_train(sc_index_t *idx, const float *x, int64_t n, int32_t niter, const int64_t *init_rows,
                   double *objective_out, void *stream);
/* The three building blocks of sc_index_train, exposed so that data-parallel training (one process
 * per GPU, rows sharded) can all-reduce between them -- SURVEY.md section 8e:
 *   kmeans_init   centroids <- x[init_rows]  (every rank passes the same rows of the same x, or calls
 *                 sc_index_set_centroids with broadcast centroids instead)
 *   kmeans_step   one assignment pass over this rank's rows; ACCUMULATES into caller-owned DEVICE buffers
 *                 sums [nlist, dim_padded] fp64, counts [nlist] int32, objective [1] fp64
 *   kmeans_update centroids <- sums / counts, empty clusters split FAISS-style; nsplit_out host, nullable */
int sc_index_kmeans_init(sc_index_t *idx, const float *x, int64_t n, const int64_t *init_rows, void *stream);
int sc_index_kmeans_step(sc_index_t *idx, const float *x, int64_t n, double *sums, int32_t *counts,
                         double *objective, void *stream);
int sc_index_kmeans_update(sc_index_t *idx, const double *sums, const int32_t *counts, int32_t *nsplit_out,
                           void *stream);
int sc_index_set_centroids(sc_index_t *idx, const float *centroids, int32_t nlist, void *stream);
int sc_index_get_centroids(sc_index_t *idx, float *out /* [nlist,dim] host or device */, void *stream);

/* list id of the best centroid per row (FAISS quantizer->assign); out_list host or device [n] */
int sc_index_assign(sc_index_t *idx, const float *x, int64_t n, int32_t *out_list, void *stream);
/* coarse pass only: the nprobe best lists per query, best first (out host or device [nq,nprobe]) */
int sc_index_probe(sc_index_t *idx, const float *q, int64_t nq, int32_t nprobe, int32_t *out_lists,
                   float *out_scores /* nullable */, void *stream);

/* -- insert.  Replaces Collection.upsert([...,vectors,...]) at milvus_store.py:128-130
 *    [FAISS IndexIVFFlat::add_with_ids].  repo_tags / lang_tags may be NULL (all zero). -------- */
int sc_index_add(sc_index_t *idx, const float *x, const int64_t *ids, const uint32_t *repo_tags,
                 const uint8_t *lang_tags, int64_t n, void *stream);
/* same, with the list of every row supplied by the caller (parity runs / sharded insert) */
int sc_index_add_preassigned(sc_index_t *idx, const float *x, const int64_t *ids, const uint32_t *repo_tags,
                             const uint8_t *lang_tags, const int32_t *lists, int64_t n, void *stream);
/* tombstone rows by id (upsert = remove + add; Collection.upsert replaces by primary key,
 * milvus_store.py:128).  n_removed_out (host, nullable) receives how many rows matched. */
int sc_index_remove_ids(sc_index_t *idx, const int64_t *ids, int64_t n, int64_t *n_removed_out, void *stream);

/* drop the tombstoned slots of every list in place (row order inside a list is kept) and return the emptied pages to
 * the index's free list; pages_freed_out (host, nullable).  Replaces Milvus' background segment compaction [EXT]
 * behind the same Collection.upsert calls (milvus_store.py:128-130). */
int sc_index_compact(sc_index_t *idx, int64_t *pages_freed_out, void *stream);

/* -- search.  Replaces Collection.search(data=[vector], param={"metric_type":"IP","params":
 *    {"nprobe":16}}, limit=top_k) at milvus_store.py:141-147 [FAISS IndexIVFFlat::search].
 *    out_dist [nq,k] raw inner product or squared L2, best first; out_ids [nq,k], -1 = no result
 *    (then dist = -FLT_MAX for IP, +FLT_MAX for L2).  filter may be NULL.  k <= 2048,
 *    nprobe is clamped to nlist. -------------------------------------------------------------- */
int sc_index_search(sc_index_t *idx, const float *q, int64_t nq, int32_t k, int32_t nprobe,
                    const sc_filter_t *filter, float *out_dist, int64_t *out_ids, void *stream);
/* same, probing caller-supplied lists (host or device [nq,nprobe]; -1 entries are skipped) */
int sc_index_search_preassigned(sc_index_t *idx, const float *q, int64_t nq, int32_t k, int32_t nprobe,
                                const int32_t *lists, const sc_filter_t *filter, float *out_dist,
                                int64_t *out_ids, void *stream);

/* -- cross-shard reduce.  Replaces the Milvus proxy's reduce over segments/shards [EXT]:
 *    part_dist/part_ids [parts,nq,kin] (device) -> out [nq,k] best first; ids < 0 are ignored. -- */
int sc_merge_topk(const float *part_dist, const int64_t *part_ids, int32_t parts, int64_t nq, int32_t kin,
                  int32_t k, int32_t metric, float *out_dist, int64_t *out_ids, int32_t device, void *stream);

/* -- fused cross-GPU exchange (one process per GPU; NVLink / NVSwitch peer memory).  Replaces the Milvus
 *    proxy's scatter of a search to the query nodes and its gather + reduce of their partial results
 *    [EXT] behind the same Collection.search call (milvus_store.py:141-147).
 *    peer_buffers[p] (HOST array of `world` DEVICE pointers, valid on `device`): rank p's exchange buffer as
 *    mapped into this process -- e.g. torch.distributed._symmetric_memory buffer_ptrs, cudaIpcOpenMemHandle
 *    or cuMemMap of a fabric handle; peer_buffers[rank] is this rank's own buffer.  All buffers have
 *    buffer_bytes bytes, are zero-filled before the first use, and stay alive until sc_exchange_destroy.
 *    Every rank must make the same sequence of sc_index_search_sharded calls (same nq, k, nprobe). --------- */
typedef struct sc_exchange sc_exchange_t; /* opaque */
int sc_exchange_create(int32_t rank, int32_t world, const void *const *peer_buffers, int64_t buffer_bytes,
                       int32_t device, sc_exchange_t **out);
int sc_exchange_destroy(sc_exchange_t *ex);
/* timed_out (host, nullable): 1 when a wait for a peer gave up (results of that step are invalid);
 * epoch (host, nullable): number of exchange steps issued.  Synchronises the device. */
int sc_exchange_status(sc_exchange_t *ex, int32_t *timed_out, int64_t *epoch);
/* the same flag WITHOUT synchronising (a pinned host word the waiting kernel stores to): 1 once a finished step of this
 * rank gave up waiting for a peer, or a step failed after its epoch had moved.  sc_index_search_sharded refuses to run
 * on such an exchange (SC_ERR_STATE): destroy it on every rank, barrier, create a new one. */
int sc_exchange_poll(sc_exchange_t *ex, int32_t *timed_out);
/* how long the waiting kernels spin for a peer before giving up (default 10 s) */
int sc_exchange_set_timeout_ms(sc_exchange_t *ex, int64_t ms);
/* One search step of a row-sharded index: `idx` holds this rank's rows, every rank passes the same queries.
 * lists == NULL: the coarse pass is split over the ranks (rank r ranks the centroids for its 1/world of the
 * batch and STORES the probe rows into every peer's table); the top-k epilogue STORES this rank's partial
 * result into every peer's gather slot and publishes a flag; the merge kernel waits for the flags of all peers
 * and writes the merged [nq,k] result to out_dist / out_ids (identical on every rank).  No collective call. */
int sc_index_search_sharded(sc_index_t *idx, sc_exchange_t *ex, const float *q, int64_t nq, int32_t k, int32_t nprobe,
                            const int32_t *lists, const sc_filter_t *filter, float *out_dist, int64_t *out_ids,
                            void *stream);

/* -- introspection / export (persistence, CPU baseline, tests) -------------------------------- */
int sc_index_stats(sc_index_t *idx, sc_stats_t *out);
int sc_index_list_sizes(sc_index_t *idx, int32_t *out_host /* [nlist], slots incl. tombstones */);
/* copy one list out (host or device buffers, each nullable): vectors [len,dim], ids, tags, where
 * tags = (removed<<31)|(repo<<8)|lang.  len_out (host) receives the slot count. cap = rows the
 * buffers can hold. */
int sc_index_export_list(sc_index_t *idx, int32_t list, int64_t cap, float *vecs, int64_t *ids, uint32_t *tags,
                         int64_t *len_out, void *stream);
/* lists [list_begin, list_end) back to back in slot order (tombstoned slots included, see tags); off_out
 * [list_end - list_begin + 1] (host, nullable) = exclusive prefix of the slot counts; cap = rows the buffers hold.
 * One call per ~GB instead of one per list: snapshots, re-training, the CPU baseline. */
int sc_index_export_lists(sc_index_t *idx, int32_t list_begin, int32_t list_end, int64_t cap, float *vecs, int64_t *ids,
                          uint32_t *tags, int64_t *off_out, void *stream);
int sc_index_set_profiling(sc_index_t *idx, int32_t enabled);
int sc_index_last_search_times(sc_index_t *idx, sc_search_times_t *out);
/* tuning knobs (tests / bench; the defaults are the measured best):
 *   "scratch_bytes"  search scratch ceiling (>= 1 MiB, default 8 GiB): larger batches run in equal passes
 *   "scan_mode"      0 = automatic (list-major from nq*nprobe >= nlist/8 up to dim 1024 (nlist/4 for batches of <= 16), nlist/2 above, 3/4 nlist with a filter), 1 = query-major,
 *                    2 = list-major
 *   "scan_variant"   0..4 rows x loads in flight of the query-major scan
 *   "lists_cfg"      tile items of the list-major scan: 0 = tcgen05 where it applies (inner product, dim % 32 == 0; list
 *                    rows as a tensor-memory operand), 1 / 2 = exact-fp32 FFMA tiles (64- / 32-float stages), 3 / 5 = the
 *                    earlier tcgen05 tile kernels with both operands in shared memory (v1 / v2),
 *                    4 = 0 with the 8-query page scan on mma.sync (parity-green, measured slower)
 *   "lists_fork"     1 = tile items on a side stream next to the page scans
 *   "tile_rem"       4 = list-major: remainders of 5..16 queries per list become tcgen05 tile items as well (default 0: 9..16
 *                    only, and only when enough lists have them): +5.6 % on the clustered headline, -3.5 % on the iid one
 *   "mq_fused"       1 = the list-major page scans (lists probed by 1..4 / 5..16 queries) in one launch in which every
 *                    warp works on both buckets (measured slower than the default two launches)
 *   "coarse_impl"    0 = tcgen05 3xTF32 contraction, 1 = fp32 SIMT;  "tc_variant" 0 = 256x256, 1 = 128x256 tiles
 *   "small_coarse"   1 (default) = streamed fp32 coarse kernel for batches of <= 16 queries
 *   "fuse_plan"      1 (default) = batches of <= 16 queries: probe selection and pair plan in one launch
 *   "pdl"            1 (default) = ... and the step's kernels chained by programmatic dependent launch
 *   "debug_canary"   1 = every scratch buffer allocated from now on (process-wide) sits between two 256-byte guards, without
 *                    growth slack;  "check_canaries" (value = fewest guarded buffers expected) returns SC_ERR_STATE when a
 *                    kernel wrote outside one -- the stand-in for compute-sanitizer memcheck where that is unavailable
 *   "plan_epoch"     tests: launch counter of the pair plan's look-back words (22-bit wrap)
 *   "add_chunk_rows" tests: rows per insert chunk (0 = automatic);  "fail_add_after" tests: the n-th insert chunk from
 *                    now fails after its slots were claimed (the index must stay consistent) */
int sc_index_set_param(sc_index_t *idx, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* SEMCODE_IVF_H */
