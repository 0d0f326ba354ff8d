// Shared declarations for libsemcode_ivf (sm_100a only).
//
// HBM layout of one index (see DESIGN.md "Data layout"):
//   centroids   [nlist, ds] fp32, ds = dim rounded up to 4 floats (16-byte rows)
//   list pages  fixed 32-row pages drawn from slabs; slab s holds `pages_per_slab` pages:
//                 vec  [pages_per_slab*32, ds] fp32
//                 ids  [pages_per_slab*32]     int64
//                 tags [pages_per_slab*32]     uint32 = removed<<31 | repo<<8 | lang
//               unused slots keep tags = 0xFFFFFFFF (removed bit set) so they never match.
//   page table  CSR: pt_off[nlist+1] (int32), pt[total_pages] (int32 page ids), list_len[nlist]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace sc {

constexpr int kPageRows = 32;
constexpr int kMaxSlabs = 1024;
constexpr uint32_t kTagRemoved = 0x80000000u;
constexpr uint32_t kTagRepoMax = (1u << 23) - 1;
constexpr int kMaxK = 2048;

struct SlabTable {
    float *vec[kMaxSlabs];
    int64_t *ids[kMaxSlabs];
    uint32_t *tags[kMaxSlabs];
};

__host__ __device__ __forceinline__ uint32_t make_tag(uint32_t repo, uint32_t lang) {
    return ((repo & kTagRepoMax) << 8) | (lang & 0xffu);
}

// ---- device helpers -----------------------------------------------------------------------
#ifdef __CUDACC__

// streaming 128-bit load: read-only path, do not allocate in L1 (list rows are read once)
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch (small-batch path): a kernel launched with launch_pdl(..., true) may become resident while
// its predecessor in the stream still runs; pdl_wait() blocks until that predecessor has completed and its stores are
// visible, so every global access of the kernel must come after it.  pdl_launch_dependents() lets the NEXT kernel in the
// stream do the same with respect to this one.  Both are no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// order-preserving float -> uint key (larger float => larger key). NaN -> 0 (below -inf).
constexpr uint32_t kKeyNegInf = 0x007fffffu;  // key of -inf; keys <= this are "no result"
__device__ __forceinline__ uint32_t f2key(float f) {
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

#endif  // __CUDACC__

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- kernel launchers (host API, defined in the .cu files) ---------------------------------

// C[M,N] = A[M,K] . B[N,K]^T ; if bnorm != nullptr: C = 2*C - bnorm[n]   (all fp32, K % 4 == 0)
cudaError_t launch_gemm_nt(const float *A, int64_t M, const float *B, int N, int K, const float *bnorm, float *C,
                           cudaStream_t st);
// ---- tensor-core contraction (gemm_tc.cu): tcgen05 kind::tf32, 3xTF32 split, TMA-fed -----------------
// hi = tf32(x), lo = tf32(x - hi); n = number of floats (multiple of 4)
cudaError_t launch_split_tf32(const float *x, int64_t n, float *hi, float *lo, cudaStream_t st);
// C[M,N] = alpha * A.B^T - bias[n]   (A, B given as hi/lo pairs, row stride K floats, K % 4 == 0)
cudaError_t launch_gemm_tc_scores(const float *ahi, const float *alo, int64_t M, const float *bhi, const float *blo, int N,
                                  int K, float alpha, const float *bias, float *C, int num_sms, cudaStream_t st);
// per row of A: best_idx = n_base + argmax_n (alpha * A.B^T - bias), best_val = that maximum; merge != 0 folds the
// values already stored in best_val / best_idx in (for centroid slabs); ties -> lowest index
cudaError_t launch_gemm_tc_argmax(const float *ahi, const float *alo, int64_t M, const float *bhi, const float *blo, int N,
                                  int K, float alpha, const float *bias, float *best_val, int32_t *best_idx, int n_base,
                                  int merge, int num_sms, unsigned long long *packed, cudaStream_t st);
// packed != nullptr: 256x256 tiles; every launch atomicMax-merges (key << 32 | ~index) into packed[M] (zeroed by the
// caller before the first slab) and launch_unpack_argmax writes best_val / best_idx afterwards.  packed == nullptr:
// 128x256 tiles, results straight into best_val / best_idx (merge flag as above).
cudaError_t launch_unpack_argmax(const unsigned long long *packed, int64_t M, float *best_val, int32_t *best_idx, cudaStream_t st);

// C[nq,N] = alpha * q.B^T - bias[n] for nq <= 16 (one warp per row of B); cudaErrorNotSupported when it does not fit
cudaError_t launch_coarse_small(const float *q, int64_t nq, const float *cent, int nlist, int ds, float alpha, const float *bias,
                                float *scores, int num_sms, cudaStream_t st);
// out[r] = sum_k x[r,k]^2
cudaError_t launch_row_norms(const float *x, int64_t rows, int ds, float *out, cudaStream_t st);
// per row: index of the largest score (ties -> lowest index) and the score
cudaError_t launch_argmax_rows(const float *scores, int64_t M, int N, int32_t *out_idx, float *out_val,
                               cudaStream_t st);
// out[n,ds] = zero-padded copy of in[n,d]
cudaError_t launch_pad_rows(const float *in, int64_t n, int d, int ds, float *out, cudaStream_t st);

// top-k of each row of scores[M,N] (largest first) -> idx [M,k] int32, val [M,k] (nullable)
cudaError_t launch_select_rows(const float *scores, int64_t M, int N, int k, int32_t *out_idx, float *out_val,
                               cudaStream_t st);

struct FilterDev {
    uint32_t flags;  // bit0: language filter, bit1: repo filter
    uint32_t lang_bits[8];
    const uint32_t *repo_bits;
    uint32_t n_repo_bits;
};

#ifdef __CUDACC__
// fused scalar predicate of the list scans: tombstone | language bitmap | repo bitmap
__device__ __forceinline__ bool filter_pass(const FilterDev &f, uint32_t tag) {
    if (tag & kTagRemoved) return false;
    if (f.flags & 1u) {
        const uint32_t lang = tag & 0xffu;
        if (!((f.lang_bits[lang >> 5] >> (lang & 31u)) & 1u)) return false;
    }
    if (f.flags & 2u) {
        const uint32_t repo = (tag >> 8) & kTagRepoMax;
        if (repo >= f.n_repo_bits) return false;
        if (!((__ldg(f.repo_bits + (repo >> 5)) >> (repo & 31u)) & 1u)) return false;
    }
    return true;
}
#endif

struct ScanArgs {
    const float *q;  // [nq, ds]
    int ds;
    int metric;
    int nprobe;
    int64_t npairs;           // nq * nprobe
    const int32_t *probe;     // [npairs] list ids (-1 = skip)
    const int64_t *page_off;  // [npairs+1] exclusive prefix of pages per pair
    const int32_t *list_len;  // [nlist]
    const int32_t *pt_off;    // [nlist+1]
    const int32_t *pt;        // page ids
    const SlabTable *slabs;
    int slab_shift;
    float *cand;  // [page_off[npairs]*32] similarity to maximise, -inf for dead slots
    int64_t max_cand;  // host-side upper bound of the candidate slots of ONE query (sizes the selection's CTAs)
    FilterDev filt;
    const void *slab_maps;  // one 128-byte TMA tensor map per slab (encode_slab_map), or nullptr
};

// ---- cross-GPU exchange over peer-mapped memory (NVLink / NVSwitch), one process per GPU -------------------
// The top-k epilogue of rank r stores its partial result straight into every peer's gather slot r and the
// last CTA publishes the step's epoch to every peer's flag word r; the consumer kernel on each peer spins on
// its own (local) flag words.  No collective library call, no extra copy: the exchange rides on the stores
// of the kernel that produced the data.
constexpr int kMaxPeers = 8;
struct PeerSignal {
    int world;                               // 0 = no exchange (plain local output)
    unsigned int *done;                      // local CTA counter, zero between launches
    unsigned long long *flag[kMaxPeers];     // peer p's flag word for THIS rank
    unsigned long long epoch;
};
struct PeerWait {
    int world;                               // 0 = nothing to wait for
    const unsigned long long *flags;         // local flag words [world], written by the peers
    unsigned long long epoch;
    unsigned int *status;                    // local; set to 1 when a peer did not arrive within timeout_ns
    unsigned long long timeout_ns;
};
struct PeerTopk {
    float *d[kMaxPeers];                     // peer p's [nq, k] slot of THIS rank
    int64_t *i[kMaxPeers];
};
struct PeerRows {
    int32_t *p[kMaxPeers];                   // peer p's probe table [nq, nprobe]
    int64_t row0;                            // first row of THIS rank's slice
};

// small batches (M <= kPlanTailMaxQ, k <= 128): launch_select_rows and launch_plan_pairs in ONE launch (select.cu).
// ws: kPlanTailWords 64-bit words of scratch, zero when first used (the kernel leaves them ready for the next launch)
constexpr int kPlanTailMaxQ = 16;
constexpr int kPlanTailWords = 1 + kPlanTailMaxQ;
cudaError_t launch_select_rows_plan(const float *scores, int64_t M, int N, int k, int32_t *out_idx, const int32_t *list_len,
                                    int32_t nlist, int64_t *page_off, unsigned long long *ws, unsigned long long *rows_total,
                                    cudaStream_t st, bool pdl);

// pair plan (scan.cu): page_off [npairs + 1] = exclusive prefix of the pages of every (query, list) pair, one launch
// (single-pass scan, decoupled look-back).  look: plan_pairs_look_words(npairs) 64-bit words of scratch, zero when
// first used; epoch: differs from the recent launches on the same scratch (22 bits are used: the caller counts up and
// zeroes the scratch when the count wraps).  rows_total (optional): += rows of every pair's list
size_t plan_pairs_look_words(int64_t npairs);
cudaError_t launch_plan_pairs(const int32_t *probe, int64_t npairs, const int32_t *list_len, int32_t nlist,
                              int64_t *page_off, unsigned long long *look, uint32_t epoch,
                              unsigned long long *rows_total, cudaStream_t st);
// exclusive prefix sum of in[n] -> out[n+1] (out[n] = total), one CTA (inputs are nlist long)
cudaError_t launch_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out, cudaStream_t st);
cudaError_t launch_scan_pages(const ScanArgs &a, int variant, int num_sms, int *launches, cudaStream_t st, bool pdl = false);

// list-major scan (scan_lists.cu, scan_mq.cu): the probed lists are read ONCE per batch and scored against every
// query that probes them.  Scratch (all device, caller-sized): cnt | cursor | counters | agg adjacent (one memset;
// 2 * nlist + 4 is even, so agg is 8-byte aligned when cnt is),
// n32 [nlist], lq_off / off32 / pg8off / pg4off [nlist+1], lq [npairs].  Writes the same candidate layout as
// launch_scan_pages.
struct ListPlan {
    int32_t nlist;
    int32_t *cnt, *cursor;             // [nlist] queries per list, fill cursor
    int32_t *counters;                 // [4] (directly after cursor) 0: work counter of the FFMA tile kernel, 1: lists probed by > 8 queries,
                                       //     2: ticket counter of the plan kernel
    unsigned long long *agg;           // [list_plan_ctas(nlist)][4] (directly after counters) per-CTA totals of the plan's four prefix sums
    int32_t chunk;                     // queries per tile item: 32 (FFMA tiles) or 64 (tcgen05 tiles)
    float *qsplit;                     // tcgen05 tiles: 2 x [nq, ds] tf32 terms (hi, lo) of the queries (scratch)
    float *bstage;                     // scan_lists_ts.cu: per-CTA query staging slots (scan_lists_ts_stage_bytes), or nullptr
    int32_t tile_rem;                  // 4 = remainders of 5..16 queries go to tile items (set_param "tile_rem"); 0 = 9..16 only
    int32_t mq_fused;                  // 1 = the two page-scan buckets in one launch (set_param "mq_fused"; measured slower: off)
    int32_t *n32;                      // [nlist] tile items (of `chunk` queries) per list
    int32_t *lq_off, *off32;           // [nlist+1] exclusive prefixes of cnt / n32 (chunk == 64: of n32 x 128-row tiles)
    int32_t *pg8off, *pg4off;          // [nlist+1] exclusive prefixes of the page x pass units of the two page scans
    int32_t *lq;                       // [npairs] pair ids grouped by list
    // optional fork/join: the (few, long) tile items run on a side stream while the page scans fill the GPU
    cudaStream_t side[2];
    cudaEvent_t ev_fork, ev_join[2];
    unsigned long long *unique_rows;   // optional: += rows of every list probed at least once
};
int list_plan_ctas(int32_t nlist);
cudaError_t launch_scan_lists(const ScanArgs &a, const ListPlan &p, int cfg, int num_sms, int *launches, cudaStream_t st);
// tile items on the tensor cores (scan_lists_tc.cu): inner product, ds % 32 == 0, p.chunk == 64
cudaError_t launch_scan_lists_tc(const ScanArgs &a, const ListPlan &p, int variant, int num_sms, cudaStream_t st);
// the same items with the list rows as a tensor-memory operand (scan_lists_ts.cu); needs a.slab_maps
cudaError_t launch_scan_lists_ts(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st);
size_t scan_lists_ts_stage_bytes(int ds, int num_sms);
// out = 2 x [n4] float4: the tf32 terms hi = tf32(q), lo = tf32(q - hi) of the query rows
cudaError_t launch_split_queries(const float *q, int64_t n4, float *out, int num_sms, cudaStream_t st);
// TMA tensor map (128 bytes, host memory) of one list slab: fp32 [rows, ds], boxes of one page x 32 floats, 128 B swizzle
cudaError_t encode_slab_map(void *map128, const float *base, int64_t rows, int ds);
// final top-k over the candidates of each query + id translation
cudaError_t launch_select_candidates(const ScanArgs &a, int64_t nq, int k, float *out_dist, int64_t *out_ids,
                                     cudaStream_t st, bool pdl = false);
cudaError_t launch_merge_topk(const float *part_dist, const int64_t *part_ids, int parts, int64_t nq, int kin, int k,
                              int metric, float *out_dist, int64_t *out_ids, cudaStream_t st);
// exchange variants (select.cu): the same selections, writing into every peer / waiting for every peer
cudaError_t launch_select_rows_peers(const float *scores, int64_t M, int N, int k, const PeerRows &rows,
                                     const PeerSignal &sig, cudaStream_t st);
cudaError_t launch_select_candidates_peers(const ScanArgs &a, int64_t nq, int k, const PeerTopk &out, const PeerSignal &sig,
                                           cudaStream_t st);
cudaError_t launch_peer_signal(const PeerSignal &sig, cudaStream_t st);  // a rank with an empty slice still publishes
cudaError_t launch_peer_wait(const PeerWait &wait, cudaStream_t st);     // orders the stream after the peers' stores

// list maintenance
// bad[0] += rows with a list id outside [0, nlist), bad[1] += rows with a repo tag above kTagRepoMax (repo nullable)
cudaError_t launch_count_positions(const int32_t *assign, const uint32_t *repo, int64_t n, int32_t nlist, int32_t *list_len,
                                   int32_t *pos, int32_t *bad, cudaStream_t st);
cudaError_t launch_page_need(const int32_t *len_old, const int32_t *len_new, int32_t nlist, int32_t *need,
                             int32_t *npg_new, cudaStream_t st);
// the batch's new pages: numbers < nfree come from free_pages[] (returned by compaction), the rest from pool_top up
cudaError_t launch_rebuild_pt(const int32_t *pt_off_old, const int32_t *pt_old, const int32_t *pt_off_new,
                              int32_t *pt_new, const int32_t *need_off, int32_t pool_top, const int32_t *free_pages,
                              int32_t nfree, int32_t nlist, cudaStream_t st);
// compaction: squeeze the tombstoned slots out of every list in place (list_len shrinks), then rebuild the page
// table keeping ceil(len / 32) pages per list and appending the others to free_pages at *free_cursor
cudaError_t launch_compact_lists(int32_t nlist, int32_t *list_len, const int32_t *pt_off, const int32_t *pt,
                                 const SlabTable *slabs, int slab_shift, int ds, int num_sms, cudaStream_t st);
cudaError_t launch_pages_of_len(const int32_t *len, int32_t nlist, int32_t *npg, cudaStream_t st);
cudaError_t launch_compact_pt(const int32_t *pt_off_old, const int32_t *pt_old, const int32_t *pt_off_new, int32_t *pt_new,
                              int32_t nlist, int32_t *free_pages, int32_t *free_cursor, cudaStream_t st);
// lists [l0, l0 + nl) back to back into vecs / ids / tags (nullable); off [nl + 1] device exclusive prefix of their slots
cudaError_t launch_export_range(const int32_t *pt_off, const int32_t *pt, int32_t l0, int32_t nl, const int64_t *off, int64_t rows,
                                int ds, int d_out, const SlabTable *slabs, int slab_shift, float *vecs, int64_t *ids, uint32_t *tags,
                                int num_sms, cudaStream_t st);
cudaError_t launch_scatter_rows(const float *x, const int64_t *ids, const uint32_t *repo, const uint8_t *lang,
                                const int32_t *assign, const int32_t *pos, int64_t n, int ds, const int32_t *pt_off,
                                const int32_t *pt, const SlabTable *slabs, int slab_shift, cudaStream_t st);
cudaError_t launch_remove_ids(const int64_t *sorted_ids, int64_t nrm, const SlabTable *slabs, int slab_shift,
                              int64_t npages, unsigned long long *count, cudaStream_t st);
cudaError_t launch_export_list(const int32_t *pt, int32_t pt_begin, int32_t len, int ds, int d_out,
                               const SlabTable *slabs, int slab_shift, float *vecs, int64_t *ids, uint32_t *tags,
                               cudaStream_t st);

// k-means
cudaError_t launch_kmeans_accumulate(const float *x, int64_t n, int ds, const int32_t *assign, const float *best,
                                     int metric, double *sums, int32_t *counts, double *objective, cudaStream_t st);
cudaError_t launch_kmeans_finalize(const double *sums, const int32_t *counts, int32_t nlist, int ds, float *centroids,
                                   cudaStream_t st);
cudaError_t launch_split_centroid(float *centroids, int ds, int32_t ci, int32_t cj, cudaStream_t st);
cudaError_t launch_gather_rows(const float *x, const int64_t *rows, int64_t n, int ds, float *out, cudaStream_t st);

}  // namespace sc
