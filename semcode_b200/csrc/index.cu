// Host side of libsemcode_ivf.so: the C ABI declared in include/semcode_ivf.h.
//
// One sc_index = one IVF_FLAT index resident on one B200: replicated centroids, paged inverted
// lists (common.cuh), and stream-ordered scratch.  It replaces what Milvus does server-side behind
// reference src/semcode/storage/milvus_store.py:75-84 (create_index), :128-130 (upsert -> train /
// add) and :141-147 (search).  There is no CPU path in here: every entry point launches the
// sm_100a kernels of this directory or fails.
#include <float.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <queue>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/semcode_ivf.h"
#include "common.cuh"

using namespace sc;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cudaGetLastError();                                                                      \
            return fail(e_ == cudaErrorMemoryAllocation ? SC_ERR_OOM : SC_ERR_CUDA, "%s: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                          \
        }                                                                                            \
    } while (0)

#define SC(call)               \
    do {                       \
        int r_ = (call);       \
        if (r_ != SC_OK) return r_; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
namespace {

// Debug guards (set_param "debug_canary"): every scratch buffer allocated while the switch is on gets a 256-byte guard in
// front and one right behind the bytes that were asked for (no growth slack), filled with 0xA5; set_param "check_canaries"
// verifies all of them.  compute-sanitizer is closed on the GPU pool this was developed on (profiles/
// r2_sanitizer_closed_on_pool.log); this catches what memcheck would catch in the buffers the plan / scan / select kernels
// index with computed offsets: writes before the start or past the end.
static bool g_canary = false;
constexpr size_t kGuardBytes = 256;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    size_t guard = 0;  // bytes of guard in front of p (and behind p + cap); 0 = no guards
    unsigned gen = 0;  // counts (re)allocations: "did the contents survive?" (a new block may come back at the old address)
    cudaError_t reserve(size_t bytes) {
        if (bytes == 0) bytes = 256;
        if (bytes <= cap && (guard != 0) == g_canary) return cudaSuccess;
        release();  // cudaFree waits for in-flight work that may still use the buffer
        ++gen;
        if (g_canary) {
            const size_t want = (bytes + 15) & ~(size_t)15;
            char *base = nullptr;
            cudaError_t e = cudaMalloc(&base, want + 2 * kGuardBytes);
            if (e != cudaSuccess) return e;
            e = cudaMemset(base, 0xA5, kGuardBytes);
            if (e == cudaSuccess) e = cudaMemset(base + kGuardBytes + want, 0xA5, kGuardBytes);
            if (e != cudaSuccess) {
                cudaFree(base);
                return e;
            }
            p = base + kGuardBytes;
            cap = want;
            guard = kGuardBytes;
            return cudaSuccess;
        }
        size_t want = bytes + bytes / 8;
        want = (want + 255) & ~(size_t)255;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = (bytes + 255) & ~(size_t)255;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(static_cast<char *>(p) - guard);
        p = nullptr;
        cap = 0;
        guard = 0;
    }
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct Slab {
    float *vec;
    int64_t *ids;
    uint32_t *tags;
};

// is `p` a device (or managed) pointer usable on `device`?  host otherwise
bool is_device_ptr(const void *p, int device) {
    if (p == nullptr) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (at.type == cudaMemoryTypeDevice) return at.device == device;
    return at.type == cudaMemoryTypeManaged;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// NVTX ranges (tracing).  One range per C-ABI call and, inside a search, one per phase -- the host-side
// enqueue of that phase, which a timeline tool (Nsight Systems) correlates with the kernels launched under it.  With no tool
// attached a push/pop is a null function-pointer check.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
struct NvtxPhases {  // consecutive phases of one call: next() closes the running phase and opens another
    bool open = false;
    void next(const char *name) {
        if (open) nvtxRangePop();
        open = name != nullptr;
        if (open) nvtxRangePushA(name);
    }
    ~NvtxPhases() {
        if (open) nvtxRangePop();
    }
};

// probe[q, j] = j : "every list", used when nprobe >= nlist (exhaustive search needs no ranking)
__global__ void iota_rows_kernel(int32_t *p, int64_t rows, int32_t n) {
    const int64_t total = rows * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = (int32_t)(i % n);
}

// bad[0] += bytes of the two guards of one buffer that no longer hold 0xA5
__global__ void check_guard_kernel(const unsigned char *front, const unsigned char *back, int n, unsigned int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && front[i] != 0xA5) atomicAdd(bad, 1u);
    if (i < n && back[i] != 0xA5) atomicAdd(bad + 1, 1u);
}

__global__ void fill_u32_kernel(uint32_t *p, int64_t n, uint32_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

}  // namespace

// Scratch of ONE call in flight.  Stream ordered: ev_done marks the end of the slot's last user, so the next user may be on
// another stream.
constexpr int kSlots = 2;
struct Scratch {
    DevBuf s_q, s_scores, s_probe, s_pageoff, s_cand, s_outd, s_outi, s_repobits;
    DevBuf s_x, s_xpad, s_ids, s_repo, s_lang, s_assign, s_best, s_pos, s_lenold, s_need, s_npg, s_needoff, s_bad;
    DevBuf s_sums, s_counts, s_obj, s_rows, s_rm, s_cnt, s_ahi, s_alo, s_lplan, s_scan, s_packed, s_qsplit, s_bstage, s_ptail;
    cudaEvent_t ev_done = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};  // fork/join streams of the list-major scan
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    uint32_t plan_epoch = 0;  // launch counter of the pair plan's look-back words (s_scan)
    bool busy = false, used = false;
    void *last_stream = nullptr;  // stream of the slot's last search
    unsigned long long last_use = 0;

    // profiling of the last search that ran in this slot
    std::vector<cudaEvent_t> prof_ev;  // 6 events per chunk: t0 coarse | select | plan | scan | topk t5
    unsigned long long *prof_rows = nullptr;  // device counter
    int64_t prof_pages = 0;
    int prof_scan_launches = 0, prof_total_launches = 0;

    template <class F>
    void each_buf(F f) {
        for (DevBuf *b : {&s_q, &s_scores, &s_probe, &s_pageoff, &s_cand, &s_outd, &s_outi, &s_repobits, &s_x, &s_xpad, &s_ids, &s_repo,
                          &s_lang, &s_assign, &s_best, &s_pos, &s_lenold, &s_need, &s_npg, &s_needoff, &s_bad, &s_sums, &s_counts, &s_obj,
                          &s_rows, &s_rm, &s_cnt, &s_ahi, &s_alo, &s_lplan, &s_scan, &s_packed, &s_qsplit, &s_bstage, &s_ptail})
            f(b);
    }
};

// the slot of the call running on this host thread (set by ReadGuard / WriteGuard for the duration of a C-ABI call)
static thread_local Scratch *tl_scr = nullptr;
static thread_local bool tl_writer = false;

// ------------------------------------------------------------------------------------------------
// the index object
// ------------------------------------------------------------------------------------------------
struct sc_index {
    int dim = 0, ds = 0, metric = 0, nlist = 0, device = 0, num_sms = 148;
    bool trained = false;
    float *centroids = nullptr;  // [nlist, ds]
    float *cnorm = nullptr;      // [nlist]  |c|^2 (L2 only)
    float *cent_hi = nullptr;    // [nlist, ds] tf32(c)           } 3xTF32 operands of the tcgen05
    float *cent_lo = nullptr;    // [nlist, ds] tf32(c - cent_hi) } coarse contraction (gemm_tc.cu)
    int coarse_impl = 0;         // 0 = tcgen05 3xTF32, 1 = fp32 SIMT (exact-fp32 reference kernel)
    int small_coarse = 1;        // batches of <= 16 rows use coarse_small_kernel
    int fuse_plan = 1;           // batches of <= 16 rows: probe selection + pair plan in one launch
    int pdl = 1;                 // ... and the step's kernels chained by programmatic dependent launch
    int tile_rem = 0;            // 4 = list-major: remainders of 5..16 queries become tcgen05 tile items (0: 9..16 only)
    int mq_fused = 0;            // 1 = list-major page scans (4-query and 8-query bucket) in ONE launch: measured slower, see scan_mq.cu
    int tc_variant = 0;          // fused argmax tile: 0 = 256x256 (64 B swizzle), 1 = 128x256 (128 B swizzle)

    // paged lists
    int slab_shift = 0;
    std::vector<Slab> slabs;
    SlabTable *h_tab = nullptr;  // host mirror
    SlabTable *d_tab = nullptr;
    uint8_t *d_maps = nullptr;   // [kMaxSlabs][128] TMA tensor maps of the slabs' vectors (scan_lists_ts.cu); nullptr = unavailable
    int32_t pool_top = 0;          // pages handed out
    int32_t *free_pages = nullptr; // pages returned by compaction; the LAST nfree entries are free (taken from the end)
    int32_t nfree = 0, free_cap = 0;
    int64_t add_chunk_rows = 0;    // tests: rows per add chunk (0 = automatic)
    int32_t fail_add_after = 0;    // tests: the n-th add chunk from now fails after its slots were claimed
    int32_t *list_len = nullptr;   // [nlist] slots used (incl. tombstones)
    int32_t *pt_off = nullptr;     // [nlist+1]
    int32_t *pt_off_alt = nullptr; // double buffer
    int32_t *pt = nullptr;
    int32_t *pt_alt = nullptr;
    int64_t pt_cap = 0, pt_alt_cap = 0;
    std::vector<int32_t> h_len;          // host mirror of list_len
    std::vector<int64_t> h_bound_prefix; // [nlist+1] pages of the j largest lists
    int32_t h_max_pages = 0;
    int64_t ntotal = 0, nremoved = 0;

    // per-call scratch lives in slots (struct Scratch below): searches take any free slot under the shared side of the
    // lock, so kSlots of them run at once on different streams; everything else runs alone on slot 0
    Scratch scr[kSlots];
    int64_t scratch_budget = (int64_t)8 << 30;  // search scratch ceiling (candidates dominate); sc_index_set_param("scratch_bytes")
    int scan_variant = 0;
    int lists_cfg = 0;  // list-major tile items: 0 = tcgen05 where it applies (scan_lists_ts.cu: list rows from tensor memory), 1 = FFMA tiles, 2 = FFMA tiles with 32-float stages, 3 / 5 = the shared-memory-operand tcgen05 kernels (scan_lists_tc.cu v1 / v2), 4 = 0 with the 8-query page scan on mma.sync
    int scan_mode = 0;  // 0 = auto, 1 = query-major (scan.cu), 2 = list-major (scan_lists.cu)
    int lists_fork = 0;  // measured slower on C2 (tile CTAs pin shared memory the page scan needs): off by default

    bool profiling = false;  // per-phase events of searches (kept in the slot the search ran in)
    unsigned long long use_clock = 0;
    int last_slot = 0;       // slot of the search that finished last: what sc_index_last_search_times reports

    // reader / writer lock with slot hand-out (host side); the device side is ordered by events: a search waits for its
    // slot's previous user and for the last writer, a writer waits for every slot and the last writer
    std::condition_variable cv;
    int readers = 0, writers_waiting = 0;
    bool writer = false;
    cudaEvent_t ev_write = nullptr;
    std::mutex mu;
};

// Host-side entry discipline of the C ABI.  WriteGuard: alone on the handle (waits for running searches, blocks new ones),
// slot 0.  ReadGuard: shared with other searches, each on its own free slot (waits while kSlots searches are in the library).
// Both only cover the time the call spends ENQUEUEING work; on the device the slots and the writers are ordered by events
// (begin_call / end_call).
struct WriteGuard {
    sc_index *ix;
    explicit WriteGuard(sc_index *i) : ix(i) {
        std::unique_lock<std::mutex> lk(ix->mu);
        ++ix->writers_waiting;
        ix->cv.wait(lk, [&] { return !ix->writer && ix->readers == 0; });
        --ix->writers_waiting;
        ix->writer = true;
        tl_scr = &ix->scr[0];
        tl_writer = true;
    }
    ~WriteGuard() {
        {
            std::lock_guard<std::mutex> lk(ix->mu);
            ix->writer = false;
        }
        tl_scr = nullptr;
        tl_writer = false;
        ix->cv.notify_all();
    }
    WriteGuard(const WriteGuard &) = delete;
    WriteGuard &operator=(const WriteGuard &) = delete;
};

struct ReadGuard {
    sc_index *ix;
    int slot = -1;
    ReadGuard(sc_index *i, void *stream) : ix(i) {
        std::unique_lock<std::mutex> lk(ix->mu);
        ix->cv.wait(lk, [&] {
            if (ix->writer || ix->writers_waiting > 0) return false;  // writers first: a stream of searches cannot starve an insert
            for (int k = 0; k < kSlots; ++k)
                if (!ix->scr[k].busy) return true;
            return false;
        });
        // A caller keeps the slot its stream used last (one stream: one slot, no second set of scratch buffers; one host
        // thread alternating between two streams: two slots, so its searches overlap on the device); otherwise the free
        // slot that has rested longest.
        int same = -1, oldest = -1;
        for (int k = 0; k < kSlots; ++k) {
            const Scratch &c = ix->scr[k];
            if (c.busy) continue;
            if (same < 0 && c.used && c.last_stream == stream) same = k;
            if (oldest < 0 || c.last_use < ix->scr[oldest].last_use) oldest = k;
        }
        slot = same >= 0 ? same : oldest;
        Scratch &sl = ix->scr[slot];
        sl.busy = true;
        sl.used = true;
        sl.last_stream = stream;
        sl.last_use = ++ix->use_clock;
        ++ix->readers;
        tl_scr = &sl;
        tl_writer = false;
    }
    ~ReadGuard() {
        {
            std::lock_guard<std::mutex> lk(ix->mu);
            ix->scr[slot].busy = false;
            --ix->readers;
            ix->last_slot = slot;
        }
        tl_scr = nullptr;
        ix->cv.notify_all();
    }
    ReadGuard(const ReadGuard &) = delete;
    ReadGuard &operator=(const ReadGuard &) = delete;
};

// ------------------------------------------------------------------------------------------------
// cross-GPU exchange (one per rank): peer-mapped symmetric buffers, no collective library on the data path
// ------------------------------------------------------------------------------------------------
// Layout of every rank's buffer (identical on all ranks):
//   [0, 512)      probe flags  [world] u64: flag p = last epoch whose probe rows rank p has stored here
//   [512, 1024)   top-k flags  [world] u64: flag p = last epoch whose partial top-k rank p has stored here
//   [4096, ...)   two halves (epoch parity), each: probe table [nq, nprobe] i32 | part_dist [world, nq, k] f32 |
//                 part_ids [world, nq, k] i64
// Double buffering by epoch parity is enough: a rank can only run one step ahead of its slowest peer, because
// finishing step s needs every peer's step-s partials, which a peer stores after it finished reading step s-1.
struct sc_exchange {
    int rank = 0, world = 1, device = 0;
    char *peer[kMaxPeers] = {nullptr};
    size_t bytes = 0;
    unsigned long long epoch = 0;
    unsigned int *done = nullptr;      // [2] CTA counters (probe select, top-k select)
    unsigned int *status = nullptr;    // device view of h_status
    unsigned int *h_status = nullptr;  // [1] pinned, mapped: 1 = a wait timed out (the waiting kernel stores it; the host
                                       //     reads it without synchronising anything)
    bool broken = false;               // a step failed after its epoch moved: ranks no longer agree
    unsigned long long timeout_ns = 10ull * 1000 * 1000 * 1000;
};
constexpr size_t kExHeader = 4096;

namespace {

struct ExLayout {
    size_t probes, part_d, part_i, end;  // byte offsets from the start of the buffer
};
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
// offsets of the three regions for this epoch's half; `end` > bytes means "does not fit"
ExLayout exchange_layout(const sc_exchange *ex, int64_t nq, int np, int k) {
    const size_t half = ((ex->bytes - kExHeader) / 2) & ~(size_t)255;
    ExLayout L;
    L.probes = kExHeader + (size_t)(ex->epoch & 1ull) * half;
    L.part_d = L.probes + align256((size_t)nq * np * 4);
    L.part_i = L.part_d + align256((size_t)ex->world * nq * k * 4);
    L.end = L.part_i + align256((size_t)ex->world * nq * k * 8);
    if (L.end - L.probes > half) L.end = ex->bytes + 1;
    return L;
}
PeerSignal make_signal(const sc_exchange *ex, int kind) {
    PeerSignal s;
    memset(&s, 0, sizeof(s));
    s.world = ex->world;
    s.done = ex->done + kind;
    s.epoch = ex->epoch;
    for (int p = 0; p < ex->world; ++p)
        s.flag[p] = reinterpret_cast<unsigned long long *>(ex->peer[p] + 512 * kind) + ex->rank;
    return s;
}
PeerWait make_wait(const sc_exchange *ex, int kind) {
    PeerWait w;
    memset(&w, 0, sizeof(w));
    w.world = ex->world;
    w.flags = reinterpret_cast<const unsigned long long *>(ex->peer[ex->rank] + 512 * kind);
    w.epoch = ex->epoch;
    w.status = ex->status;
    w.timeout_ns = ex->timeout_ns;
    return w;
}

size_t page_bytes(const sc_index *ix) { return (size_t)kPageRows * ix->ds * sizeof(float); }

int ensure_slabs(sc_index *ix, int64_t pages_needed, cudaStream_t st) {
    const int64_t pps = (int64_t)1 << ix->slab_shift;
    bool grew = false;
    while ((int64_t)ix->slabs.size() * pps < pages_needed) {
        if ((int)ix->slabs.size() >= kMaxSlabs) return fail(SC_ERR_OOM, "slab table full (%d slabs)", kMaxSlabs);
        Slab s{nullptr, nullptr, nullptr};
        const int64_t rows = pps * kPageRows;
        cudaError_t e = cudaMalloc(&s.vec, (size_t)rows * ix->ds * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&s.ids, (size_t)rows * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMalloc(&s.tags, (size_t)rows * sizeof(uint32_t));
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (s.vec) cudaFree(s.vec);
            if (s.ids) cudaFree(s.ids);
            if (s.tags) cudaFree(s.tags);
            return fail(SC_ERR_OOM, "cannot allocate list slab %zu (%lld rows x %d floats): %s", ix->slabs.size(),
                        (long long)rows, ix->ds, cudaGetErrorString(e));
        }
        // unused slots never match any predicate
        fill_u32_kernel<<<ix->num_sms * 4, 256, 0, st>>>(s.tags, rows, 0xFFFFFFFFu);
        CU(cudaGetLastError());
        const int i = (int)ix->slabs.size();
        ix->h_tab->vec[i] = s.vec;
        ix->h_tab->ids[i] = s.ids;
        ix->h_tab->tags[i] = s.tags;
        ix->slabs.push_back(s);
        grew = true;
        if (ix->d_maps && ix->ds % 32 == 0) {  // tensor map of the slab for the TMA-fed tile kernel
            alignas(64) uint8_t m[128];
            if (encode_slab_map(m, s.vec, rows, ix->ds) == cudaSuccess) {
                CU(cudaMemcpyAsync(ix->d_maps + (size_t)i * 128, m, 128, cudaMemcpyHostToDevice, st));
                CU(cudaStreamSynchronize(st));  // `m` dies at scope exit
            } else {
                cudaGetLastError();
                cudaFree(ix->d_maps);  // no driver entry point: the shared-memory tile kernel keeps serving
                ix->d_maps = nullptr;
            }
        }
    }
    if (grew) CU(cudaMemcpyAsync(ix->d_tab, ix->h_tab, sizeof(SlabTable), cudaMemcpyHostToDevice, st));
    return SC_OK;
}

void free_lists(sc_index *ix) {
    for (auto &s : ix->slabs) {
        cudaFree(s.vec);
        cudaFree(s.ids);
        cudaFree(s.tags);
    }
    ix->slabs.clear();
    ix->pool_top = 0;
    ix->nfree = 0;
    ix->ntotal = 0;
    ix->nremoved = 0;
}

void refresh_bounds(sc_index *ix) {
    std::vector<int32_t> pages(ix->nlist);
    int32_t mx = 0;
    for (int i = 0; i < ix->nlist; ++i) {
        pages[i] = (ix->h_len[i] + kPageRows - 1) / kPageRows;
        mx = std::max(mx, pages[i]);
    }
    std::sort(pages.begin(), pages.end(), [](int32_t a, int32_t b) { return a > b; });
    ix->h_bound_prefix.assign(ix->nlist + 1, 0);
    for (int i = 0; i < ix->nlist; ++i) ix->h_bound_prefix[i + 1] = ix->h_bound_prefix[i] + pages[i];
    ix->h_max_pages = mx;
}

// device pointer to `count` elements of T: `p` itself when it already lives on the device,
// otherwise a stream-ordered copy in `buf`
template <typename T>
int stage(sc_index *ix, const T *p, size_t count, DevBuf &buf, cudaStream_t st, const T **out) {
    if (p == nullptr) {
        *out = nullptr;
        return SC_OK;
    }
    if (is_device_ptr(p, ix->device)) {
        *out = p;
        return SC_OK;
    }
    CU(buf.reserve(count * sizeof(T)));
    CU(cudaMemcpyAsync(buf.p, p, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *out = buf.as<T>();
    return SC_OK;
}

// rows [n, dim] (host or device) -> device rows [n, ds], 16-byte aligned
int stage_rows(sc_index *ix, const float *x, int64_t n, DevBuf &raw, DevBuf &padded, cudaStream_t st,
               const float **out) {
    const float *xd = nullptr;
    SC(stage(ix, x, (size_t)n * ix->dim, raw, st, &xd));
    if (ix->ds == ix->dim && ((uintptr_t)xd & 15) == 0) {
        *out = xd;
        return SC_OK;
    }
    CU(padded.reserve((size_t)n * ix->ds * sizeof(float)));
    CU(launch_pad_rows(xd, n, ix->dim, ix->ds, padded.as<float>(), st));
    *out = padded.as<float>();
    return SC_OK;
}

int begin_call(sc_index *ix, cudaStream_t st) {
    // order this call on the device after the previous user of its slot and after the last writer (a writer: after every
    // slot), even when the caller switched streams
    if (tl_writer) {
        for (int k = 0; k < kSlots; ++k) CU(cudaStreamWaitEvent(st, ix->scr[k].ev_done, 0));
    } else {
        CU(cudaStreamWaitEvent(st, tl_scr->ev_done, 0));
    }
    CU(cudaStreamWaitEvent(st, ix->ev_write, 0));
    return SC_OK;
}

int end_call(sc_index *ix, cudaStream_t st) {
    CU(cudaEventRecord(tl_scr->ev_done, st));
    if (tl_writer) CU(cudaEventRecord(ix->ev_write, st));
    return SC_OK;
}

int require_trained(const sc_index *ix) {
    if (!ix->trained) return fail(SC_ERR_STATE, "index has no centroids: call sc_index_train or sc_index_set_centroids first");
    return SC_OK;
}

// everything derived from the centroids: |c|^2 and the tf32 hi/lo split
int update_cnorm(sc_index *ix, cudaStream_t st) {
    CU(launch_row_norms(ix->centroids, ix->nlist, ix->ds, ix->cnorm, st));
    CU(launch_split_tf32(ix->centroids, (int64_t)ix->nlist * ix->ds, ix->cent_hi, ix->cent_lo, st));
    return SC_OK;
}

bool use_tc(const sc_index *ix) { return ix->coarse_impl == 0 && ix->ds >= 32; }
// kernels coarse_scores launches for m rows (the launch count reported with the search times)
int coarse_launches(const sc_index *ix, int64_t m) { return (m <= 16 && ix->small_coarse) ? 1 : (use_tc(ix) ? 2 : 1); }

// scores[m, nlist] = similarity to maximise (IP: x.c ; L2: 2 x.c - |c|^2) of device rows xd[m, ds]
int coarse_scores(sc_index *ix, const float *xd, int64_t m, float *scores, cudaStream_t st) {
    const bool l2 = ix->metric == SC_METRIC_L2;
    if (m <= 16 && ix->small_coarse) {  // tiny batches: stream the centroid table once, exact fp32
        const cudaError_t e = launch_coarse_small(xd, m, ix->centroids, ix->nlist, ix->ds, l2 ? 2.f : 1.f,
                                                  l2 ? ix->cnorm : nullptr, scores, ix->num_sms, st);
        if (e == cudaSuccess) return SC_OK;
        if (e != cudaErrorNotSupported) CU(e);
        cudaGetLastError();
    }
    if (!use_tc(ix)) {
        CU(launch_gemm_nt(xd, m, ix->centroids, ix->nlist, ix->ds, l2 ? ix->cnorm : nullptr, scores, st));
        return SC_OK;
    }
    CU(tl_scr->s_ahi.reserve((size_t)m * ix->ds * 4));
    CU(tl_scr->s_alo.reserve((size_t)m * ix->ds * 4));
    CU(launch_split_tf32(xd, m * ix->ds, tl_scr->s_ahi.as<float>(), tl_scr->s_alo.as<float>(), st));
    CU(launch_gemm_tc_scores(tl_scr->s_ahi.as<float>(), tl_scr->s_alo.as<float>(), m, ix->cent_hi, ix->cent_lo, ix->nlist, ix->ds,
                             l2 ? 2.f : 1.f, l2 ? ix->cnorm : nullptr, scores, ix->num_sms, st));
    return SC_OK;
}

// rows per coarse chunk so that the [rows, nlist] similarity tile stays within ~1/4 of the budget
int64_t coarse_chunk_rows(const sc_index *ix, int64_t n) {
    const int64_t budget = std::max<int64_t>(ix->scratch_budget / 4, (int64_t)64 << 20);
    int64_t rows = budget / ((int64_t)ix->nlist * 4);
    rows = std::max<int64_t>(128, (rows / 128) * 128);
    return std::min<int64_t>(rows, std::max<int64_t>(n, 1));
}

// assign[i] = argbest centroid of xd[i] (device rows [n, ds]); best[i] = its similarity
int coarse_assign(sc_index *ix, const float *xd, int64_t n, int32_t *assign, float *best, cudaStream_t st) {
    if (use_tc(ix) && n >= 4096) {
        // fused contraction + argmax on the tensor cores: the [n, nlist] matrix is never written.
        // Centroids are swept in slabs whose hi/lo copies (~48 MB) stay L2-resident while every row
        // tile passes; the running best is carried across slabs in best/assign.
        const bool l2 = ix->metric == SC_METRIC_L2;
        CU(tl_scr->s_ahi.reserve((size_t)n * ix->ds * 4));
        CU(tl_scr->s_alo.reserve((size_t)n * ix->ds * 4));
        CU(launch_split_tf32(xd, n * ix->ds, tl_scr->s_ahi.as<float>(), tl_scr->s_alo.as<float>(), st));
        float *bv = best;
        if (!bv) {
            CU(tl_scr->s_best.reserve((size_t)n * 4));
            bv = tl_scr->s_best.as<float>();
        }
        int64_t slab = (((int64_t)48 << 20) / ((int64_t)ix->ds * 8)) / 256 * 256;
        slab = std::max<int64_t>(256, std::min<int64_t>(slab, ix->nlist));
        unsigned long long *packed = nullptr;
        if (ix->tc_variant == 0) {  // 256x256 tiles: per-row best merged across CTAs and slabs with atomicMax
            CU(tl_scr->s_packed.reserve((size_t)n * 8));
            packed = tl_scr->s_packed.as<unsigned long long>();
            CU(cudaMemsetAsync(packed, 0, (size_t)n * 8, st));
        }
        for (int64_t c0 = 0; c0 < ix->nlist; c0 += slab) {
            const int nc = (int)std::min<int64_t>(slab, ix->nlist - c0);
            CU(launch_gemm_tc_argmax(tl_scr->s_ahi.as<float>(), tl_scr->s_alo.as<float>(), n, ix->cent_hi + c0 * ix->ds,
                                     ix->cent_lo + c0 * ix->ds, nc, ix->ds, l2 ? 2.f : 1.f, l2 ? ix->cnorm + c0 : nullptr, bv,
                                     assign, (int)c0, c0 > 0 ? 1 : 0, ix->num_sms, packed, st));
        }
        if (packed) CU(launch_unpack_argmax(packed, n, bv, assign, st));
        return SC_OK;
    }
    const int64_t ch = coarse_chunk_rows(ix, n);
    CU(tl_scr->s_scores.reserve((size_t)ch * ix->nlist * sizeof(float)));
    for (int64_t s = 0; s < n; s += ch) {
        const int64_t m = std::min(ch, n - s);
        SC(coarse_scores(ix, xd + s * ix->ds, m, tl_scr->s_scores.as<float>(), st));
        CU(launch_argmax_rows(tl_scr->s_scores.as<float>(), m, ix->nlist, assign + s, best ? best + s : nullptr, st));
    }
    return SC_OK;
}

// append rows that already sit on the device ([n, ds] rows, device ids/tags/lists)
int add_device_rows_unguarded(sc_index *ix, const float *xd, const int64_t *ids_d, const uint32_t *repo_d, const uint8_t *lang_d,
                              const int32_t *lists_d, int64_t n, cudaStream_t st, bool *bumped) {
    const int nlist = ix->nlist;
    CU(tl_scr->s_pos.reserve((size_t)n * 4));
    CU(tl_scr->s_lenold.reserve((size_t)nlist * 4));
    CU(tl_scr->s_need.reserve((size_t)nlist * 4));
    CU(tl_scr->s_npg.reserve((size_t)nlist * 4));
    CU(tl_scr->s_needoff.reserve((size_t)(nlist + 1) * 4));
    CU(tl_scr->s_bad.reserve(16));
    CU(cudaMemcpyAsync(tl_scr->s_lenold.p, ix->list_len, (size_t)nlist * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(tl_scr->s_bad.p, 0, 16, st));
    *bumped = true;  // from here on the device list lengths are ahead of the page table until the scatter is queued
    CU(launch_count_positions(lists_d, repo_d, n, nlist, ix->list_len, tl_scr->s_pos.as<int32_t>(), tl_scr->s_bad.as<int32_t>(), st));
    CU(launch_page_need(tl_scr->s_lenold.as<int32_t>(), ix->list_len, nlist, tl_scr->s_need.as<int32_t>(),
                        tl_scr->s_npg.as<int32_t>(), st));
    CU(launch_exclusive_scan_i32(tl_scr->s_need.as<int32_t>(), nlist, tl_scr->s_needoff.as<int32_t>(), st));
    CU(launch_exclusive_scan_i32(tl_scr->s_npg.as<int32_t>(), nlist, ix->pt_off_alt, st));
    int32_t h_new = 0, h_total = 0, h_bad[2] = {0, 0};
    CU(cudaMemcpyAsync(&h_new, tl_scr->s_needoff.as<int32_t>() + nlist, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&h_total, ix->pt_off_alt + nlist, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_bad, tl_scr->s_bad.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h_bad[0] != 0) return fail(SC_ERR_INVALID, "%d rows carry a list id outside [0, %d)", h_bad[0], nlist);
    if (h_bad[1] != 0) return fail(SC_ERR_INVALID, "%d rows carry a repo tag above %u", h_bad[1], kTagRepoMax);
    if (ix->fail_add_after > 0 && --ix->fail_add_after == 0)  // tests: a device allocation failure in a later chunk
        return fail(SC_ERR_OOM, "injected allocation failure (fail_add_after)");
    // pages come from the free list first (compaction returns pages there), then from the top of the pool
    const int32_t from_free = std::min<int32_t>(h_new, ix->nfree);
    const int32_t from_top = h_new - from_free;
    if ((int64_t)ix->pool_top + from_top > (int64_t)INT32_MAX / 2) return fail(SC_ERR_OOM, "page id space exhausted");
    SC(ensure_slabs(ix, (int64_t)ix->pool_top + from_top, st));
    if (h_total > ix->pt_alt_cap) {
        const int64_t want = std::max<int64_t>((int64_t)h_total + h_total / 2, 1024);
        if (ix->pt_alt) cudaFree(ix->pt_alt);
        ix->pt_alt = nullptr;
        ix->pt_alt_cap = 0;
        CU(cudaMalloc(&ix->pt_alt, (size_t)want * 4));
        ix->pt_alt_cap = want;
    }
    CU(launch_rebuild_pt(ix->pt_off, ix->pt, ix->pt_off_alt, ix->pt_alt, tl_scr->s_needoff.as<int32_t>(), ix->pool_top,
                         ix->free_pages ? ix->free_pages + (ix->nfree - from_free) : nullptr, from_free, nlist, st));
    CU(launch_scatter_rows(xd, ids_d, repo_d, lang_d, lists_d, tl_scr->s_pos.as<int32_t>(), n, ix->ds, ix->pt_off_alt, ix->pt_alt,
                           ix->d_tab, ix->slab_shift, st));
    // committed: every launch is queued, only now does the host state move
    std::swap(ix->pt, ix->pt_alt);
    std::swap(ix->pt_cap, ix->pt_alt_cap);
    std::swap(ix->pt_off, ix->pt_off_alt);
    ix->pool_top += from_top;
    ix->nfree -= from_free;
    ix->ntotal += n;
    *bumped = false;
    return SC_OK;
}

// A failure between the slot claim (count_positions bumps the device list lengths) and the scatter must not leave
// the lengths ahead of the page table: plan_pairs sizes the scan from the DEVICE lengths, so candidates would be
// written past the scratch the host sized from its own mirror.  Roll the lengths back on every such exit.
int add_device_rows(sc_index *ix, const float *xd, const int64_t *ids_d, const uint32_t *repo_d, const uint8_t *lang_d,
                    const int32_t *lists_d, int64_t n, cudaStream_t st) {
    bool bumped = false;
    const int rc = add_device_rows_unguarded(ix, xd, ids_d, repo_d, lang_d, lists_d, n, st, &bumped);
    if (rc != SC_OK && bumped) {
        const std::string why = g_err;
        cudaGetLastError();
        if (cudaMemcpyAsync(ix->list_len, tl_scr->s_lenold.p, (size_t)ix->nlist * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            cudaGetLastError();
            return fail(rc, "%s; AND the list lengths could not be restored: reset the index", why.c_str());
        }
        g_err = why;
    }
    return rc;
}

int sync_host_lengths(sc_index *ix, cudaStream_t st) {
    CU(cudaMemcpyAsync(ix->h_len.data(), ix->list_len, (size_t)ix->nlist * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    refresh_bounds(ix);
    return SC_OK;
}

int add_chunks(sc_index *ix, const float *x, const int64_t *ids, const uint32_t *repo, const uint8_t *lang,
               const int32_t *lists, int64_t n, cudaStream_t st) {
    // bounded staging: ~256 MB of rows per pass
    int64_t chunk = std::max<int64_t>(1024, ((int64_t)256 << 20) / ((int64_t)ix->ds * 4));
    if (is_device_ptr(x, ix->device)) chunk = std::max<int64_t>(chunk, (int64_t)1 << 18);
    if (ix->add_chunk_rows > 0) chunk = ix->add_chunk_rows;
    chunk = std::min(chunk, n);
    for (int64_t s = 0; s < n; s += chunk) {
        const int64_t m = std::min(chunk, n - s);
        const float *xd = nullptr;
        const int64_t *ids_d = nullptr;
        const uint32_t *repo_d = nullptr;
        const uint8_t *lang_d = nullptr;
        const int32_t *lists_d = nullptr;
        SC(stage_rows(ix, x + s * ix->dim, m, tl_scr->s_x, tl_scr->s_xpad, st, &xd));
        SC(stage(ix, ids + s, (size_t)m, tl_scr->s_ids, st, &ids_d));
        SC(stage(ix, repo ? repo + s : nullptr, (size_t)m, tl_scr->s_repo, st, &repo_d));
        SC(stage(ix, lang ? lang + s : nullptr, (size_t)m, tl_scr->s_lang, st, &lang_d));
        if (lists) {
            SC(stage(ix, lists + s, (size_t)m, tl_scr->s_assign, st, &lists_d));
        } else {
            CU(tl_scr->s_assign.reserve((size_t)m * 4));
            SC(coarse_assign(ix, xd, m, tl_scr->s_assign.as<int32_t>(), nullptr, st));
            lists_d = tl_scr->s_assign.as<int32_t>();
        }
        SC(add_device_rows(ix, xd, ids_d, repo_d, lang_d, lists_d, m, st));
    }
    return SC_OK;
}

int add_impl(sc_index *ix, const float *x, const int64_t *ids, const uint32_t *repo, const uint8_t *lang,
             const int32_t *lists, int64_t n, cudaStream_t st) {
    if (n < 0) return fail(SC_ERR_INVALID, "n < 0");
    if (n == 0) return SC_OK;
    if (x == nullptr || ids == nullptr) return fail(SC_ERR_INVALID, "x and ids must not be NULL");
    SC(require_trained(ix));
    CU(cudaDeviceSynchronize());  // no search may still be reading the page table we are about to swap
    const int rc = add_chunks(ix, x, ids, repo, lang, lists, n, st);
    // the chunks that went in before a failure stay committed: the host mirror of the list lengths (scratch sizing of
    // the searches) must follow them on EVERY exit
    const std::string why = g_err;
    const int rs = sync_host_lengths(ix, st);
    if (rc != SC_OK) {
        g_err = why;
        return rc;
    }
    return rs;
}

int build_filter(sc_index *ix, const sc_filter_t *filt, cudaStream_t st, FilterDev *out) {
    FilterDev f;
    memset(&f, 0, sizeof(f));
    if (filt != nullptr) {
        if (filt->n_langs < 0 || filt->n_repos < 0) return fail(SC_ERR_INVALID, "negative filter length");
        if (filt->n_langs > 0) {
            if (!filt->lang_tags) return fail(SC_ERR_INVALID, "filter lang_tags is NULL");
            f.flags |= 1u;
            for (int i = 0; i < filt->n_langs; ++i) f.lang_bits[filt->lang_tags[i] >> 5] |= 1u << (filt->lang_tags[i] & 31);
        }
        if (filt->n_repos > 0) {
            if (!filt->repo_tags) return fail(SC_ERR_INVALID, "filter repo_tags is NULL");
            f.flags |= 2u;
            uint32_t mx = 0;
            for (int i = 0; i < filt->n_repos; ++i) {
                if (filt->repo_tags[i] > kTagRepoMax) return fail(SC_ERR_INVALID, "repo tag %u exceeds %u", filt->repo_tags[i], kTagRepoMax);
                mx = std::max(mx, filt->repo_tags[i]);
            }
            const uint32_t nbits = mx + 1;
            std::vector<uint32_t> bits((nbits + 31) / 32, 0u);
            for (int i = 0; i < filt->n_repos; ++i) bits[filt->repo_tags[i] >> 5] |= 1u << (filt->repo_tags[i] & 31);
            CU(tl_scr->s_repobits.reserve(bits.size() * 4));
            CU(cudaMemcpyAsync(tl_scr->s_repobits.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));  // `bits` dies at scope exit
            f.repo_bits = tl_scr->s_repobits.as<uint32_t>();
            f.n_repo_bits = nbits;
        }
    }
    *out = f;
    return SC_OK;
}

void clear_prof(Scratch *sl) {
    for (auto e : sl->prof_ev) cudaEventDestroy(e);
    sl->prof_ev.clear();
    sl->prof_scan_launches = 0;
    sl->prof_total_launches = 0;
}

int prof_mark(sc_index *ix, cudaStream_t st) {
    if (!ix->profiling) return SC_OK;
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    CU(cudaEventRecord(e, st));
    tl_scr->prof_ev.push_back(e);
    return SC_OK;
}

// ex != nullptr: this index is one shard of a row-sharded index.  The coarse pass (when `lists` is NULL) is split
// over the ranks and the probe rows are stored into every peer's table; the top-k epilogue stores this rank's
// partial result into every peer's gather slot and the merge kernel waits for the peers' flags (select.cu).
int search_impl(sc_index *ix, const float *q, int64_t nq, int k, int nprobe, const int32_t *lists,
                const sc_filter_t *filt, float *out_dist, int64_t *out_ids, cudaStream_t st, sc_exchange *ex = nullptr) {
    if (nq < 0) return fail(SC_ERR_INVALID, "nq < 0");
    if (k < 1 || k > kMaxK) return fail(SC_ERR_INVALID, "k must be in [1, %d]", kMaxK);
    if (nprobe < 1) return fail(SC_ERR_INVALID, "nprobe must be >= 1");
    SC(require_trained(ix));
    if (nq == 0 && !ex) return SC_OK;
    if (nq == 0) return fail(SC_ERR_INVALID, "an exchange step needs nq >= 1 on every rank");
    if (!q || !out_dist || !out_ids) return fail(SC_ERR_INVALID, "q / out_dist / out_ids must not be NULL");
    const int np = lists ? nprobe : std::min(nprobe, ix->nlist);
    const bool all_lists = !lists && np == ix->nlist;
    if (!lists && !all_lists && np > kMaxK)
        return fail(SC_ERR_INVALID, "nprobe %d: values above %d are supported only as nprobe >= nlist (exhaustive)", np, kMaxK);
    const bool outd_dev = is_device_ptr(out_dist, ix->device), outi_dev = is_device_ptr(out_ids, ix->device);

    SC(begin_call(ix, st));
    FilterDev fdev;
    SC(build_filter(ix, filt, st, &fdev));
    tl_scr->prof_scan_launches = 0;  // launch counts describe the last call, profiled or not
    tl_scr->prof_total_launches = 0;
    if (ix->profiling) {
        clear_prof(tl_scr);
        if (!tl_scr->prof_rows) CU(cudaMalloc(&tl_scr->prof_rows, 16));
        CU(cudaMemsetAsync(tl_scr->prof_rows, 0, 16, st));
    }

    // worst-case pages one query can touch -> candidate scratch per query
    int64_t pb = lists ? (int64_t)np * ix->h_max_pages : ix->h_bound_prefix[std::min(np, ix->nlist)];
    pb = std::max<int64_t>(pb, 1);
    const int64_t per_query = ((lists || all_lists) ? 0 : (int64_t)ix->nlist * 4) + (int64_t)np * 20 + pb * kPageRows * 4 +
                              (int64_t)ix->ds * 4 + (int64_t)k * 12;
    int64_t nqc = std::max<int64_t>(1, ix->scratch_budget / per_query);
    ExLayout exl{};
    if (ex) {
        if (ex->device != ix->device) return fail(SC_ERR_INVALID, "exchange and index live on different devices");
        // an exchange step is ONE pass on every rank (the ranks' list lengths differ, so a per-rank chunking
        // rule would desynchronise them): the scratch budget is not applied here, the caller bounds nq
        nqc = nq;
        exl = exchange_layout(ex, nq, np, k);
        if (exl.end > ex->bytes)  // depends on (nq, nprobe, k, world) only: every rank fails alike, before the epoch moves
            return fail(SC_ERR_INVALID, "exchange buffer of %zu bytes is too small for nq=%lld nprobe=%d k=%d world=%d",
                        ex->bytes, (long long)nq, np, k, ex->world);
    }
    if (nqc >= nq) {
        nqc = nq;  // one pass
    } else {
        // equal passes: the list-major scan amortises a list over the queries of ONE pass
        const int64_t passes = (nq + nqc - 1) / nqc;
        nqc = (nq + passes - 1) / passes;
        if (!lists && nqc >= 128) nqc = std::min<int64_t>(((nqc + 127) / 128) * 128, nq);  // whole GEMM tiles
    }

    const int64_t npairs_max = nqc * np;
    if (!lists && !all_lists) CU(tl_scr->s_scores.reserve((size_t)nqc * ix->nlist * 4));
    if (!lists) CU(tl_scr->s_probe.reserve((size_t)npairs_max * 4));
    CU(tl_scr->s_pageoff.reserve((size_t)(npairs_max + 1) * 8));
    {   // look-back words of the pair plan: zero when (re)allocated and when the 22-bit epoch wraps
        const unsigned before = tl_scr->s_scan.gen;
        CU(tl_scr->s_scan.reserve(plan_pairs_look_words(npairs_max) * 8));
        if (tl_scr->s_scan.gen != before) {
            CU(cudaMemsetAsync(tl_scr->s_scan.p, 0, tl_scr->s_scan.cap, st));
            tl_scr->plan_epoch = 0;
        }
    }
    CU(tl_scr->s_cand.reserve((size_t)nqc * pb * kPageRows * 4));
    if (!outd_dev) CU(tl_scr->s_outd.reserve((size_t)nqc * k * 4));
    if (!outi_dev) CU(tl_scr->s_outi.reserve((size_t)nqc * k * 8));
    if (ex) {
        // The step's epoch moves only now, after the large scratch reservations: a rank that fails above (out of memory)
        // has not published anything and its peers time out cleanly.  A failure further down leaves this rank's epoch ahead
        // of what it published: sc_index_search_sharded marks the exchange broken so that the next call fails loudly.
        ex->epoch += 1;  // every rank makes the same calls, so the epochs agree
        exl = exchange_layout(ex, nq, np, k);
    }

    for (int64_t s = 0; s < nq; s += nqc) {
        const int64_t m = std::min(nqc, nq - s);
        const int64_t npairs = m * np;
        const float *qd = nullptr;
        SC(stage_rows(ix, q + s * ix->dim, m, tl_scr->s_q, tl_scr->s_xpad, st, &qd));
        const int32_t *probe = nullptr;
        NvtxPhases nv;
        nv.next("search:coarse");
        bool planned = false;  // the pair plan came with the probe selection
        // small batches without per-phase events: each kernel of the step is launched as a programmatic dependent of the one
        // before it (resident early, blocked in griddepcontrol.wait), which hides the launch latency of the 4-kernel chain
        const bool pdl = ix->pdl && !ix->profiling;
        SC(prof_mark(ix, st));
        if (lists) {
            SC(stage(ix, lists + s * np, (size_t)npairs, tl_scr->s_probe, st, &probe));
            SC(prof_mark(ix, st));
        } else if (all_lists) {
            SC(prof_mark(ix, st));
            iota_rows_kernel<<<ix->num_sms * 4, 256, 0, st>>>(tl_scr->s_probe.as<int32_t>(), m, np);
            CU(cudaGetLastError());
            probe = tl_scr->s_probe.as<int32_t>();
            tl_scr->prof_total_launches += 1;
        } else if (ex) {
            // this rank ranks the centroids for its 1/world of the batch and stores the rows into every peer's table
            const int64_t per = (m + ex->world - 1) / ex->world;
            const int64_t lo = std::min<int64_t>(m, (int64_t)ex->rank * per), hi = std::min<int64_t>(m, lo + per);
            if (hi > lo) SC(coarse_scores(ix, qd + lo * ix->ds, hi - lo, tl_scr->s_scores.as<float>(), st));
            SC(prof_mark(ix, st));
            nv.next("search:select+exchange");
            PeerRows rows;
            memset(&rows, 0, sizeof(rows));
            for (int p = 0; p < ex->world; ++p) rows.p[p] = reinterpret_cast<int32_t *>(ex->peer[p] + exl.probes);
            rows.row0 = lo;
            CU(launch_select_rows_peers(tl_scr->s_scores.as<float>(), hi - lo, ix->nlist, np, rows, make_signal(ex, 0), st));
            CU(launch_peer_wait(make_wait(ex, 0), st));
            probe = reinterpret_cast<const int32_t *>(ex->peer[ex->rank] + exl.probes);
            tl_scr->prof_total_launches += (hi > lo ? coarse_launches(ix, hi - lo) : 0) + 1 + 1;  // coarse, select + scatter, wait
        } else if (m <= kPlanTailMaxQ && np <= 128 && ix->fuse_plan) {
            // small batches: probe selection and pair plan share one launch
            SC(coarse_scores(ix, qd, m, tl_scr->s_scores.as<float>(), st));
            SC(prof_mark(ix, st));
            nv.next("search:select+plan");
            {
                const unsigned before = tl_scr->s_ptail.gen;
                CU(tl_scr->s_ptail.reserve((size_t)kPlanTailWords * 8));
                if (tl_scr->s_ptail.gen != before) CU(cudaMemsetAsync(tl_scr->s_ptail.p, 0, tl_scr->s_ptail.cap, st));
            }
            CU(launch_select_rows_plan(tl_scr->s_scores.as<float>(), m, ix->nlist, np, tl_scr->s_probe.as<int32_t>(), ix->list_len, ix->nlist,
                                       tl_scr->s_pageoff.as<int64_t>(), tl_scr->s_ptail.as<unsigned long long>(),
                                       ix->profiling ? tl_scr->prof_rows : nullptr, st, pdl));
            probe = tl_scr->s_probe.as<int32_t>();
            tl_scr->prof_total_launches += coarse_launches(ix, m) + 1 - 1;  // select + plan in one; no separate plan launch below
            planned = true;
        } else {
            SC(coarse_scores(ix, qd, m, tl_scr->s_scores.as<float>(), st));
            SC(prof_mark(ix, st));
            nv.next("search:select");
            CU(launch_select_rows(tl_scr->s_scores.as<float>(), m, ix->nlist, np, tl_scr->s_probe.as<int32_t>(), nullptr, st));
            probe = tl_scr->s_probe.as<int32_t>();
            tl_scr->prof_total_launches += coarse_launches(ix, m) + 1;
        }
        SC(prof_mark(ix, st));
        nv.next("search:plan");
        if (!planned) {
            if (((tl_scr->plan_epoch + 1) & 0x3fffffu) == 0) {  // epochs 1 .. 2^22 - 2, then start over on zeroed words
                CU(cudaMemsetAsync(tl_scr->s_scan.p, 0, tl_scr->s_scan.cap, st));
                tl_scr->plan_epoch = 0;
            }
            CU(launch_plan_pairs(probe, npairs, ix->list_len, ix->nlist, tl_scr->s_pageoff.as<int64_t>(),
                                 tl_scr->s_scan.as<unsigned long long>(), ++tl_scr->plan_epoch, ix->profiling ? tl_scr->prof_rows : nullptr, st));
        }
        SC(prof_mark(ix, st));
        nv.next("search:scan");
        ScanArgs a;
        memset(&a, 0, sizeof(a));
        a.q = qd;
        a.ds = ix->ds;
        a.metric = ix->metric;
        a.nprobe = np;
        a.npairs = npairs;
        a.probe = probe;
        a.page_off = tl_scr->s_pageoff.as<int64_t>();
        a.list_len = ix->list_len;
        a.pt_off = ix->pt_off;
        a.pt = ix->pt;
        a.slabs = ix->d_tab;
        a.slab_shift = ix->slab_shift;
        a.cand = tl_scr->s_cand.as<float>();
        a.max_cand = pb * kPageRows;
        a.filt = fdev;
        a.slab_maps = ix->d_maps;
        // large batches re-probe the same lists: read each list once and score it against all its queries
        // (auto: when a list is probed 0.5x or more on average (0.25x up to dim 1024) -- 79 % or fewer of the pair passes hit
        //  a distinct list -- and lists hold at least a page.  Measured list-major / query-major step time at 0.125x / 0.25x /
        //  0.5x / 1x / 2x / 4x on C2 (dim 768): 1.01 / 0.95 / 0.85 / 0.69 / 0.48 / 0.32
        //  (profiles/r1_scan_mode_crossover.md); at dim 2048, where the 4-query page scan needs four slices, 0.25x is
        //  a loss (1.11) and 0.5x a gain (0.82).  With a scalar filter few rows per page are live and the scans are
        //  bound by per-page work rather than bytes: 10M x 2048, 5 % selectivity, whole step list-major / query-major
        //  at 0.5x: 0.92 / 0.70 ms, at 1x: 1.42 / 1.52 ms -> 0.75x, the threshold of the first version)
        const bool long_lists = ix->ntotal + ix->nremoved >= (int64_t)kPageRows * ix->nlist;
        // (round 2: 0.25x up to dim 1024 -- 0.95 on the iid set, and on clustered data, where the probes of a batch pile up on
        //  the hot lists, far better: 10M x 768 clustered, nq 256 / nprobe 16 = 0.25x: 5.93 ms query-major, nprobe 32: 4.01 ms
        //  list-major)
        //  ... and 0.125x for batches of more than 16 queries: neutral on the iid set (1.01), while the same clustered set at
        //  nq 256 / nprobe 8 takes 3.56 ms query-major against 2.95 ms list-major at nprobe 16)
        const int64_t lm_min_pairs = fdev.flags != 0 ? (3 * (int64_t)ix->nlist + 3) / 4
                                     : (ix->ds <= 1024 ? ((int64_t)ix->nlist + (m > 16 ? 7 : 3)) / (m > 16 ? 8 : 4) : ((int64_t)ix->nlist + 1) / 2);
        const bool list_major = ix->ds >= 128 && npairs <= (int64_t)INT32_MAX &&
                                (ix->scan_mode == 2 || (ix->scan_mode == 0 && long_lists && npairs >= lm_min_pairs));
        if (list_major) {
            const size_t nl = (size_t)ix->nlist;
            const size_t agg_words = (size_t)list_plan_ctas(ix->nlist) * 8;  // 4 x u64 per plan CTA
            const size_t words = 3 * nl + 4 + agg_words + 4 * (nl + 1) + (size_t)npairs + 16;
            CU(tl_scr->s_lplan.reserve(words * 4));
            int32_t *w = tl_scr->s_lplan.as<int32_t>();
            ListPlan lp;
            lp.nlist = ix->nlist;
            // tile items on the tensor cores (scan_lists_tc.cu): inner product and whole 32-float k-blocks only;
            // lists_cfg 1 / 2 keep the exact-fp32 FFMA tiles (scan_lists.cu), which also serve L2 and other dims
            const bool tc_tiles = ix->lists_cfg != 1 && ix->lists_cfg != 2 && ix->metric == SC_METRIC_IP && ix->ds % 32 == 0;
            lp.chunk = tc_tiles ? 64 : 32;
            lp.qsplit = nullptr;
            lp.bstage = nullptr;
            lp.mq_fused = ix->mq_fused;
            lp.tile_rem = ix->tile_rem;
            if (tc_tiles) {
                const bool ts = ix->lists_cfg != 5 && ix->lists_cfg != 3 && ix->d_maps != nullptr &&
                                npairs * (int64_t)(ix->ds / 32) < ((int64_t)1 << 30);
                if (!ts) {
                    CU(tl_scr->s_qsplit.reserve((size_t)m * ix->ds * 4 * 2));
                    lp.qsplit = tl_scr->s_qsplit.as<float>();
                }
                if (ts) {
                    const unsigned before = tl_scr->s_bstage.gen;
                    CU(tl_scr->s_bstage.reserve(scan_lists_ts_stage_bytes(ix->ds, ix->num_sms)));
                    if (tl_scr->s_bstage.gen != before) CU(cudaMemsetAsync(tl_scr->s_bstage.p, 0, tl_scr->s_bstage.cap, st));  // padded rows are read (never stored): keep them finite
                    lp.bstage = tl_scr->s_bstage.as<float>();
                }
            }
            lp.cnt = w;
            lp.cursor = w + nl;
            lp.counters = w + 2 * nl;
            lp.agg = reinterpret_cast<unsigned long long *>(lp.counters + 4);
            lp.n32 = lp.counters + 4 + agg_words;
            lp.lq_off = lp.n32 + nl;
            lp.off32 = lp.lq_off + nl + 1;
            lp.pg8off = lp.off32 + nl + 1;
            lp.pg4off = lp.pg8off + nl + 1;
            lp.lq = lp.pg4off + nl + 1;
            lp.unique_rows = ix->profiling ? tl_scr->prof_rows + 1 : nullptr;
            for (int i = 0; i < 2; ++i) {
                lp.side[i] = ix->lists_fork ? tl_scr->side[i] : nullptr;
                lp.ev_join[i] = tl_scr->ev_join[i];
            }
            lp.ev_fork = tl_scr->ev_fork;
            CU(launch_scan_lists(a, lp, ix->lists_cfg, ix->num_sms, &tl_scr->prof_scan_launches, st));
        } else {
            CU(launch_scan_pages(a, ix->scan_variant, ix->num_sms, &tl_scr->prof_scan_launches, st, pdl && planned));
        }
        SC(prof_mark(ix, st));
        nv.next(ex ? "search:topk+exchange+merge" : "search:topk");
        float *od = outd_dev ? out_dist + s * k : tl_scr->s_outd.as<float>();
        int64_t *oi = outi_dev ? out_ids + s * k : tl_scr->s_outi.as<int64_t>();
        if (ex) {
            PeerTopk pk;
            memset(&pk, 0, sizeof(pk));
            const size_t slot = (size_t)ex->rank * m * k;
            for (int p = 0; p < ex->world; ++p) {
                pk.d[p] = reinterpret_cast<float *>(ex->peer[p] + exl.part_d) + slot;
                pk.i[p] = reinterpret_cast<int64_t *>(ex->peer[p] + exl.part_i) + slot;
            }
            CU(launch_select_candidates_peers(a, m, k, pk, make_signal(ex, 1), st));
            char *mine = ex->peer[ex->rank];
            // The wait is its own one-warp kernel: a merge whose every CTA spins would fill the SMs while it waits, and with two
            // steps in flight per rank (two streams, two exchanges) the peers' producers of the OTHER step could then be
            // locked out on every rank at once.
            CU(launch_peer_wait(make_wait(ex, 1), st));
            CU(launch_merge_topk(reinterpret_cast<const float *>(mine + exl.part_d),
                                 reinterpret_cast<const int64_t *>(mine + exl.part_i), ex->world, m, k, k, ix->metric, od, oi, st));
            tl_scr->prof_total_launches += 2;
        } else {
            CU(launch_select_candidates(a, m, k, od, oi, st, pdl && planned && !list_major));
        }
        SC(prof_mark(ix, st));
        tl_scr->prof_total_launches += 2;  // pair plan, top-k (the scan launchers count their own)
        if (!outd_dev) CU(cudaMemcpyAsync(out_dist + s * k, od, (size_t)m * k * 4, cudaMemcpyDeviceToHost, st));
        if (!outi_dev) CU(cudaMemcpyAsync(out_ids + s * k, oi, (size_t)m * k * 8, cudaMemcpyDeviceToHost, st));
    }
    SC(end_call(ix, st));
    if (!outd_dev || !outi_dev) CU(cudaStreamSynchronize(st));
    return SC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char *sc_last_error(void) { return g_err.c_str(); }

int sc_abi_version(void) { return SC_ABI_VERSION; }

int sc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int sc_index_create(int32_t dim, int32_t metric, int32_t nlist, int32_t device, sc_index_t **out) {
    if (!out) return fail(SC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (dim < 1 || dim > 65536) return fail(SC_ERR_INVALID, "dim must be in [1, 65536]");
    if (metric != SC_METRIC_IP && metric != SC_METRIC_L2) return fail(SC_ERR_INVALID, "unknown metric %d", metric);
    if (nlist < 1 || nlist > (1 << 24)) return fail(SC_ERR_INVALID, "nlist must be in [1, 2^24]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SC_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(SC_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    int major = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (major != 10) return fail(SC_ERR_CUDA, "device %d has compute capability %d.x; kernels are built for sm_100a only", device, major);
    DeviceGuard g(device);
    if (!g.ok) return fail(SC_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    sc_index *ix = new sc_index();
    ix->dim = dim;
    ix->ds = (dim + 3) & ~3;
    ix->metric = metric;
    ix->nlist = nlist;
    ix->device = device;
    ix->num_sms = sms;
    // slab = ~256 MB of vectors, at least 64 pages, at most 2^16 pages
    int shift = 6;
    while (shift < 16 && ((size_t)2 << shift) * page_bytes(ix) <= ((size_t)256 << 20)) ++shift;
    if (const char *env = getenv("SEMCODE_SLAB_SHIFT")) {  // experiments: pages per slab = 2^shift
        const int v = atoi(env);
        if (v >= 6 && v <= 16) shift = v;
    }
    ix->slab_shift = shift;
    ix->h_tab = new SlabTable();
    memset(ix->h_tab, 0, sizeof(SlabTable));
    ix->h_len.assign(nlist, 0);
    ix->h_bound_prefix.assign(nlist + 1, 0);
    auto cleanup = [&](int code) {
        sc_index_destroy(ix);
        return code;
    };
    cudaError_t e = cudaMalloc(&ix->centroids, (size_t)nlist * ix->ds * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->cnorm, (size_t)nlist * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->cent_hi, (size_t)nlist * ix->ds * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->cent_lo, (size_t)nlist * ix->ds * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->d_tab, sizeof(SlabTable));
    if (e == cudaSuccess) e = cudaMalloc(&ix->d_maps, (size_t)kMaxSlabs * 128);
    if (e == cudaSuccess) e = cudaMalloc(&ix->list_len, (size_t)nlist * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->pt_off, (size_t)(nlist + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->pt_off_alt, (size_t)(nlist + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->pt, 1024 * 4);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->ev_write, cudaEventDisableTiming);
    for (int k = 0; k < kSlots && e == cudaSuccess; ++k) {
        Scratch &sl = ix->scr[k];
        e = cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaStreamCreateWithFlags(&sl.side[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_join[i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventRecord(sl.ev_done, 0);
    }
    if (e == cudaSuccess) e = cudaEventRecord(ix->ev_write, 0);
    if (e == cudaSuccess) e = cudaMemset(ix->list_len, 0, (size_t)nlist * 4);
    if (e == cudaSuccess) e = cudaMemset(ix->pt_off, 0, (size_t)(nlist + 1) * 4);
    if (e == cudaSuccess) e = cudaMemset(ix->d_tab, 0, sizeof(SlabTable));
    if (e == cudaSuccess) e = cudaMemset(ix->centroids, 0, (size_t)nlist * ix->ds * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cleanup(fail(e == cudaErrorMemoryAllocation ? SC_ERR_OOM : SC_ERR_CUDA, "index allocation failed: %s",
                            cudaGetErrorString(e)));
    }
    ix->pt_cap = 1024;
    *out = ix;
    return SC_OK;
}

int sc_index_destroy(sc_index_t *ix) {
    if (!ix) return SC_OK;
    DeviceGuard g(ix->device);
    cudaDeviceSynchronize();
    free_lists(ix);
    for (int k = 0; k < kSlots; ++k) {
        Scratch &sl = ix->scr[k];
        clear_prof(&sl);
        sl.each_buf([](DevBuf *b) { b->release(); });
        if (sl.prof_rows) cudaFree(sl.prof_rows);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
        for (int i = 0; i < 2; ++i) {
            if (sl.ev_join[i]) cudaEventDestroy(sl.ev_join[i]);
            if (sl.side[i]) cudaStreamDestroy(sl.side[i]);
        }
    }
    if (ix->ev_write) cudaEventDestroy(ix->ev_write);
    for (void *p : {(void *)ix->centroids, (void *)ix->cnorm, (void *)ix->cent_hi, (void *)ix->cent_lo, (void *)ix->d_tab, (void *)ix->d_maps, (void *)ix->list_len, (void *)ix->pt_off,
                    (void *)ix->pt_off_alt, (void *)ix->pt, (void *)ix->pt_alt, (void *)ix->free_pages})
        if (p) cudaFree(p);
    delete ix->h_tab;
    cudaGetLastError();
    delete ix;
    return SC_OK;
}

int sc_index_reset(sc_index_t *ix) {
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    CU(cudaDeviceSynchronize());
    free_lists(ix);
    memset(ix->h_tab, 0, sizeof(SlabTable));
    CU(cudaMemset(ix->list_len, 0, (size_t)ix->nlist * 4));
    CU(cudaMemset(ix->pt_off, 0, (size_t)(ix->nlist + 1) * 4));
    std::fill(ix->h_len.begin(), ix->h_len.end(), 0);
    refresh_bounds(ix);
    return SC_OK;
}

int sc_index_set_centroids(sc_index_t *ix, const float *centroids, int32_t nlist, void *stream) {
    if (!ix || !centroids) return fail(SC_ERR_INVALID, "NULL argument");
    if (nlist != ix->nlist) return fail(SC_ERR_INVALID, "nlist %d does not match the index (%d)", nlist, ix->nlist);
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (ix->ntotal > 0) return fail(SC_ERR_STATE, "cannot replace centroids of a non-empty index (reset it first)");
    CU(cudaDeviceSynchronize());
    const float *cd = nullptr;
    SC(stage(ix, centroids, (size_t)nlist * ix->dim, tl_scr->s_x, st, &cd));
    CU(launch_pad_rows(cd, nlist, ix->dim, ix->ds, ix->centroids, st));
    SC(update_cnorm(ix, st));
    CU(cudaStreamSynchronize(st));
    ix->trained = true;
    return SC_OK;
}

int sc_index_get_centroids(sc_index_t *ix, float *out, void *stream) {
    if (!ix || !out) return fail(SC_ERR_INVALID, "NULL argument");
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    SC(require_trained(ix));
    SC(begin_call(ix, st));
    const bool dev = is_device_ptr(out, ix->device);
    CU(cudaMemcpy2DAsync(out, (size_t)ix->dim * 4, ix->centroids, (size_t)ix->ds * 4, (size_t)ix->dim * 4, ix->nlist,
                         dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    SC(end_call(ix, st));
    if (!dev) CU(cudaStreamSynchronize(st));
    return SC_OK;
}

// ---- k-means ------------------------------------------------------------------------------------
int sc_index_kmeans_init(sc_index_t *ix, const float *x, int64_t n, const int64_t *init_rows, void *stream) {
    if (!ix || !x || !init_rows) return fail(SC_ERR_INVALID, "NULL argument");
    if (n < ix->nlist) return fail(SC_ERR_INVALID, "need at least nlist=%d training rows, got %lld", ix->nlist, (long long)n);
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (ix->ntotal > 0) return fail(SC_ERR_STATE, "cannot retrain a non-empty index (reset it first)");
    CU(cudaDeviceSynchronize());
    std::vector<int64_t> rows(ix->nlist);
    if (is_device_ptr(init_rows, ix->device)) {
        CU(cudaMemcpy(rows.data(), init_rows, (size_t)ix->nlist * 8, cudaMemcpyDeviceToHost));
    } else {
        memcpy(rows.data(), init_rows, (size_t)ix->nlist * 8);
    }
    for (int i = 0; i < ix->nlist; ++i)
        if (rows[i] < 0 || rows[i] >= n) return fail(SC_ERR_INVALID, "init_rows[%d]=%lld outside [0,%lld)", i, (long long)rows[i], (long long)n);
    if (is_device_ptr(x, ix->device)) {
        if (ix->ds == ix->dim && ((uintptr_t)x & 15) == 0) {
            CU(tl_scr->s_rows.reserve((size_t)ix->nlist * 8));
            CU(cudaMemcpyAsync(tl_scr->s_rows.p, rows.data(), (size_t)ix->nlist * 8, cudaMemcpyHostToDevice, st));
            CU(launch_gather_rows(x, tl_scr->s_rows.as<int64_t>(), ix->nlist, ix->ds, ix->centroids, st));
        } else {
            for (int i = 0; i < ix->nlist; ++i)
                CU(launch_pad_rows(x + rows[i] * ix->dim, 1, ix->dim, ix->ds, ix->centroids + (int64_t)i * ix->ds, st));
        }
    } else {
        std::vector<float> c((size_t)ix->nlist * ix->ds, 0.f);
        for (int i = 0; i < ix->nlist; ++i) memcpy(&c[(size_t)i * ix->ds], x + rows[i] * ix->dim, (size_t)ix->dim * 4);
        CU(cudaMemcpyAsync(ix->centroids, c.data(), c.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    SC(update_cnorm(ix, st));
    CU(cudaStreamSynchronize(st));
    ix->trained = true;
    return SC_OK;
}

// one assignment pass over x[n, dim]: sums[nlist, ds] (fp64), counts[nlist], objective[1] are
// ACCUMULATED into (device buffers owned by the caller, so that ranks can all-reduce them)
int sc_index_kmeans_step(sc_index_t *ix, const float *x, int64_t n, double *sums, int32_t *counts, double *objective,
                         void *stream) {
    NvtxRange nvtx_call("sc_index_kmeans_step");
    if (!ix || !sums || !counts || !objective) return fail(SC_ERR_INVALID, "NULL argument");
    if (n < 0) return fail(SC_ERR_INVALID, "n < 0");
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    SC(require_trained(ix));
    if (!is_device_ptr(sums, ix->device) || !is_device_ptr(counts, ix->device) || !is_device_ptr(objective, ix->device))
        return fail(SC_ERR_INVALID, "sums / counts / objective must be device buffers");
    if (n == 0) return SC_OK;
    if (!x) return fail(SC_ERR_INVALID, "x is NULL");
    SC(begin_call(ix, st));
    // host rows are staged 256 MB at a time; device-resident rows are used in place, in larger chunks
    // (2048 row tiles = ~14 waves of the persistent tensor-core kernel, so the tail wave is small)
    int64_t chunk = std::max<int64_t>(1024, ((int64_t)256 << 20) / ((int64_t)ix->ds * 4));
    if (is_device_ptr(x, ix->device)) chunk = std::max<int64_t>(chunk, (int64_t)1 << 18);
    chunk = std::min(chunk, n);
    CU(tl_scr->s_assign.reserve((size_t)chunk * 4));
    CU(tl_scr->s_best.reserve((size_t)chunk * 4));
    for (int64_t s = 0; s < n; s += chunk) {
        const int64_t m = std::min(chunk, n - s);
        const float *xd = nullptr;
        SC(stage_rows(ix, x + s * ix->dim, m, tl_scr->s_x, tl_scr->s_xpad, st, &xd));
        SC(coarse_assign(ix, xd, m, tl_scr->s_assign.as<int32_t>(), tl_scr->s_best.as<float>(), st));
        CU(launch_kmeans_accumulate(xd, m, ix->ds, tl_scr->s_assign.as<int32_t>(), tl_scr->s_best.as<float>(), ix->metric, sums,
                                    counts, objective, st));
    }
    SC(end_call(ix, st));
    return SC_OK;
}

// centroids <- sums / counts, then FAISS-style split of empty clusters (largest donor first,
// ties to the lowest index, +-1/1024 perturbation).  nsplit_out (host, nullable).
int sc_index_kmeans_update(sc_index_t *ix, const double *sums, const int32_t *counts, int32_t *nsplit_out,
                           void *stream) {
    NvtxRange nvtx_call("sc_index_kmeans_update");
    if (!ix || !sums || !counts) return fail(SC_ERR_INVALID, "NULL argument");
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    SC(require_trained(ix));
    if (ix->ntotal + ix->nremoved > 0)  // the lists were assigned under the current centroids (as kmeans_init / set_centroids refuse)
        return fail(SC_ERR_STATE, "cannot move the centroids of a non-empty index (reset it first)");
    if (!is_device_ptr(sums, ix->device) || !is_device_ptr(counts, ix->device))
        return fail(SC_ERR_INVALID, "sums / counts must be device buffers");
    SC(begin_call(ix, st));
    CU(launch_kmeans_finalize(sums, counts, ix->nlist, ix->ds, ix->centroids, st));
    std::vector<int32_t> hc(ix->nlist);
    CU(cudaMemcpyAsync(hc.data(), counts, (size_t)ix->nlist * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // donor = currently largest cluster (ties -> lowest index), as oracle/ivf_numpy.py::split_empty_clusters;
    // a heap keeps this O((empties + nlist) log nlist) instead of one argmax scan per empty cluster
    int nsplit = 0;
    std::vector<int64_t> c64(hc.begin(), hc.end());
    typedef std::pair<int64_t, int> Ent;  // (count, -index): max-heap on count, then on lower index
    std::priority_queue<Ent> heap;
    for (int j = 0; j < ix->nlist; ++j)
        if (c64[j] >= 2) heap.push(Ent(c64[j], -j));
    for (int ci = 0; ci < ix->nlist; ++ci) {
        if (c64[ci] != 0) continue;
        while (!heap.empty() && heap.top().first != c64[-heap.top().second]) heap.pop();  // stale entries
        if (heap.empty()) break;
        const int cj = -heap.top().second;
        if (c64[cj] < 2) break;
        heap.pop();
        CU(launch_split_centroid(ix->centroids, ix->ds, ci, cj, st));
        c64[ci] = c64[cj] / 2;
        c64[cj] -= c64[ci];
        if (c64[ci] >= 2) heap.push(Ent(c64[ci], -ci));
        if (c64[cj] >= 2) heap.push(Ent(c64[cj], -cj));
        ++nsplit;
    }
    SC(update_cnorm(ix, st));
    SC(end_call(ix, st));
    CU(cudaStreamSynchronize(st));
    if (nsplit_out) *nsplit_out = nsplit;
    return SC_OK;
}

int sc_index_train(sc_index_t *ix, const float *x, int64_t n, int32_t niter, const int64_t *init_rows,
                   double *objective_out, void *stream) {
    NvtxRange nvtx_call("sc_index_train");
    if (!ix || !x || !init_rows) return fail(SC_ERR_INVALID, "NULL argument");
    if (niter < 0) return fail(SC_ERR_INVALID, "niter < 0");
    cudaStream_t st = (cudaStream_t)stream;
    const float *xd = x;
    float *owned = nullptr;
    {
        DeviceGuard g(ix->device);
        if (!is_device_ptr(x, ix->device)) {
            // keep the training set resident for all iterations
            cudaError_t e = cudaMalloc(&owned, (size_t)n * ix->dim * 4);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(SC_ERR_OOM, "cannot stage %lld training rows on the device: %s", (long long)n, cudaGetErrorString(e));
            }
            e = cudaMemcpyAsync(owned, x, (size_t)n * ix->dim * 4, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) {
                cudaFree(owned);
                return fail(SC_ERR_CUDA, "H2D copy of the training set failed: %s", cudaGetErrorString(e));
            }
            xd = owned;
        }
    }
    int rc = sc_index_kmeans_init(ix, xd, n, init_rows, stream);
    double *sums = nullptr;
    int32_t *counts = nullptr;
    double *obj = nullptr;
    {
        DeviceGuard g(ix->device);
        if (rc == SC_OK) {
            cudaError_t e = cudaMalloc(&sums, (size_t)ix->nlist * ix->ds * 8);
            if (e == cudaSuccess) e = cudaMalloc(&counts, (size_t)ix->nlist * 4);
            if (e == cudaSuccess) e = cudaMalloc(&obj, 8);
            if (e != cudaSuccess) {
                cudaGetLastError();
                rc = fail(SC_ERR_OOM, "k-means accumulators: %s", cudaGetErrorString(e));
            }
        }
        for (int it = 0; rc == SC_OK && it < niter; ++it) {
            cudaMemsetAsync(sums, 0, (size_t)ix->nlist * ix->ds * 8, st);
            cudaMemsetAsync(counts, 0, (size_t)ix->nlist * 4, st);
            cudaMemsetAsync(obj, 0, 8, st);
            rc = sc_index_kmeans_step(ix, xd, n, sums, counts, obj, stream);
            if (rc != SC_OK) break;
            if (objective_out) {
                cudaError_t e = cudaMemcpyAsync(objective_out + it, obj, 8, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) {
                    rc = fail(SC_ERR_CUDA, "objective readback: %s", cudaGetErrorString(e));
                    break;
                }
            }
            rc = sc_index_kmeans_update(ix, sums, counts, nullptr, stream);
        }
        cudaStreamSynchronize(st);
        if (sums) cudaFree(sums);
        if (counts) cudaFree(counts);
        if (obj) cudaFree(obj);
        if (owned) cudaFree(owned);
    }
    return rc;
}

// ---- coarse quantizer -----------------------------------------------------------------------------
int sc_index_assign(sc_index_t *ix, const float *x, int64_t n, int32_t *out_list, void *stream) {
    NvtxRange nvtx_call("sc_index_assign");
    if (!ix || !out_list) return fail(SC_ERR_INVALID, "NULL argument");
    if (n < 0) return fail(SC_ERR_INVALID, "n < 0");
    ReadGuard lk(ix, stream);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    SC(require_trained(ix));
    if (n == 0) return SC_OK;
    if (!x) return fail(SC_ERR_INVALID, "x is NULL");
    SC(begin_call(ix, st));
    const bool dev = is_device_ptr(out_list, ix->device);
    int64_t chunk = std::max<int64_t>(1024, ((int64_t)256 << 20) / ((int64_t)ix->ds * 4));
    chunk = std::min(chunk, n);
    if (!dev) CU(tl_scr->s_assign.reserve((size_t)chunk * 4));
    for (int64_t s = 0; s < n; s += chunk) {
        const int64_t m = std::min(chunk, n - s);
        const float *xd = nullptr;
        SC(stage_rows(ix, x + s * ix->dim, m, tl_scr->s_x, tl_scr->s_xpad, st, &xd));
        int32_t *dst = dev ? out_list + s : tl_scr->s_assign.as<int32_t>();
        SC(coarse_assign(ix, xd, m, dst, nullptr, st));
        if (!dev) CU(cudaMemcpyAsync(out_list + s, dst, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    }
    SC(end_call(ix, st));
    if (!dev) CU(cudaStreamSynchronize(st));
    return SC_OK;
}

int sc_index_probe(sc_index_t *ix, const float *q, int64_t nq, int32_t nprobe, int32_t *out_lists, float *out_scores,
                   void *stream) {
    NvtxRange nvtx_call("sc_index_probe");
    if (!ix || !out_lists) return fail(SC_ERR_INVALID, "NULL argument");
    if (nq < 0 || nprobe < 1) return fail(SC_ERR_INVALID, "bad nq / nprobe");
    ReadGuard lk(ix, stream);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    SC(require_trained(ix));
    if (nprobe > ix->nlist) return fail(SC_ERR_INVALID, "nprobe %d exceeds nlist %d (outputs are [nq, nprobe])", nprobe, ix->nlist);
    if (nprobe > kMaxK) return fail(SC_ERR_INVALID, "sc_index_probe ranks at most %d lists per query", kMaxK);
    if (nq == 0) return SC_OK;
    if (!q) return fail(SC_ERR_INVALID, "q is NULL");
    SC(begin_call(ix, st));
    const bool ldev = is_device_ptr(out_lists, ix->device);
    const bool sdev = out_scores ? is_device_ptr(out_scores, ix->device) : true;
    const int64_t ch = coarse_chunk_rows(ix, nq);
    CU(tl_scr->s_scores.reserve((size_t)ch * ix->nlist * 4));
    if (!ldev) CU(tl_scr->s_probe.reserve((size_t)ch * nprobe * 4));
    if (out_scores && !sdev) CU(tl_scr->s_best.reserve((size_t)ch * nprobe * 4));
    for (int64_t s = 0; s < nq; s += ch) {
        const int64_t m = std::min(ch, nq - s);
        const float *qd = nullptr;
        SC(stage_rows(ix, q + s * ix->dim, m, tl_scr->s_q, tl_scr->s_xpad, st, &qd));
        SC(coarse_scores(ix, qd, m, tl_scr->s_scores.as<float>(), st));
        int32_t *ol = ldev ? out_lists + s * nprobe : tl_scr->s_probe.as<int32_t>();
        float *os = out_scores ? (sdev ? out_scores + s * nprobe : tl_scr->s_best.as<float>()) : nullptr;
        CU(launch_select_rows(tl_scr->s_scores.as<float>(), m, ix->nlist, nprobe, ol, os, st));
        if (!ldev) CU(cudaMemcpyAsync(out_lists + s * nprobe, ol, (size_t)m * nprobe * 4, cudaMemcpyDeviceToHost, st));
        if (out_scores && !sdev)
            CU(cudaMemcpyAsync(out_scores + s * nprobe, os, (size_t)m * nprobe * 4, cudaMemcpyDeviceToHost, st));
    }
    SC(end_call(ix, st));
    if (!ldev || !sdev) CU(cudaStreamSynchronize(st));
    return SC_OK;
}

// ---- insert / remove ------------------------------------------------------------------------------
int sc_index_add(sc_index_t *ix, const float *x, const int64_t *ids, const uint32_t *repo_tags,
                 const uint8_t *lang_tags, int64_t n, void *stream) {
    NvtxRange nvtx_call("sc_index_add");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    return add_impl(ix, x, ids, repo_tags, lang_tags, nullptr, n, (cudaStream_t)stream);
}

int sc_index_add_preassigned(sc_index_t *ix, const float *x, const int64_t *ids, const uint32_t *repo_tags,
                             const uint8_t *lang_tags, const int32_t *lists, int64_t n, void *stream) {
    NvtxRange nvtx_call("sc_index_add_preassigned");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    if (!lists && n > 0) return fail(SC_ERR_INVALID, "lists is NULL");
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    return add_impl(ix, x, ids, repo_tags, lang_tags, lists, n, (cudaStream_t)stream);
}

int sc_index_remove_ids(sc_index_t *ix, const int64_t *ids, int64_t n, int64_t *n_removed_out, void *stream) {
    NvtxRange nvtx_call("sc_index_remove_ids");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    if (n < 0) return fail(SC_ERR_INVALID, "n < 0");
    if (n_removed_out) *n_removed_out = 0;
    if (n == 0) return SC_OK;
    if (!ids) return fail(SC_ERR_INVALID, "ids is NULL");
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaDeviceSynchronize());
    std::vector<int64_t> h((size_t)n);
    if (is_device_ptr(ids, ix->device)) {
        CU(cudaMemcpy(h.data(), ids, (size_t)n * 8, cudaMemcpyDeviceToHost));
    } else {
        memcpy(h.data(), ids, (size_t)n * 8);
    }
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    CU(tl_scr->s_rm.reserve(h.size() * 8));
    CU(tl_scr->s_cnt.reserve(8));
    CU(cudaMemcpyAsync(tl_scr->s_rm.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(tl_scr->s_cnt.p, 0, 8, st));
    CU(launch_remove_ids(tl_scr->s_rm.as<int64_t>(), (int64_t)h.size(), ix->d_tab, ix->slab_shift, ix->pool_top,
                         tl_scr->s_cnt.as<unsigned long long>(), st));
    unsigned long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, tl_scr->s_cnt.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ix->ntotal -= (int64_t)cnt;
    ix->nremoved += (int64_t)cnt;
    if (n_removed_out) *n_removed_out = (int64_t)cnt;
    return SC_OK;
}

// ---- search ---------------------------------------------------------------------------------------
int sc_index_search(sc_index_t *ix, const float *q, int64_t nq, int32_t k, int32_t nprobe, const sc_filter_t *filter,
                    float *out_dist, int64_t *out_ids, void *stream) {
    NvtxRange nvtx_call("sc_index_search");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    ReadGuard lk(ix, stream);
    DeviceGuard g(ix->device);
    return search_impl(ix, q, nq, k, nprobe, nullptr, filter, out_dist, out_ids, (cudaStream_t)stream);
}

int sc_index_search_preassigned(sc_index_t *ix, const float *q, int64_t nq, int32_t k, int32_t nprobe,
                                const int32_t *lists, const sc_filter_t *filter, float *out_dist, int64_t *out_ids,
                                void *stream) {
    NvtxRange nvtx_call("sc_index_search_preassigned");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    if (!lists && nq > 0) return fail(SC_ERR_INVALID, "lists is NULL");
    ReadGuard lk(ix, stream);
    DeviceGuard g(ix->device);
    return search_impl(ix, q, nq, k, nprobe, lists, filter, out_dist, out_ids, (cudaStream_t)stream);
}

int sc_exchange_create(int32_t rank, int32_t world, const void *const *peer_buffers, int64_t buffer_bytes, int32_t device,
                       sc_exchange_t **out) {
    if (!out) return fail(SC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(SC_ERR_INVALID, "rank %d / world %d: world must be in [1, %d]", rank, world, kMaxPeers);
    if (!peer_buffers) return fail(SC_ERR_INVALID, "peer_buffers is NULL");
    if (buffer_bytes < (int64_t)(kExHeader + (64 << 10))) return fail(SC_ERR_INVALID, "exchange buffers must hold at least 68 KiB");
    DeviceGuard g(device);
    if (!g.ok) return fail(SC_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    for (int p = 0; p < world; ++p) {
        if (!peer_buffers[p] || ((uintptr_t)peer_buffers[p] & 255)) return fail(SC_ERR_INVALID, "peer buffer %d is NULL or not 256-byte aligned", p);
    }
    sc_exchange *ex = new sc_exchange();
    ex->rank = rank;
    ex->world = world;
    ex->device = device;
    ex->bytes = (size_t)buffer_bytes;
    for (int p = 0; p < world; ++p) ex->peer[p] = (char *)peer_buffers[p];
    unsigned int *w = nullptr;
    cudaError_t e = cudaMalloc(&w, 16);
    if (e == cudaSuccess) e = cudaMemset(w, 0, 16);
    if (e == cudaSuccess) e = cudaHostAlloc(&ex->h_status, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) {
        *ex->h_status = 0;
        e = cudaHostGetDevicePointer(&ex->status, ex->h_status, 0);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (w) cudaFree(w);
        if (ex->h_status) cudaFreeHost(ex->h_status);
        delete ex;
        return fail(SC_ERR_CUDA, "exchange counters: %s", cudaGetErrorString(e));
    }
    ex->done = w;
    *out = ex;
    return SC_OK;
}

int sc_exchange_destroy(sc_exchange_t *ex) {
    if (!ex) return SC_OK;
    DeviceGuard g(ex->device);
    if (ex->done) cudaFree(ex->done);
    if (ex->h_status) cudaFreeHost(ex->h_status);
    delete ex;
    return SC_OK;
}

int sc_exchange_status(sc_exchange_t *ex, int32_t *timed_out, int64_t *epoch) {
    if (!ex) return fail(SC_ERR_INVALID, "exchange is NULL");
    DeviceGuard g(ex->device);
    CU(cudaDeviceSynchronize());  // every step issued so far has finished: the word is final
    if (timed_out) *timed_out = (*(volatile unsigned int *)ex->h_status != 0 || ex->broken) ? 1 : 0;
    if (epoch) *epoch = (int64_t)ex->epoch;
    return SC_OK;
}

// the same word WITHOUT synchronising: 1 as soon as a finished step of this rank gave up waiting for a peer (or a step
// failed midway).  Cheap enough to call before every step.
int sc_exchange_poll(sc_exchange_t *ex, int32_t *timed_out) {
    if (!ex || !timed_out) return fail(SC_ERR_INVALID, "NULL argument");
    *timed_out = (*(volatile unsigned int *)ex->h_status != 0 || ex->broken) ? 1 : 0;
    return SC_OK;
}

int sc_exchange_set_timeout_ms(sc_exchange_t *ex, int64_t ms) {
    if (!ex) return fail(SC_ERR_INVALID, "exchange is NULL");
    if (ms < 1) return fail(SC_ERR_INVALID, "timeout must be >= 1 ms");
    ex->timeout_ns = (unsigned long long)ms * 1000000ull;
    return SC_OK;
}

int sc_index_search_sharded(sc_index_t *ix, sc_exchange_t *ex, const float *q, int64_t nq, int32_t k, int32_t nprobe,
                            const int32_t *lists, const sc_filter_t *filter, float *out_dist, int64_t *out_ids, void *stream) {
    NvtxRange nvtx_call("sc_index_search_sharded");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    if (!ex) return fail(SC_ERR_INVALID, "exchange is NULL");
    ReadGuard lk(ix, stream);
    DeviceGuard g(ix->device);
    if (ex->broken || *(volatile unsigned int *)ex->h_status != 0)
        return fail(SC_ERR_STATE, "the exchange is unusable: %s.  Results since then are invalid; destroy the exchange on every "
                                  "rank, barrier, and create a new one", ex->broken ? "an earlier step failed midway on this rank"
                                                                                    : "a peer did not arrive within the timeout");
    const unsigned long long epoch0 = ex->epoch;
    const int rc = search_impl(ix, q, nq, k, nprobe, lists, filter, out_dist, out_ids, (cudaStream_t)stream, ex);
    if (rc != SC_OK && ex->epoch != epoch0) ex->broken = true;  // this rank's epoch is ahead of what it published
    return rc;
}

int sc_merge_topk(const float *part_dist, const int64_t *part_ids, int32_t parts, int64_t nq, int32_t kin, int32_t k,
                  int32_t metric, float *out_dist, int64_t *out_ids, int32_t device, void *stream) {
    NvtxRange nvtx_call("sc_merge_topk");
    if (!part_dist || !part_ids || !out_dist || !out_ids) return fail(SC_ERR_INVALID, "NULL argument");
    if (parts < 1 || kin < 1 || k < 1 || k > kMaxK || nq < 0) return fail(SC_ERR_INVALID, "bad parts / kin / k / nq");
    if (metric != SC_METRIC_IP && metric != SC_METRIC_L2) return fail(SC_ERR_INVALID, "unknown metric %d", metric);
    DeviceGuard g(device);
    if (!g.ok) return fail(SC_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    if (!is_device_ptr(part_dist, device) || !is_device_ptr(part_ids, device) || !is_device_ptr(out_dist, device) ||
        !is_device_ptr(out_ids, device))
        return fail(SC_ERR_INVALID, "sc_merge_topk works on device buffers only");
    CU(launch_merge_topk(part_dist, part_ids, parts, nq, kin, k, metric, out_dist, out_ids, (cudaStream_t)stream));
    return SC_OK;
}

// ---- compaction -------------------------------------------------------------------------------------
// Squeeze the tombstoned slots out of every list in place and hand the emptied pages to the free list (the next
// inserts take them before the pool grows).  Milvus compacts segments in the background [EXT]; here the host wrapper
// calls this when nremoved / (ntotal + nremoved) passes its threshold.  Searches return the same rows before and after
// (slot order inside a list is kept, so exact-tie order does not change either).
int sc_index_compact(sc_index_t *ix, int64_t *pages_freed_out, void *stream) {
    NvtxRange nvtx_call("sc_index_compact");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    if (pages_freed_out) *pages_freed_out = 0;
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (ix->nremoved == 0) return SC_OK;
    CU(cudaDeviceSynchronize());  // no search may still be reading the lists we are about to rewrite
    const int nlist = ix->nlist;
    int32_t h_old = 0;
    CU(cudaMemcpy(&h_old, ix->pt_off + nlist, 4, cudaMemcpyDeviceToHost));
    // free list: room for every page that may come back, the current entries kept at the front
    if (ix->nfree + h_old > ix->free_cap) {
        const int32_t want = ix->nfree + h_old + 1024;
        int32_t *nf = nullptr;
        CU(cudaMalloc(&nf, (size_t)want * 4));
        if (ix->nfree) CU(cudaMemcpy(nf, ix->free_pages, (size_t)ix->nfree * 4, cudaMemcpyDeviceToDevice));
        if (ix->free_pages) cudaFree(ix->free_pages);
        ix->free_pages = nf;
        ix->free_cap = want;
    }
    if (h_old > ix->pt_alt_cap) {
        if (ix->pt_alt) cudaFree(ix->pt_alt);
        ix->pt_alt = nullptr;
        ix->pt_alt_cap = 0;
        CU(cudaMalloc(&ix->pt_alt, (size_t)std::max<int32_t>(h_old, 1024) * 4));
        ix->pt_alt_cap = std::max<int32_t>(h_old, 1024);
    }
    CU(tl_scr->s_npg.reserve((size_t)nlist * 4));
    CU(tl_scr->s_bad.reserve(16));
    CU(launch_compact_lists(nlist, ix->list_len, ix->pt_off, ix->pt, ix->d_tab, ix->slab_shift, ix->ds, ix->num_sms, st));
    CU(launch_pages_of_len(ix->list_len, nlist, tl_scr->s_npg.as<int32_t>(), st));
    CU(launch_exclusive_scan_i32(tl_scr->s_npg.as<int32_t>(), nlist, ix->pt_off_alt, st));
    CU(cudaMemcpyAsync(tl_scr->s_bad.p, &ix->nfree, 4, cudaMemcpyHostToDevice, st));  // cursor starts behind the current entries
    CU(launch_compact_pt(ix->pt_off, ix->pt, ix->pt_off_alt, ix->pt_alt, nlist, ix->free_pages, tl_scr->s_bad.as<int32_t>(), st));
    int32_t h_free = 0;
    CU(cudaMemcpyAsync(&h_free, tl_scr->s_bad.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::swap(ix->pt, ix->pt_alt);
    std::swap(ix->pt_cap, ix->pt_alt_cap);
    std::swap(ix->pt_off, ix->pt_off_alt);
    if (pages_freed_out) *pages_freed_out = h_free - ix->nfree;
    ix->nfree = h_free;
    ix->nremoved = 0;
    SC(sync_host_lengths(ix, st));
    return SC_OK;
}

// ---- introspection ----------------------------------------------------------------------------------
int sc_index_stats(sc_index_t *ix, sc_stats_t *out) {
    if (!ix || !out) return fail(SC_ERR_INVALID, "NULL argument");
    WriteGuard lk(ix);
    memset(out, 0, sizeof(*out));
    out->dim = ix->dim;
    out->dim_padded = ix->ds;
    out->metric = ix->metric;
    out->nlist = ix->nlist;
    out->device = ix->device;
    out->trained = ix->trained ? 1 : 0;
    out->ntotal = ix->ntotal;
    out->nremoved = ix->nremoved;
    out->npages = ix->pool_top;
    out->nfree_pages = ix->nfree;
    const int64_t rows = ((int64_t)ix->slabs.size() << ix->slab_shift) * kPageRows;
    out->bytes_lists = rows * ((int64_t)ix->ds * 4 + 12);
    int64_t sb = 0;
    for (const DevBuf *b : {&tl_scr->s_q, &tl_scr->s_scores, &tl_scr->s_probe, &tl_scr->s_pageoff, &tl_scr->s_cand,
                            &tl_scr->s_outd, &tl_scr->s_outi, &tl_scr->s_repobits, &tl_scr->s_x, &tl_scr->s_xpad, &tl_scr->s_ids, &tl_scr->s_repo,
                            &tl_scr->s_lang, &tl_scr->s_assign, &tl_scr->s_best, &tl_scr->s_pos, &tl_scr->s_lenold, &tl_scr->s_need, &tl_scr->s_npg,
                            &tl_scr->s_needoff, &tl_scr->s_bad, &tl_scr->s_sums, &tl_scr->s_counts, &tl_scr->s_obj, &tl_scr->s_rows, &tl_scr->s_rm,
                            &tl_scr->s_cnt, &tl_scr->s_ahi, &tl_scr->s_alo, &tl_scr->s_lplan, &tl_scr->s_scan, &tl_scr->s_packed, &tl_scr->s_qsplit, &tl_scr->s_bstage})
        sb += (int64_t)b->cap;
    out->bytes_scratch = sb;
    int32_t mx = 0, mn = ix->nlist > 0 ? INT32_MAX : 0;
    for (int32_t v : ix->h_len) {
        mx = std::max(mx, v);
        mn = std::min(mn, v);
    }
    out->max_list_len = mx;
    out->min_list_len = mn;
    return SC_OK;
}

int sc_index_list_sizes(sc_index_t *ix, int32_t *out_host) {
    if (!ix || !out_host) return fail(SC_ERR_INVALID, "NULL argument");
    WriteGuard lk(ix);
    memcpy(out_host, ix->h_len.data(), (size_t)ix->nlist * 4);
    return SC_OK;
}

int sc_index_export_list(sc_index_t *ix, int32_t list, int64_t cap, float *vecs, int64_t *ids, uint32_t *tags,
                         int64_t *len_out, void *stream) {
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    if (list < 0 || list >= ix->nlist) return fail(SC_ERR_INVALID, "list %d outside [0,%d)", list, ix->nlist);
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t len = ix->h_len[list];
    if (len_out) *len_out = len;
    if (len == 0 || (!vecs && !ids && !tags)) return SC_OK;
    if (cap < len) return fail(SC_ERR_INVALID, "list %d holds %d rows, buffers hold %lld", list, len, (long long)cap);
    SC(begin_call(ix, st));
    int32_t pt_begin = 0;
    CU(cudaMemcpyAsync(&pt_begin, ix->pt_off + list, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const bool vdev = vecs ? is_device_ptr(vecs, ix->device) : true;
    const bool idev = ids ? is_device_ptr(ids, ix->device) : true;
    const bool tdev = tags ? is_device_ptr(tags, ix->device) : true;
    float *vd = vecs;
    int64_t *idd = ids;
    uint32_t *td = tags;
    if (vecs && !vdev) {
        CU(tl_scr->s_x.reserve((size_t)len * ix->dim * 4));
        vd = tl_scr->s_x.as<float>();
    }
    if (ids && !idev) {
        CU(tl_scr->s_ids.reserve((size_t)len * 8));
        idd = tl_scr->s_ids.as<int64_t>();
    }
    if (tags && !tdev) {
        CU(tl_scr->s_repo.reserve((size_t)len * 4));
        td = tl_scr->s_repo.as<uint32_t>();
    }
    CU(launch_export_list(ix->pt, pt_begin, len, ix->ds, ix->dim, ix->d_tab, ix->slab_shift, vd, idd, td, st));
    if (vecs && !vdev) CU(cudaMemcpyAsync(vecs, vd, (size_t)len * ix->dim * 4, cudaMemcpyDeviceToHost, st));
    if (ids && !idev) CU(cudaMemcpyAsync(ids, idd, (size_t)len * 8, cudaMemcpyDeviceToHost, st));
    if (tags && !tdev) CU(cudaMemcpyAsync(tags, td, (size_t)len * 4, cudaMemcpyDeviceToHost, st));
    SC(end_call(ix, st));
    CU(cudaStreamSynchronize(st));
    return SC_OK;
}

// lists [list_begin, list_end) back to back, slot order (tombstoned slots included: the caller reads the tags):
// off_out [list_end - list_begin + 1] (host) = exclusive prefix of the slot counts; buffers host or device, nullable
int sc_index_export_lists(sc_index_t *ix, int32_t list_begin, int32_t list_end, int64_t cap, float *vecs, int64_t *ids,
                          uint32_t *tags, int64_t *off_out, void *stream) {
    NvtxRange nvtx_call("sc_index_export_lists");
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    if (list_begin < 0 || list_end > ix->nlist || list_begin > list_end)
        return fail(SC_ERR_INVALID, "list range [%d, %d) outside [0, %d]", list_begin, list_end, ix->nlist);
    WriteGuard lk(ix);
    DeviceGuard g(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t nl = list_end - list_begin;
    std::vector<int64_t> off((size_t)nl + 1, 0);
    for (int32_t i = 0; i < nl; ++i) off[i + 1] = off[i] + ix->h_len[list_begin + i];
    if (off_out) memcpy(off_out, off.data(), off.size() * 8);
    const int64_t rows = off[nl];
    if (rows == 0 || (!vecs && !ids && !tags)) return SC_OK;
    if (cap < rows) return fail(SC_ERR_INVALID, "lists [%d, %d) hold %lld rows, buffers hold %lld", list_begin, list_end, (long long)rows, (long long)cap);
    SC(begin_call(ix, st));
    const bool vdev = vecs ? is_device_ptr(vecs, ix->device) : true;
    const bool idev = ids ? is_device_ptr(ids, ix->device) : true;
    const bool tdev = tags ? is_device_ptr(tags, ix->device) : true;
    float *vd = vecs;
    int64_t *idd = ids;
    uint32_t *td = tags;
    if (vecs && !vdev) {
        CU(tl_scr->s_x.reserve((size_t)rows * ix->dim * 4));
        vd = tl_scr->s_x.as<float>();
    }
    if (ids && !idev) {
        CU(tl_scr->s_ids.reserve((size_t)rows * 8));
        idd = tl_scr->s_ids.as<int64_t>();
    }
    if (tags && !tdev) {
        CU(tl_scr->s_repo.reserve((size_t)rows * 4));
        td = tl_scr->s_repo.as<uint32_t>();
    }
    CU(tl_scr->s_rows.reserve(off.size() * 8));
    CU(cudaMemcpyAsync(tl_scr->s_rows.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
    CU(launch_export_range(ix->pt_off, ix->pt, list_begin, nl, tl_scr->s_rows.as<int64_t>(), rows, ix->ds, ix->dim, ix->d_tab, ix->slab_shift,
                           vd, idd, td, ix->num_sms, st));
    if (vecs && !vdev) CU(cudaMemcpyAsync(vecs, vd, (size_t)rows * ix->dim * 4, cudaMemcpyDeviceToHost, st));
    if (ids && !idev) CU(cudaMemcpyAsync(ids, idd, (size_t)rows * 8, cudaMemcpyDeviceToHost, st));
    if (tags && !tdev) CU(cudaMemcpyAsync(tags, td, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
    SC(end_call(ix, st));
    CU(cudaStreamSynchronize(st));  // `off` dies at scope exit
    return SC_OK;
}

int sc_index_set_profiling(sc_index_t *ix, int32_t enabled) {
    if (!ix) return fail(SC_ERR_INVALID, "idx is NULL");
    WriteGuard lk(ix);
    ix->profiling = enabled != 0;
    if (!ix->profiling) {
        DeviceGuard g(ix->device);
        for (int k = 0; k < kSlots; ++k) clear_prof(&ix->scr[k]);
    }
    return SC_OK;
}

int sc_index_last_search_times(sc_index_t *ix, sc_search_times_t *out) {
    if (!ix || !out) return fail(SC_ERR_INVALID, "NULL argument");
    WriteGuard lk(ix);
    tl_scr = &ix->scr[ix->last_slot];  // the search that finished last
    DeviceGuard g(ix->device);
    memset(out, 0, sizeof(*out));
    if (!ix->profiling || tl_scr->prof_ev.empty()) return fail(SC_ERR_STATE, "profiling is off or no search has run");
    if (tl_scr->prof_ev.size() % 6 != 0) return fail(SC_ERR_STATE, "incomplete profile (a search failed midway)");
    CU(cudaEventSynchronize(tl_scr->prof_ev.back()));
    for (size_t c = 0; c < tl_scr->prof_ev.size(); c += 6) {
        float ms[5];
        for (int i = 0; i < 5; ++i) CU(cudaEventElapsedTime(&ms[i], tl_scr->prof_ev[c + i], tl_scr->prof_ev[c + i + 1]));
        out->coarse_ms += ms[0];
        out->probe_select_ms += ms[1];
        out->plan_ms += ms[2];
        out->scan_ms += ms[3];
        out->topk_ms += ms[4];
    }
    float tot = 0.f;
    CU(cudaEventElapsedTime(&tot, tl_scr->prof_ev.front(), tl_scr->prof_ev.back()));
    out->total_ms = tot;
    unsigned long long rows[2] = {0, 0};
    if (tl_scr->prof_rows) CU(cudaMemcpy(rows, tl_scr->prof_rows, 16, cudaMemcpyDeviceToHost));
    out->scanned_rows = (int64_t)rows[0];
    out->unique_rows = (int64_t)rows[1];
    out->scan_launches = tl_scr->prof_scan_launches;
    out->total_launches = tl_scr->prof_total_launches + tl_scr->prof_scan_launches;
    return SC_OK;
}

int sc_index_set_param(sc_index_t *ix, const char *name, int64_t value) {
    if (!ix || !name) return fail(SC_ERR_INVALID, "NULL argument");
    WriteGuard lk(ix);
    if (strcmp(name, "scratch_bytes") == 0) {
        if (value < ((int64_t)1 << 20)) return fail(SC_ERR_INVALID, "scratch_bytes must be >= 1 MiB");
        ix->scratch_budget = value;
        return SC_OK;
    }
    if (strcmp(name, "scan_variant") == 0) {
        if (value < 0 || value > 4) return fail(SC_ERR_INVALID, "scan_variant must be in [0,4]");
        ix->scan_variant = (int)value;
        return SC_OK;
    }
    if (strcmp(name, "lists_fork") == 0) {
        ix->lists_fork = value != 0;
        return SC_OK;
    }
    if (strcmp(name, "lists_cfg") == 0) {
        if (value < 0 || value > 5) return fail(SC_ERR_INVALID, "lists_cfg must be in [0,5]");
        ix->lists_cfg = (int)value;
        return SC_OK;
    }
    if (strcmp(name, "scan_mode") == 0) {
        if (value < 0 || value > 2) return fail(SC_ERR_INVALID, "scan_mode: 0 = auto, 1 = query-major, 2 = list-major");
        ix->scan_mode = (int)value;
        return SC_OK;
    }
    if (strcmp(name, "plan_epoch") == 0) {  // tests: put the pair plan's launch counter next to its 22-bit wrap
        if (value < 0 || value >= (1 << 22)) return fail(SC_ERR_INVALID, "plan_epoch must be in [0, 2^22)");
        for (int k = 0; k < kSlots; ++k) ix->scr[k].plan_epoch = (uint32_t)value;
        return SC_OK;
    }
    if (strcmp(name, "add_chunk_rows") == 0) {  // tests: rows per add chunk (0 = automatic)
        if (value < 0) return fail(SC_ERR_INVALID, "add_chunk_rows must be >= 0");
        ix->add_chunk_rows = value;
        return SC_OK;
    }
    if (strcmp(name, "fail_add_after") == 0) {  // tests: the value-th add chunk from now fails after claiming its slots
        if (value < 0 || value > INT32_MAX) return fail(SC_ERR_INVALID, "fail_add_after must be >= 0");
        ix->fail_add_after = (int32_t)value;
        return SC_OK;
    }
    if (strcmp(name, "small_coarse") == 0) {
        ix->small_coarse = value != 0;
        return SC_OK;
    }
    if (strcmp(name, "debug_canary") == 0) {  // guards around every scratch buffer allocated from now on (process-wide switch)
        DeviceGuard g(ix->device);
        CU(cudaDeviceSynchronize());
        g_canary = value != 0;
        for (int k = 0; k < kSlots; ++k) ix->scr[k].each_buf([](DevBuf *b) { b->release(); });  // re-allocated on next use
        return SC_OK;
    }
    if (strcmp(name, "check_canaries") == 0) {  // SC_ERR_STATE when a kernel wrote outside a guarded scratch buffer
        DeviceGuard g(ix->device);
        CU(cudaDeviceSynchronize());
        unsigned int *bad = nullptr;
        CU(cudaMalloc(&bad, 8));
        int checked = 0, which = -1, slot_bad = -1;
        unsigned int hb[2] = {0, 0}, first[2] = {0, 0};
        for (int k = 0; k < kSlots; ++k) {
            int idx = 0;
            ix->scr[k].each_buf([&](DevBuf *b) {
                const int me = idx++;
                if (!b->p || !b->guard) return;
                cudaMemset(bad, 0, 8);
                const unsigned char *base = static_cast<const unsigned char *>(b->p);
                check_guard_kernel<<<1, (int)kGuardBytes>>>(base - b->guard, base + b->cap, (int)kGuardBytes, bad);
                cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
                ++checked;
                if ((hb[0] || hb[1]) && which < 0) {
                    which = me;
                    slot_bad = k;
                    first[0] = hb[0];
                    first[1] = hb[1];
                }
            });
        }
        cudaFree(bad);
        CU(cudaGetLastError());
        if (which >= 0)
            return fail(SC_ERR_STATE, "scratch buffer #%d of slot %d was written out of bounds: %u guard bytes in front, %u behind", which,
                        slot_bad, first[0], first[1]);
        if (value > 0 && checked < value) return fail(SC_ERR_STATE, "only %d guarded buffers exist (expected at least %lld)", checked, (long long)value);
        return SC_OK;
    }
    if (strcmp(name, "tile_rem") == 0) {
        if (value != 0 && value != 4) return fail(SC_ERR_INVALID, "tile_rem: 0 (remainders of 9..16 queries may become tile items) or 4 (5..16)");
        ix->tile_rem = (int)value;
        return SC_OK;
    }
    if (strcmp(name, "mq_fused") == 0) {  // list-major page scans: 1 = both buckets in one launch (measured slower), 0 = two launches
        ix->mq_fused = value != 0;
        return SC_OK;
    }
    if (strcmp(name, "pdl") == 0) {  // small batches: programmatic dependent launch of the step's kernels (default on)
        ix->pdl = value != 0;
        return SC_OK;
    }
    if (strcmp(name, "fuse_plan") == 0) {  // batches of <= 16 queries: probe selection + pair plan in one launch (default on)
        ix->fuse_plan = value != 0;
        return SC_OK;
    }
    if (strcmp(name, "tc_variant") == 0) {
        if (value != 0 && value != 1) return fail(SC_ERR_INVALID, "tc_variant: 0 = 256x256 tiles, 1 = 128x256 tiles");
        ix->tc_variant = (int)value;
        return SC_OK;
    }
    if (strcmp(name, "coarse_impl") == 0) {
        if (value != 0 && value != 1) return fail(SC_ERR_INVALID, "coarse_impl: 0 = tcgen05 3xTF32, 1 = fp32 SIMT");
        ix->coarse_impl = (int)value;
        return SC_OK;
    }
    return fail(SC_ERR_INVALID, "unknown parameter '%s'", name);
}

}  // extern "C"
