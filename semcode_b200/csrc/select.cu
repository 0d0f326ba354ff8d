// K6: per-query top-k selection, one CTA per query.
//   k <= 128: ONE pass over the candidates with a threshold.  Every thread takes the maximum of the values dealt to it
//   (16-byte loads, the first 32 values stay in registers); the warp's ceil(k / 16)-th best thread maximum, minimised
//   over the 16 warps, is a bound T with at least k candidates >= T and -- candidates being dealt round-robin -- only a
//   few more (54 .. 130 of 16 384 for k = 10 .. 32).  The threads whose maximum reaches T put their values >= T into
//   shared memory; up to 128 gathered pairs are ranked by counting (no sort, no barrier ladder), more by a bitonic
//   sort.  More than 2048 values >= T (never seen outside the tests) falls back to the radix select.
//   (v1 always ran the radix select -- 4 reads of the candidates, 1.4 of 14 ms at nq 4096 / nprobe 128; v2 kept 4 pairs
//   per thread and sorted the 512 thread maxima: one CTA took 20 us for 16 384 scores -- 41 k cycles for 2 900
//   instructions per warp, a chain of 8 load rounds, 45 sort stages and 32 insertion networks -- which was 45 of the
//   72 us of an nq = 1 search and the second wave of every N = 8 step.)
//   Otherwise: 3-pass radix select (11/11/10 bits) over an order-preserving key finds the k-th best similarity
//   exactly, the winners are compacted into shared memory.
// Either way the winners are bitonic-sorted (key descending, candidate index ascending) and translated to ids.  Replaces FAISS HeapResultHandler / the Milvus segment reduce behind
// Collection.search(..., limit=top_k)  (reference src/semcode/storage/milvus_store.py:141-147).
//
// The same kernel serves three candidate sources:
//   RowsSrc   coarse similarities [nq, nlist]        -> the nprobe best lists per query (K1 tail)
//   CandSrc   candidate similarities written by K5   -> final (dist, id) per query
//   MergeSrc  partial results of several shards      -> merged (dist, id) (K7 tail)
#include <float.h>

#include "common.cuh"

namespace sc {

namespace {

constexpr int SEL_T = 512;
// Short candidate lists (an 8-way shard's 3072 candidates per query, a merge of 8 x k partials): a CTA's time is latency,
// not work, so 128-thread CTAs -- 8 resident per SM instead of 2 -- finish 1024 queries in one wave instead of 3.5.
constexpr int SEL_T_SMALL = 128;
constexpr uint32_t SEL_SMALL_MAX_N = 8192;
constexpr int SEL_BINS = 2048;

struct SelShared {
    uint32_t hist[SEL_BINS];
    unsigned long long pairs[kMaxK];
    uint32_t warp_tot[SEL_T / 32];
    uint32_t sel_bin, need, eq_total, cnt_gt, cnt_eq, base;
    uint32_t fast_cnt, fast_nq;
};
constexpr int SEL_FAST_K = 128;   // largest k of the one-pass selection
constexpr int SEL_RV = 8;         // 16-byte chunks per thread that stay in registers between the two phases
constexpr int SEL_RANK_MAX = 128; // gathered pairs ranked by counting; more are sorted

__device__ __forceinline__ unsigned long long pack_pair(uint32_t key, uint32_t idx) {
    return ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - idx);
}

// ---- exchange over peer-mapped memory ----------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Every thread of every CTA has issued its (remote) stores.  The last CTA to arrive publishes the epoch to
// every peer: fence.sys by each writer -> CTA barrier -> device-scope counter -> fence.sys -> release stores.
__device__ __forceinline__ void peer_signal(const PeerSignal &s) {
    if (s.world <= 0) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(s.done, 1u);
        if (prev == gridDim.x - 1) {
            atomicExch(s.done, 0u);  // ready for the next launch on this stream
            __threadfence_system();
            for (int p = 0; p < s.world; ++p) st_release_sys_u64(s.flag[p], s.epoch);
        }
    }
}

// Spin (thread p on peer p's flag word) until every peer has published `epoch`; bounded by timeout_ns so that a
// missing peer becomes an error status instead of a hung GPU.
__device__ __forceinline__ void peer_wait(const PeerWait &w) {
    if (w.world <= 0) return;
    if ((int)threadIdx.x < w.world) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys_u64(w.flags + threadIdx.x) < w.epoch) {
            if (global_timer_ns() - t0 > w.timeout_ns) {
                *(volatile unsigned int *)w.status = 1u;  // pinned host word (zero-copy): the host polls it without a sync
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

__global__ void peer_signal_kernel(PeerSignal s) { peer_signal(s); }
__global__ void peer_wait_kernel(PeerWait w) { peer_wait(w); }

// ---- candidate sources ----------------------------------------------------------------------
struct RowsSrc {
    const float *s;
    int N;
    int k;
    int32_t *out_idx;
    float *out_val;
    int world;       // > 0: write the row into every peer's probe table instead of out_idx
    PeerRows peers;
    struct View {
        const float *p;
        uint32_t n;
        __device__ __forceinline__ float load(uint32_t i) const { return __ldg(p + i); }
        __device__ __forceinline__ bool vec() const { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
        __device__ __forceinline__ float4 load4(uint32_t i4) const { return __ldg(reinterpret_cast<const float4 *>(p) + i4); }
    };
    __device__ __forceinline__ View view(int64_t q) const { return View{s + q * (int64_t)N, (uint32_t)N}; }
    __device__ __forceinline__ void emit(int64_t q, int j, bool valid, float score, uint32_t idx) const {
        if (world > 0) {
            const int64_t o = (peers.row0 + q) * k + j;
            for (int p = 0; p < world; ++p) peers.p[p][o] = valid ? (int32_t)idx : -1;
            return;
        }
        out_idx[q * k + j] = valid ? (int32_t)idx : -1;
        if (out_val) out_val[q * k + j] = valid ? score : -INFINITY;
    }
};

struct CandSrc {
    ScanArgs a;
    int k;
    float *out_dist;
    int64_t *out_ids;
    int world;       // > 0: store the result into every peer's gather slot of this rank instead of out_*
    PeerTopk peers;
    __device__ __forceinline__ void put(int64_t o, float d, int64_t id) const {
        if (world > 0) {
            for (int p = 0; p < world; ++p) {
                peers.d[p][o] = d;
                peers.i[p][o] = id;
            }
            return;
        }
        out_dist[o] = d;
        out_ids[o] = id;
    }
    struct View {
        const float *p;
        uint32_t n;
        __device__ __forceinline__ float load(uint32_t i) const { return __ldg(p + i); }
        __device__ __forceinline__ bool vec() const { return true; }  // whole pages of 32 floats
        __device__ __forceinline__ float4 load4(uint32_t i4) const { return __ldg(reinterpret_cast<const float4 *>(p) + i4); }
    };
    __device__ __forceinline__ View view(int64_t q) const {
        const int64_t b = a.page_off[q * a.nprobe], e = a.page_off[(q + 1) * a.nprobe];
        return View{a.cand + b * kPageRows, (uint32_t)((e - b) * kPageRows)};
    }
    __device__ __forceinline__ void emit(int64_t q, int j, bool valid, float score, uint32_t idx) const {
        const bool ip = a.metric == 0;
        if (!valid) {
            put(q * k + j, ip ? -FLT_MAX : FLT_MAX, -1);
            return;
        }
        const int64_t w = a.page_off[q * a.nprobe] + (idx >> 5);
        const int r = idx & 31;
        // last pair i in [q*nprobe, (q+1)*nprobe) with page_off[i] <= w
        int64_t lo = q * a.nprobe, hi = (q + 1) * a.nprobe;  // invariant: page_off[lo] <= w < page_off[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (a.page_off[mid] <= w)
                lo = mid;
            else
                hi = mid;
        }
        const int32_t l = a.probe[lo];
        const int32_t page = a.pt[a.pt_off[l] + (int32_t)(w - a.page_off[lo])];
        const int slab = page >> a.slab_shift;
        const int64_t slot = (int64_t)(page & ((1 << a.slab_shift) - 1)) * kPageRows + r;
        put(q * k + j, ip ? score : -score, a.slabs->ids[slab][slot]);
    }
};

struct MergeSrc {
    const float *pd;
    const int64_t *pi;
    int parts;
    int64_t nq;
    int kin;
    int k;
    int metric;
    float *out_dist;
    int64_t *out_ids;
    __device__ __forceinline__ int64_t off(int64_t q, uint32_t i) const {
        const int p = i / kin, j = i - p * kin;
        return ((int64_t)p * nq + q) * kin + j;
    }
    struct View {
        const MergeSrc *m;
        int64_t q;
        uint32_t n;
        __device__ __forceinline__ float load(uint32_t i) const {
            const int64_t o = m->off(q, i);
            if (__ldcg(m->pi + o) < 0) return -INFINITY;  // L2 loads: the partials may come from peers' stores
            const float d = __ldcg(m->pd + o);
            return m->metric == 0 ? d : -d;
        }
        __device__ __forceinline__ bool vec() const { return false; }
        __device__ __forceinline__ float4 load4(uint32_t) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
    };
    __device__ __forceinline__ View view(int64_t q) const { return View{this, q, (uint32_t)(parts * kin)}; }
    __device__ __forceinline__ void emit(int64_t q, int j, bool valid, float, uint32_t idx) const {
        if (!valid) {
            out_dist[q * k + j] = metric == 0 ? -FLT_MAX : FLT_MAX;
            out_ids[q * k + j] = -1;
            return;
        }
        const int64_t o = off(q, idx);
        out_dist[q * k + j] = __ldcg(pd + o);
        out_ids[q * k + j] = __ldcg(pi + o);
    }
};

// inclusive warp scan
template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

static_assert(SEL_FAST_K <= SEL_T && SEL_RANK_MAX <= SEL_T_SMALL && SEL_BINS % SEL_T_SMALL == 0,
              "one-pass selection: one gathered pair per thread in the ranking step; radix bins dealt evenly to the threads");

// bitonic sort (descending) of pairs[0, P), P a power of two, by the whole CTA
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long *pairs, uint32_t P) {
    const uint32_t tid = threadIdx.x;
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (P >> 1); t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = pairs[lo], y = pairs[hi];
                if ((x < y) == desc) {
                    pairs[lo] = y;
                    pairs[hi] = x;
                }
            }
            __syncthreads();
        }
    }
}

// One pass over the view.  Returns true with every candidate >= T (at least kk of them when the view holds that many,
// unsorted) in sh.pairs[0, *count); false (uniformly for the CTA) when more than kMaxK candidates reached T.
template <int NT, class View>
__device__ __forceinline__ bool select_one_pass(const View &view, uint32_t n, uint32_t kk, SelShared &sh, uint32_t *count) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float NEG = -INFINITY;
    const bool vec = view.vec();
    const uint32_t n4 = vec ? (n >> 2) : 0u;  // 16-byte chunks; the tail (and a view without vector loads) goes scalar
    if (tid == 0) {
        sh.fast_cnt = 0;
        sh.fast_nq = 0;
    }

    // phase 1: the maximum key of the values dealt to this thread
    float4 r[SEL_RV];
#pragma unroll
    for (int j = 0; j < SEL_RV; ++j) {
        const uint32_t i4 = tid + (uint32_t)j * NT;
        r[j] = i4 < n4 ? view.load4(i4) : make_float4(NEG, NEG, NEG, NEG);
    }
    uint32_t mx = 0;
#pragma unroll
    for (int j = 0; j < SEL_RV; ++j)
        mx = max(max(mx, max(f2key(r[j].x), f2key(r[j].y))), max(f2key(r[j].z), f2key(r[j].w)));
    {
        uint32_t i4 = tid + (uint32_t)SEL_RV * NT;
        for (; i4 + 3 * NT < n4; i4 += 4 * NT) {  // four loads in flight per thread
            const float4 v0 = view.load4(i4), v1 = view.load4(i4 + NT), v2 = view.load4(i4 + 2 * NT), v3 = view.load4(i4 + 3 * NT);
            mx = max(max(mx, max(f2key(v0.x), f2key(v0.y))), max(f2key(v0.z), f2key(v0.w)));
            mx = max(max(mx, max(f2key(v1.x), f2key(v1.y))), max(f2key(v1.z), f2key(v1.w)));
            mx = max(max(mx, max(f2key(v2.x), f2key(v2.y))), max(f2key(v2.z), f2key(v2.w)));
            mx = max(max(mx, max(f2key(v3.x), f2key(v3.y))), max(f2key(v3.z), f2key(v3.w)));
        }
        for (; i4 < n4; i4 += NT) {
            const float4 v = view.load4(i4);
            mx = max(max(mx, max(f2key(v.x), f2key(v.y))), max(f2key(v.z), f2key(v.w)));
        }
        uint32_t i = 4 * n4 + tid;
        for (; i + 3 * NT < n; i += 4 * NT) {
            const float v0 = view.load(i), v1 = view.load(i + NT), v2 = view.load(i + 2 * NT), v3 = view.load(i + 3 * NT);
            mx = max(max(mx, max(f2key(v0), f2key(v1))), max(f2key(v2), f2key(v3)));
        }
        for (; i < n; i += NT) mx = max(mx, f2key(view.load(i)));
    }

    // T: every warp holds at least jstar thread maxima >= its jstar-th best one, so the minimum of those over the warps
    // has at least (NT / 32) * jstar >= kk candidates at or above it.  A view that fits the gather buffer needs no bound.
    uint32_t T = kKeyNegInf + 1u;
    if (n > (uint32_t)kMaxK) {
        const uint32_t jstar = (kk + NT / 32 - 1) / (NT / 32);
        uint32_t v = mx, tw = 0;
        for (uint32_t j = 0; j < jstar; ++j) {
            tw = __reduce_max_sync(0xffffffffu, v);
            const uint32_t holders = __ballot_sync(0xffffffffu, v == tw);
            if (lane == (uint32_t)(__ffs(holders) - 1)) v = 0;  // one holder leaves the pool per round
        }
        if (lane == 0) sh.warp_tot[warp] = tw;
        __syncthreads();
        uint32_t t = sh.warp_tot[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) t = min(t, sh.warp_tot[w]);
        T = max(T, t);
    } else {
        __syncthreads();  // fast_cnt is zero
    }

    // phase 2: gather the values >= T.  Only the few threads whose maximum reached T can hold one.  Their first SEL_RV chunks are
    // still in registers; what they were dealt beyond that is re-read -- by the WHOLE CTA, one (qualifying thread, chunk) pair per
    // thread, so it is one round trip.  (First version: each qualifying thread re-read its own chunks one after the other:
    // 31 dependent loads at 82k candidates -- 30 of a CTA's 42 us, half of the warp lanes idle over the whole kernel by ncu.)
    auto put = [&](float f, uint32_t idx) {
        const uint32_t key = f2key(f);
        if (key >= T) {
            const uint32_t s = atomicAdd(&sh.fast_cnt, 1u);
            if (s < (uint32_t)kMaxK) sh.pairs[s] = pack_pair(key, idx);
        }
    };
    const uint32_t chunks_per_thread = (n4 + NT - 1) / NT;  // chunks dealt to a thread (the last one may fall off the end)
    const bool reread = chunks_per_thread > (uint32_t)SEL_RV;
    if (mx >= T) {
#pragma unroll
        for (int j = 0; j < SEL_RV; ++j) {
            const uint32_t i4 = tid + (uint32_t)j * NT;
            if (i4 < n4) {
                put(r[j].x, 4 * i4);
                put(r[j].y, 4 * i4 + 1);
                put(r[j].z, 4 * i4 + 2);
                put(r[j].w, 4 * i4 + 3);
            }
        }
        if (reread) {
            const uint32_t slot = atomicAdd(&sh.fast_nq, 1u);
            if (slot < (uint32_t)SEL_BINS) sh.hist[slot] = tid;  // the histogram is idle on this path
        }
    }
    if (reread) {
        __syncthreads();
        const uint32_t nqual = sh.fast_nq;
        if (nqual > (uint32_t)SEL_BINS) {  // (cannot happen with NT <= SEL_BINS threads; kept as the overflow rule)
            *count = kMaxK + 1;
            return false;
        }
        const uint32_t per = chunks_per_thread - (uint32_t)SEL_RV;  // chunks to re-read per qualifying thread
        for (uint32_t w = tid; w < nqual * per; w += NT) {
            const uint32_t i4 = sh.hist[w / per] + ((uint32_t)SEL_RV + w % per) * NT;
            if (i4 < n4) {
                const float4 v = view.load4(i4);
                put(v.x, 4 * i4);
                put(v.y, 4 * i4 + 1);
                put(v.z, 4 * i4 + 2);
                put(v.w, 4 * i4 + 3);
            }
        }
    }
    // values outside the 16-byte chunks (the tail; everything, for a view without vector loads): every thread checks its own
    {
        uint32_t i = 4 * n4 + tid;
        for (; i + 3 * NT < n; i += 4 * NT) {
            const float v0 = view.load(i), v1 = view.load(i + NT), v2 = view.load(i + 2 * NT), v3 = view.load(i + 3 * NT);
            put(v0, i);
            put(v1, i + NT);
            put(v2, i + 2 * NT);
            put(v3, i + 3 * NT);
        }
        for (; i < n; i += NT) put(view.load(i), i);
    }
    __syncthreads();
    *count = sh.fast_cnt;
    return sh.fast_cnt <= (uint32_t)kMaxK;
}

template <int NT, class Src>
__device__ __forceinline__ void select_topk_body(const Src &src, int k, SelShared &sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const typename Src::View view = src.view(q);
    const uint32_t n = view.n;
    const uint32_t kk = (uint32_t)k < n ? (uint32_t)k : n;

    if (kk == 0) {
        for (int j = tid; j < k; j += NT) src.emit(q, j, false, 0.f, 0);
        return;
    }

    uint32_t n_sort = 0;  // pairs in sh.pairs to sort; the first min(n_sort, kk) are the result
    bool selected = false;
    if (kk <= (uint32_t)SEL_FAST_K) {
        selected = select_one_pass<NT>(view, n, kk, sh, &n_sort);
        __syncthreads();  // everybody has read the verdict before the radix path reuses the shared fields
    }
    if (!selected) {
        uint32_t prefix = 0, mask = 0, need = kk;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
            const uint32_t nb = pass == 2 ? 1024u : 2048u;
            for (int b = tid; b < SEL_BINS; b += NT) sh.hist[b] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < n; i += NT) {
                const uint32_t key = f2key(view.load(i));
                if ((key & mask) == prefix) atomicAdd(&sh.hist[(key >> shift) & (nb - 1)], 1u);
            }
            __syncthreads();
            // thread t owns bins BPT t .. BPT t + BPT - 1; find the bin (from the top) that holds the need-th element
            constexpr int BPT = SEL_BINS / NT;
            uint32_t c[BPT];
            uint32_t local = 0;
#pragma unroll
            for (int b = 0; b < BPT; ++b) {
                c[b] = sh.hist[BPT * tid + b];
                local += c[b];
            }
            const uint32_t incl = warp_incl_scan(local, lane);
            if (lane == 31) sh.warp_tot[warp] = incl;
            __syncthreads();
            uint32_t wprefix = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) {
                const uint32_t t = sh.warp_tot[w];
                if (w < warp) wprefix += t;
                total += t;
            }
            uint32_t running = total - (wprefix + incl);  // elements in bins above this thread's bins
#pragma unroll
            for (int b = BPT - 1; b >= 0; --b) {
                if (running < need && running + c[b] >= need) {
                    sh.sel_bin = BPT * tid + b;
                    sh.need = need - running;
                    sh.eq_total = c[b];
                }
                running += c[b];
            }
            __syncthreads();
            prefix |= sh.sel_bin << shift;
            mask |= (nb - 1) << shift;
            need = sh.need;
            __syncthreads();
        }

        const uint32_t T = prefix;           // key of the kk-th best candidate
        const uint32_t n_gt = kk - need;     // candidates strictly better than T
        const uint32_t eq_total = sh.eq_total;
        const bool take_eq = T > kKeyNegInf;  // keys <= key(-inf) are "no result"
        const bool ordered = take_eq && eq_total > need;
        if (tid == 0) {
            sh.cnt_gt = 0;
            sh.cnt_eq = 0;
            sh.base = 0;
        }
        __syncthreads();
        for (uint32_t i = tid; i < n; i += NT) {
            const uint32_t key = f2key(view.load(i));
            if (key > T) {
                const uint32_t s = atomicAdd(&sh.cnt_gt, 1u);
                sh.pairs[s] = pack_pair(key, i);
            } else if (key == T && take_eq && !ordered) {
                const uint32_t s = atomicAdd(&sh.cnt_eq, 1u);
                if (s < need) sh.pairs[n_gt + s] = pack_pair(key, i);
            }
        }
        __syncthreads();
        if (ordered) {
            // more candidates tie with the k-th than there is room for: keep the lowest indices
            for (uint32_t c0 = 0; c0 < n; c0 += NT) {
                const uint32_t base = sh.base;
                if (base >= need) break;
                const uint32_t i = c0 + tid;
                const bool flag = i < n && f2key(view.load(i)) == T;
                const uint32_t bal = __ballot_sync(0xffffffffu, flag);
                if (lane == 0) sh.warp_tot[warp] = __popc(bal);
                __syncthreads();
                uint32_t wprefix = 0, total = 0;
#pragma unroll
                for (int w = 0; w < NT / 32; ++w) {
                    const uint32_t t = sh.warp_tot[w];
                    if (w < warp) wprefix += t;
                    total += t;
                }
                const uint32_t rank = base + wprefix + __popc(bal & ((1u << lane) - 1u));
                if (flag && rank < need) sh.pairs[n_gt + rank] = pack_pair(T, i);
                __syncthreads();
                if (tid == 0) sh.base = base + total;
                __syncthreads();
            }
        }
        n_sort = n_gt + (take_eq ? need : 0u);

    }
    const uint32_t n_valid = n_sort < kk ? n_sort : kk;

    if (n_sort <= (uint32_t)SEL_RANK_MAX) {
        // few winners: the rank of a pair is the number of pairs above it (pairs are distinct: the index is part of them)
        if ((uint32_t)tid < n_sort) {
            const unsigned long long mine = sh.pairs[tid];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n_sort; ++j) rank += sh.pairs[j] > mine ? 1u : 0u;
            if (rank < n_valid) src.emit(q, (int)rank, true, key2f((uint32_t)(mine >> 32)), 0xffffffffu - (uint32_t)mine);
        }
        for (int j = (int)n_valid + tid; j < k; j += NT) src.emit(q, j, false, 0.f, 0);
        return;
    }

    // bitonic sort (descending) of the winners
    uint32_t P = 1;
    while (P < n_sort) P <<= 1;
    for (uint32_t i = n_sort + tid; i < P; i += NT) sh.pairs[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(sh.pairs, P);
    for (int j = tid; j < k; j += NT) {
        if ((uint32_t)j < n_valid) {
            const unsigned long long p = sh.pairs[j];
            src.emit(q, j, true, key2f((uint32_t)(p >> 32)), 0xffffffffu - (uint32_t)p);
        } else {
            src.emit(q, j, false, 0.f, 0);
        }
    }
}

// wait: peers whose stores this kernel consumes (merge); sig: peers that consume this kernel's stores
// NT = threads per CTA: SEL_T, or SEL_T_SMALL for short candidate lists (see launch_select)
template <class Src, int NT>
// (512 threads: no minimum -- 64 registers, two CTAs per SM; capping at 40 registers for three made the probe selection 44 -> 76 us.  128 threads:
//  eight CTAs per SM at 64 registers: 8-way shard top-k 22 -> 16 us.)
__global__ void __launch_bounds__(NT, NT == 512 ? 0 : 8) select_topk_kernel(Src src, int k, PeerWait wait, PeerSignal sig) {
    __shared__ SelShared sh;
    pdl_launch_dependents();
    pdl_wait();
    peer_wait(wait);
    select_topk_body<NT>(src, k, sh);
    peer_signal(sig);
}

// ---- small batches: probe selection and pair plan in one launch ------------------------------------------------------
// One CTA per query ranks the centroids (as above), then turns its own probe row into page counts and their exclusive
// prefix; the CTA that finishes last (a ticket) adds the per-query bases and writes the grand total -- what
// plan_pairs_kernel (scan.cu) computes with its own launch and look-back chain.  For nq <= kPlanTailMaxQ a launch boundary
// costs more than the plan itself (5.7 of the 53 us of an nq = 1 search).
//   ws: [0] ticket (u32, zero between launches) | [1 ..] per-query page totals (i64), kPlanTailWords 64-bit words
__global__ void __launch_bounds__(SEL_T) select_rows_plan_kernel(RowsSrc src, int k, const int32_t *__restrict__ list_len, int32_t nlist,
                                                                  int64_t *__restrict__ page_off, unsigned long long *__restrict__ ws,
                                                                  unsigned long long *__restrict__ rows_total) {
    __shared__ SelShared sh;
    __shared__ uint32_t s_last;
    pdl_launch_dependents();
    pdl_wait();
    select_topk_body<SEL_T>(src, k, sh);
    __syncthreads();  // the CTA's own stores of the probe row are visible to the CTA
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x, nq = gridDim.x;
    int32_t pages = 0, len = 0;
    if (tid < k) {
        const int32_t l = __ldcg(src.out_idx + q * k + tid);
        if (l >= 0 && l < nlist) {
            len = __ldg(list_len + l);
            pages = (len + kPageRows - 1) / kPageRows;
        }
    }
    // exclusive prefix over the k <= 128 entries (warps 0..3)
    int32_t incl = pages;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh.warp_tot[warp] = (uint32_t)incl;
    if (rows_total != nullptr && warp < (k + 31) / 32) {
        unsigned long long r = (unsigned long long)len;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0 && r) atomicAdd(rows_total, r);
    }
    __syncthreads();
    int64_t wpre = 0, total = 0;
    for (int w = 0; w < (k + 31) / 32; ++w) {
        const int64_t t = sh.warp_tot[w];
        if (w < warp) wpre += t;
        total += t;
    }
    int64_t *qtot = reinterpret_cast<int64_t *>(ws + 1);
    if (nq == 1) {
        if (tid < k) page_off[tid] = wpre + incl - pages;
        if (tid == 0) page_off[k] = total;
        return;
    }
    if (tid < k) page_off[q * k + tid] = wpre + incl - pages;  // offsets inside this query's pairs; the base comes last
    if (tid == 0) qtot[q] = total;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(reinterpret_cast<unsigned int *>(ws), 1u) == (unsigned int)(nq - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int64_t base = 0;  // thread t serves entries t, t + SEL_T, ...: query = entry / k
    for (int64_t e = tid; e < nq * k; e += SEL_T) {
        const int64_t qq = e / k;
        base = 0;
        for (int64_t j = 0; j < qq; ++j) base += __ldcg(qtot + j);
        if (qq > 0) page_off[e] = __ldcg(page_off + e) + base;
    }
    if (tid == 0) {
        int64_t all = 0;
        for (int64_t j = 0; j < nq; ++j) all += __ldcg(qtot + j);
        page_off[nq * k] = all;
        *reinterpret_cast<unsigned int *>(ws) = 0u;  // ready for the next launch on this scratch
    }
}

// ---- exclusive prefix sums (single CTA; inputs are at most nq*nprobe or nlist long) -----------
template <typename T>
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const T *__restrict__ in, int64_t n, T *__restrict__ out) {
    __shared__ T warp_tot[32];
    __shared__ T blk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T carry = 0;
    constexpr int IPT = 4;
    for (int64_t base = 0; base < n; base += 1024 * IPT) {
        const int64_t i0 = base + (int64_t)tid * IPT;
        T v[IPT];
        T local = 0;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            v[j] = (i0 + j < n) ? in[i0 + j] : (T)0;
            local += v[j];
        }
        const T incl = warp_incl_scan(local, lane);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const T t = warp_tot[lane];
            const T ti = warp_incl_scan(t, lane);
            warp_tot[lane] = ti - t;
            if (lane == 31) blk_total = ti;
        }
        __syncthreads();
        T run = carry + warp_tot[warp] + incl - local;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            if (i0 + j < n) out[i0 + j] = run;
            run += v[j];
        }
        carry += blk_total;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

}  // namespace

// n_max: an upper bound of the candidates per query (picks the CTA size)
template <class Src>
static cudaError_t launch_select(const Src &src, int64_t nq, int k, uint64_t n_max, const PeerWait &wait, const PeerSignal &sig,
                                 cudaStream_t st, bool pdl) {
    if (n_max <= SEL_SMALL_MAX_N && nq >= 64)
        return launch_pdl(select_topk_kernel<Src, SEL_T_SMALL>, dim3((unsigned)nq), dim3(SEL_T_SMALL), 0, st, pdl, src, k, wait, sig);
    return launch_pdl(select_topk_kernel<Src, SEL_T>, dim3((unsigned)nq), dim3(SEL_T), 0, st, pdl, src, k, wait, sig);
}

cudaError_t launch_select_rows(const float *scores, int64_t M, int N, int k, int32_t *out_idx, float *out_val,
                               cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    RowsSrc src{scores, N, k, out_idx, out_val, 0, PeerRows{}};
    return launch_select(src, M, k, (uint64_t)N, PeerWait{}, PeerSignal{}, st, false);
}

cudaError_t launch_select_rows_plan(const float *scores, int64_t M, int N, int k, int32_t *out_idx, const int32_t *list_len,
                                    int32_t nlist, int64_t *page_off, unsigned long long *ws, unsigned long long *rows_total,
                                    cudaStream_t st, bool pdl) {
    if (M <= 0 || M > kPlanTailMaxQ || k > SEL_FAST_K) return cudaErrorInvalidValue;
    RowsSrc src{scores, N, k, out_idx, nullptr, 0, PeerRows{}};
    return launch_pdl(select_rows_plan_kernel, dim3((unsigned)M), dim3(SEL_T), 0, st, pdl, src, k, list_len, nlist, page_off, ws,
                      rows_total);
}

cudaError_t launch_select_rows_peers(const float *scores, int64_t M, int N, int k, const PeerRows &rows,
                                     const PeerSignal &sig, cudaStream_t st) {
    if (M <= 0) return launch_peer_signal(sig, st);
    RowsSrc src{scores, N, k, nullptr, nullptr, sig.world, rows};
    return launch_select(src, M, k, (uint64_t)N, PeerWait{}, sig, st, false);
}

cudaError_t launch_peer_signal(const PeerSignal &sig, cudaStream_t st) {
    peer_signal_kernel<<<1, 32, 0, st>>>(sig);
    return cudaGetLastError();
}

cudaError_t launch_peer_wait(const PeerWait &wait, cudaStream_t st) {
    peer_wait_kernel<<<1, 32, 0, st>>>(wait);
    return cudaGetLastError();
}

cudaError_t launch_select_candidates(const ScanArgs &a, int64_t nq, int k, float *out_dist, int64_t *out_ids,
                                     cudaStream_t st, bool pdl) {
    if (nq <= 0) return cudaSuccess;
    CandSrc src{a, k, out_dist, out_ids, 0, PeerTopk{}};
    return launch_select(src, nq, k, (uint64_t)a.max_cand, PeerWait{}, PeerSignal{}, st, pdl);
}

cudaError_t launch_select_candidates_peers(const ScanArgs &a, int64_t nq, int k, const PeerTopk &out, const PeerSignal &sig,
                                           cudaStream_t st) {
    if (nq <= 0) return launch_peer_signal(sig, st);
    CandSrc src{a, k, nullptr, nullptr, sig.world, out};
    return launch_select(src, nq, k, (uint64_t)a.max_cand, PeerWait{}, sig, st, false);
}

cudaError_t launch_merge_topk(const float *part_dist, const int64_t *part_ids, int parts, int64_t nq, int kin, int k,
                              int metric, float *out_dist, int64_t *out_ids, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    MergeSrc src{part_dist, part_ids, parts, nq, kin, k, metric, out_dist, out_ids};
    return launch_select(src, nq, k, (uint64_t)parts * (uint64_t)kin, PeerWait{}, PeerSignal{}, st, false);
}

cudaError_t launch_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out, cudaStream_t st) {
    exclusive_scan_kernel<int32_t><<<1, 1024, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

}  // namespace sc
