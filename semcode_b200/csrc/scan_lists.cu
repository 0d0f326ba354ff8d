// K5b: list-major inverted-list scan for large query batches.
//
// The query-major kernel (scan.cu) streams a probed list once per (query, list) pair; with
// nq * nprobe >> nlist every list is probed many times (32x at nq 4096, nprobe 128, nlist 16384)
// and the batch becomes a stack of small dense contractions.  Here the (query, list) pairs are
// inverted to (list -> queries), each probed list is read from HBM ONCE and scored against all the
// queries that probe it in exact fp32 (FFMA, direct forms sum(x*q) / sum((x-q)^2): same 1e-5 relative
// bar as the query-major path, which tensor-core TF32 would not hold for L2).  Replaces the same
// FAISS IVFFlatScanner::scan_codes loop (reference src/semcode/storage/milvus_store.py:141-147).
//
// Work item = (list, chunk of <= QT queries); CTAs pull items from an atomic counter.  Per item the
// CTA walks the list in tiles of 128 rows x 64 floats through a cp.async ring (16-byte copies,
// zero-fill for ragged edges) and keeps a register tile of RPT rows x 8 queries per thread, as packed
// fp32 pairs (FFMA2):
//   QT = 32  lists probed by many queries   FP32-pipe bound (measured ceiling ~37 TFLOP/s on B200)
//   QT = 8   lists probed by <= 8 queries   HBM bound
// The result lands in the same per-pair candidate layout the query-major kernel writes, so the
// top-k selection (select.cu) is unchanged.
// Algorithmic bytes: sum over DISTINCT probed lists of len * 4 * dim (compulsory), per chunk item.
#include "common.cuh"

namespace sc {

namespace {

constexpr int NT = 256;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// ---- planning: invert (query, list) pairs into per-list query groups ----------------------------------
// cnt[l] = queries probing list l; *over8 = lists probed by more than 8 queries (the plan's tile-item threshold)
__global__ void count_list_pairs_kernel(const int32_t *__restrict__ probe, int64_t npairs, const int32_t *__restrict__ list_len,
                                        int32_t nlist, int32_t *__restrict__ cnt, int32_t *__restrict__ over8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int32_t l = probe[i];
    if (l >= 0 && l < nlist && list_len[l] > 0) {
        const int32_t before = atomicAdd(cnt + l, 1);
        if (before == 8) atomicAdd(over8, 1);
        if (before == 4) atomicAdd(over8 + 2, 1);  // counters[3]: lists probed by more than 4 queries
    }
}

// Plans every list.  A list probed by c queries becomes c / chunk tile items of `chunk` queries (32 for the
// FFMA tiles, 64 for the tcgen05 tiles) plus a remainder:
//   rem > T      one more (ragged) tile item; T = 16 for the FFMA tiles, 8 for the tcgen05 tiles
//   rem 5..T     ceil(rem / 8) passes of the 8-query page scan
//   rem 1..4     one pass of the 4-query page scan
// and four exclusive prefix sums are produced in the same sweep: lq_off (queries per list), off32 (tile items; for
// the tcgen05 tiles: items x 128-row tiles, the unit its CTAs share out in equal ranges),
// pg8off / pg4off (page x pass units of the two page scans).
// (First version: seven launches; second: ONE CTA for all lists -- 64k warp instructions on a single SM, 47 us at
// 16384 lists.)  Now 1024 lists per CTA and a single-pass scan: a CTA takes a ticket (its position in scheduling
// order, so it only waits for CTAs that are already running), publishes its four totals as self-validating words
// (1 << 32 | total) in `agg` (zeroed by the plan's one memset) and adds up the words of ALL its predecessors (16
// CTAs at 16384 lists: one load per lane; no chaining, every CTA publishes before it waits).  9 us at 16384 lists.
constexpr int PL_T = 256, PL_IPT = 4, PL_BLK = PL_T * PL_IPT;

__device__ __forceinline__ unsigned long long pl_ld(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(PL_T) plan_lists_kernel(const int32_t *__restrict__ cnt, const int32_t *__restrict__ list_len,
                                                          int32_t nlist, int32_t chunk, int32_t min_items, int32_t tile_rem,
                                                          int32_t *__restrict__ counters, unsigned long long *__restrict__ agg,
                                                          int32_t *__restrict__ n32, int32_t *__restrict__ lq_off,
                                                          int32_t *__restrict__ off32, int32_t *__restrict__ pg8off,
                                                          int32_t *__restrict__ pg4off,
                                                          unsigned long long *__restrict__ unique_rows) {
    __shared__ int32_t warp_tot[4][PL_T / 32];
    __shared__ int32_t s_pre[4];
    __shared__ int32_t s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(counters + 2, 1);
    __syncthreads();
    const int32_t b = s_ticket;
    // a ragged tile item costs ~2.8 list reads of time on the FFMA tiles (two passes of 8 are cheaper up to 16
    // queries) but about 1.6 on the tcgen05 tiles (cheaper than two passes from 9 queries on) -- provided there are
    // enough items to fill the GPU: an item is walked by ONE CTA (~100 us), so a handful of them is a pure tail
    // tile_rem = 4 (set_param "tile_rem"): remainders of 5..16 queries become tile items too when enough lists have them
    const int32_t rem_tile = chunk != 64 ? 16 : ((tile_rem == 4 && counters[3] >= min_items) ? 4 : (counters[1] >= min_items ? 8 : 16));
    const int32_t i0 = b * PL_BLK + tid * PL_IPT;
    int32_t v[4][PL_IPT];
    int32_t local[4] = {0, 0, 0, 0};
    unsigned long long rows = 0;
#pragma unroll
    for (int j = 0; j < PL_IPT; ++j) {
        const int32_t l = i0 + j;
        int32_t c = 0, a = 0, u8 = 0, u4 = 0, tiles = 1;
        if (l < nlist) {
            c = cnt[l];
            const int32_t len = list_len[l];
            const int32_t pages = (len + kPageRows - 1) / kPageRows;
            if (chunk == 64) tiles = (len + 127) / 128;
            if (c > 0) rows += (unsigned long long)len;
            a = c / chunk;
            const int32_t rem = c - a * chunk;
            if (rem > rem_tile)
                ++a;
            else if (rem > 4)
                u8 = ((rem + 7) / 8) * pages;
            else if (rem > 0)
                u4 = pages;
            n32[l] = a;
        }
        v[0][j] = c;
        v[1][j] = a * tiles;  // FFMA tiles: items; tcgen05 tiles: (item, 128-row tile) units
        v[2][j] = u8;
        v[3][j] = u4;
#pragma unroll
        for (int t = 0; t < 4; ++t) local[t] += v[t][j];
    }
    int32_t incl[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int32_t x = local[t];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        incl[t] = x;
        if (lane == 31) warp_tot[t][warp] = x;
    }
    __syncthreads();
    int32_t wpre[4], total[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        wpre[t] = 0;
        total[t] = 0;
#pragma unroll
        for (int w = 0; w < PL_T / 32; ++w) {
            const int32_t x = warp_tot[t][w];
            if (w < warp) wpre[t] += x;
            total[t] += x;
        }
    }
    if (warp == 0) {
        if (lane < 4) {
            int32_t mine = total[0];
#pragma unroll
            for (int t = 1; t < 4; ++t)
                if (lane == t) mine = total[t];
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(agg + 4 * (size_t)b + lane),
                         "l"((1ull << 32) | (unsigned long long)(uint32_t)mine)
                         : "memory");
        }
        int32_t pre[4] = {0, 0, 0, 0};
        for (int32_t j = lane; j < b; j += 32) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                unsigned long long w;
                do {
                    w = pl_ld(agg + 4 * (size_t)j + t);
                } while ((w >> 32) == 0);
                pre[t] += (int32_t)(uint32_t)w;
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pre[t] += __shfl_xor_sync(0xffffffffu, pre[t], o);
            if (lane == 0) s_pre[t] = pre[t];
        }
    }
    __syncthreads();
    int32_t *const outs[4] = {lq_off, off32, pg8off, pg4off};
    const bool last = b == (int32_t)gridDim.x - 1 && tid == PL_T - 1;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int32_t run = s_pre[t] + wpre[t] + incl[t] - local[t];
#pragma unroll
        for (int j = 0; j < PL_IPT; ++j) {
            if (i0 + j < nlist) outs[t][i0 + j] = run;
            run += v[t][j];
        }
        if (last) outs[t][nlist] = s_pre[t] + total[t];
    }
    if (unique_rows != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
        if (lane == 0 && rows) atomicAdd(unique_rows, rows);
    }
}

__global__ void fill_list_pairs_kernel(const int32_t *__restrict__ probe, int64_t npairs, const int32_t *__restrict__ list_len,
                                       int32_t nlist, const int32_t *__restrict__ lq_off, int32_t *__restrict__ cursor,
                                       int32_t *__restrict__ lq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int32_t l = probe[i];
    if (l >= 0 && l < nlist && list_len[l] > 0) lq[lq_off[l] + atomicAdd(cursor + l, 1)] = (int32_t)i;
}

// ---- the tile kernel ------------------------------------------------------------------------------------
// 256 threads = 8 warps.  Lanes run over ROWS (row = lane + 32*i), so an x fragment load is a conflict-free
// 512-byte LDS.128 and a q fragment load is a single-wavefront broadcast; each thread owns RPT rows x 8
// queries.  The k range of a stage is split over KS warp groups (split-K inside the CTA) so that all
// 8 warps have work on one 128-row tile; the KS partial tiles are summed through shared memory once
// per row tile.
//   QT = 32: KS = 2, RPT = 4 (warp = 128 rows x 8 queries), 5.3 FFMA per shared-memory wavefront
//   QT =  8: KS = 4, RPT = 2 (warp =  64 rows x 8 queries), HBM-bound regime
template <int QT, int BKX, int NS, int RB, int RPT_>
struct TileCfg {
    static constexpr int QPT = QT < 8 ? QT : 8;       // queries per thread
    static constexpr int QG = QT / QPT;               // query groups
    static constexpr int RPT = RPT_;                  // rows per thread
    static constexpr int RG = RB / (32 * RPT);        // row groups
    static constexpr int WPK = QG * RG;               // warps per k-split
    static constexpr int KS = 8 / WPK;                // k-splits
    static constexpr int LD = BKX + 4;                // padded row stride (floats): rows shift 4 banks
    static constexpr int K4S = (BKX / 4) / KS;        // k4-steps per warp per stage
    static constexpr int RING = 4;                    // row-table ring (tiles)
};

template <int QT, int BKX, int NS, int RB, int RPT>
struct TileSmem {
    using C = TileCfg<QT, BKX, NS, RB, RPT>;
    static_assert(C::KS >= 1 && C::K4S >= 1 && C::KS * C::WPK == 8 && (256 % (BKX / 4)) == 0, "bad tile configuration");
    float xs[NS][RB][C::LD];
    float qs[NS][QT][C::LD];
    float red[C::KS > 1 ? C::KS - 1 : 1][QT][RB];
    const float *rowptr[C::RING][RB];
    uint8_t live[C::RING][RB];
    const float *qptr[QT];
    int64_t cbase[QT];
    int32_t item;
};

// last index i in [0, n) with off[i] <= v   (off is an exclusive prefix: off[n] = total)
__device__ __forceinline__ int32_t owner_of(const int32_t *__restrict__ off, int32_t n, int32_t v) {
    int32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (off[mid] <= v)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// packed fp32 pairs (sm_100 FFMA2): two exact fp32 FMAs per issued instruction
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r, m1;
    asm("mov.b64 %0, {%1, %1};" : "=l"(m1) : "f"(-1.0f));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(b), "l"(m1), "l"(a));  // a - b, exactly rounded
    return r;
}
__device__ __forceinline__ float sum2(unsigned long long v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}

template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int QT, int BKX, int NS, int RB, int RPT, bool L2>
__global__ void __launch_bounds__(NT, 2) scan_lists_kernel(const ScanArgs a, const ListPlan p, int which) {
    using C = TileCfg<QT, BKX, NS, RB, RPT>;
    using SM = TileSmem<QT, BKX, NS, RB, RPT>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ks = warp / C::WPK;                   // k-split of this warp
    const int qg = (warp % C::WPK) % C::QG;         // query group (8 queries)
    const int rg = (warp % C::WPK) / C::QG;         // row group
    const int row0 = rg * 32 * C::RPT + lane;       // rows row0 + 32*i
    const int32_t *item_off = p.off32;
    const int32_t total = item_off[p.nlist];
    const int ds = a.ds;
    const int KB = (ds + BKX - 1) / BKX;
    const int slab_mask = (1 << a.slab_shift) - 1;
    constexpr int CPR = BKX / 4;  // 16-byte chunks per row segment

    for (;;) {
        __syncthreads();  // previous item fully consumed (smem tables reused)
        if (tid == 0) sm.item = atomicAdd(p.counters + which, 1);
        __syncthreads();
        const int32_t item = sm.item;
        if (item >= total) break;
        const int32_t l = owner_of(item_off, p.nlist, item);
        const int32_t chunk = item - item_off[l];
        const int32_t qbase = p.lq_off[l] + 32 * chunk;
        const int32_t nqi = min(QT, p.lq_off[l + 1] - qbase);
        const int32_t len = a.list_len[l];
        const int32_t ptbase = a.pt_off[l];
        const int32_t slots = ((len + kPageRows - 1) / kPageRows) * kPageRows;  // candidate slots of this list
        const int ntiles = (len + RB - 1) / RB;
        const int S = ntiles * KB;

        if (tid < QT) {
            if (tid < nqi) {
                const int32_t pair = p.lq[qbase + tid];
                sm.qptr[tid] = a.q + (int64_t)(pair / a.nprobe) * ds;
                sm.cbase[tid] = a.page_off[pair] * kPageRows;
            } else {
                sm.qptr[tid] = a.q;
                sm.cbase[tid] = -1;
            }
        }
        auto prep_rows = [&](int tile) {  // row pointers + fused predicate for one 128-row tile
            if (tid < RB) {
                const int32_t r = tile * RB + tid;
                const float *ptr = a.q;  // any valid address; never dereferenced when r >= len
                bool ok = false;
                if (r < len) {
                    const int32_t page = __ldg(a.pt + ptbase + (r >> 5));
                    const int slab = page >> a.slab_shift;
                    const int64_t slot = (int64_t)(page & slab_mask) * kPageRows + (r & 31);
                    ptr = a.slabs->vec[slab] + slot * ds;
                    ok = filter_pass(a.filt, __ldg(a.slabs->tags[slab] + slot));
                }
                sm.rowptr[tile % C::RING][tid] = ptr;
                sm.live[tile % C::RING][tid] = ok ? 1 : 0;
            }
        };
        auto issue = [&](int s) {  // cp.async the operands of flattened stage s into ring slot s % NS
            const int tile = s / KB, kb = s - tile * KB;
            const int buf = s % NS;
            const int k0 = kb * BKX;
            const int c = tid % CPR;  // 16-byte chunk within the row segment
            const bool kin = k0 + c * 4 < ds;
            constexpr int RPI = NT / CPR;  // rows covered per pass
#pragma unroll
            for (int i = 0; i < (RB + RPI - 1) / RPI; ++i) {
                const int r = tid / CPR + RPI * i;
                if (RB % RPI == 0 || r < RB) {
                    const bool valid = kin && (tile * RB + r < len);
                    cp_async16(&sm.xs[buf][r][c * 4], sm.rowptr[tile % C::RING][r] + k0 + c * 4, valid);
                }
            }
#pragma unroll
            for (int i = 0; i < (QT * CPR + NT - 1) / NT; ++i) {
                const int idx = tid + i * NT;
                if (idx < QT * CPR) {
                    const int j = idx / CPR;
                    cp_async16(&sm.qs[buf][j][c * 4], sm.qptr[j] + k0 + c * 4, kin && j < nqi);
                }
            }
        };

        // prologue: row tables for every tile the first NS stages touch, then NS-1 stages in flight
        int prepped = -1;
        {
            const int need = min(ntiles - 1, (NS - 1) / KB + 1);
            for (int t = 0; t <= need; ++t) prep_rows(t);
            prepped = need;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < NS - 1; ++s) {
            if (s < S) issue(s);
            cp_async_commit();
        }

        // packed accumulators: acc2[i][j] = (sum over even k, sum over odd k) -> one FFMA2 per two products
        unsigned long long acc2[C::RPT][C::QPT];
        auto k4_step = [&](int buf, int k4) {
            ulonglong2 xv[C::RPT];
#pragma unroll
            for (int i = 0; i < C::RPT; ++i)
                xv[i] = *reinterpret_cast<const ulonglong2 *>(&sm.xs[buf][row0 + 32 * i][k4 * 4]);
#pragma unroll
            for (int j = 0; j < C::QPT; ++j) {
                const ulonglong2 qv = *reinterpret_cast<const ulonglong2 *>(&sm.qs[buf][qg * C::QPT + j][k4 * 4]);
#pragma unroll
                for (int i = 0; i < C::RPT; ++i) {
                    if (L2) {
                        const unsigned long long d0 = sub2(xv[i].x, qv.x), d1 = sub2(xv[i].y, qv.y);
                        fma2(acc2[i][j], d0, d0);
                        fma2(acc2[i][j], d1, d1);
                    } else {
                        fma2(acc2[i][j], xv[i].x, qv.x);
                        fma2(acc2[i][j], xv[i].y, qv.y);
                    }
                }
            }
        };
        for (int s = 0; s < S; ++s) {
            const int tile = s / KB, kb = s - tile * KB;
            const int buf = s % NS;
            cp_async_wait<NS - 2>();
            __syncthreads();  // stage s visible to all; ring slot (s-1) % NS and older row tables are free
            if (s + NS - 1 < S) issue(s + NS - 1);
            cp_async_commit();
            {  // row table of the tile whose first stage is issued in the NEXT iteration
                const int tn = min(ntiles - 1, (s + NS) / KB);
                if (tn > prepped) {
                    prep_rows(tn);
                    prepped = tn;
                }
            }
            if (kb == 0) {
#pragma unroll
                for (int i = 0; i < C::RPT; ++i)
#pragma unroll
                    for (int j = 0; j < C::QPT; ++j) acc2[i][j] = 0ull;
            }
            const int kv = (min(ds - kb * BKX, BKX)) >> 2;  // valid k4-steps of this stage
            if (kv == BKX / 4) {
#pragma unroll
                for (int t = 0; t < C::K4S; ++t) k4_step(buf, ks * C::K4S + t);
            } else {  // ragged last stage (dim not a multiple of the stage width)
#pragma unroll 1
                for (int t = 0; t < C::K4S; ++t)
                    if (ks * C::K4S + t < kv) k4_step(buf, ks * C::K4S + t);
            }
            if (kb == KB - 1) {  // tile finished: fold even/odd and the k-splits, one candidate per (row slot, query)
                if (ks > 0) {
#pragma unroll
                    for (int j = 0; j < C::QPT; ++j)
#pragma unroll
                        for (int i = 0; i < C::RPT; ++i) sm.red[ks - 1][qg * C::QPT + j][row0 + 32 * i] = sum2(acc2[i][j]);
                }
                __syncthreads();
                if (ks == 0) {
#pragma unroll
                    for (int j = 0; j < C::QPT; ++j) {
                        const int64_t cb = sm.cbase[qg * C::QPT + j];
                        if (cb < 0) continue;
#pragma unroll
                        for (int i = 0; i < C::RPT; ++i) {
                            const int rr = row0 + 32 * i;
                            const int32_t r = tile * RB + rr;
                            if (r < slots) {
                                float v = sum2(acc2[i][j]);
#pragma unroll
                                for (int h = 0; h < C::KS - 1; ++h) v += sm.red[h][qg * C::QPT + j][rr];
                                const bool ok = r < len && sm.live[tile % C::RING][rr];
                                a.cand[cb + r] = ok ? (L2 ? -v : v) : -INFINITY;
                            }
                        }
                    }
                }
                // sm.red is next written at the end of the following tile, >= KB >= 2 barriers from here
            }
        }
        cp_async_wait<0>();
    }
}

template <int QT, int BKX, int NS, int RB, int RPT>
cudaError_t launch_lists_variant(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st) {
    const size_t smem = sizeof(TileSmem<QT, BKX, NS, RB, RPT>);
    const int which = 0;
    const int per_sm = smem * 2 + 2048 <= 227 * 1024 ? 2 : 1;
    cudaError_t e;
    if (a.metric == 1) {
        auto kern = scan_lists_kernel<QT, BKX, NS, RB, RPT, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<num_sms * per_sm, NT, smem, st>>>(a, p, which);
    } else {
        auto kern = scan_lists_kernel<QT, BKX, NS, RB, RPT, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<num_sms * per_sm, NT, smem, st>>>(a, p, which);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_scan_mq(const ScanArgs &a, const ListPlan &p, int bucket, int cfg, const int32_t *pgoff, int num_sms, cudaStream_t st);
cudaError_t launch_scan_mq_both(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st);

int list_plan_ctas(int32_t nlist) { return (nlist + PL_BLK - 1) / PL_BLK; }

// p.chunk == 64 (set by the caller): tcgen05 tiles; else FFMA tiles, cfg 2 = their 32-float-stage variant
cudaError_t launch_scan_lists(const ScanArgs &a, const ListPlan &p, int cfg, int num_sms, int *launches, cudaStream_t st) {
    if (a.npairs <= 0) return cudaSuccess;
    if (a.npairs > (int64_t)INT32_MAX) return cudaErrorInvalidValue;
    cudaError_t e;
    const unsigned pb = (unsigned)((a.npairs + 255) / 256);
    // cnt | cursor | counters | agg are adjacent: one memset
    const unsigned plan_ctas = (unsigned)list_plan_ctas(p.nlist);
    if ((e = cudaMemsetAsync(p.cnt, 0, (size_t)(2 * p.nlist + 4) * 4 + (size_t)plan_ctas * 32, st)) != cudaSuccess) return e;
    count_list_pairs_kernel<<<pb, 256, 0, st>>>(a.probe, a.npairs, a.list_len, p.nlist, p.cnt, p.counters + 1);
    plan_lists_kernel<<<plan_ctas, PL_T, 0, st>>>(p.cnt, a.list_len, p.nlist, p.chunk, 2 * num_sms, p.tile_rem, p.counters, p.agg, p.n32, p.lq_off, p.off32,
                                                  p.pg8off, p.pg4off, p.unique_rows);
    fill_list_pairs_kernel<<<pb, 256, 0, st>>>(a.probe, a.npairs, a.list_len, p.nlist, p.lq_off, p.cursor, p.lq);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // Three consumers of the plan.  The 32-query tile kernel holds the FP32-bound items (lists probed by many
    // queries); the two page scans hold the HBM-bound remainders, every warp an equal share of the pages, so they
    // have no tail however few lists a bucket holds.  With a side stream the tile kernel is launched first and the
    // page scans fill the SMs it leaves free.
    const bool fork = p.side[0] != nullptr;
    cudaStream_t s32 = fork ? p.side[0] : st;
    if (fork) {
        if ((e = cudaEventRecord(p.ev_fork, st)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(s32, p.ev_fork, 0)) != cudaSuccess) return e;
    }
    if (p.chunk == 64) {  // tcgen05 tiles (index.cu picked chunk = 64 only when the kernel applies)
        // default: list rows as a tensor-memory operand (scan_lists_ts.cu; index.cu hands over the staging area when
        // it applies); cfg 5 / 3: the shared-memory-operand kernels (scan_lists_tc.cu v2 / v1)
        if (p.bstage != nullptr && a.slab_maps != nullptr)
            e = launch_scan_lists_ts(a, p, num_sms, s32);
        else
            e = launch_scan_lists_tc(a, p, cfg == 3 ? 1 : 0, num_sms, s32);
        if (e != cudaSuccess) return e;
    } else if (cfg == 2) {
        if ((e = launch_lists_variant<32, 32, 3, 128, 4>(a, p, num_sms, s32)) != cudaSuccess) return e;
    } else {
        if ((e = launch_lists_variant<32, 64, 2, 128, 4>(a, p, num_sms, s32)) != cudaSuccess) return e;
    }
    if (fork && (e = cudaEventRecord(p.ev_join[0], s32)) != cudaSuccess) return e;
    const bool both = cfg != 4 && p.mq_fused != 0;  // one launch for both page-scan buckets (an option: measured slower)
    if (both) {
        if ((e = launch_scan_mq_both(a, p, num_sms, st)) != cudaSuccess) return e;
    } else {
        if ((e = launch_scan_mq(a, p, 1, cfg, p.pg8off, num_sms, st)) != cudaSuccess) return e;
        if ((e = launch_scan_mq(a, p, 0, cfg, p.pg4off, num_sms, st)) != cudaSuccess) return e;
    }
    if (fork && (e = cudaStreamWaitEvent(st, p.ev_join[0], 0)) != cudaSuccess) return e;
    if (launches) *launches += (p.chunk == 64 && p.bstage == nullptr ? 7 : 6) - (both ? 1 : 0);  // the shared-memory-operand kernel splits the queries first
    return cudaSuccess;
}

}  // namespace sc
