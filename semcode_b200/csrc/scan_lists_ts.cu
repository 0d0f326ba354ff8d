// K5e: list-major tile items on the tensor cores with the LIST ROWS AS A TENSOR-MEMORY OPERAND (inner product).
//
// scan_lists_tc.cu (K5d) moves 120 KB through the shared-memory / LSU data path per 16 KB k-block of list rows (cp.async
// fill 16, query fill 16, splitter read + lo write 32, MMA operand reads 56) and stops at 3.1 TB/s of HBM; its real
// pace-setter turned out to be the LSU: cp.async.128 sustains ~16-20 B/cycle/SM, and one warp issued every query
// copy of a stage.  Here nothing on the streaming path goes through the LSU and the row operand never returns to
// shared memory:
//
//   HBM --TMA box (32 rows x 128 B, 128-byte swizzle, one per list page)--> raw ring in shared memory        16 KB
//       --converter warps: LDS.128 row-per-lane (conflict-free under the swizzle)--> registers                16 KB
//       --hi = x & ~0x1fff, lo = tf32(x - hi)--> tcgen05.st into a TMEM operand slot (lane = row)              0
//   D[128 rows, queries] += A[TMEM] . B[shared]^T    (tcgen05.mma, A from tensor memory: only B is read)
//
// The query tile is only as wide as the item needs (N = queries rounded up to 16, not 64) and arrives as ONE TMA box per
// k-block: a stager warp gathers the item's query rows once (splitting them into tf32 hi / lo terms on the way) into a
// per-CTA staging slot in global memory (L2-resident: 2 x 128 rows x dim per CTA), laid out [hi rows ; lo rows], so a
// k-block of the tile is a plain 2-D box of 2 N rows x 128 B.  (Tried for the query rows: a cp.async warp -- 1800
// cycles per stage with naive addressing, 970 with everything hoisted, still the slowest role; TMA tile::gather4, four
// arbitrary rows per instruction, parity-green -- but a TMA instruction costs its issuing thread ~70 cycles, so 2 N / 4
// of them per stage are slower still.)
// Shared-memory traffic per 16 KB of list rows: 72 KB (N = 64) / 52 KB (N = 32) instead of 120 KB, none of it LSU stores.
//
// fp32 accuracy as in K5d: three tf32 terms (hi.hi + hi.lo + lo.hi), the first two as ONE MMA against the query tile
// [B_hi ; B_lo] (N doubled), and per tile two accumulator sets by k-step parity ([hh | cross] each), summed by the
// epilogue in fp32 -- the tensor core's truncating accumulation stays two short chains (see scan_lists_tc.cu, NACC).
//
// CTA = 16 warps, one per SM; every role walks the same contiguous range of (list, chunk, 128-row tile) units.  The
// warp scheduler prefers the highest warp id of a sub-partition, so the roles on the critical path come last:
//   warps 0-7   converters: warp w serves TMEM lane quarter w % 4 (= page w % 4 of the tile); two sets take the k-blocks
//               round-robin, 4 TMEM operand slots of 64 columns (hi | lo); each warp also requests its own page boxes
//               (TMA, 32 rows x 128 B, NR k-blocks ahead): the row stream has eight issuers and no producer warp
//   warps 8-11  epilogue: tcgen05.ld, fused tag predicate, coalesced candidate stores (same layout as the other scans)
//   warp 12     query producer (one elected lane): per k-block one TMA box [2 N rows x 32 floats] from the staging slot
//   warps 13-14 MMA issuers (one elected lane each): issuer p owns the k-blocks of parity p of every tile and ITS
//               accumulator set, so the order of the additions into every accumulator is fixed and results are reproducible
//   warp 15     stager: gathers + splits the NEXT item's query rows into the other staging slot
// TMEM: columns [0, 256) accumulators (2 k-block parities x [hh | cross]; two buffers when N <= 32), [256, 512) four operand slots.
// Replaces the same FAISS IVFFlatScanner::scan_codes loop (reference src/semcode/storage/milvus_store.py:141-147).
// Bound: HBM (each list once per 64 queries).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace sc {

namespace {

using namespace tcu;

constexpr int RAW_TILE = TM * TK * 4;       // 16 KB: four 4 KB boxes (one per page of the tile)
constexpr int B_SLOT = 2 * TN * TK * 4;     // 16 KB: up to 64 hi rows + 64 lo rows
constexpr int NSETS = 2;                    // converter sets (4 warps each), k-blocks dealt round-robin
constexpr int NR = 6;                       // raw stages (96 KB in flight per SM), a multiple of NSETS
constexpr int NS = 4;                       // TMEM operand slots (hi | lo, 64 columns each)
constexpr int NB = 6;                       // query ring slots
constexpr int ND = 12;                      // k-block-done barriers: a multiple of NS and NB
constexpr int ACC_COLS = 256;
constexpr int A_SLOT_COLS = 64;
constexpr int TMEM_COLS_TS = 512;
constexpr int W_EPI = 4 * NSETS, W_QPROD = W_EPI + 4, W_ISSUE = W_QPROD + 1, W_STAGER = W_ISSUE + 2;  // first warp of each role
constexpr int NSTAGERS = 1;
constexpr int NT_TS = (W_STAGER + NSTAGERS) * 32;  // 16 warps: 128 registers per thread
constexpr int SMEM_TS = NR * RAW_TILE + NB * B_SLOT + 1024 /*align*/ + 2048 /*barriers, tables*/;

// L2 policies of the two streams (the encodings CUTLASS passes as TMA cache hints): the list rows are read once, the
// query tiles once per tile of the item
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull, kEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *map, uint32_t bar, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

// wait-time profile (SEMCODE_TS_PROF=1): cycles summed over all CTAs, one warp (lane 0) per role
//   0-2 query producer: query slot free, staging slot filled, total     4-6 converter (warp 0): raw full, slot free, total
//   7-10 issuer 0: accumulators empty, A ready, B ready, total     11-12 stager: staging slot consumed, total
//   13-14 epilogue warp 0: accumulators full, total     15 kernel total (thread 0)
__device__ unsigned long long g_ts_prof[16];

struct BMaps {
    CUtensorMap m[4];  // the staging area as a [rows, ds] tensor with boxes of 32 / 64 / 96 / 128 rows x 32 floats
};

// Which accumulator buffer(s) a tile uses -- every role steps through the same sequence.  A tile of N <= 32 queries needs
// 2 sets x [hh | cross] x 32 = 128 columns and alternates between the two halves of the accumulator region, so its
// epilogue overlaps the next tile's MMAs; a wider tile takes both halves.
struct AccSched {
    int next = 0;
    uint32_t uses[2] = {0u, 0u};
    __device__ __forceinline__ int pick(int npad) {
        if (npad > 32) return 3;
        const int m = 1 << next;
        next ^= 1;
        return m;
    }
    // first column of accumulator set `set` (k-block parity) of the tile that owns buffer mask m
    __device__ __forceinline__ static uint32_t col(int m, int set) { return m == 3 ? (uint32_t)(set * 128) : (uint32_t)((m >> 1) * 128 + set * 64); }
};

template <bool PROF>
__global__ void __launch_bounds__(NT_TS, 1) scan_lists_ts_kernel(const __grid_constant__ BMaps bmaps, const ScanArgs a, const ListPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *ringR = smem;                      // NR x RAW_TILE
    uint8_t *ringB = ringR + NR * RAW_TILE;     // NB x B_SLOT
    uint64_t *bars = reinterpret_cast<uint64_t *>(ringB + NB * B_SLOT);
    // bars: raw full [NR] (4 converter warps + TMA bytes) | A ready [NS] (4 converter warps) | B ready [NB] (TMA bytes) |
    //       k-block done [ND] (commit of the issuer that owns it: frees TMEM slot s % NS and query slot s % NB) |
    //       accumulators full [2] (2 commits), empty [2] (4 epilogue warps) | staging slot filled [2] (stager), consumed [2]
    //       (both issuers)
    constexpr int NBARS = NR + NS + NB + ND + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    int64_t *cbE = reinterpret_cast<int64_t *>(bars + NBARS + 2);  // [TN] candidate bases of the epilogue's current item
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto rfull_bar = [&](int s) { return bar0 + 8u * s; };
    auto aready_bar = [&](int s) { return bar0 + 8u * (NR + s); };
    auto bready_bar = [&](int s) { return bar0 + 8u * (NR + NS + s); };
    auto done_bar = [&](int s) { return bar0 + 8u * (NR + NS + NB + s); };
    auto accfull_bar = [&](int s) { return bar0 + 8u * (NR + NS + NB + ND + s); };
    auto accempty_bar = [&](int s) { return bar0 + 8u * (NR + NS + NB + ND + 2 + s); };
    auto staged_bar = [&](int s) { return bar0 + 8u * (NR + NS + NB + ND + 4 + s); };
    auto bfree_bar = [&](int s) { return bar0 + 8u * (NR + NS + NB + ND + 6 + s); };

    if (p.off32[p.nlist] == 0) return;  // no tile items in this batch (uniform: nothing has been set up yet)
    unsigned long long pw[3] = {0, 0, 0};
    const long long t_start = PROF ? clock64() : 0;
    auto pwait = [&](uint32_t bar, uint32_t parity, int which) {
        if (PROF) {
            const long long t0 = clock64();
            mbar_wait(bar, parity);
            pw[which] += (unsigned long long)(clock64() - t0);
        } else {
            mbar_wait(bar, parity);
        }
    };
    auto pflush = [&](int base, int n) {
        if (PROF) {
            for (int i = 0; i < n; ++i) atomicAdd(&g_ts_prof[base + i], pw[i]);
            atomicAdd(&g_ts_prof[base + n], (unsigned long long)(clock64() - t_start));
        }
    };
    // k-block s may overwrite a slot of a ring of depth D once k-block s - D has been multiplied
    auto wait_done = [&](int32_t s, int depth, int which) {
        if (s >= depth) pwait(done_bar((s - depth) % ND), ((uint32_t)((s - depth) / ND)) & 1u, which);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NR; ++s) mbar_init(rfull_bar(s), 4);
        for (int s = 0; s < NS; ++s) mbar_init(aready_bar(s), 4);
        for (int s = 0; s < NB; ++s) mbar_init(bready_bar(s), 1);
        for (int s = 0; s < ND; ++s) mbar_init(done_bar(s), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(accfull_bar(s), 2);
            mbar_init(accempty_bar(s), 4);
            mbar_init(staged_bar(s), NSTAGERS);
            mbar_init(bfree_bar(s), 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == W_ISSUE) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS_TS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int32_t total = p.off32[p.nlist];
    const int32_t u0 = (int32_t)((int64_t)blockIdx.x * total / gridDim.x);  // balanced contiguous ranges of units
    const int32_t u1 = (int32_t)(((int64_t)blockIdx.x + 1) * total / gridDim.x);
    const int KB = a.ds / TK;  // launcher guarantees ds % 32 == 0 and (u1 - u0) * KB < 2^31
    const int slab_mask = (1 << a.slab_shift) - 1;

    if (warp == W_QPROD) {
        // ---------------- query producer (one elected lane): per k-block one box [2 N rows x 32 floats] from the item's staging slot ----------------
        if (elect_one()) {
            int32_t s = 0;
            int n = -1, npad = 16, row0 = 0;
            const CUtensorMap *bm = &bmaps.m[0];
            UnitCursor cur;
            for (cur.start(a, p, u0, u1); cur.valid; cur.next_unit(a, p)) {
                if (cur.new_chunk) {
                    cur.new_chunk = false;
                    ++n;
                    npad = (cur.nqi + 15) & ~15;
                    bm = &bmaps.m[(npad >> 4) - 1];
                    row0 = ((int)blockIdx.x * 2 + (n & 1)) * (2 * TN);
                    pwait(staged_bar(n & 1), ((uint32_t)(n >> 1)) & 1u, 1);
                }
                for (int kb = 0; kb < KB; ++kb, ++s) {
                    const int slot = s % NB;
                    wait_done(s, NB, 0);
                    mbar_expect_tx(bready_bar(slot), (uint32_t)npad * 256u);
                    tma_load_2d(smem_u32(ringB + slot * B_SLOT), bm, bready_bar(slot), kb * TK, row0, kEvictLast);
                }
            }
            pflush(0, 2);
        }
    } else if (warp >= W_STAGER) {
        // ---------------- stager: the next item's query rows, split into tf32 terms, [hi rows ; lo rows] ----------------
        // 3 rows x 6 float4 per lane in flight per pass (9 KB per warp); the copy of an item runs one item ahead of its use.
        // Loads and stores carry the L2 evict_last policy: the queries and the staging slots (~0.4 MB per CTA) are re-read for
        // every tile of the item while ~100 MB of list rows stream through the L2 between two uses.
        const int ds4 = a.ds >> 2;
        float4 *slot_base = reinterpret_cast<float4 *>(p.bstage) + (size_t)blockIdx.x * 2 * (2 * TN) * ds4;
        const float4 *q4 = reinterpret_cast<const float4 *>(a.q);
        uint64_t keep;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
        int n = 0;
        UnitCursor cur;
        cur.start(a, p, u0, u1);
        while (cur.valid) {
            const int npad = (cur.nqi + 15) & ~15;
            // lane l keeps the query rows of columns l and l + 32 of the item
            int32_t qa = 0, qb = 0;
            if (lane < cur.nqi) qa = p.lq[cur.qbase + lane] / a.nprobe;
            if (lane + 32 < cur.nqi) qb = p.lq[cur.qbase + lane + 32] / a.nprobe;
            if (n >= 2) pwait(bfree_bar(n & 1), (((uint32_t)(n >> 1)) & 1u) ^ 1u, 0);  // the item two back has been consumed
            float4 *hi_rows = slot_base + (size_t)(n & 1) * (2 * TN) * ds4;
            float4 *lo_rows = hi_rows + (size_t)npad * ds4;
            for (int j0 = 0; j0 < cur.nqi; j0 += 3) {  // rows past the item keep stale (finite) values: their columns are never stored
                const float4 *src[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int j = min(j0 + r, cur.nqi - 1);
                    const int32_t qi = __shfl_sync(0xffffffffu, j < 32 ? qa : qb, j & 31);
                    src[r] = q4 + (size_t)qi * ds4;
                }
                for (int c0 = 0; c0 < ds4; c0 += 192) {
                    float4 v[3][6];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int t = 0; t < 6; ++t) {
                            const int c = c0 + t * 32 + lane;
                            if (c < ds4)
                                asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                                             : "=f"(v[r][t].x), "=f"(v[r][t].y), "=f"(v[r][t].z), "=f"(v[r][t].w)
                                             : "l"(src[r] + c), "l"(keep));
                        }
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const int j = j0 + r;
                        if (j >= cur.nqi) break;
#pragma unroll
                        for (int t = 0; t < 6; ++t) {
                            const int c = c0 + t * 32 + lane;
                            if (c < ds4) {
                                float4 h, l;
                                h.x = to_tf32(v[r][t].x);
                                h.y = to_tf32(v[r][t].y);
                                h.z = to_tf32(v[r][t].z);
                                h.w = to_tf32(v[r][t].w);
                                l.x = to_tf32(v[r][t].x - h.x);
                                l.y = to_tf32(v[r][t].y - h.y);
                                l.z = to_tf32(v[r][t].z - h.z);
                                l.w = to_tf32(v[r][t].w - h.w);
                                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(hi_rows + (size_t)j * ds4 + c),
                                             "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w), "l"(keep)
                                             : "memory");
                                asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(lo_rows + (size_t)j * ds4 + c),
                                             "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w), "l"(keep)
                                             : "memory");
                            }
                        }
                    }
                }
            }
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy global writes -> visible to the TMA reads
            __syncwarp();
            if (lane == 0) mbar_arrive(staged_bar(n & 1));
            ++n;
            cur.new_chunk = false;  // skip the remaining tiles of this item
            do {
                cur.next_unit(a, p);
            } while (cur.valid && !cur.new_chunk);
        }
        if (warp == W_STAGER && lane == 0) pflush(11, 1);
    } else if (warp >= W_ISSUE) {
        // ---------------- MMA issuers: issuer `par` multiplies the k-blocks of parity `par` of every tile (all four k-steps, all
        // three terms) into ITS accumulator set, so the order of the additions into every accumulator is fixed ----------------
        const int par = warp - W_ISSUE;
        if (elect_one()) {
            int32_t sbase = 0;
            AccSched acc;
            UnitCursor cur;
            int n = -1;
            for (cur.start(a, p, u0, u1); cur.valid; cur.next_unit(a, p), sbase += KB) {
                if (cur.new_chunk) {
                    // this issuer is past the B-ready wait of every k-block of its parity of the previous item: once both
                    // issuers say so, every TMA read of that item's staging slot has completed and the slot may be rewritten
                    cur.new_chunk = false;
                    if (n >= 0) mbar_arrive(bfree_bar(n & 1));
                    ++n;
                }
                const int npad = (cur.nqi + 15) & ~15;
                const uint32_t idesc_fold = umma_idesc_tf32(TM, 2 * npad);
                const uint32_t idesc_lo = umma_idesc_tf32(TM, npad);
                const int m = acc.pick(npad);
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if ((m >> i) & 1) pwait(accempty_bar(i), (acc.uses[i] & 1u) ^ 1u, 0);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + AccSched::col(m, par);
                for (int kb = par; kb < KB; kb += 2) {
                    const int32_t s = sbase + kb;
                    const int aslot = s % NS, bslot = s % NB;
                    pwait(aready_bar(aslot), ((uint32_t)(s / NS)) & 1u, 1);
                    pwait(bready_bar(bslot), ((uint32_t)(s / NB)) & 1u, 2);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + (uint32_t)(ACC_COLS + aslot * A_SLOT_COLS);
                    const uint64_t db = umma_desc_sw128(smem_u32(ringB + bslot * B_SLOT));
#pragma unroll
                    for (int k8 = 0; k8 < TK / 8; ++k8) {
                        const uint64_t off = (uint64_t)((k8 * 8 * 4) >> 4);
                        umma_tf32_ts(tmem_d, a_hi + (uint32_t)(k8 * 8), db + off, idesc_fold, (kb != par || k8 != 0) ? 1u : 0u);
                        umma_tf32_ts(tmem_d + (uint32_t)npad, a_hi + (uint32_t)(32 + k8 * 8), db + off, idesc_lo, 1u);
                    }
                    umma_commit(done_bar(s % ND));
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if ((m >> i) & 1) {
                        umma_commit(accfull_bar(i));
                        acc.uses[i] += 1;
                    }
                }
            }
            if (par == 0) pflush(7, 3);
        }
    } else if (warp < W_EPI) {
        // ---------------- converters: warp (set, quarter) converts page `quarter` of the k-blocks s = set (mod NSETS) and, while
        // its TMEM stores drain, refills that quarter of the raw slot with the k-block NR further on (one elected lane: one TMA
        // box).  An asynchronous-copy or mbarrier instruction costs its issuing thread ~80-100 cycles, so a single producer
        // thread (5 boxes, 2 waits, 2 expect_tx per k-block) cannot feed the ring -- twelve warps can, and no "slot empty"
        // handshake is needed. ----------------
        const int set = warp >> 2, quarter = warp & 3;
        const uint32_t lane_off = (uint32_t)(quarter * 4096 + lane * 128);
        const uint32_t x7 = (uint32_t)(lane & 7);
        const uint32_t tq = ((uint32_t)(quarter * 32) << 16);
        const uint8_t *maps = reinterpret_cast<const uint8_t *>(a.slab_maps);
        const int32_t nstages = (u1 - u0) * KB;
        // load cursor: (unit, k-block) of the next k-block this warp has to request
        UnitCursor cl;
        cl.start(a, p, u0, u1);
        int kbl = set;
        while (cl.valid && kbl >= KB) {
            kbl -= KB;
            cl.next_unit(a, p);
        }
        int32_t l_unit = -1;
        const void *l_map = maps;
        int l_row = -1;
        // `dep` is 0 at run time but computed from the words the caller has just read out of the slot: the copy's
        // destination address depends on it, so the TMA instruction cannot be issued before those LDS have returned
        // (ptxas is free to sink the arithmetic that consumes them below the copy otherwise -- seen as a few wrong rows per
        // thousand queries)
        auto request = [&](int32_t s, uint32_t dep) {  // k-block s (this warp's set) into raw slot s % NR, quarter `quarter`
            if (cl.u != l_unit) {        // new tile: where does its page `quarter` live?
                l_unit = cl.u;
                l_row = -1;
                const int32_t npages = (cl.len + kPageRows - 1) / kPageRows;
                if (cl.tile * 4 + quarter < npages) {
                    const int32_t page = __ldg(a.pt + cl.ptbase + cl.tile * 4 + quarter);
                    l_map = maps + (size_t)(page >> a.slab_shift) * 128;
                    l_row = (page & slab_mask) * kPageRows;
                }
            }
            if (elect_one()) {
                const int rslot = s % NR;
                if (l_row >= 0) {
                    mbar_expect_tx(rfull_bar(rslot), 4096u);
                    tma_load_2d(smem_u32(ringR + rslot * RAW_TILE) + quarter * 4096 + dep, l_map, rfull_bar(rslot), kbl * TK, l_row, kEvictFirst);
                } else {
                    mbar_arrive(rfull_bar(rslot));  // the tile ends before this page: nothing to load
                }
            }
            kbl += NSETS;
            while (cl.valid && kbl >= KB) {
                kbl -= KB;
                cl.next_unit(a, p);
            }
        };
        static_assert(NR % NSETS == 0, "the warp that converts k-block s refills its slot with k-block s + NR");
        const uint32_t rt_zero = (uint32_t)(p.nlist >> 31);  // 0, but not to the compiler
        for (int32_t s = set; s < nstages && s < set + NR; s += NSETS) request(s, 0u);
        for (int32_t s = set; s < nstages; s += NSETS) {
            const int rslot = s % NR;
            const int slot = s % NS;
            const uint32_t ta = tmem_base + tq + (uint32_t)(ACC_COLS + slot * A_SLOT_COLS);
            pwait(rfull_bar(rslot), ((uint32_t)(s / NR)) & 1u, 0);
            const uint32_t src = smem_u32(ringR + rslot * RAW_TILE) + lane_off;
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[4 * c]), "=r"(v[4 * c + 1]), "=r"(v[4 * c + 2]), "=r"(v[4 * c + 3])
                             : "r"(src + (((uint32_t)c ^ x7) << 4)));
            // hi = the 19 bits the tensor core reads of x; lo = x - hi, exact in fp32 and at most 2^-10 |x| (the tensor core
            // reads its upper 19 bits, so what is lost is below 2^-20 |x|): two ALU instructions per element on the
            // sub-partition that owns the quarter
            uint32_t h[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) h[i] = v[i] & 0xffffe000u;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) - __uint_as_float(h[i]));
            // every word of the raw quarter has been consumed by an instruction (not merely loaded): refill the quarter now --
            // the row stream must not wait for the MMAs -- then wait for the TMEM operand slot
            __syncwarp();
            if (s + NR < nstages)
                request(s + NR, (v[0] | v[4] | v[8] | v[12] | v[16] | v[20] | v[24] | v[28]) & rt_zero);
            wait_done(s, NS, 1);
            tc_fence_after();
            tmem_st32(ta, h);
            tmem_st32(ta + 32u, v);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aready_bar(slot));
        }
        if (warp == 0 && lane == 0) pflush(4, 2);
    } else {
        // ---------------- epilogue (warps W_EPI .. W_EPI + 3) ----------------
        const int quarter = warp & 3;
        const int et = threadIdx.x - W_EPI * 32;  // 0..127
        const int row = quarter * 32 + lane;
        AccSched acc;
        UnitCursor I;
        for (I.start(a, p, u0, u1); I.valid; I.next_unit(a, p)) {
            if (I.new_chunk) {  // same decision in all four warps
                I.new_chunk = false;
                asm volatile("bar.sync 1, 128;" ::: "memory");  // everybody is done with the previous chunk's bases
                if (et < TN) cbE[et] = et < I.nqi ? a.page_off[p.lq[I.qbase + et]] * kPageRows : -1;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const int npad = (I.nqi + 15) & ~15;
            const int32_t slots = ((I.len + kPageRows - 1) / kPageRows) * kPageRows;
            const int32_t r = I.tile * TM + row;
            bool live = false;
            if (r < I.len) {
                const int32_t page = __ldg(a.pt + I.ptbase + (r >> 5));
                live = filter_pass(a.filt, __ldg(a.slabs->tags[page >> a.slab_shift] + (int64_t)(page & slab_mask) * kPageRows + (r & 31)));
            }
            const int m = acc.pick(npad);
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if ((m >> i) & 1) pwait(accfull_bar(i), acc.uses[i] & 1u, 0);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + AccSched::col(m, 0);
            const uint32_t t1 = tmem_base + ((uint32_t)(quarter * 32) << 16) + AccSched::col(m, 1);
            // drain the accumulators into registers first (one row x up to 64 queries per thread) and hand them back to the
            // issuers; the candidate stores that follow go to one region per query -- a TLB miss each, ~10k cycles per tile
            float res[TN];
#pragma unroll
            for (int g = 0; g < TN / 16; ++g) {
                if (g * 16 < I.nqi) {
                    uint32_t t[16];
                    tmem_ld16_nowait(t0 + (uint32_t)(npad + g * 16), t);  // cross terms, even k-blocks
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) res[g * 16 + j] = __uint_as_float(t[j]);
                    tmem_ld16_nowait(t0 + (uint32_t)(g * 16), t);         // hi.hi, even k-blocks
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) res[g * 16 + j] += __uint_as_float(t[j]);
                    if (KB > 1) {  // (a one-k-block tile never touches the odd set)
                        tmem_ld16_nowait(t1 + (uint32_t)(npad + g * 16), t);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) res[g * 16 + j] += __uint_as_float(t[j]);
                        tmem_ld16_nowait(t1 + (uint32_t)(g * 16), t);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) res[g * 16 + j] += __uint_as_float(t[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if ((m >> i) & 1) {
                    if (lane == 0) mbar_arrive(accempty_bar(i));
                    acc.uses[i] += 1;
                }
            }
            if (r < slots) {
#pragma unroll
                for (int jj = 0; jj < TN; ++jj)
                    if (jj < I.nqi) a.cand[cbE[jj] + r] = live ? res[jj] : -INFINITY;
            }
        }
        if (warp == W_EPI && lane == 0) pflush(13, 1);
    }

    tc_fence_before();
    __syncthreads();
    if (PROF && threadIdx.x == 0) atomicAdd(&g_ts_prof[15], (unsigned long long)(clock64() - t_start));
    if (warp == W_ISSUE) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS_TS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn_ts() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

}  // namespace

// 128-byte tensor map of one list slab: fp32 [rows, ds], boxes of one page (32 rows) x 32 floats, 128-byte swizzle
cudaError_t encode_slab_map(void *map128, const float *base, int64_t rows, int ds) {
    static_assert(sizeof(CUtensorMap) == 128, "slab maps are stored as 128-byte records");
    EncodeTiledFn fn = encode_fn_ts();
    if (!fn) return cudaErrorNotSupported;
    if (ds < TK) return cudaErrorInvalidValue;
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)ds, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ds * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)kPageRows};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    memcpy(map128, &m, 128);
    return cudaSuccess;
}

// items: (list, chunk of 64 queries) from p.off32 (plan_lists_kernel with chunk = 64); p.bstage: num_sms x 2 staging
// slots of 128 rows x ds floats; a.slab_maps: one tensor map per slab (encode_slab_map)
size_t scan_lists_ts_stage_bytes(int ds, int num_sms) { return (size_t)num_sms * 2 * (2 * TN) * ds * sizeof(float); }

cudaError_t launch_scan_lists_ts(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st) {
    if (a.metric != 0 || (a.ds % TK) != 0 || p.chunk != TN || p.bstage == nullptr || a.slab_maps == nullptr) return cudaErrorNotSupported;
    if (a.npairs * (int64_t)(a.ds / TK) >= ((int64_t)1 << 30)) return cudaErrorNotSupported;  // 32-bit k-block counters per CTA (index.cu checks first)
    EncodeTiledFn fn = encode_fn_ts();
    if (!fn) return cudaErrorNotSupported;
    BMaps bm;
    const cuuint64_t dims[2] = {(cuuint64_t)a.ds, (cuuint64_t)num_sms * 2 * (2 * TN)};
    const cuuint64_t strides[1] = {(cuuint64_t)a.ds * 4};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 4; ++i) {
        const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)(32 * (i + 1))};
        if (fn(&bm.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.bstage, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    static const bool prof = getenv("SEMCODE_TS_PROF") != nullptr;
    auto kern = prof ? scan_lists_ts_kernel<true> : scan_lists_ts_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TS);
    if (e != cudaSuccess) return e;
    kern<<<num_sms, NT_TS, SMEM_TS, st>>>(bm, a, p);
    return cudaGetLastError();
}

}  // namespace sc

// debug: read and reset the wait-time profile of scan_lists_ts_kernel<true> (see g_ts_prof)
extern "C" int scdbg_ts_prof(unsigned long long *out16) {
    unsigned long long zero[16] = {0};
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(out16, sc::g_ts_prof, sizeof(zero)) != cudaSuccess) return -2;
    if (cudaMemcpyToSymbol(sc::g_ts_prof, zero, sizeof(zero)) != cudaSuccess) return -2;
    return 0;
}
