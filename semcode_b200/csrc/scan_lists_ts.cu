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
//   warps 0-7   converters: warp w serves TMEM lane quarter w % 4 (= page w % 4 of the tile); the two sets alternate
//               k-blocks, 4 TMEM operand slots of 64 columns (hi | lo); each warp also requests its own page boxes
//               (TMA, 32 rows x 128 B, NR k-blocks ahead), so the row stream has eight issuers and no producer warp
//   warps 8-11  epilogue: tcgen05.ld, fused tag predicate, coalesced candidate stores (same layout as the other scans)
//   warp 12     query producer (one lane): per k-block one TMA box [2 N rows x 32 floats] from the item's staging slot
//   warps 13-14 MMA issuers by k-step parity (one lane each); each owns its accumulator set, so the order of the
//               additions into every accumulator is fixed and results are reproducible
//   warp 15     stager: gathers + splits the NEXT item's query rows into the other staging slot
// TMEM: columns [0, 256) accumulators (2 parities x [hh | cross] x 64), [256, 512) four operand slots.
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
constexpr int NR = 8;                       // raw stages (128 KB in flight per SM at most)
constexpr int NS = 4;                       // operand slots: TMEM A slots and shared-memory B slots advance together
constexpr int ACC_COLS = 256;
constexpr int A_SLOT_COLS = 64;
constexpr int TMEM_COLS_TS = 512;
constexpr int NT_TS = 16 * 32;
constexpr int SMEM_TS = NR * RAW_TILE + NS * B_SLOT + 1024 /*align*/ + 2048 /*barriers, tables*/;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// wait-time profile (SEMCODE_TS_PROF=1): cycles summed over all CTAs, one warp (lane 0) per role
//   0-2 query producer: operand slot free, staging slot filled, total     4-6 converter (warp 0): raw full, slot free, total
//   7-10 issuer 0: accumulators empty, A ready, B ready, total     11-12 stager: staging slot consumed, total
//   13-14 epilogue warp 0: accumulators full, total     15 kernel total (thread 0)
__device__ unsigned long long g_ts_prof[16];

struct BMaps {
    CUtensorMap m[4];  // the staging area as a [rows, ds] tensor with boxes of 32 / 64 / 96 / 128 rows x 32 floats
};

template <bool PROF>
__global__ void __launch_bounds__(NT_TS, 1) scan_lists_ts_kernel(const __grid_constant__ BMaps bmaps, const ScanArgs a, const ListPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *ringR = smem;                      // NR x RAW_TILE
    uint8_t *ringB = ringR + NR * RAW_TILE;     // NS x B_SLOT
    uint64_t *bars = reinterpret_cast<uint64_t *>(ringB + NS * B_SLOT);
    // bars: [0,NR) raw full (4 converter warps + TMA tx)   [NR,2NR) unused   then per slot: A ready (4 converter warps),
    //       B ready (TMA tx), slot free (one commit per issuer); then accumulators full (2 commits), empty (4 warps),
    //       staging slot filled x 2 (stager), staging slot consumed x 2 (issuer 0)
    constexpr int NBARS = 2 * NR + 3 * NS + 2 + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    int64_t *cbE = reinterpret_cast<int64_t *>(bars + NBARS + 2);  // [TN] candidate bases of the epilogue's current item
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto rfull_bar = [&](int s) { return bar0 + 8u * s; };
    auto aready_bar = [&](int s) { return bar0 + 8u * (2 * NR + s); };
    auto bready_bar = [&](int s) { return bar0 + 8u * (2 * NR + NS + s); };
    auto sfree_bar = [&](int s) { return bar0 + 8u * (2 * NR + 2 * NS + s); };
    const uint32_t accfull_bar = bar0 + 8u * (2 * NR + 3 * NS), accempty_bar = accfull_bar + 8u;
    auto staged_bar = [&](int s) { return accfull_bar + 16u + 8u * s; };
    auto bfree_bar = [&](int s) { return accfull_bar + 32u + 8u * s; };

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
    if (threadIdx.x == 0) {
        for (int s = 0; s < NR; ++s) {
            mbar_init(rfull_bar(s), 4);  // one arrival (+ 4 KB of TMA bytes) per converter warp of the set
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(aready_bar(s), 4);
            mbar_init(bready_bar(s), 1);
            mbar_init(sfree_bar(s), 2);
        }
        mbar_init(accfull_bar, 2);
        mbar_init(accempty_bar, 4);
        for (int s = 0; s < 2; ++s) {
            mbar_init(staged_bar(s), 1);
            mbar_init(bfree_bar(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS_TS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int32_t total = p.off32[p.nlist];
    const int32_t per_cta = (int32_t)(((int64_t)total + gridDim.x - 1) / gridDim.x);
    const int32_t u0 = (int32_t)min((int64_t)total, (int64_t)blockIdx.x * per_cta);
    const int32_t u1 = (int32_t)min((int64_t)total, (int64_t)u0 + per_cta);
    const int KB = a.ds / TK;  // launcher guarantees ds % 32 == 0
    const int slab_mask = (1 << a.slab_shift) - 1;

    if (warp == 12) {
        // ---------------- query producer (one lane): per k-block one box [2 N rows x 32 floats] from the item's staging slot ----------------
        if (lane == 0) {
            int s = 0, n = -1, npad = 16, row0 = 0;
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
                    const int slot = s % NS;
                    pwait(sfree_bar(slot), (((uint32_t)(s / NS)) & 1u) ^ 1u, 0);
                    mbar_expect_tx(bready_bar(slot), (uint32_t)npad * 256u);
                    tma_load_2d(smem_u32(ringB + slot * B_SLOT), bm, bready_bar(slot), kb * TK, row0);
                }
            }
            pflush(0, 2);
        }
    } else if (warp == 15) {
        // ---------------- stager: the next item's query rows, split into tf32 terms, [hi rows ; lo rows] ----------------
        // three rows in flight per pass (12 x 512 bytes per warp): the copy of an item (<= 64 rows) takes a few thousand
        // cycles, an item lasts tens of thousands
        const int ds4 = a.ds >> 2;
        float4 *slot_base = reinterpret_cast<float4 *>(p.bstage) + (size_t)blockIdx.x * 2 * (2 * TN) * ds4;
        const float4 *q4 = reinterpret_cast<const float4 *>(a.q);
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
                for (int c0 = 0; c0 < ds4; c0 += 128) {
                    float4 v[3][4];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int c = c0 + t * 32 + lane;
                            if (c < ds4) v[r][t] = __ldg(src[r] + c);
                        }
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const int j = j0 + r;
                        if (j >= cur.nqi) break;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
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
                                hi_rows[(size_t)j * ds4 + c] = h;
                                lo_rows[(size_t)j * ds4 + c] = l;
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
        if (lane == 0) pflush(11, 1);
    } else if (warp >= 13) {
        // ---------------- MMA issuers: warp 2 = even k-steps -> accumulators [0, 128), warp 3 = odd -> [128, 256) ----------------
        const int par = warp - 13;
        if (lane == 0) {
            int s = 0;
            uint32_t acc_phase = 0;
            UnitCursor cur;
            int n = -1;
            for (cur.start(a, p, u0, u1); cur.valid; cur.next_unit(a, p)) {
                if (cur.new_chunk) {  // every k-block of the previous item has landed (its B-ready waits are behind us)
                    cur.new_chunk = false;
                    if (par == 0 && n >= 0) mbar_arrive(bfree_bar(n & 1));
                    ++n;
                }
                const int npad = (cur.nqi + 15) & ~15;
                const uint32_t idesc_fold = umma_idesc_tf32(TM, 2 * npad);
                const uint32_t idesc_lo = umma_idesc_tf32(TM, npad);
                pwait(accempty_bar, acc_phase ^ 1u, 0);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(par * 128);
                for (int kb = 0; kb < KB; ++kb, ++s) {
                    const int slot = s % NS;
                    const uint32_t ph = ((uint32_t)(s / NS)) & 1u;
                    pwait(aready_bar(slot), ph, 1);
                    pwait(bready_bar(slot), ph, 2);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + (uint32_t)(ACC_COLS + slot * A_SLOT_COLS);
                    const uint64_t db = umma_desc_sw128(smem_u32(ringB + slot * B_SLOT));
#pragma unroll
                    for (int ks = 0; ks < TK / 8; ks += 2) {
                        const int k8 = ks + par;
                        const uint64_t off = (uint64_t)((k8 * 8 * 4) >> 4);
                        umma_tf32_ts(tmem_d, a_hi + (uint32_t)(k8 * 8), db + off, idesc_fold, (kb | ks) != 0 ? 1u : 0u);
                        umma_tf32_ts(tmem_d + (uint32_t)npad, a_hi + (uint32_t)(32 + k8 * 8), db + off, idesc_lo, 1u);
                    }
                    umma_commit(sfree_bar(slot));
                }
                umma_commit(accfull_bar);
                acc_phase ^= 1u;
            }
            if (par == 0) pflush(7, 3);
        }
    } else if (warp <= 7) {
        // ---------------- converters: warp (set, quarter) converts page `quarter` of the k-blocks s = set (mod 2) and, as soon as
        // it holds a k-block in registers, refills that quarter of the raw slot with the k-block NR further on (lane 0: one TMA
        // box).  The row stream is thus issued by eight warps -- a single producer thread needs ~1100 cycles for the five TMA
        // boxes, two barrier waits and the bookkeeping of a stage -- and needs no "slot empty" handshake. ----------------
        const int set = warp >> 2, quarter = warp & 3;
        const uint32_t lane_off = (uint32_t)(quarter * 4096 + lane * 128);
        const uint32_t x7 = (uint32_t)(lane & 7);
        const uint32_t tq = ((uint32_t)(quarter * 32) << 16);
        const uint8_t *maps = reinterpret_cast<const uint8_t *>(a.slab_maps);
        const int64_t nstages = (int64_t)(u1 - u0) * KB;
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
        auto request = [&](int64_t s) {  // k-block s (this warp's parity) into raw slot s % NR, quarter `quarter`
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
            if (lane == 0) {
                const int rslot = (int)(s % NR);
                if (l_row >= 0) {
                    mbar_expect_tx(rfull_bar(rslot), 4096u);
                    tma_load_2d(smem_u32(ringR + rslot * RAW_TILE) + quarter * 4096, l_map, rfull_bar(rslot), kbl * TK, l_row);
                } else {
                    mbar_arrive(rfull_bar(rslot));  // the tile ends before this page: nothing to load
                }
            }
            kbl += 2;
            while (cl.valid && kbl >= KB) {
                kbl -= KB;
                cl.next_unit(a, p);
            }
        };
        for (int64_t s = set; s < nstages && s < set + NR; s += 2) request(s);
        for (int64_t s = set; s < nstages; s += 2) {
            const int rslot = (int)(s % NR);
            pwait(rfull_bar(rslot), ((uint32_t)(s / NR)) & 1u, 0);
            const uint32_t src = smem_u32(ringR + rslot * RAW_TILE) + lane_off;
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[4 * c]), "=r"(v[4 * c + 1]), "=r"(v[4 * c + 2]), "=r"(v[4 * c + 3])
                             : "r"(src + (((uint32_t)c ^ x7) << 4)));
            uint32_t h[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) h[i] = v[i] & 0xffffe000u;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(to_tf32(__uint_as_float(v[i]) - __uint_as_float(h[i])));
            __syncwarp();  // every lane holds its row: the quarter may be overwritten
            if (s + NR < nstages) request(s + NR);
            const int slot = (int)(s % NS);
            pwait(sfree_bar(slot), (((uint32_t)(s / NS)) & 1u) ^ 1u, 1);
            tc_fence_after();
            const uint32_t ta = tmem_base + tq + (uint32_t)(ACC_COLS + slot * A_SLOT_COLS);
            tmem_st32(ta, h);
            tmem_st32(ta + 32u, v);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aready_bar(slot));
        }
        if (warp == 0 && lane == 0) pflush(4, 2);
    } else {
        // ---------------- epilogue (the last four warps) ----------------
        const int quarter = warp & 3;
        const int et = threadIdx.x - 8 * 32;  // 0..127
        const int row = quarter * 32 + lane;
        uint32_t acc_phase = 0;
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
            pwait(accfull_bar, acc_phase, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int h = 0; h * 32 < I.nqi; ++h) {
                float v[32], w[32], x[32];
                tmem_ld32(taddr + (uint32_t)(npad + h * 32), v);        // cross terms, even k-steps
                tmem_ld32(taddr + (uint32_t)(128 + npad + h * 32), w);  // cross terms, odd k-steps
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = v[j] + w[j];
                tmem_ld32(taddr + (uint32_t)(h * 32), v);               // hi.hi, even
                tmem_ld32(taddr + (uint32_t)(128 + h * 32), w);         // hi.hi, odd
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (v[j] + w[j]) + x[j];
                if (r < slots) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int jj = h * 32 + j;
                        if (jj < I.nqi) a.cand[cbE[jj] + r] = live ? v[j] : -INFINITY;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty_bar);
            acc_phase ^= 1u;
        }
        if (warp == 8 && lane == 0) pflush(13, 1);
    }

    tc_fence_before();
    __syncthreads();
    if (PROF && threadIdx.x == 0) atomicAdd(&g_ts_prof[15], (unsigned long long)(clock64() - t_start));
    if (warp == 13) {
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
