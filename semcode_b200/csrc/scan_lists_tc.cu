// K5d: list-major scan on the tensor cores, for lists probed by MANY queries of the batch (inner product).
//
// With nq * nprobe >> nlist a probed list meets tens of queries (32 on average at nq 4096, nprobe 128, nlist
// 16384): scoring it is a [rows x dim] x [dim x queries] contraction, and the exact-fp32 FFMA tiles of
// scan_lists.cu top out at the FP32 pipe (~37 TFLOP/s measured) -- 21 ms per batch where HBM needs 4.7 ms.
// Here the same work item (list, chunk of <= 64 queries) runs on tcgen05:
//
//   D[128 rows, 64 queries] (fp32, TMEM) += A[128 x 32] . B[64 x 32]^T      kind::tf32, per 32-float k-block
//
// fp32 accuracy comes from splitting every operand into two tf32 terms ON THE FLY (the lists stay plain fp32 in
// HBM, nothing is stored twice): hi = tf32(x), lo = tf32(x - hi); product = hi.hi + hi.lo + lo.hi (3 MMAs; a third
// term changed no digit of the result -- the accuracy limit is the accumulation, see NACC below).
// The query rows are split once per batch into two global arrays (a few MB); list rows are split by the loader
// warps between their cp.async landing and the MMA.
//
// CTA = 10 warps, one CTA per SM, items dealt round-robin (no counter: every role derives the same sequence):
//   warps 0-7  A loader/splitter: cp.async 16-byte copies of the raw A k-block (rows through the page table)
//              straight into a 128-byte-swizzled K-major tile of the raw ring, 5 stages in flight; when a stage
//              has landed each thread splits the 4 float4 it copied (hi in place, lo into the lo ring),
//              fence.proxy.async, one mbarrier arrival per warp
//   warp 8     B loader: cp.async of the pre-split query k-blocks (hi | lo) into the B ring, 3 stages in flight
//   warps 9-11 MMA issuers (one lane each), one per product term -- hi.hi, hi.lo, lo.hi -- 4 tcgen05.mma 128x64x8 per
//              stage each into their own accumulators.  (One issuer for all 12 was the bottleneck: a 128x64x8
//              MMA occupies the tensor pipe for 32 cycles but costs ~100 cycles of uniform-register traffic to
//              issue.)  tcgen05.commit by each issuer frees the stage / publishes the tile (two tiles x four
//              64-column accumulators = all 512 TMEM columns; see NACC below)
//   warps 12-15 epilogue: tcgen05.ld of the finished 128x64 tile (sum of its accumulators), fused tag predicate,
//              one coalesced 128-byte store per (query, 32 rows) into the per-pair candidate layout shared with
//              the other scans
// Replaces the same FAISS IVFFlatScanner::scan_codes loop (reference src/semcode/storage/milvus_store.py:141-147).
// Bound: HBM (each list once per 64 queries) once the tensor pipe has >= 6x headroom over FFMA.
#include "tc_common.cuh"

namespace sc {

namespace {

using namespace tcu;

constexpr int A_TILE = TM * TK * 4;  // 16 KB
constexpr int B_TILE = TN * TK * 4;  // 8 KB
// The tensor core adds every 128x64x8 product block into the fp32 accumulator with truncation, so ONE accumulation
// chain of 3 x dim/8 additions carries a biased error of ~1e-6 on unit vectors (measured: 1.3e-6 at dim 768, 2.1e-6 at
// dim 3072, the same with 2 or 3 tf32 terms) -- above the 1e-5 relative bar for scores of ~0.1.  Each tile therefore
// keeps NACC = 4 accumulators: #0 and #1 take the small cross terms (hi.lo, lo.hi: 2^-11 of the result), #2 and #3
// take the hi.hi blocks of the even and odd k-steps; the epilogue adds the four in fp32 (round to nearest).  Shorter
// chains over smaller partial sums cut the truncation error ~10x, to the level of an FFMA chain (measured 1.8e-7
// at dim 768, 2.6e-7 at dim 3072).
constexpr int NACC = 4;
constexpr int TMEM_COLS_TC = 512;  // two tiles x NACC 64-column accumulators

// Shared memory: three rings, so that the bytes in flight from HBM are not tied to the operand staging.
//   R  NR x 16 KB  raw A k-blocks (cp.async target); split in place into the hi operand; freed by the stage's MMAs
//   L   2 x 16 KB  lo operand of A
//   B  NB x 16 KB  hi | lo k-blocks of the (pre-split) queries, loaded by their own warp
// NR - 2 = 5 raw stages (80 KB per SM, 11.8 MB per GPU) are in flight while one is split and one is multiplied.
// (ring depths are template parameters of the kernel: NR raw, NL lo, NB query stages)
constexpr int smem_tc(int nr, int nl, int nb) { return nr * A_TILE + nl * A_TILE + nb * 2 * B_TILE + 1024 /*align*/ + 1024 /*barriers, tables*/; }
constexpr int NAW = 8;        // A loader/splitter warps
constexpr int NMW = 3;        // MMA issuer warps (one lane each), one per product term
constexpr int NT_TC2 = (NAW + 1 + NMW + 4) * 32;  // + 1 B loader warp, 4 epilogue warps

// out[s][i] = s-th tf32 term of x[i]
constexpr int NSPLIT = 2;
__global__ void split_rows_kernel(const float4 *__restrict__ x, int64_t n4, float4 *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const float in[4] = {v.x, v.y, v.z, v.w};
        float t[NSPLIT][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float rem = in[e];
#pragma unroll
            for (int s = 0; s < NSPLIT; ++s) {
                t[s][e] = to_tf32(rem);
                rem -= t[s][e];
            }
        }
#pragma unroll
        for (int s = 0; s < NSPLIT; ++s) out[(int64_t)s * n4 + i] = make_float4(t[s][0], t[s][1], t[s][2], t[s][3]);
    }
}

// V2 (default): the raw fp32 k-block IS the hi operand (the tensor core reads the top 19 bits of a tf32 operand, i.e.
// hi = x with the low 13 mantissa bits dropped; the splitter only writes lo = tf32(x - hi)), and hi.hi + hi.lo are ONE
// 128-column MMA against the query tile [B_hi ; B_lo] (64 + 64 rows, contiguous in the B ring).  Per 32-float stage the
// shared-memory traffic drops from 152 KB to 120 KB and the MMA count from 12 to 8.  Two issuers, one per k-step
// parity, each the only writer of its accumulator set [hh | cross] (fold MMA, then lo.hi into the cross columns):
// the order of additions into every accumulator is fixed, so results are reproducible run to run.
template <bool V2, int NR, int NL, int NB>
__global__ void __launch_bounds__(NT_TC2, 1) scan_lists_tc_kernel(const ScanArgs a, const ListPlan p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *ringR = smem;                                // NR x A_TILE: raw -> hi
    uint8_t *ringL = ringR + NR * A_TILE;                 // NL x A_TILE: lo
    uint8_t *ringB = ringL + NL * A_TILE;                 // NB x (hi | lo) B_TILE
    uint64_t *bars = reinterpret_cast<uint64_t *>(ringB + NB * 2 * B_TILE);
    // bars: [0,NR) stage done (one commit per issuer; frees R slot s%NR, L slot s%NL, B slot s%NB)   [NR,NR+NL) A ready (NAW warps)
    //       [NR+NL,NR+NL+NB) B ready (1 warp)   then 2 tile full (commit), 2 tile empty (4 epilogue warps)
    constexpr int NBARS = NR + NL + NB + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBARS);
    int64_t *cbE = reinterpret_cast<int64_t *>(bars + NBARS + 2);  // [TN] candidate bases of the epilogue's current item
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto done_bar = [&](int s) { return bar0 + 8u * s; };
    auto aready_bar = [&](int s) { return bar0 + 8u * (NR + s); };
    auto bready_bar = [&](int s) { return bar0 + 8u * (NR + NL + s); };
    auto tfull_bar = [&](int acc) { return bar0 + 8u * (NR + NL + NB + acc); };
    auto tempty_bar = [&](int acc) { return bar0 + 8u * (NR + NL + NB + 2 + acc); };
    // the MMAs of stage x (>= 0) have completed: everything they read may be overwritten
    auto wait_stage_done = [&](int x) {
        if (x >= 0) mbar_wait(done_bar(x % NR), ((uint32_t)(x / NR)) & 1u);
    };

    if (threadIdx.x == 0) {
        constexpr int NISS = V2 ? 2 : NMW;  // issuers that commit per stage / per tile
        for (int s = 0; s < NR; ++s) mbar_init(done_bar(s), NISS);
        for (int s = 0; s < NL; ++s) mbar_init(aready_bar(s), NAW);
        for (int s = 0; s < NB; ++s) mbar_init(bready_bar(s), 1);
        for (int acc = 0; acc < 2; ++acc) {
            mbar_init(tfull_bar(acc), NISS);
            mbar_init(tempty_bar(acc), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == NAW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS_TC)
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
    const int KB = a.ds / TK;  // launcher guarantees ds % 32 == 0
    const int slab_mask = (1 << a.slab_shift) - 1;

    if (warp < NAW) {
        // ---------------- A loader / splitter ----------------
        constexpr int CPT = TM * 8 / (NAW * 32);  // 16-byte cells per thread per stage (4)
        constexpr int RSTEP = NAW * 4;            // rows between a thread's cells (32)
        const int t = threadIdx.x;  // 0..NAW*32-1
        const int c = t & 7;        // 16-byte chunk of the 128-byte k-block
        const int r0 = t >> 3;      // rows r0 + RSTEP*i
        uint32_t cell[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) cell[i] = swz(r0 + RSTEP * i, c);
        UnitCursor cur;
        cur.start(a, p, u0, u1);
        int kb = 0;
        const float *rowp[CPT];
        int issued = 0, done = 0;
        constexpr int AHEAD = NR - 2;  // raw stages in flight

        auto issue = [&]() {
            if (kb == 0) {
                // (a bulk L2 prefetch of the next tile's pages was tried here: it doubled the DRAM reads -- 65 GB
                //  instead of 33 GB per batch, L2 hit rate 12 % -- because the lines were evicted before use)
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    const int32_t r = cur.tile * TM + r0 + RSTEP * i;
                    rowp[i] = nullptr;
                    if (r < cur.len) {
                        const int32_t page = __ldg(a.pt + cur.ptbase + (r >> 5));
                        rowp[i] = a.slabs->vec[page >> a.slab_shift] + ((int64_t)(page & slab_mask) * kPageRows + (r & 31)) * a.ds;
                    }
                }
            }
            wait_stage_done(issued - NR);
            const uint32_t sbase = smem_u32(ringR + (issued % NR) * A_TILE);
            const int k0 = kb * TK + c * 4;
#pragma unroll
            for (int i = 0; i < CPT; ++i)
                cp_async16_zfill(sbase + cell[i], rowp[i] ? (const void *)(rowp[i] + k0) : (const void *)a.q, rowp[i] != nullptr);
            ++issued;
            if (++kb == KB) {
                kb = 0;
                cur.next_unit(a, p);
            }
        };

#pragma unroll 1
        for (int s = 0; s < AHEAD; ++s) {
            if (cur.valid) issue();
            cp_async_commit_group();
        }
#pragma unroll 1
        while (done < issued) {
            cp_async_wait_group<AHEAD - 1>();  // this thread's copies of stage `done` have landed
            wait_stage_done(done - NL);         // the lo slot is free
            const uint32_t sr = smem_u32(ringR + (done % NR) * A_TILE);
            const uint32_t sl = smem_u32(ringL + (done % NL) * A_TILE);
            // three phases (all loads, all arithmetic, all stores) so that the cells do not serialise on aliasing
            float4 v[CPT], h[CPT], l[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(sr + cell[i]));
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                if (V2) {  // hi = what the tensor core reads of the raw word
                    h[i].x = __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u);
                    h[i].y = __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u);
                    h[i].z = __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u);
                    h[i].w = __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u);
                } else {
                    h[i].x = to_tf32(v[i].x);
                    h[i].y = to_tf32(v[i].y);
                    h[i].z = to_tf32(v[i].z);
                    h[i].w = to_tf32(v[i].w);
                }
                l[i].x = to_tf32(v[i].x - h[i].x);
                l[i].y = to_tf32(v[i].y - h[i].y);
                l[i].z = to_tf32(v[i].z - h[i].z);
                l[i].w = to_tf32(v[i].w - h[i].w);
            }
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                if (!V2)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sr + cell[i]), "f"(h[i].x), "f"(h[i].y), "f"(h[i].z), "f"(h[i].w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sl + cell[i]), "f"(l[i].x), "f"(l[i].y), "f"(l[i].z), "f"(l[i].w) : "memory");
            }
            fence_async_smem();  // generic-proxy writes (cp.async + the stores above) -> visible to the MMA's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(aready_bar(done % NL));
            ++done;
            if (cur.valid) issue();
            cp_async_commit_group();
        }
        cp_async_wait_group<0>();
    } else if (warp == NAW) {
        // ---------------- B loader: 64 query rows x 8 chunks x (hi | lo) per stage = 32 copies per lane ----------------
        const int c = lane & 7, r0 = lane >> 3;  // rows r0 + 4*i, i < 16
        const int64_t qstride = (a.npairs / a.nprobe) * (int64_t)a.ds;  // floats per split array
        UnitCursor cur;
        cur.start(a, p, u0, u1);
        int kb = 0;
        int64_t qoff[16];
        int issued = 0, done = 0;
        constexpr int BAHEAD = NB - 1;
        auto issue = [&]() {
            if (cur.new_chunk) {
                cur.new_chunk = false;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = r0 + 4 * i;
                    qoff[i] = -1;
                    if (j < cur.nqi) qoff[i] = (int64_t)(p.lq[cur.qbase + j] / a.nprobe) * a.ds;
                }
            }
            wait_stage_done(issued - NB);
            const uint32_t sbase = smem_u32(ringB + (issued % NB) * 2 * B_TILE);
            const int k0 = kb * TK + c * 4;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t o = swz(r0 + 4 * i, c);
                const bool ok = qoff[i] >= 0;
                cp_async16_zfill(sbase + o, ok ? (const void *)(p.qsplit + qoff[i] + k0) : (const void *)a.q, ok);
                cp_async16_zfill(sbase + B_TILE + o, ok ? (const void *)(p.qsplit + qstride + qoff[i] + k0) : (const void *)a.q, ok);
            }
            ++issued;
            if (++kb == KB) {
                kb = 0;
                cur.next_unit(a, p);
            }
        };
#pragma unroll 1
        for (int s = 0; s < BAHEAD; ++s) {
            if (cur.valid) issue();
            cp_async_commit_group();
        }
#pragma unroll 1
        while (done < issued) {
            cp_async_wait_group<BAHEAD - 1>();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bready_bar(done % NB));
            ++done;
            if (cur.valid) issue();
            cp_async_commit_group();
        }
        cp_async_wait_group<0>();
    } else if (warp <= NAW + NMW) {
        // ---------------- MMA issuers: term 0 = hi.hi -> accumulators 2 / 3 (even / odd k-steps), term 1 = hi.lo -> 0,
        //                  term 2 = lo.hi -> 1 ----------------
        const int term = warp - (NAW + 1);
        if (lane == 0 && (!V2 || term < 2)) {
            constexpr uint32_t idesc = umma_idesc_tf32(TM, TN);
            constexpr uint32_t idesc_fold = umma_idesc_tf32(TM, 2 * TN);
            int sc = 0, acc = 0, sR = 0, sL = 0, sB = 0;
            uint32_t acc_phase = 0, phL = 0, phB = 0;
            {
                for (int32_t unit = u0; unit < u1; ++unit) {  // the issuers need no list fields: one tile per unit
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_t = tmem_base + (uint32_t)(acc * NACC * TN);
                    for (int kb = 0; kb < KB; ++kb, ++sc) {
                        mbar_wait(aready_bar(sL), phL);
                        mbar_wait(bready_bar(sB), phB);
                        tc_fence_after();
                        if (V2) {
                            // issuer `term` owns the k-steps of parity `term` and the accumulator set [hh | cross] at
                            // columns term * 128: fold MMA (N = 128: hi.hi | hi.lo), then lo.hi into the cross half
                            const uint64_t dah = umma_desc_sw128(smem_u32(ringR + sR * A_TILE));
                            const uint64_t dal = umma_desc_sw128(smem_u32(ringL + sL * A_TILE));
                            const uint64_t db = umma_desc_sw128(smem_u32(ringB + sB * 2 * B_TILE));
                            const uint32_t tmem_d = tmem_t + (uint32_t)(term * 2 * TN);
#pragma unroll
                            for (int ks = 0; ks < TK / 8; ks += 2) {
                                const uint64_t off = (uint64_t)(((ks + term) * 8 * 4) >> 4);
                                umma_tf32(tmem_d, dah + off, db + off, idesc_fold, (kb | ks) != 0 ? 1u : 0u);
                                umma_tf32(tmem_d + (uint32_t)TN, dal + off, db + off, idesc, 1u);
                            }
                        } else {
                        const uint32_t aop = term == 2 ? smem_u32(ringL + sL * A_TILE) : smem_u32(ringR + sR * A_TILE);
                        const uint32_t bop = smem_u32(ringB + sB * 2 * B_TILE) + (term == 1 ? B_TILE : 0);
                        const uint64_t da = umma_desc_sw128(aop), db = umma_desc_sw128(bop);
                        if (term == 0) {
#pragma unroll
                            for (int ks = 0; ks < TK / 8; ++ks) {
                                const uint64_t off = (uint64_t)((ks * 8 * 4) >> 4);
                                umma_tf32(tmem_t + (uint32_t)((2 + (ks & 1)) * TN), da + off, db + off, idesc, (kb != 0 || ks >= 2) ? 1u : 0u);
                            }
                        } else {
                            const uint32_t tmem_d = tmem_t + (uint32_t)((term - 1) * TN);
#pragma unroll
                            for (int ks = 0; ks < TK / 8; ++ks) {
                                const uint64_t off = (uint64_t)((ks * 8 * 4) >> 4);
                                umma_tf32(tmem_d, da + off, db + off, idesc, (kb | ks) != 0 ? 1u : 0u);
                            }
                        }
                        }
                        umma_commit(done_bar(sR));
                        if (++sR == NR) sR = 0;
                        if (++sL == NL) {
                            sL = 0;
                            phL ^= 1u;
                        }
                        if (++sB == NB) {
                            sB = 0;
                            phB ^= 1u;
                        }
                    }
                    umma_commit(tfull_bar(acc));
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ---------------- epilogue (the last four warps) ----------------
        const int quarter = warp & 3;
        const int et = threadIdx.x - (NAW + 1 + NMW) * 32;  // 0..127
        const int row = quarter * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        UnitCursor I;
        for (I.start(a, p, u0, u1); I.valid; I.next_unit(a, p)) {
            if (I.new_chunk) {  // same decision in all four warps
                I.new_chunk = false;
                asm volatile("bar.sync 1, 128;" ::: "memory");  // everybody is done with the previous chunk's bases
                if (et < TN) cbE[et] = et < I.nqi ? a.page_off[p.lq[I.qbase + et]] * kPageRows : -1;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const int32_t slots = ((I.len + kPageRows - 1) / kPageRows) * kPageRows;
            {
                const int tile = I.tile;
                const int32_t r = tile * TM + row;
                bool live = false;
                if (r < I.len) {
                    const int32_t page = __ldg(a.pt + I.ptbase + (r >> 5));
                    live = filter_pass(a.filt, __ldg(a.slabs->tags[page >> a.slab_shift] + (int64_t)(page & slab_mask) * kPageRows + (r & 31)));
                }
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * NACC * TN);
#pragma unroll 1
                for (int h = 0; h < TN / 32; ++h) {
                    float v[32], w[32];
                    float x[32];
                    // V1: 0, 1 = the two cross terms, 2, 3 = hi.hi of the even / odd k-steps
                    // V2: 1, 3 = cross terms (hi.lo + lo.hi) of the even / odd k-steps, 0, 2 = hi.hi
                    tmem_ld32(taddr + (uint32_t)((V2 ? 1 : 0) * TN + h * 32), v);
                    tmem_ld32(taddr + (uint32_t)((V2 ? 3 : 1) * TN + h * 32), w);
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = v[j] + w[j];
                    tmem_ld32(taddr + (uint32_t)((V2 ? 0 : 2) * TN + h * 32), v);
                    tmem_ld32(taddr + (uint32_t)((V2 ? 2 : 3) * TN + h * 32), w);
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
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NAW + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS_TC) : "memory");
    }
}

}  // namespace

cudaError_t launch_split_queries(const float *q, int64_t n4, float *out, int num_sms, cudaStream_t st) {
    if (n4 <= 0) return cudaSuccess;
    const int64_t want = (n4 + 255) / 256;
    split_rows_kernel<<<(unsigned)(want < num_sms * 8 ? want : num_sms * 8), 256, 0, st>>>(reinterpret_cast<const float4 *>(q), n4,
                                                                                          reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

// items: (list, chunk of 64 queries) from p.off32 (plan_lists_kernel with chunk = 64); p.qsplit holds 2 x [nq, ds]
// variant 1: the first version (rounded hi stored in place, three 64-column MMAs per k-step, three issuers)
cudaError_t launch_scan_lists_tc(const ScanArgs &a, const ListPlan &p, int variant, int num_sms, cudaStream_t st) {
    if (a.metric != 0 || (a.ds % TK) != 0 || p.chunk != TN || p.qsplit == nullptr) return cudaErrorNotSupported;
    cudaError_t e = launch_split_queries(a.q, (a.npairs / a.nprobe) * (int64_t)a.ds / 4, p.qsplit, num_sms, st);
    if (e != cudaSuccess) return e;
    // ring depths 7 / 2 / 4 (raw, lo, query stages); 7/3/3, 6/3/4 and 8/2/3 measured the same 13.0 ms per batch at
    // nq 4096 / nprobe 128: the kernel is bound by shared-memory bandwidth (ncu: LSU + tensor-core wavefronts = 71 % of
    // the data pipe, tensor pipe 26 %, DRAM 38 %), not by the depth of its pipeline
    auto kern = variant == 1 ? scan_lists_tc_kernel<false, 7, 2, 4> : scan_lists_tc_kernel<true, 7, 2, 4>;
    const int smem = smem_tc(7, 2, 4);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<num_sms, NT_TC2, smem, st>>>(a, p);
    return cudaGetLastError();
}

}  // namespace sc
