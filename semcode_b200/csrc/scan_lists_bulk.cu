// K5b, HBM-bound tile: lists probed by <= 8 queries of the batch, streamed with bulk async copies.
//
// Same work items and candidate layout as scan_lists.cu (its <8, ...> variant), different data path:
// a producer warp walks the list 4 rows at a time and issues one `cp.async.bulk` per row (contiguous
// dim*4 bytes, 3 KB at dim 768; one 12 KB copy per group measured SLOWER: 8.9 vs 6.4 ms) plus one for the 4 row tags into a shared-memory ring, completion
// signalled on per-slot mbarriers.  Each of the 8 consumer warps takes every 8th row group, scores
// its 4 rows against the item's <= 8 queries (held in shared memory, double-buffered across items)
// and hands the slot back -- no CTA-wide barrier, no per-thread copy instructions and no global
// loads sit between HBM and the FMAs, so up to 13 row groups (160 KB per SM) are in flight.
//
// Consumer mapping (one warp = one 4-row group): lane = (row r = lane & 3, k-part kq = lane >> 2) reads
// float4 number 8t + kq of its row: rows are padded by 8 floats so that a quarter-warp touches 8
// distinct bank groups (conflict-free LDS.128); the query fragment is one 128-byte broadcast
// wavefront.  Each lane keeps 8 packed fp32-pair accumulators (FFMA2), folded over kq with shuffles.
// Arithmetic is exact fp32 in the direct forms -- same parity bar as scan.cu.
// Requires dim_padded % 128 == 0 and dim_padded <= 1024.
// STATUS: selectable with lists_cfg = 3.  On B200 it streams at the same ~4.2 TB/s as the cp.async variant
// (6.4 vs 6.1 ms on C2, nq 4096, nprobe 8), so the cp.async kernel stays the default; see DESIGN.md.
// Algorithmic bytes: rows of the item's list x 4 x dim, once per item.
#include "common.cuh"

namespace sc {

namespace {

constexpr int RT = 4;            // rows per ring slot (one consumer warp)
constexpr int NCW = 8;           // consumer warps
constexpr int NTHR = (NCW + 1) * 32;
constexpr int MAXSLOTS = 24;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r, m1;
    asm("mov.b64 %0, {%1, %1};" : "=l"(m1) : "f"(-1.0f));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(b), "l"(m1), "l"(a));
    return r;
}
__device__ __forceinline__ float sum2(unsigned long long v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
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

struct BulkCtl {  // fixed-size control block at the start of dynamic shared memory
    unsigned long long full[MAXSLOTS], empty[MAXSLOTS];  // ring slots
    unsigned long long qfull[2], qempty[2];              // the item's query block, double-buffered
    long long cbase[2][8];
    uint32_t tags[MAXSLOTS][RT];
    int32_t item[2][4];  // [buffer]{item id or -1, list, len, unused}
};
static_assert(sizeof(BulkCtl) % 16 == 0, "the query block behind BulkCtl must stay 16-byte aligned");

// dynamic smem layout: BulkCtl | q blocks [2][8][ds + 8] | ring [nslots][RT][ds + 8]
template <bool L2>
__global__ void __launch_bounds__(NTHR, 1) scan_lists8_bulk_kernel(const ScanArgs a, const ListPlan p, int nslots) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    BulkCtl &ctl = *reinterpret_cast<BulkCtl *>(smem_raw);
    const int ds = a.ds;
    const int ld = ds + 8;  // padded row stride (floats): consecutive rows shift two 16-byte bank groups
    float *qs = reinterpret_cast<float *>(smem_raw + sizeof(BulkCtl));
    float *ring = qs + 2 * 8 * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *item_off = p.off8;
    const int32_t total = item_off[p.nlist];
    const int slab_mask = (1 << a.slab_shift) - 1;

    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) {
            mbar_init(smem_u32(&ctl.full[s]), 1);
            mbar_init(smem_u32(&ctl.empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&ctl.qfull[b]), 1);
            mbar_init(smem_u32(&ctl.qempty[b]), NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // Both sides count row groups with one global counter G: slot = G % nslots, its phase parity =
    // (G / nslots) & 1, consumer warp = G % 8 (kept as running counters: no 64-bit divisions in the loops).
    if (warp == NCW) {
        // ================= producer warp =================
        int slot = 0;
        uint32_t par = 0;
        for (int it = 0;; ++it) {
            const int qb = it & 1;
            int32_t item = 0;
            if (lane == 0) item = atomicAdd(p.counters + 1, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            const bool done = item >= total;
            int32_t l = 0, len = 0, ptbase = 0, nqi = 0, qbase = 0;
            if (!done) {
                l = owner_of(item_off, p.nlist, item);
                qbase = p.lq_off[l] + 32 * p.n32[l];
                nqi = min(8, p.lq_off[l + 1] - qbase);
                len = a.list_len[l];
                ptbase = a.pt_off[l];
            }
            // the query block of iteration it-2 must have been released by every consumer warp
            mbar_wait(smem_u32(&ctl.qempty[qb]), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            const float *qsrc = a.q;
            if (lane == 0) {
                ctl.item[qb][0] = done ? -1 : item;
                ctl.item[qb][1] = l;
                ctl.item[qb][2] = len;
            }
            if (lane < 8) {
                long long cb = -1;
                if (!done && lane < nqi) {
                    const int32_t pair = p.lq[qbase + lane];
                    qsrc = a.q + (int64_t)(pair / a.nprobe) * ds;
                    cb = a.page_off[pair] * kPageRows;
                }
                ctl.cbase[qb][lane] = cb;
            }
            __syncwarp();  // table stores happen-before lane 0's releasing arrive
            if (lane == 0) mbar_expect_tx(smem_u32(&ctl.qfull[qb]), done ? 0u : (uint32_t)nqi * ds * 4u);
            __syncwarp();
            if (!done && lane < nqi) bulk_g2s(smem_u32(qs + ((size_t)qb * 8 + lane) * ld), qsrc, (uint32_t)ds * 4u, smem_u32(&ctl.qfull[qb]));
            if (done) break;
            // stream the list.  Page metadata is fetched 32 pages at a time into registers (one page per
            // lane), so the row-group loop itself issues no global loads.
            const int npages = (len + kPageRows - 1) / kPageRows;
            for (int pbase = 0; pbase < npages; pbase += 32) {
                const float *my_vec = nullptr;
                const uint32_t *my_tag = nullptr;
                if (pbase + lane < npages) {
                    const int32_t page = __ldg(a.pt + ptbase + pbase + lane);
                    const int slab = page >> a.slab_shift;
                    const int64_t slot0 = (int64_t)(page & slab_mask) * kPageRows;
                    my_vec = a.slabs->vec[slab] + slot0 * ds;
                    my_tag = a.slabs->tags[slab] + slot0;
                }
                const int pcount = min(32, npages - pbase);
                for (int pg = 0; pg < pcount; ++pg) {
                    const float *vecp = reinterpret_cast<const float *>(__shfl_sync(0xffffffffu, (unsigned long long)my_vec, pg));
                    const uint32_t *tagp = reinterpret_cast<const uint32_t *>(__shfl_sync(0xffffffffu, (unsigned long long)my_tag, pg));
                    const int32_t prow0 = (pbase + pg) * kPageRows;
                    const int prows = min(kPageRows, len - prow0);
                    for (int r0 = 0; r0 < prows; r0 += RT) {
                        const int rows = min(RT, prows - r0);
                        mbar_wait(smem_u32(&ctl.empty[slot]), par ^ 1u);
                        if (lane == 0) mbar_expect_tx(smem_u32(&ctl.full[slot]), (uint32_t)rows * ds * 4u + 16u);
                        __syncwarp();
                        if (lane < rows)
                            bulk_g2s(smem_u32(ring + ((size_t)slot * RT + lane) * ld), vecp + (int64_t)(r0 + lane) * ds, (uint32_t)ds * 4u,
                                     smem_u32(&ctl.full[slot]));
                        else if (lane == RT)  // the 4 tags of the group: 16 contiguous bytes inside the page
                            bulk_g2s(smem_u32(&ctl.tags[slot][0]), tagp + r0, 16u, smem_u32(&ctl.full[slot]));
                        if (++slot == nslots) {
                            slot = 0;
                            par ^= 1u;
                        }
                    }
                }
            }
        }
    } else {
        // ================= consumer warps =================
        const int r = lane & 3, kq = lane >> 2;
        const int steps = ds / 32;  // float4 per lane: 8 k-parts interleaved at float4 granularity
        // this warp owns the row groups G = warp, warp + 8, ... of the global sequence (across items)
        long long Gbase = 0, G = warp;
        int slot = warp;  // nslots >= 8 > warp
        uint32_t par = 0;
        for (int it = 0;; ++it) {
            const int qb = it & 1;
            mbar_wait(smem_u32(&ctl.qfull[qb]), (uint32_t)(it >> 1) & 1u);
            const int32_t item = ctl.item[qb][0];
            if (item < 0) break;
            const int32_t len = ctl.item[qb][2];
            const int32_t slots_total = ((len + kPageRows - 1) / kPageRows) * kPageRows;
            const long long ngroups = (len + RT - 1) / RT;  // pages hold 32 rows = 8 whole groups
            const float *qblk = qs + (size_t)qb * 8 * ld + kq * 4;
            for (; G < Gbase + ngroups; G += NCW) {
                const long long g = G - Gbase;
                unsigned long long acc2[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc2[j] = 0ull;
                mbar_wait(smem_u32(&ctl.full[slot]), par);
                const float *xrow = ring + ((size_t)slot * RT + r) * ld + kq * 4;
                const uint32_t tag = ctl.tags[slot][r];
#pragma unroll 2
                for (int t = 0; t < steps; ++t) {
                    const ulonglong2 xv = *reinterpret_cast<const ulonglong2 *>(xrow + t * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const ulonglong2 qv = *reinterpret_cast<const ulonglong2 *>(qblk + (size_t)j * ld + t * 32);
                        if (L2) {
                            const unsigned long long d0 = sub2(xv.x, qv.x), d1 = sub2(xv.y, qv.y);
                            fma2(acc2[j], d0, d0);
                            fma2(acc2[j], d1, d1);
                        } else {
                            fma2(acc2[j], xv.x, qv.x);
                            fma2(acc2[j], xv.y, qv.y);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl.empty[slot]));  // slot back to the producer
                // fold the 8 k-parts (lanes 4, 8, 16 apart); lanes with kq == 0 own one row each
                const int32_t row = (int32_t)g * RT + r;
                const bool ok = filter_pass(a.filt, tag);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v = sum2(acc2[j]);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if ((lane >> 2) == 0 && row < len) {
                        const long long cb = ctl.cbase[qb][j];
                        if (cb >= 0) a.cand[cb + row] = ok ? (L2 ? -v : v) : -INFINITY;
                    }
                }
                slot += NCW;
                if (slot >= nslots) {
                    slot -= nslots;
                    par ^= 1u;
                }
            }
            // padding slots of the last page (warp 0)
            if (warp == 0) {
                for (int j = 0; j < 8; ++j) {
                    const long long cb = ctl.cbase[qb][j];
                    if (cb < 0) continue;
                    for (int32_t row = len + lane; row < slots_total; row += 32) a.cand[cb + row] = -INFINITY;
                }
            }
            Gbase += ngroups;
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ctl.qempty[qb]));  // this query block may be replaced
        }
    }
}

}  // namespace

// returns cudaErrorNotSupported when the shape does not fit this kernel (caller falls back)
cudaError_t launch_scan_lists8_bulk(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st) {
    const int ds = a.ds;
    if (ds % 128 != 0 || ds > 1024) return cudaErrorNotSupported;
    const size_t fixed = sizeof(BulkCtl) + (size_t)2 * 8 * (ds + 8) * 4;
    const size_t per_slot = (size_t)RT * (ds + 8) * 4;
    const size_t budget = 227 * 1024 - 1024;
    if (fixed + 8 * per_slot > budget) return cudaErrorNotSupported;
    int nslots = (int)((budget - fixed) / per_slot);
    if (nslots > MAXSLOTS) nslots = MAXSLOTS;
    const size_t smem = fixed + (size_t)nslots * per_slot;
    cudaError_t e;
    if (a.metric == 1) {
        auto kern = scan_lists8_bulk_kernel<true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        kern<<<num_sms, NTHR, smem, st>>>(a, p, nslots);
    } else {
        auto kern = scan_lists8_bulk_kernel<false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        kern<<<num_sms, NTHR, smem, st>>>(a, p, nslots);
    }
    return cudaGetLastError();
}

}  // namespace sc
