// K5: the nprobe inverted-list scan -- the HBM-bound hot kernel.
//
// Replaces FAISS IVFFlatScanner::scan_codes (+ knowhere BitsetView filtering) behind
// Collection.search(data=[vector], param={"metric_type":"IP","params":{"nprobe":16}}, ...)
// at reference src/semcode/storage/milvus_store.py:141-147.
//
// Work decomposition: the (query, probed list) pairs are flattened into 32-row *pages*
// (page_off = exclusive prefix of pages per pair).  The kernel is persistent: every warp owns a
// contiguous, equally sized range of pages, so each warp streams the same number of bytes no
// matter how skewed the list lengths are.  Per page a warp
//   1. reads the 32 row tags with one coalesced 128-byte load and evaluates the fused predicate
//      (tombstone | language bitmap | repo bitmap) -> ballot of live rows;
//   2. streams the live rows with 128-bit ld.global.nc.L1::no_allocate loads, R rows x U float4
//      per lane in flight (all independent), against the query slice staged in shared memory;
//   3. butterfly-reduces the R partial sums and writes one similarity per row into the candidate
//      array (dead slots get -inf).  Top-k selection is a separate kernel (select.cu).
// Algorithmic bytes per live row: 4*ds (vector) + 4 (tag); the 4-byte candidate write and the
// winners' 8-byte ids are the only other traffic.  Roofline: HBM bandwidth.
#include "common.cuh"

namespace sc {

namespace {

// ---- pair plan: pages per (query, list) pair and their exclusive prefix in ONE launch -------------------
// (first version: one kernel for the page counts + a three-launch two-level scan.)  Single-pass scan with
// decoupled look-back: a CTA takes a ticket (its position in scheduling order, so a CTA only ever waits for CTAs
// that are already running), publishes the total of its 1024 pairs, adds up its predecessors' words until it
// meets one that already carries an inclusive prefix, and publishes its own inclusive prefix.  One 64-bit word
// per CTA holds value, state and the launch's epoch together -- [epoch:22 | state:2 | pages:40] -- so a word is
// valid on its own: no fence, no flag array, and no memset between launches (the epoch changes; the host zeroes
// the words when the epoch wraps).  The CTA holding the last ticket writes the grand total and resets the ticket.
constexpr int PP_T = 256, PP_IPT = 4, PP_BLK = PP_T * PP_IPT;
constexpr int PP_VAL_BITS = 40;  // pages of a batch: the candidate array (128 B per page) must fit in HBM

__device__ __forceinline__ unsigned long long pp_pack(uint32_t epoch, uint32_t state, int64_t v) {
    return ((unsigned long long)epoch << (PP_VAL_BITS + 2)) | ((unsigned long long)state << PP_VAL_BITS) | (unsigned long long)v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(PP_T) plan_pairs_kernel(const int32_t *__restrict__ probe, int64_t npairs,
                                                          const int32_t *__restrict__ list_len, int32_t nlist,
                                                          int64_t *__restrict__ page_off, unsigned long long *__restrict__ look,
                                                          uint32_t epoch, unsigned long long *__restrict__ rows_total) {
    __shared__ int64_t warp_tot[PP_T / 32];
    __shared__ int64_t s_pre;
    __shared__ uint32_t s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned int *ticket = reinterpret_cast<unsigned int *>(look);  // word 0: ticket counter; words 1..: CTA states
    unsigned long long *state = look + 1;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t b = s_ticket;
    const int64_t i0 = (int64_t)b * PP_BLK + (int64_t)tid * PP_IPT;
    int32_t v[PP_IPT];
    int64_t local = 0;
    unsigned long long rows = 0;
#pragma unroll
    for (int j = 0; j < PP_IPT; ++j) {
        int32_t pages = 0;
        if (i0 + j < npairs) {
            const int32_t l = probe[i0 + j];
            if (l >= 0 && l < nlist) {
                const int32_t len = __ldg(list_len + l);
                pages = (len + kPageRows - 1) / kPageRows;
                rows += (unsigned long long)len;
            }
        }
        v[j] = pages;
        local += pages;
    }
    int64_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int64_t wpre = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PP_T / 32; ++w) {
        const int64_t t = warp_tot[w];
        if (w < warp) wpre += t;
        total += t;
    }
    if (warp == 0) {
        int64_t pre = 0;
        if (b == 0) {
            if (lane == 0) st_relaxed_u64(state, pp_pack(epoch, 2, total));
        } else {
            if (lane == 0) st_relaxed_u64(state + b, pp_pack(epoch, 1, total));
            const unsigned long long vmask = (1ull << PP_VAL_BITS) - 1;
            for (int64_t j = (int64_t)b - 1;; j -= 32) {  // lane 0 looks at the nearest predecessor
                const int64_t idx = j - lane;
                unsigned long long s = pp_pack(epoch, 2, 0);  // before CTA 0: an inclusive prefix of 0
                if (idx >= 0) {
                    do {
                        s = ld_relaxed_u64(state + idx);
                    } while ((uint32_t)(s >> (PP_VAL_BITS + 2)) != epoch || ((s >> PP_VAL_BITS) & 3) == 0);
                }
                const uint32_t inc = __ballot_sync(0xffffffffu, ((s >> PP_VAL_BITS) & 3) == 2);
                const int stop = inc ? (__ffs(inc) - 1) : 31;  // nearest word that is an inclusive prefix
                int64_t val = lane <= stop ? (int64_t)(s & vmask) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                pre += val;
                if (inc) break;
            }
            if (lane == 0) st_relaxed_u64(state + b, pp_pack(epoch, 2, pre + total));
        }
        if (lane == 0) s_pre = pre;
    }
    __syncthreads();
    int64_t run = s_pre + wpre + incl - local;
#pragma unroll
    for (int j = 0; j < PP_IPT; ++j) {
        if (i0 + j < npairs) page_off[i0 + j] = run;
        run += v[j];
    }
    if (b == gridDim.x - 1 && tid == PP_T - 1) {
        page_off[npairs] = s_pre + total;
        *ticket = 0;  // every ticket of this launch has been taken
    }
    if (rows_total != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
        if (lane == 0 && rows) atomicAdd(rows_total, rows);
    }
}

template <bool L2>
__device__ __forceinline__ float accum4(float acc, const float4 &x, const float4 &q) {
    if (L2) {
        const float a = x.x - q.x, b = x.y - q.y, c = x.z - q.z, d = x.w - q.w;
        acc = fmaf(a, a, acc);
        acc = fmaf(b, b, acc);
        acc = fmaf(c, c, acc);
        acc = fmaf(d, d, acc);
    } else {
        acc = fmaf(x.x, q.x, acc);
        acc = fmaf(x.y, q.y, acc);
        acc = fmaf(x.z, q.z, acc);
        acc = fmaf(x.w, q.w, acc);
    }
    return acc;
}

// R rows in flight per warp, U float4 per lane per row per chunk. EXACT: ds4 % (32*U) == 0.
template <int R, int U, bool L2, bool EXACT, int MINB>
__global__ void __launch_bounds__(256, MINB) scan_pages_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) float4 qsmem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int ds4 = a.ds >> 2;
    float4 *qs = qsmem + (size_t)warp * ds4;
    pdl_launch_dependents();
    pdl_wait();

    const int64_t W = a.page_off[a.npairs];
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int64_t gw = (int64_t)blockIdx.x * wpb + warp;
    // small batches: fewer pages than warps -> split every page into 2 or 4 row ranges so that all SMs stream
    const int sub_shift = (W * 4 <= nwarps) ? 2 : ((W * 2 <= nwarps) ? 1 : 0);
    const int part_rows = kPageRows >> sub_shift;
    const int64_t units = W << sub_shift;
    const int64_t u0 = gw * units / nwarps;  // balanced contiguous ranges (sizes differ by at most one unit)
    const int64_t u1 = (gw + 1) * units / nwarps;
    if (u0 >= u1) return;
    const int64_t w0 = u0 >> sub_shift;

    // pair that owns page w0: last i with page_off[i] <= w0 (pairs with no pages are skipped)
    int64_t lo = 0, hi = a.npairs;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (a.page_off[mid] <= w0)
            lo = mid;
        else
            hi = mid;
    }
    int64_t pair = lo;
    int64_t pair_start = a.page_off[pair];
    int64_t pair_end = a.page_off[pair + 1];
    int64_t cur_q = -1;
    int32_t len = 0, ptbase = 0;
    bool fresh = true;
    const int slab_mask = (1 << a.slab_shift) - 1;

    for (int64_t u = u0; u < u1; ++u) {
        const int64_t w = u >> sub_shift;
        const uint32_t part = (uint32_t)(u & ((1 << sub_shift) - 1));
        const uint32_t part_mask = (part_rows == 32 ? 0xffffffffu : ((1u << part_rows) - 1u)) << (part * part_rows);
        while (w >= pair_end) {
            ++pair;
            pair_start = pair_end;
            pair_end = a.page_off[pair + 1];
            fresh = true;
        }
        if (fresh) {
            fresh = false;
            const int32_t l = a.probe[pair];
            len = a.list_len[l];
            ptbase = a.pt_off[l];
            const int64_t qi = pair / a.nprobe;
            if (qi != cur_q) {
                cur_q = qi;
                __syncwarp();
                const float4 *qg = reinterpret_cast<const float4 *>(a.q + qi * (int64_t)a.ds);
                for (int c = lane; c < ds4; c += 32) qs[c] = __ldg(qg + c);
                __syncwarp();
            }
        }
        const int32_t j = (int32_t)(w - pair_start);
        const int32_t page = __ldg(a.pt + ptbase + j);
        const int slab = page >> a.slab_shift;
        const int64_t slot0 = (int64_t)(page & slab_mask) * kPageRows;
        const int rows = min(kPageRows, len - j * kPageRows);
        const uint32_t tag = __ldg(a.slabs->tags[slab] + slot0 + lane);
        const bool live = lane < rows && filter_pass(a.filt, tag);
        uint32_t m = __ballot_sync(0xffffffffu, live) & part_mask;
        float *cpage = a.cand + w * kPageRows;
        if (!live && ((part_mask >> lane) & 1u)) cpage[lane] = -INFINITY;
        const float4 *vbase = reinterpret_cast<const float4 *>(a.slabs->vec[slab]) + slot0 * ds4;

        while (m) {
            int row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                row[r] = m ? (__ffs(m) - 1) : -1;
                m &= m - 1;  // 0 & (0-1) == 0
            }
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = 0.f;
            for (int c0 = 0; c0 < ds4; c0 += 32 * U) {
                float4 x[R][U];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 *rp = vbase + (int64_t)(row[r] < 0 ? row[0] : row[r]) * ds4 + c0 + lane;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (row[r] >= 0 && (EXACT || c0 + lane + 32 * u < ds4))
                            x[r][u] = ld_stream_f4(rp + 32 * u);
                        else
                            x[r][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (EXACT || c0 + lane + 32 * u < ds4) qv = qs[c0 + lane + 32 * u];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = accum4<L2>(acc[r], x[r][u], qv);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float s = warp_sum(acc[r]);
                if (lane == 0 && row[r] >= 0) cpage[row[r]] = L2 ? -s : s;
            }
        }
    }
}

template <int R, int U, bool L2, bool EXACT, int MINB>
cudaError_t launch_variant(const ScanArgs &a, int num_sms, cudaStream_t st, bool pdl) {
    auto kern = scan_pages_kernel<R, U, L2, EXACT, MINB>;
    // shared memory: one query copy per warp; shrink the CTA when the query is large
    const size_t per_warp = (size_t)a.ds * sizeof(float);
    int wpb = 8;
    const size_t budget = (size_t)(200 * 1024) / MINB;
    while (wpb > 1 && per_warp * wpb > budget) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = num_sms * MINB;
    return launch_pdl(kern, dim3(grid), dim3(wpb * 32), smem, st, pdl, a);
}

template <int R, int U, int MINB>
cudaError_t launch_rum(const ScanArgs &a, int num_sms, cudaStream_t st, bool pdl) {
    const bool exact = ((a.ds >> 2) % (32 * U)) == 0;
    if (a.metric == 1)
        return exact ? launch_variant<R, U, true, true, MINB>(a, num_sms, st, pdl)
                     : launch_variant<R, U, true, false, MINB>(a, num_sms, st, pdl);
    return exact ? launch_variant<R, U, false, true, MINB>(a, num_sms, st, pdl)
                 : launch_variant<R, U, false, false, MINB>(a, num_sms, st, pdl);
}

}  // namespace

// page_off [npairs + 1]: exclusive prefix of the pages of each (query, list) pair; look: plan_pairs_look_words(npairs)
// 64-bit words of device scratch that were zero when first used, epoch: a 22-bit value that differs from the last
// launches on the same scratch (the caller counts up and zeroes the scratch when it wraps)
size_t plan_pairs_look_words(int64_t npairs) { return (size_t)((npairs + PP_BLK - 1) / PP_BLK) + 1; }

cudaError_t launch_plan_pairs(const int32_t *probe, int64_t npairs, const int32_t *list_len, int32_t nlist,
                              int64_t *page_off, unsigned long long *look, uint32_t epoch,
                              unsigned long long *rows_total, cudaStream_t st) {
    if (npairs <= 0) return cudaMemsetAsync(page_off, 0, sizeof(int64_t), st);
    plan_pairs_kernel<<<(unsigned)((npairs + PP_BLK - 1) / PP_BLK), PP_T, 0, st>>>(probe, npairs, list_len, nlist, page_off,
                                                                                  look, epoch & 0x3fffffu, rows_total);
    return cudaGetLastError();
}

// variant: 0 = auto, 1 = R2/U6 x2 CTAs, 2 = R4/U6 x1 CTA, 3 = R2/U4 x2 CTAs, 4 = R4/U4 x2 CTAs
cudaError_t launch_scan_pages(const ScanArgs &a, int variant, int num_sms, int *launches, cudaStream_t st, bool pdl) {
    if (a.npairs <= 0) return cudaSuccess;
    if (launches) *launches += 1;
    const int ds4 = a.ds >> 2;
    if (variant == 0) variant = (ds4 % 192 == 0) ? 1 : 3;
    switch (variant) {
        case 1:
            return launch_rum<2, 6, 2>(a, num_sms, st, pdl);
        case 2:
            return launch_rum<4, 6, 1>(a, num_sms, st, pdl);
        case 3:
            return launch_rum<2, 4, 2>(a, num_sms, st, pdl);
        case 4:
            return launch_rum<4, 4, 2>(a, num_sms, st, pdl);
        default:
            return cudaErrorInvalidValue;
    }
}

}  // namespace sc
