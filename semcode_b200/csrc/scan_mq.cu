// K5c: multi-query page scan -- the HBM-bound end of the list-major path.
//
// A list whose remainder group holds 1..MQ queries of the batch is streamed ONCE per group, exactly the way
// the query-major kernel (scan.cu) streams it -- persistent warps with equal page ranges, 128-bit
// ld.global.nc.L1::no_allocate loads, R rows x U float4 per lane in flight, fused tag predicate -- and every
// row is scored against all MQ queries of the group before the registers are recycled.
//
//   bucket 0: MQ = 4, R = 2, U <= 6   lists with 1..4 (remaining) queries, one pass
//   bucket 1: MQ = 8, R = 4, U <= 3   lists with 5..16 (remaining) queries, one or two passes of 8
//
// The vector dimension is cut into SLICES of 32*U float4 (<= 768 floats for MQ = 4, <= 384 for MQ = 8), so the
// per-warp query stage is MQ * 32*U * 16 B <= 12 KB whatever the dimension is (16 warps per SM at dim 3072 as
// at dim 768).  A warp walks its pages in groups (consecutive pages of one list and pass): per (group, slice) it
// stages the MQ query slices from L2 once, streams that slice of every live row of every page of the group, reduces
// the R*MQ partial sums (MQ = 8: one transposing butterfly, 31 shuffles for 32 sums instead of 160, into a per-warp
// [MQ][32] shared-memory page accumulator) and writes the page's 32 x MQ values to the candidate array with
// coalesced 128-byte stores -- adding the earlier slices' partial sums, which it left there itself.
// Tried for MQ = 8 (4.3 TB/s: 47 % issue-active): packed fp32 pairs (fma.rn.f32x2) with R = 2 and a register
// prefetch of the next rows -- half the FMA issue slots but twice the query LDS per byte: the same 460 us.
// Replaces the same FAISS IVFFlatScanner::scan_codes loop (reference src/semcode/storage/milvus_store.py:141-147).
// Algorithmic bytes: rows of each list x 4 x dim, once per pass (compulsory for one pass).
#include "common.cuh"

namespace sc {

namespace {

template <bool L2>
__device__ __forceinline__ float mq_accum4(float acc, const float4 &x, const float4 &q) {
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

// Sum each of v[0..N) over the 32 lanes; returns, in lane L, the total of v[L / (32 / N)]  (N = 8 or 32).
// Halving exchange: at offset o the upper lanes keep the upper half of the live values and send the lower half.
template <int N>
__device__ __forceinline__ float reduce_transpose(float (&v)[N], int lane) {
    int o = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    float r = v[0];
#pragma unroll
    for (int oo = (32 / N) >> 1; oo > 0; oo >>= 1) r += __shfl_xor_sync(0xffffffffu, r, oo);
    return r;
}

constexpr int kTotLd = 33;  // page accumulator [MQ][33]: row-major over rows, one pad word per query

// per-warp shared memory: query slices [MQ][32*U] float4 | candidate bases [MQ] int64 | query pointers [MQ] |
// (MQ = 8 only) page accumulator [MQ][33] float.  With MQ = 4 the page accumulator lives in registers (lane = row):
// two CTAs of 8 warps then need 2 x 97.5 KB, which leaves the 196 KB carve-out and 60 KB of L1 for the streaming
// loads in flight (measured: the same kernel with 2 x 101 KB -> 228 KB carve-out, 28 KB of L1, streams 11 % slower)
template <int MQ, int U>
__host__ __device__ constexpr size_t mq_warp_floats() {
    return ((size_t)MQ * 32 * U * 4 + (size_t)MQ * 4 + (MQ > 4 ? (size_t)MQ * kTotLd : 0) + 3) / 4 * 4;
}

// pgoff [nlist+1]: exclusive prefix of the units (pages x passes) of the lists handled here (0 for the others)
// SO (slice-outer): the order described above.  SO = false walks a group page by page and every page slice by slice
// (queries restaged per (page, slice), the page's sums stay in registers / shared memory across the slices, ONE
// candidate store per page): chosen when a scalar filter is active -- with 5 % of the rows live the streaming is
// short and the per-(page, slice) read-modify-write of the candidates costs more than the restaging
// (10M x 2048, 5 % selectivity, nprobe 64: 0.93 vs 1.15 ms).  With one slice the two orders coincide.
// The work of ONE warp on ONE bucket: `qs` is the warp's own shared memory (mq_warp_floats<MQ, U>() floats).
template <int MQ, int R, int U, bool L2, bool EXACT, bool SO>
__device__ __forceinline__ void scan_mq_warp(const ScanArgs &a, const ListPlan &p, const int32_t *__restrict__ pgoff, float4 *qs) {
    constexpr int N = R * MQ;
    constexpr int SL4 = 32 * U;  // float4 per slice
    static_assert(N == 8 || N == 32, "reduce_transpose covers 8 or 32 partial sums");
    constexpr bool SMEM_TOT = MQ > 4;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int ds4 = a.ds >> 2;
    const int nslices = (ds4 + SL4 - 1) / SL4;
    int64_t *cbs = reinterpret_cast<int64_t *>(qs + MQ * SL4);                    // [MQ] candidate base or -1
    const float4 **qgs = reinterpret_cast<const float4 **>(cbs + MQ);             // [MQ] query row or nullptr
    float *tot = reinterpret_cast<float *>(qgs + MQ);                             // [MQ][kTotLd]

    const int32_t W = pgoff[p.nlist];
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int64_t gw = (int64_t)blockIdx.x * wpb + warp;
    // balanced contiguous ranges: warp g takes units [g W / nwarps, (g + 1) W / nwarps) -- sizes differ by at most one and,
    // when there are fewer units than warps, the busy warps are spread over all SMs (ranges of ceil(W / nwarps) units left
    // half of the warps idle on an 8-way shard: `scan_mq<8>` 100 us for 0.26 GB)
    const int32_t w0 = (int32_t)(gw * (int64_t)W / nwarps);
    const int32_t w1 = (int32_t)((gw + 1) * (int64_t)W / nwarps);
    if (w0 >= w1) return;

    // list that owns unit w0: last l with pgoff[l] <= w0 (lists without units are skipped)
    int32_t lo = 0, hi = p.nlist;
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (pgoff[mid] <= w0)
            lo = mid;
        else
            hi = mid;
    }
    int32_t l = lo;
    int32_t l_start = pgoff[l], l_end = pgoff[l + 1];
    const int slab_mask = (1 << a.slab_shift) - 1;
    // which (row, query) total this lane receives from reduce_transpose
    const int my_idx = lane / (32 / N);
    const int my_r = my_idx / MQ, my_j = my_idx % MQ;
    const bool my_owner = (lane % (32 / N)) == 0;

    // The warp's units are walked in GROUPS: the run of consecutive pages of one (list, pass) inside [w0, w1).
    // Per group the query slices are staged ONCE per slice and every page of the group is streamed against them
    // (slice-outer, page-inner); with more than one slice the per-row partial sums of a page travel through the
    // candidate array itself (4 B per row and query, L2-resident, re-read by the lane that wrote them) instead of
    // restaging 12 KB of queries per (page, slice) as the first version did.
    int32_t w = w0;
    while (w < w1) {
        while (w >= l_end) {
            ++l;
            l_start = l_end;
            l_end = pgoff[l + 1];
        }
        const int32_t len = a.list_len[l];
        const int32_t ptbase = a.pt_off[l];
        const int32_t npg = (len + kPageRows - 1) / kPageRows;
        const int32_t ul = w - l_start;
        const int32_t pass = ul / npg;
        const int32_t jpage0 = ul - pass * npg;
        const int32_t wend = min(w1, l_start + (pass + 1) * npg);
        const int32_t G = wend - w;
        {
            const int32_t qbase = p.lq_off[l] + p.chunk * p.n32[l] + MQ * pass;
            const int nqi = min(MQ, p.lq_off[l + 1] - qbase);
            __syncwarp();  // the previous group's readers are done with cbs / qgs
            if (lane < MQ) {
                int64_t cb = -1;
                const float4 *qg = nullptr;
                if (lane < nqi) {
                    const int32_t pair = p.lq[qbase + lane];
                    cb = a.page_off[pair] * kPageRows;
                    qg = reinterpret_cast<const float4 *>(a.q + (int64_t)(pair / a.nprobe) * a.ds);
                }
                cbs[lane] = cb;
                qgs[lane] = qg;
            }
            __syncwarp();
        }
        // stage the MQ query slices of slice s
        auto stage = [&](int s) {
            const int c0 = s * SL4;
            __syncwarp();  // the previous readers are done with qs
#pragma unroll
            for (int j = 0; j < MQ; ++j) {
                const float4 *qg = qgs[j];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + lane + 32 * u;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (qg != nullptr && (EXACT || c < ds4)) v = __ldg(qg + c);
                    qs[j * SL4 + lane + 32 * u] = v;
                }
            }
        };
        // stream slice s of the live rows of one page against the staged queries; sums are ADDED to treg / tot
        auto stream = [&](int s, const float4 *vbase, uint32_t live_mask, float(&treg)[MQ]) {
            const int c0 = s * SL4;
            uint32_t m = live_mask;
            while (m) {
                int row[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    row[r] = m ? (__ffs(m) - 1) : -1;
                    m &= m - 1;
                }
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
                float acc[N];
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] = 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int j = 0; j < MQ; ++j) {
                        const float4 qv = qs[j * SL4 + lane + 32 * u];
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r * MQ + j] = mq_accum4<L2>(acc[r * MQ + j], x[r][u], qv);
                    }
                }
                if (SMEM_TOT) {
                    const float t = reduce_transpose<N>(acc, lane);
                    int myrow = row[0];
#pragma unroll
                    for (int r = 1; r < R; ++r)
                        if (my_r == r) myrow = row[r];
                    if (my_owner && myrow >= 0) tot[my_j * kTotLd + myrow] += t;  // one owner per (row, query): no race
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int j = 0; j < MQ; ++j) {
                            const float t = warp_sum(acc[r * MQ + j]);
                            if (lane == row[r]) treg[j] += t;
                        }
                }
            }
        };
        struct PageRef {
            const float4 *vbase;
            uint32_t live_mask;
            int64_t poff;
            bool live;
        };
        auto open_page = [&](int32_t jpage) {
            const int32_t page = __ldg(a.pt + ptbase + jpage);
            const int slab = page >> a.slab_shift;
            const int64_t slot0 = (int64_t)(page & slab_mask) * kPageRows;
            const int rows = min(kPageRows, len - jpage * kPageRows);
            const uint32_t tag = __ldg(a.slabs->tags[slab] + slot0 + lane);
            PageRef pr;
            pr.live = lane < rows && filter_pass(a.filt, tag);
            pr.live_mask = __ballot_sync(0xffffffffu, pr.live);
            pr.poff = (int64_t)jpage * kPageRows;
            pr.vbase = reinterpret_cast<const float4 *>(a.slabs->vec[slab]) + slot0 * ds4;
            return pr;
        };
        if (SO || nslices == 1) {
            for (int s = 0; s < nslices; ++s) {
                stage(s);
                for (int32_t g = 0; g < G; ++g) {
                    const PageRef pr = open_page(jpage0 + g);
                    float treg[MQ];  // MQ = 4: this lane's row of the page, one total per query
#pragma unroll
                    for (int j = 0; j < MQ; ++j) {
                        treg[j] = 0.f;
                        if (SMEM_TOT) tot[j * kTotLd + lane] = 0.f;
                    }
                    __syncwarp();  // qs and the zeroed tot visible to the whole warp
                    stream(s, pr.vbase, pr.live_mask, treg);
                    if (SMEM_TOT) __syncwarp();  // the page's totals are complete
                    const bool first = s == 0, final = s == nslices - 1;
#pragma unroll
                    for (int j = 0; j < MQ; ++j) {
                        const int64_t cb = cbs[j];
                        if (cb >= 0) {
                            float v = SMEM_TOT ? tot[j * kTotLd + lane] : treg[j];
                            float *dst = a.cand + cb + pr.poff + lane;
                            if (!first) v += *dst;  // this lane's own partial sum of the earlier slices
                            if (final) v = pr.live ? (L2 ? -v : v) : -INFINITY;
                            *dst = v;
                        }
                    }
                    if (SMEM_TOT) __syncwarp();  // tot is re-zeroed for the next page
                }
            }
        } else {
            for (int32_t g = 0; g < G; ++g) {
                const PageRef pr = open_page(jpage0 + g);
                float treg[MQ];
#pragma unroll
                for (int j = 0; j < MQ; ++j) {
                    treg[j] = 0.f;
                    if (SMEM_TOT) tot[j * kTotLd + lane] = 0.f;
                }
                for (int s = 0; s < nslices; ++s) {
                    stage(s);
                    __syncwarp();  // qs (and, for s = 0, the zeroed tot) visible to the whole warp
                    stream(s, pr.vbase, pr.live_mask, treg);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < MQ; ++j) {
                    const int64_t cb = cbs[j];
                    if (cb >= 0) {
                        const float v = SMEM_TOT ? tot[j * kTotLd + lane] : treg[j];
                        a.cand[cb + pr.poff + lane] = pr.live ? (L2 ? -v : v) : -INFINITY;
                    }
                }
                __syncwarp();  // tot is re-zeroed for the next page
            }
        }
        w = wend;
    }
}

template <int MQ, int R, int U, bool L2, bool EXACT, bool SO>
__global__ void __launch_bounds__(256, 2)
    scan_mq_kernel(const ScanArgs a, const ListPlan p, const int32_t *__restrict__ pgoff) {
    extern __shared__ __align__(16) float4 qsmem[];
    constexpr size_t WF = mq_warp_floats<MQ, U>();  // floats per warp, 16-byte multiple
    scan_mq_warp<MQ, R, U, L2, EXACT, SO>(a, p, pgoff, qsmem + (size_t)(threadIdx.x >> 5) * (WF / 4));
}

// Both buckets in ONE launch (set_param "mq_fused" = 1; NOT the default).  Every warp takes its 1/nwarps of the 4-query bucket
// AND of the 8-query bucket; odd warps start with the 8-query bucket, even warps end with it.  The idea: the 8-query scan is
// issue-bound (47 % issue-active at 4.5 TB/s), the 4-query scan memory-bound (84 % of DRAM peak at 37 % issue-active), so side
// by side on an SM they should fill each other's idle resource and save the tail + ramp between two launches.  Measured: it
// loses -- headline scan 4.77 vs 4.10 ms, 8-way shard 0.671 vs 0.603 ms, clustered 4.98 vs 4.69 ms.  The shared per-warp
// stage is the 8-query one (13.2 KB: 2 x 108 KB per SM -> the 228 KB carve-out and 28 KB of L1, the configuration in which the
// 4-query scan alone already streamed 11 % slower), and the 8-query warps' shared-memory traffic now competes with the
// 4-query warps' loads for the same L1 pipe for the whole launch instead of 10 % of the step.
template <int U4, bool E4, int U8, bool E8, bool L2, bool SO>
__global__ void __launch_bounds__(256, 2)
    scan_mq_both_kernel(const ScanArgs a, const ListPlan p, const int32_t *__restrict__ pg4off, const int32_t *__restrict__ pg8off) {
    extern __shared__ __align__(16) float4 qsmem[];
    constexpr size_t WF4 = mq_warp_floats<4, U4>(), WF8 = mq_warp_floats<8, U8>();
    constexpr size_t WF = WF4 > WF8 ? WF4 : WF8;
    const int warp = threadIdx.x >> 5;
    float4 *qs = qsmem + (size_t)warp * (WF / 4);
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
        if ((phase == 0) == ((warp & 1) != 0))
            scan_mq_warp<8, 4, U8, L2, E8, SO>(a, p, pg8off, qs);
        else
            scan_mq_warp<4, 2, U4, L2, E4, SO>(a, p, pg4off, qs);
        __syncwarp();
    }
}

// ---- 8-query page scan on the (legacy) warp-level tensor cores (lists_cfg = 4; measured slower, see launch_scan_mq) ----
// The scalar 8-query scan above is issue-bound (47 % issue-active at 4.3 TB/s: 384 FMAs, 24 LDS.128 and a 31-shuffle
// butterfly per 6 KB).  Here a page is two 16-row MMA tiles and the 8 queries are exactly one n-tile of
// mma.sync.m16n8k8 (tf32 x tf32 -> fp32): the A fragments come straight from global memory in the fragment layout
// (lane (g, t) loads the float4 at row g / g + 8, columns 16*kc + 4t of each tile: 8 rows x 64 B per instruction,
// every 32-byte sector used once), the k index of the MMA is a permutation of those 16 columns that A and B share,
// and the MMA does the reduction over the dimension -- no shuffles.  fp32 accuracy as in the tcgen05 tiles:
// x = hi + lo with hi = x & 0xffffe000 (exact in tf32) and lo = x - hi, product = hi.hi + hi.lo + lo.hi, hi.hi in
// separate accumulators for the two k-steps of a chunk; the split of both operands happens in registers.
// Inner product, dim % 16 == 0, no scalar filter (a filtered page has few live rows: the scalar kernel streams
// only those).  Same units, groups, candidate layout and slice-outer order as scan_mq_kernel.
constexpr int MM_SLF = 384;            // floats per slice
constexpr int MM_QLD = MM_SLF + 16;    // padded query row (floats): consecutive queries shift by 64 B -> the
                                       // quarter-warp's two query rows of an LDS.128 hit disjoint banks
constexpr size_t MM_WARP_BYTES = (size_t)8 * MM_QLD * 4 + 8 * 8 + 8 * 8;

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }

__global__ void __launch_bounds__(256, 2)
    scan_mq8_mma_kernel(const ScanArgs a, const ListPlan p, const int32_t *__restrict__ pgoff) {
    extern __shared__ __align__(16) float4 qsmem[];
    constexpr int MQ = 8;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ds4 = a.ds >> 2;
    const int nslices = (a.ds + MM_SLF - 1) / MM_SLF;
    float4 *qs = reinterpret_cast<float4 *>(reinterpret_cast<char *>(qsmem) + (size_t)warp * MM_WARP_BYTES);  // [8][MM_QLD / 4]
    int64_t *cbs = reinterpret_cast<int64_t *>(qs + MQ * (MM_QLD / 4));
    const float4 **qgs = reinterpret_cast<const float4 **>(cbs + MQ);

    const int32_t W = pgoff[p.nlist];
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int64_t gw = (int64_t)blockIdx.x * wpb + warp;
    // balanced contiguous ranges: warp g takes units [g W / nwarps, (g + 1) W / nwarps) -- sizes differ by at most one and,
    // when there are fewer units than warps, the busy warps are spread over all SMs (ranges of ceil(W / nwarps) units left
    // half of the warps idle on an 8-way shard: `scan_mq<8>` 100 us for 0.26 GB)
    const int32_t w0 = (int32_t)(gw * (int64_t)W / nwarps);
    const int32_t w1 = (int32_t)((gw + 1) * (int64_t)W / nwarps);
    if (w0 >= w1) return;
    int32_t lo = 0, hi = p.nlist;  // list that owns unit w0: last l with pgoff[l] <= w0
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (pgoff[mid] <= w0)
            lo = mid;
        else
            hi = mid;
    }
    int32_t l = lo;
    int32_t l_start = pgoff[l], l_end = pgoff[l + 1];
    const int slab_mask = (1 << a.slab_shift) - 1;

    int32_t w = w0;
    while (w < w1) {
        while (w >= l_end) {
            ++l;
            l_start = l_end;
            l_end = pgoff[l + 1];
        }
        const int32_t len = a.list_len[l];
        const int32_t ptbase = a.pt_off[l];
        const int32_t npg = (len + kPageRows - 1) / kPageRows;
        const int32_t ul = w - l_start;
        const int32_t pass = ul / npg;
        const int32_t jpage0 = ul - pass * npg;
        const int32_t wend = min(w1, l_start + (pass + 1) * npg);
        const int32_t G = wend - w;
        {
            const int32_t qbase = p.lq_off[l] + p.chunk * p.n32[l] + MQ * pass;
            const int nqi = min(MQ, p.lq_off[l + 1] - qbase);
            __syncwarp();  // the previous group's readers are done with cbs / qgs
            if (lane < MQ) {
                int64_t cb = -1;
                const float4 *qg = nullptr;
                if (lane < nqi) {
                    const int32_t pair = p.lq[qbase + lane];
                    cb = a.page_off[pair] * kPageRows;
                    qg = reinterpret_cast<const float4 *>(a.q + (int64_t)(pair / a.nprobe) * a.ds);
                }
                cbs[lane] = cb;
                qgs[lane] = qg;
            }
            __syncwarp();
        }
        const int64_t cb0 = cbs[2 * t], cb1 = cbs[2 * t + 1];  // this lane's two queries (C fragment columns 2t, 2t + 1)
        for (int s = 0; s < nslices; ++s) {
            const int c0f = s * MM_SLF;                              // first float of the slice
            const int nkc = (min(MM_SLF, a.ds - c0f)) >> 4;         // 16-column chunks in the slice
            __syncwarp();  // the previous slice's readers are done with qs
#pragma unroll
            for (int j = 0; j < MQ; ++j) {
                const float4 *qg = qgs[j];
                for (int c = lane; c < 4 * nkc; c += 32)
                    qs[j * (MM_QLD / 4) + c] = qg != nullptr ? __ldg(qg + (c0f >> 2) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
            const float4 *qrow = qs + g * (MM_QLD / 4) + t;  // query g, float4 t of every chunk
            for (int32_t gi = 0; gi < G; ++gi) {
                const int32_t jpage = jpage0 + gi;
                const int32_t page = __ldg(a.pt + ptbase + jpage);
                const int slab = page >> a.slab_shift;
                const int64_t slot0 = (int64_t)(page & slab_mask) * kPageRows;
                const int rows = min(kPageRows, len - jpage * kPageRows);
                const uint32_t tag = __ldg(a.slabs->tags[slab] + slot0 + lane);
                const bool live = lane < rows && filter_pass(a.filt, tag);
                const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
                const int64_t poff = (int64_t)jpage * kPageRows;
                // rows g, g + 8 (tile 0) and g + 16, g + 24 (tile 1), float4 t of chunk kc
                const float4 *rp = reinterpret_cast<const float4 *>(a.slabs->vec[slab]) + (slot0 + g) * ds4 + (c0f >> 2) + t;
                bool rl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) rl[i] = (live_mask >> (g + 8 * i)) & 1u;
                float acc[2][3][4];  // [tile][hh of k-step 0, hh of k-step 1, cross][C fragment]
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[mt][k][i] = 0.f;
#pragma unroll 1
                for (int kc = 0; kc < nkc; kc += 2) {
                    float4 x[2][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (rl[i] && kc + u < nkc)
                                x[u][i] = ld_stream_f4(rp + (int64_t)(8 * i) * ds4 + 4 * (kc + u));
                            else
                                x[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (kc + u < nkc) {
                            const float4 qv = qrow[4 * (kc + u)];
                            const float qf[4] = {qv.x, qv.y, qv.z, qv.w};
                            uint32_t qh[4], ql[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                qh[e] = tf32_hi(qf[e]);
                                ql[e] = tf32_lo(qf[e], qh[e]);
                            }
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt) {
                                const float4 xa = x[u][2 * mt], xb = x[u][2 * mt + 1];  // rows g + 16 mt, g + 8 + 16 mt
                                const float fa[4] = {xa.x, xa.y, xa.z, xa.w}, fb[4] = {xb.x, xb.y, xb.z, xb.w};
                                uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    ah[e] = tf32_hi(fa[e]);
                                    al[e] = tf32_lo(fa[e], ah[e]);
                                    bh[e] = tf32_hi(fb[e]);
                                    bl[e] = tf32_lo(fb[e], bh[e]);
                                }
#pragma unroll
                                for (int st = 0; st < 2; ++st) {  // k-step st: k = t <-> column 4t + 2 st, k = t + 4 <-> column 4t + 2 st + 1
                                    const int e0 = 2 * st, e1 = 2 * st + 1;
                                    mma_tf32_16x8x8(acc[mt][st], ah[e0], bh[e0], ah[e1], bh[e1], qh[e0], qh[e1]);
                                    mma_tf32_16x8x8(acc[mt][2], ah[e0], bh[e0], ah[e1], bh[e1], ql[e0], ql[e1]);
                                    mma_tf32_16x8x8(acc[mt][2], al[e0], bl[e0], al[e1], bl[e1], qh[e0], qh[e1]);
                                }
                            }
                        }
                    }
                }
                const bool first = s == 0, final = s == nslices - 1;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {  // C fragment: i = 0, 1 -> row g, queries 2t, 2t + 1; i = 2, 3 -> row g + 8
                        const int64_t cb = (i & 1) ? cb1 : cb0;
                        if (cb >= 0) {
                            const int ri = 2 * mt + (i >> 1);  // index into rl: rows g + 8 ri
                            float v = (acc[mt][0][i] + acc[mt][1][i]) + acc[mt][2][i];
                            float *dst = a.cand + cb + poff + (g + 8 * ri);
                            if (!first) v += *dst;  // this lane's own partial sum of the earlier slices
                            if (final) v = rl[ri] ? v : -INFINITY;
                            *dst = v;
                        }
                    }
            }
        }
        w = wend;
    }
}

cudaError_t launch_mq8_mma(const ScanArgs &a, const ListPlan &p, const int32_t *pgoff, int num_sms, cudaStream_t st) {
    constexpr int wpb = 8;
    constexpr size_t smem = MM_WARP_BYTES * wpb;
    static_assert(MM_WARP_BYTES % 16 == 0 && 2 * (smem + 1024) <= 227 * 1024, "two CTAs per SM must fit");
    cudaError_t e = cudaFuncSetAttribute(scan_mq8_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    scan_mq8_mma_kernel<<<num_sms * 2, wpb * 32, smem, st>>>(a, p, pgoff);
    return cudaGetLastError();
}

template <int MQ, int R, int U, bool L2, bool EXACT, bool SO>
cudaError_t launch_mq_order(const ScanArgs &a, const ListPlan &p, const int32_t *pgoff, int num_sms, cudaStream_t st) {
    auto kern = scan_mq_kernel<MQ, R, U, L2, EXACT, SO>;
    constexpr size_t per_warp = mq_warp_floats<MQ, U>() * sizeof(float);
    constexpr int wpb = 8;
    constexpr size_t smem = per_warp * wpb;
    static_assert(2 * (smem + 1024) <= 227 * 1024, "two CTAs per SM must fit");
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<num_sms * 2, wpb * 32, smem, st>>>(a, p, pgoff);
    return cudaGetLastError();
}

template <int MQ, int R, int U, bool L2, bool EXACT>
cudaError_t launch_mq_variant(const ScanArgs &a, const ListPlan &p, const int32_t *pgoff, int num_sms, cudaStream_t st) {
    // a scalar filter leaves few live rows per page: page-outer order (see SO above)
    return a.filt.flags != 0 ? launch_mq_order<MQ, R, U, L2, EXACT, false>(a, p, pgoff, num_sms, st)
                             : launch_mq_order<MQ, R, U, L2, EXACT, true>(a, p, pgoff, num_sms, st);
}

template <int MQ, int R, int U, bool EXACT>
cudaError_t launch_mq_metric(const ScanArgs &a, const ListPlan &p, const int32_t *pgoff, int num_sms, cudaStream_t st) {
    return a.metric == 1 ? launch_mq_variant<MQ, R, U, true, EXACT>(a, p, pgoff, num_sms, st)
                         : launch_mq_variant<MQ, R, U, false, EXACT>(a, p, pgoff, num_sms, st);
}

template <int U4, bool E4, int U8, bool E8>
cudaError_t launch_mq_both_cfg(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st) {
    constexpr size_t WF4 = mq_warp_floats<4, U4>(), WF8 = mq_warp_floats<8, U8>();
    constexpr size_t per_warp = (WF4 > WF8 ? WF4 : WF8) * sizeof(float);
    constexpr int wpb = 8;
    constexpr size_t smem = per_warp * wpb;
    static_assert(2 * (smem + 1024) <= 227 * 1024, "two CTAs per SM must fit");
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<num_sms * 2, wpb * 32, smem, st>>>(a, p, p.pg4off, p.pg8off);
        return cudaGetLastError();
    };
    const bool so = a.filt.flags == 0;  // a scalar filter leaves few live rows per page: page-outer order (see SO above)
    if (a.metric == 1) return so ? go(scan_mq_both_kernel<U4, E4, U8, E8, true, true>) : go(scan_mq_both_kernel<U4, E4, U8, E8, true, false>);
    return so ? go(scan_mq_both_kernel<U4, E4, U8, E8, false, true>) : go(scan_mq_both_kernel<U4, E4, U8, E8, false, false>);
}

}  // namespace

// Both page-scan buckets in one launch (same kernels' bodies, see scan_mq_both_kernel).
cudaError_t launch_scan_mq_both(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st) {
    const int ds4 = a.ds >> 2;
    if (ds4 % 192 == 0) return launch_mq_both_cfg<6, true, 3, true>(a, p, num_sms, st);  // 768, 1536, 2304, 3072, ...
    if (ds4 % 128 == 0) return launch_mq_both_cfg<4, true, 2, true>(a, p, num_sms, st);  // 512, 1024, 2048, ...
    if (ds4 % 96 == 0) return launch_mq_both_cfg<4, false, 3, true>(a, p, num_sms, st);
    if (ds4 % 64 == 0) return launch_mq_both_cfg<4, false, 2, true>(a, p, num_sms, st);
    return launch_mq_both_cfg<4, false, 3, false>(a, p, num_sms, st);
}

// bucket 0: remainders of 1..4 queries (one pass); bucket 1: remainders of 5..16 queries (passes of 8).
// pgoff [nlist+1]: exclusive prefix of the bucket's page x pass units (plan_lists_kernel)
// cfg 4: the 8-query bucket on mma.sync where it applies (inner product, dim % 16 == 0, no filter).  Parity-green but
// SLOWER than the scalar scan on B200 -- headline step 4.79 vs 4.43 ms (profiles/r1_bench40_mq8_mma.json vs
// r1_bench40_mq8_scalar.json): 12 legacy-path MMAs + 40 split instructions per 16 columns do not beat 96 FMAs -- so
// it stays an option
cudaError_t launch_scan_mq(const ScanArgs &a, const ListPlan &p, int bucket, int cfg, const int32_t *pgoff, int num_sms, cudaStream_t st) {
    const int ds4 = a.ds >> 2;
    if (bucket == 0) {
        if (ds4 % 192 == 0) return launch_mq_metric<4, 2, 6, true>(a, p, pgoff, num_sms, st);
        if (ds4 % 128 == 0) return launch_mq_metric<4, 2, 4, true>(a, p, pgoff, num_sms, st);
        return launch_mq_metric<4, 2, 4, false>(a, p, pgoff, num_sms, st);
    }
    if (cfg == 4 && a.metric == 0 && a.ds % 16 == 0 && a.filt.flags == 0) return launch_mq8_mma(a, p, pgoff, num_sms, st);
    if (ds4 % 96 == 0) return launch_mq_metric<8, 4, 3, true>(a, p, pgoff, num_sms, st);
    if (ds4 % 64 == 0) return launch_mq_metric<8, 4, 2, true>(a, p, pgoff, num_sms, st);
    return launch_mq_metric<8, 4, 3, false>(a, p, pgoff, num_sms, st);
}

}  // namespace sc
