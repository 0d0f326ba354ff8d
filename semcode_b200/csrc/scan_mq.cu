// K5c: multi-query page scan -- the HBM-bound end of the list-major path.
//
// A list that is probed by only 1..4 queries of the batch is streamed ONCE, exactly the way the
// query-major kernel (scan.cu) streams it -- persistent warps with equal page ranges, 128-bit
// ld.global.nc.L1::no_allocate loads, R rows x U float4 per lane in flight, fused tag predicate --
// and every row is scored against all of the list's queries (staged in the warp's shared-memory
// slice) before the registers are recycled.  Compared with the shared-memory tiles of scan_lists.cu
// this keeps ~100 KB of loads in flight per SM with no CTA barrier, which is what the 4.3 TB/s ceiling
// of the cp.async tiles was missing.  Results go to the same per-pair candidate layout.
// Replaces the same FAISS IVFFlatScanner::scan_codes loop (reference src/semcode/storage/milvus_store.py:141-147).
// Algorithmic bytes: rows of each list x 4 x dim, once (compulsory).
#include "common.cuh"

namespace sc {

namespace {

constexpr int MQ = 4;  // queries per pass

__global__ void mq_pages_kernel(const int32_t *__restrict__ n4, const int32_t *__restrict__ list_len, int32_t nlist,
                                int32_t *__restrict__ pages) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    pages[l] = n4[l] ? (list_len[l] + kPageRows - 1) / kPageRows : 0;
}

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

// pgoff [nlist+1]: exclusive prefix of the pages of the lists handled here (0 for the others)
template <int R, int U, bool L2, bool EXACT>
__global__ void __launch_bounds__(256, 2) scan_mq_kernel(const ScanArgs a, const ListPlan p, const int32_t *__restrict__ pgoff) {
    extern __shared__ __align__(16) float4 qsmem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int ds4 = a.ds >> 2;
    float4 *qs = qsmem + (size_t)warp * MQ * ds4;  // [MQ][ds4]

    const int32_t W = pgoff[p.nlist];
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int64_t gw = (int64_t)blockIdx.x * wpb + warp;
    const int32_t per = (int32_t)((W + nwarps - 1) / nwarps);
    const int64_t w0l = gw * per;
    if (w0l >= W) return;
    const int32_t w0 = (int32_t)w0l;
    const int32_t w1 = (w0 + per < W) ? (w0 + per) : W;

    // list that owns page w0: last l with pgoff[l] <= w0 (lists without pages are skipped)
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
    int32_t len = 0, ptbase = 0, nqi = 0;
    int64_t cb[MQ];
#pragma unroll
    for (int j = 0; j < MQ; ++j) cb[j] = -1;
    bool fresh = true;
    const int slab_mask = (1 << a.slab_shift) - 1;

    for (int32_t w = w0; w < w1; ++w) {
        while (w >= l_end) {
            ++l;
            l_start = l_end;
            l_end = pgoff[l + 1];
            fresh = true;
        }
        if (fresh) {
            fresh = false;
            len = a.list_len[l];
            ptbase = a.pt_off[l];
            const int32_t qbase = p.lq_off[l] + 32 * p.n32[l];
            nqi = min(MQ, p.lq_off[l + 1] - qbase);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < MQ; ++j) {
                cb[j] = -1;
                if (j < nqi) {
                    const int32_t pair = p.lq[qbase + j];
                    cb[j] = a.page_off[pair] * kPageRows;
                    const float4 *qg = reinterpret_cast<const float4 *>(a.q + (int64_t)(pair / a.nprobe) * a.ds);
                    for (int c = lane; c < ds4; c += 32) qs[j * ds4 + c] = __ldg(qg + c);
                } else {
                    for (int c = lane; c < ds4; c += 32) qs[j * ds4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            __syncwarp();
        }
        const int32_t jpage = w - l_start;
        const int32_t page = __ldg(a.pt + ptbase + jpage);
        const int slab = page >> a.slab_shift;
        const int64_t slot0 = (int64_t)(page & slab_mask) * kPageRows;
        const int rows = min(kPageRows, len - jpage * kPageRows);
        const uint32_t tag = __ldg(a.slabs->tags[slab] + slot0 + lane);
        const bool live = lane < rows && filter_pass(a.filt, tag);
        uint32_t m = __ballot_sync(0xffffffffu, live);
        const int64_t poff = (int64_t)jpage * kPageRows;
        if (!live) {
#pragma unroll
            for (int j = 0; j < MQ; ++j)
                if (cb[j] >= 0) a.cand[cb[j] + poff + lane] = -INFINITY;
        }
        const float4 *vbase = reinterpret_cast<const float4 *>(a.slabs->vec[slab]) + slot0 * ds4;

        while (m) {
            int row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                row[r] = m ? (__ffs(m) - 1) : -1;
                m &= m - 1;
            }
            float acc[R][MQ];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < MQ; ++j) acc[r][j] = 0.f;
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
                    const bool kin = EXACT || c0 + lane + 32 * u < ds4;
#pragma unroll
                    for (int j = 0; j < MQ; ++j) {
                        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (kin) qv = qs[j * ds4 + c0 + lane + 32 * u];
                        if (L2 && !kin) continue;  // padded lanes: x == q == 0 anyway
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r][j] = mq_accum4<L2>(acc[r][j], x[r][u], qv);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < MQ; ++j) {
                    const float s = warp_sum(acc[r][j]);
                    if (lane == 0 && row[r] >= 0 && cb[j] >= 0) a.cand[cb[j] + poff + row[r]] = L2 ? -s : s;
                }
        }
    }
}

template <int R, int U, bool L2, bool EXACT>
cudaError_t launch_mq_variant(const ScanArgs &a, const ListPlan &p, const int32_t *pgoff, int num_sms, cudaStream_t st) {
    auto kern = scan_mq_kernel<R, U, L2, EXACT>;
    const size_t per_warp = (size_t)MQ * a.ds * sizeof(float);
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > (size_t)100 * 1024) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<num_sms * 2, wpb * 32, smem, st>>>(a, p, pgoff);
    return cudaGetLastError();
}

}  // namespace

// scratch: pages [nlist], pgoff [nlist+1] (int32)
cudaError_t launch_scan_mq(const ScanArgs &a, const ListPlan &p, int32_t *pages, int32_t *pgoff, int num_sms, cudaStream_t st) {
    mq_pages_kernel<<<(p.nlist + 255) / 256, 256, 0, st>>>(p.n4, a.list_len, p.nlist, pages);
    cudaError_t e = launch_exclusive_scan_i32(pages, p.nlist, pgoff, st);
    if (e != cudaSuccess) return e;
    const int ds4 = a.ds >> 2;
    const bool l2 = a.metric == 1;
    if (ds4 % 192 == 0)
        return l2 ? launch_mq_variant<2, 6, true, true>(a, p, pgoff, num_sms, st) : launch_mq_variant<2, 6, false, true>(a, p, pgoff, num_sms, st);
    if (ds4 % 128 == 0)
        return l2 ? launch_mq_variant<2, 4, true, true>(a, p, pgoff, num_sms, st) : launch_mq_variant<2, 4, false, true>(a, p, pgoff, num_sms, st);
    return l2 ? launch_mq_variant<2, 4, true, false>(a, p, pgoff, num_sms, st) : launch_mq_variant<2, 4, false, false>(a, p, pgoff, num_sms, st);
}

}  // namespace sc
