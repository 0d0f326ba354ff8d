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
// CTA walks the list in tiles of 128 rows x 64 floats, double-buffered with cp.async (16-byte
// copies, zero-fill for ragged edges), and keeps a 4 x (QT/8) register tile per thread:
//   QT = 32  lists probed by many queries   FP32-pipe bound (16 FMA per 2 LDS.128)
//   QT = 8   lists probed by <= 8 queries   HBM bound
// The result lands in the same per-pair candidate layout the query-major kernel writes, so the
// top-k selection (select.cu) is unchanged.
// Algorithmic bytes: sum over DISTINCT probed lists of len * 4 * dim (compulsory), per chunk item.
#include "common.cuh"

namespace sc {

namespace {

constexpr int RB = 128;      // rows per tile
constexpr int BKX = 64;      // floats per k-stage
constexpr int LDS_ = BKX + 4;  // padded row stride in floats (272 B: consecutive rows shift 4 banks)
constexpr int NT = 256;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- planning: invert (query, list) pairs into per-list query groups ----------------------------------
__global__ void count_list_pairs_kernel(const int32_t *__restrict__ probe, int64_t npairs, const int32_t *__restrict__ list_len,
                                        int32_t nlist, int32_t *__restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int32_t l = probe[i];
    if (l >= 0 && l < nlist && list_len[l] > 0) atomicAdd(cnt + l, 1);
}

__global__ void plan_items_kernel(const int32_t *__restrict__ cnt, int32_t nlist, int32_t *__restrict__ n32,
                                  int32_t *__restrict__ n8) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    const int32_t c = cnt[l];
    int32_t a = c / 32, b = 0;
    const int32_t rem = c - a * 32;
    if (rem > 8)
        ++a;  // a ragged 32-chunk costs less than re-reading the list for several 8-chunks
    else if (rem > 0)
        b = 1;
    n32[l] = a;
    n8[l] = b;
}

__global__ void fill_list_pairs_kernel(const int32_t *__restrict__ probe, int64_t npairs, const int32_t *__restrict__ list_len,
                                       int32_t nlist, const int32_t *__restrict__ lq_off, int32_t *__restrict__ cursor,
                                       int32_t *__restrict__ lq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int32_t l = probe[i];
    if (l >= 0 && l < nlist && list_len[l] > 0) lq[lq_off[l] + atomicAdd(cursor + l, 1)] = (int32_t)i;
}

template <int QT>
struct TileSmem {
    float xs[2][RB][LDS_];
    float qs[2][QT][LDS_];
    const float *rowptr[2][RB];
    uint8_t live[2][RB];
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

template <int QT, bool L2>
__global__ void __launch_bounds__(NT, 2) scan_lists_kernel(const ScanArgs a, const ListPlan p, int which) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TileSmem<QT> &sm = *reinterpret_cast<TileSmem<QT> *>(smem_raw);
    constexpr int NQ = QT / 8;  // queries per thread
    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;  // query group / row group
    const int32_t *item_off = QT == 32 ? p.off32 : p.off8;
    const int32_t total = item_off[p.nlist];
    const int ds = a.ds;
    const int KB = (ds + BKX - 1) / BKX;
    const int slab_mask = (1 << a.slab_shift) - 1;

    for (;;) {
        __syncthreads();  // previous item fully consumed (smem tables reused)
        if (tid == 0) sm.item = atomicAdd(p.counters + which, 1);
        __syncthreads();
        const int32_t item = sm.item;
        if (item >= total) break;
        const int32_t l = owner_of(item_off, p.nlist, item);
        const int32_t chunk = item - item_off[l];
        const int32_t qbase = p.lq_off[l] + (QT == 32 ? 32 * chunk : 32 * p.n32[l]);
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
                sm.rowptr[tile & 1][tid] = ptr;
                sm.live[tile & 1][tid] = ok ? 1 : 0;
            }
        };
        auto issue = [&](int s) {  // cp.async the operands of flattened stage s into buffer s & 1
            const int tile = s / KB, kb = s - tile * KB;
            const int buf = s & 1;
            const int k0 = kb * BKX;
            const int c = tid & 15;  // 16-byte chunk within the 256-byte row segment
            const bool kin = k0 + c * 4 < ds;
#pragma unroll
            for (int i = 0; i < RB / 16; ++i) {
                const int r = (tid >> 4) + 16 * i;
                const bool valid = kin && (tile * RB + r < len);
                cp_async16(&sm.xs[buf][r][c * 4], sm.rowptr[tile & 1][r] + k0 + c * 4, valid);
            }
#pragma unroll
            for (int i = 0; i < (QT * 16 + NT - 1) / NT; ++i) {
                const int idx = tid + i * NT;
                if (idx < QT * 16) {
                    const int j = idx >> 4;
                    cp_async16(&sm.qs[buf][j][c * 4], sm.qptr[j] + k0 + c * 4, kin && j < nqi);
                }
            }
            cp_async_commit();
        };

        prep_rows(0);
        __syncthreads();
        issue(0);

        float acc[4][NQ];
        for (int s = 0; s < S; ++s) {
            const int tile = s / KB, kb = s - tile * KB;
            const int buf = s & 1;
            cp_async_wait_all();
            __syncthreads();  // stage s visible; everyone is done with buffer (s+1)&1 and the older row table
            if (s + 1 < S) issue(s + 1);
            if (kb == 0) {
                if (tile + 1 < ntiles) prep_rows(tile + 1);  // first needed KB-1 >= 1 iterations from now
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NQ; ++j) acc[i][j] = 0.f;
            }
#pragma unroll 4
            for (int k4 = 0; k4 < BKX / 4; ++k4) {
                float4 xv[4], qv[NQ];
#pragma unroll
                for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4 *>(&sm.xs[buf][ty + 32 * i][k4 * 4]);
#pragma unroll
                for (int j = 0; j < NQ; ++j) qv[j] = *reinterpret_cast<const float4 *>(&sm.qs[buf][tx + 8 * j][k4 * 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        if (L2) {
                            const float d0 = xv[i].x - qv[j].x, d1 = xv[i].y - qv[j].y, d2 = xv[i].z - qv[j].z,
                                        d3 = xv[i].w - qv[j].w;
                            acc[i][j] = fmaf(d0, d0, acc[i][j]);
                            acc[i][j] = fmaf(d1, d1, acc[i][j]);
                            acc[i][j] = fmaf(d2, d2, acc[i][j]);
                            acc[i][j] = fmaf(d3, d3, acc[i][j]);
                        } else {
                            acc[i][j] = fmaf(xv[i].x, qv[j].x, acc[i][j]);
                            acc[i][j] = fmaf(xv[i].y, qv[j].y, acc[i][j]);
                            acc[i][j] = fmaf(xv[i].z, qv[j].z, acc[i][j]);
                            acc[i][j] = fmaf(xv[i].w, qv[j].w, acc[i][j]);
                        }
                    }
            }
            if (kb == KB - 1) {  // tile finished: one candidate per (row slot, query)
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const int64_t cb = sm.cbase[tx + 8 * j];
                    if (cb < 0) continue;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int32_t r = tile * RB + ty + 32 * i;
                        if (r < slots) {
                            const bool ok = r < len && sm.live[tile & 1][ty + 32 * i];
                            a.cand[cb + r] = ok ? (L2 ? -acc[i][j] : acc[i][j]) : -INFINITY;
                        }
                    }
                }
            }
        }
    }
}

template <int QT>
cudaError_t launch_lists_variant(const ScanArgs &a, const ListPlan &p, int num_sms, cudaStream_t st) {
    const size_t smem = sizeof(TileSmem<QT>);
    const int which = QT == 32 ? 0 : 1;
    cudaError_t e;
    if (a.metric == 1) {
        auto kern = scan_lists_kernel<QT, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<num_sms * 2, NT, smem, st>>>(a, p, which);
    } else {
        auto kern = scan_lists_kernel<QT, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<num_sms * 2, NT, smem, st>>>(a, p, which);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_scan_lists(const ScanArgs &a, const ListPlan &p, int num_sms, int *launches, cudaStream_t st) {
    if (a.npairs <= 0) return cudaSuccess;
    if (a.npairs > (int64_t)INT32_MAX) return cudaErrorInvalidValue;
    cudaError_t e;
    const unsigned pb = (unsigned)((a.npairs + 255) / 256), lb = (unsigned)((p.nlist + 255) / 256);
    if ((e = cudaMemsetAsync(p.cnt, 0, (size_t)p.nlist * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.cursor, 0, (size_t)p.nlist * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.counters, 0, 8, st)) != cudaSuccess) return e;
    count_list_pairs_kernel<<<pb, 256, 0, st>>>(a.probe, a.npairs, a.list_len, p.nlist, p.cnt);
    plan_items_kernel<<<lb, 256, 0, st>>>(p.cnt, p.nlist, p.n32, p.n8);
    if ((e = launch_exclusive_scan_i32(p.cnt, p.nlist, p.lq_off, st)) != cudaSuccess) return e;
    if ((e = launch_exclusive_scan_i32(p.n32, p.nlist, p.off32, st)) != cudaSuccess) return e;
    if ((e = launch_exclusive_scan_i32(p.n8, p.nlist, p.off8, st)) != cudaSuccess) return e;
    fill_list_pairs_kernel<<<pb, 256, 0, st>>>(a.probe, a.npairs, a.list_len, p.nlist, p.lq_off, p.cursor, p.lq);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_lists_variant<32>(a, p, num_sms, st)) != cudaSuccess) return e;
    if ((e = launch_lists_variant<8>(a, p, num_sms, st)) != cudaSuccess) return e;
    if (launches) *launches += 8;
    return cudaSuccess;
}

}  // namespace sc
