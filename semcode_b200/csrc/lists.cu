// K4: inverted-list maintenance on paged storage.
//
// Replaces FAISS IndexIVFFlat::add_core (append raw vector + id to the list of its best centroid)
// behind Collection.upsert at reference src/semcode/storage/milvus_store.py:128-130, and the
// delete half of upsert-by-primary-key (tombstones).
//
// Lists are chains of fixed 32-row pages (see common.cuh).  An insert batch
//   1. claims a slot per row:      pos = atomicAdd(list_len[list], 1)
//   2. sizes the new pages:        need[l] = ceil(new/32) - ceil(old/32)  -> exclusive scans
//   3. rebuilds the CSR page table (old page ids kept, new ones appended from the pool top)
//   4. scatters rows, ids and tags into their slots (one warp per row, 128-bit copies).
// No existing row ever moves, so inserts cost O(batch) HBM traffic, not O(index).
#include "common.cuh"

namespace sc {

namespace {

// bad[0] += rows with a list id outside [0, nlist); bad[1] += rows whose repo tag does not fit the 23-bit field
// (make_tag would mask it and alias another repo)
__global__ void count_positions_kernel(const int32_t *__restrict__ assign, const uint32_t *__restrict__ repo, int64_t n,
                                       int32_t nlist, int32_t *__restrict__ list_len, int32_t *__restrict__ pos,
                                       int32_t *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t l = assign[i];
    if (repo != nullptr && repo[i] > kTagRepoMax) atomicAdd(bad + 1, 1);
    if (l < 0 || l >= nlist) {
        pos[i] = -1;
        atomicAdd(bad, 1);
        return;
    }
    pos[i] = atomicAdd(list_len + l, 1);
}

__global__ void page_need_kernel(const int32_t *__restrict__ len_old, const int32_t *__restrict__ len_new,
                                 int32_t nlist, int32_t *__restrict__ need, int32_t *__restrict__ npg_new) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    const int32_t po = (len_old[l] + kPageRows - 1) / kPageRows;
    const int32_t pn = (len_new[l] + kPageRows - 1) / kPageRows;
    need[l] = pn - po;
    npg_new[l] = pn;
}

// new page number m (in list order) of this insert batch -> page id: the first nfree come from the free list
// (pages returned by compaction), the rest from the top of the pool
__global__ void rebuild_pt_kernel(const int32_t *__restrict__ pt_off_old, const int32_t *__restrict__ pt_old,
                                  const int32_t *__restrict__ pt_off_new, int32_t *__restrict__ pt_new,
                                  const int32_t *__restrict__ need_off, int32_t pool_top, const int32_t *__restrict__ free_pages,
                                  int32_t nfree, int32_t nlist) {
    // one warp per list
    const int32_t l = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (l >= nlist) return;
    const int32_t ob = pt_off_old[l], had = pt_off_old[l + 1] - ob;
    const int32_t nb = pt_off_new[l], now = pt_off_new[l + 1] - nb;
    const int32_t first_new = need_off[l];
    for (int32_t j = lane; j < now; j += 32) {
        int32_t page;
        if (j < had) {
            page = pt_old[ob + j];
        } else {
            const int32_t m = first_new + (j - had);
            page = m < nfree ? free_pages[m] : pool_top + (m - nfree);
        }
        pt_new[nb + j] = page;
    }
}

// ---- compaction: drop the tombstoned slots of every list in place ---------------------------------------------------
// One CTA per list (grid-stride).  Rows are taken in batches of one row per warp: every warp reads its live row
// (a slice of 512 floats at a time), the CTA synchronises, then the rows are written to their new, lower slots.  A
// destination never lies beyond its source and never inside a later batch, so no live row is overwritten before it
// was read.  new_len[l] = live rows; the slots [new_len, old_len) become unused (tag 0xFFFFFFFF).
constexpr int CW = 16;  // warps per CTA = rows per batch
__global__ void __launch_bounds__(CW * 32) compact_lists_kernel(int32_t nlist, int32_t *__restrict__ list_len,
                                                                const int32_t *__restrict__ pt_off, const int32_t *__restrict__ pt,
                                                                const SlabTable *__restrict__ slabs, int slab_shift, int ds) {
    __shared__ int32_t s_dst[CW];
    __shared__ int32_t s_w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slab_mask = (1 << slab_shift) - 1;
    const int ds4 = ds >> 2;
    for (int32_t l = blockIdx.x; l < nlist; l += gridDim.x) {
        const int32_t len = list_len[l];
        const int32_t ptb = pt_off[l];
        __syncthreads();
        if (threadIdx.x == 0) s_w = 0;
        __syncthreads();
        auto slot_of = [&](int32_t r, int &slab) -> int64_t {
            const int32_t page = pt[ptb + (r >> 5)];
            slab = page >> slab_shift;
            return (int64_t)(page & slab_mask) * kPageRows + (r & 31);
        };
        for (int32_t j0 = 0; j0 < len; j0 += CW) {
            const int32_t j = j0 + warp;
            int slab = 0;
            int64_t slot = 0;
            bool live = false;
            uint32_t tag = 0;
            int64_t id = 0;
            if (j < len) {
                slot = slot_of(j, slab);
                tag = slabs->tags[slab][slot];
                live = (tag & kTagRemoved) == 0;
                if (live) id = slabs->ids[slab][slot];
            }
            if (lane == 0) s_dst[warp] = live ? 1 : 0;
            __syncthreads();
            int32_t before = 0, batch_live = 0;
#pragma unroll
            for (int w = 0; w < CW; ++w) {
                const int32_t f = s_dst[w];
                if (w < warp) before += f;
                batch_live += f;
            }
            const int32_t w0 = s_w;
            const int32_t dst = w0 + before;
            const bool move = live && dst != j;
            int dslab = 0;
            int64_t dslot = 0;
            if (move) dslot = slot_of(dst, dslab);
            for (int c0 = 0; c0 < ds4; c0 += 128) {  // slices of 128 float4: 4 per lane
                float4 v[4];
                if (move) {
                    const float4 *src = reinterpret_cast<const float4 *>(slabs->vec[slab] + slot * ds);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int c = c0 + t * 32 + lane;
                        if (c < ds4) v[t] = src[c];
                    }
                }
                __syncthreads();  // every row of the batch has been read for this slice
                if (move) {
                    float4 *dp = reinterpret_cast<float4 *>(slabs->vec[dslab] + dslot * ds);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int c = c0 + t * 32 + lane;
                        if (c < ds4) dp[c] = v[t];
                    }
                }
                __syncthreads();
            }
            if (move && lane == 0) {  // ids / tags were read above, before any write of this batch
                slabs->ids[dslab][dslot] = id;
                slabs->tags[dslab][dslot] = tag;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_w = w0 + batch_live;
            __syncthreads();
        }
        const int32_t nl = s_w;
        for (int32_t r = nl + threadIdx.x; r < len; r += blockDim.x) {
            int slab;
            const int64_t slot = slot_of(r, slab);
            slabs->tags[slab][slot] = 0xFFFFFFFFu;
        }
        if (threadIdx.x == 0) list_len[l] = nl;
    }
}

// after compact_lists_kernel: keep the first ceil(len / 32) pages of every list, hand the others to the free list
__global__ void compact_pt_kernel(const int32_t *__restrict__ pt_off_old, const int32_t *__restrict__ pt_old,
                                  const int32_t *__restrict__ pt_off_new, int32_t *__restrict__ pt_new, int32_t nlist,
                                  int32_t *__restrict__ free_pages, int32_t *__restrict__ free_cursor) {
    const int32_t l = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (l >= nlist) return;
    const int32_t ob = pt_off_old[l], had = pt_off_old[l + 1] - ob;
    const int32_t nb = pt_off_new[l], keep = pt_off_new[l + 1] - nb;
    for (int32_t j = lane; j < keep; j += 32) pt_new[nb + j] = pt_old[ob + j];
    if (had > keep) {
        int32_t base = 0;
        if (lane == 0) base = atomicAdd(free_cursor, had - keep);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int32_t j = keep + lane; j < had; j += 32) free_pages[base + (j - keep)] = pt_old[ob + j];
    }
}

__global__ void pages_of_len_kernel(const int32_t *__restrict__ len, int32_t nlist, int32_t *__restrict__ npg) {
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nlist) npg[l] = (len[l] + kPageRows - 1) / kPageRows;
}

// lists [l0, l1) back to back: row r of the range belongs to the list whose slice of `off` (exclusive prefix of the
// slot counts, off[0] = 0, l1 - l0 + 1 entries) holds it
__global__ void export_range_kernel(const int32_t *__restrict__ pt_off, const int32_t *__restrict__ pt, int32_t l0, int32_t nl,
                                    const int64_t *__restrict__ off, int ds, int d_out, const SlabTable *__restrict__ slabs,
                                    int slab_shift, float *__restrict__ vecs, int64_t *__restrict__ ids,
                                    uint32_t *__restrict__ tags) {
    const int64_t total = off[nl];
    const int lane = threadIdx.x & 31;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < total; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        int32_t lo = 0, hi = nl;
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (off[mid] <= r)
                lo = mid;
            else
                hi = mid;
        }
        const int32_t j = (int32_t)(r - off[lo]);
        const int32_t page = pt[pt_off[l0 + lo] + (j >> 5)];
        const int slab = page >> slab_shift;
        const int64_t slot = (int64_t)(page & ((1 << slab_shift) - 1)) * kPageRows + (j & 31);
        if (vecs) {
            const float *src = slabs->vec[slab] + slot * ds;
            for (int c = lane; c < d_out; c += 32) vecs[r * (int64_t)d_out + c] = src[c];
        }
        if (lane == 0) {
            if (ids) ids[r] = slabs->ids[slab][slot];
            if (tags) tags[r] = slabs->tags[slab][slot];
        }
    }
}

__global__ void scatter_rows_kernel(const float *__restrict__ x, const int64_t *__restrict__ ids,
                                    const uint32_t *__restrict__ repo, const uint8_t *__restrict__ lang,
                                    const int32_t *__restrict__ assign, const int32_t *__restrict__ pos, int64_t n,
                                    int ds, const int32_t *__restrict__ pt_off, const int32_t *__restrict__ pt,
                                    const SlabTable *__restrict__ slabs, int slab_shift) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int32_t p = pos[i];
    if (p < 0) return;
    const int32_t l = assign[i];
    const int32_t page = pt[pt_off[l] + p / kPageRows];
    const int slab = page >> slab_shift;
    const int64_t slot = (int64_t)(page & ((1 << slab_shift) - 1)) * kPageRows + (p % kPageRows);
    const float4 *src = reinterpret_cast<const float4 *>(x + i * (int64_t)ds);
    float4 *dst = reinterpret_cast<float4 *>(slabs->vec[slab] + slot * ds);
    for (int c = lane; c < (ds >> 2); c += 32) dst[c] = __ldg(src + c);
    if (lane == 0) {
        slabs->ids[slab][slot] = ids[i];
        slabs->tags[slab][slot] = make_tag(repo ? repo[i] : 0u, lang ? (uint32_t)lang[i] : 0u);
    }
}

__global__ void remove_ids_kernel(const int64_t *__restrict__ sorted_ids, int64_t nrm,
                                  const SlabTable *__restrict__ slabs, int slab_shift, int64_t npages,
                                  unsigned long long *__restrict__ count) {
    const int64_t total = npages * kPageRows;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t page = g / kPageRows;
        const int slab = (int)(page >> slab_shift);
        const int64_t slot = (page & ((1 << slab_shift) - 1)) * kPageRows + (g % kPageRows);
        const uint32_t tag = slabs->tags[slab][slot];
        if (tag & kTagRemoved) continue;
        const int64_t id = slabs->ids[slab][slot];
        int64_t lo = 0, hi = nrm;  // lower_bound
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (sorted_ids[mid] < id)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo < nrm && sorted_ids[lo] == id) {
            slabs->tags[slab][slot] = tag | kTagRemoved;
            atomicAdd(count, 1ull);
        }
    }
}

__global__ void export_list_kernel(const int32_t *__restrict__ pt, int32_t pt_begin, int32_t len, int ds, int d_out,
                                   const SlabTable *__restrict__ slabs, int slab_shift, float *__restrict__ vecs,
                                   int64_t *__restrict__ ids, uint32_t *__restrict__ tags) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= len) return;
    const int32_t page = pt[pt_begin + (int32_t)(r / kPageRows)];
    const int slab = page >> slab_shift;
    const int64_t slot = (int64_t)(page & ((1 << slab_shift) - 1)) * kPageRows + (r % kPageRows);
    if (vecs) {
        const float *src = slabs->vec[slab] + slot * ds;
        for (int c = lane; c < d_out; c += 32) vecs[r * (int64_t)d_out + c] = src[c];
    }
    if (lane == 0) {
        if (ids) ids[r] = slabs->ids[slab][slot];
        if (tags) tags[r] = slabs->tags[slab][slot];
    }
}

}  // namespace

cudaError_t launch_count_positions(const int32_t *assign, const uint32_t *repo, int64_t n, int32_t nlist, int32_t *list_len,
                                   int32_t *pos, int32_t *bad, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    count_positions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(assign, repo, n, nlist, list_len, pos, bad);
    return cudaGetLastError();
}

cudaError_t launch_page_need(const int32_t *len_old, const int32_t *len_new, int32_t nlist, int32_t *need,
                             int32_t *npg_new, cudaStream_t st) {
    page_need_kernel<<<(nlist + 255) / 256, 256, 0, st>>>(len_old, len_new, nlist, need, npg_new);
    return cudaGetLastError();
}

cudaError_t launch_rebuild_pt(const int32_t *pt_off_old, const int32_t *pt_old, const int32_t *pt_off_new,
                              int32_t *pt_new, const int32_t *need_off, int32_t pool_top, const int32_t *free_pages,
                              int32_t nfree, int32_t nlist, cudaStream_t st) {
    const int64_t threads = (int64_t)nlist * 32;
    rebuild_pt_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(pt_off_old, pt_old, pt_off_new, pt_new,
                                                                          need_off, pool_top, free_pages, nfree, nlist);
    return cudaGetLastError();
}

cudaError_t launch_compact_lists(int32_t nlist, int32_t *list_len, const int32_t *pt_off, const int32_t *pt,
                                 const SlabTable *slabs, int slab_shift, int ds, int num_sms, cudaStream_t st) {
    const int grid = nlist < num_sms * 4 ? nlist : num_sms * 4;
    compact_lists_kernel<<<grid, CW * 32, 0, st>>>(nlist, list_len, pt_off, pt, slabs, slab_shift, ds);
    return cudaGetLastError();
}

cudaError_t launch_pages_of_len(const int32_t *len, int32_t nlist, int32_t *npg, cudaStream_t st) {
    pages_of_len_kernel<<<(nlist + 255) / 256, 256, 0, st>>>(len, nlist, npg);
    return cudaGetLastError();
}

cudaError_t launch_compact_pt(const int32_t *pt_off_old, const int32_t *pt_old, const int32_t *pt_off_new, int32_t *pt_new,
                              int32_t nlist, int32_t *free_pages, int32_t *free_cursor, cudaStream_t st) {
    const int64_t threads = (int64_t)nlist * 32;
    compact_pt_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(pt_off_old, pt_old, pt_off_new, pt_new, nlist, free_pages,
                                                                          free_cursor);
    return cudaGetLastError();
}

cudaError_t launch_export_range(const int32_t *pt_off, const int32_t *pt, int32_t l0, int32_t nl, const int64_t *off, int64_t rows,
                                int ds, int d_out, const SlabTable *slabs, int slab_shift, float *vecs, int64_t *ids, uint32_t *tags,
                                int num_sms, cudaStream_t st) {
    if (rows <= 0 || nl <= 0) return cudaSuccess;
    const int64_t want = (rows * 32 + 255) / 256;
    const unsigned blocks = (unsigned)(want < (int64_t)num_sms * 16 ? want : (int64_t)num_sms * 16);
    export_range_kernel<<<blocks, 256, 0, st>>>(pt_off, pt, l0, nl, off, ds, d_out, slabs, slab_shift, vecs, ids, tags);
    return cudaGetLastError();
}

cudaError_t launch_scatter_rows(const float *x, const int64_t *ids, const uint32_t *repo, const uint8_t *lang,
                                const int32_t *assign, const int32_t *pos, int64_t n, int ds, const int32_t *pt_off,
                                const int32_t *pt, const SlabTable *slabs, int slab_shift, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t threads = n * 32;
    scatter_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, ids, repo, lang, assign, pos, n, ds,
                                                                            pt_off, pt, slabs, slab_shift);
    return cudaGetLastError();
}

cudaError_t launch_remove_ids(const int64_t *sorted_ids, int64_t nrm, const SlabTable *slabs, int slab_shift,
                              int64_t npages, unsigned long long *count, cudaStream_t st) {
    if (nrm <= 0 || npages <= 0) return cudaSuccess;
    const int64_t total = npages * kPageRows;
    const int64_t want = (total + 255) / 256;
    const unsigned blocks = (unsigned)(want < 148 * 32 ? want : 148 * 32);
    remove_ids_kernel<<<blocks, 256, 0, st>>>(sorted_ids, nrm, slabs, slab_shift, npages, count);
    return cudaGetLastError();
}

cudaError_t launch_export_list(const int32_t *pt, int32_t pt_begin, int32_t len, int ds, int d_out,
                               const SlabTable *slabs, int slab_shift, float *vecs, int64_t *ids, uint32_t *tags,
                               cudaStream_t st) {
    if (len <= 0) return cudaSuccess;
    const int64_t threads = (int64_t)len * 32;
    export_list_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(pt, pt_begin, len, ds, d_out, slabs,
                                                                           slab_shift, vecs, ids, tags);
    return cudaGetLastError();
}

}  // namespace sc
