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

__global__ void count_positions_kernel(const int32_t *__restrict__ assign, int64_t n, int32_t nlist,
                                       int32_t *__restrict__ list_len, int32_t *__restrict__ pos,
                                       int32_t *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t l = assign[i];
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

__global__ void rebuild_pt_kernel(const int32_t *__restrict__ pt_off_old, const int32_t *__restrict__ pt_old,
                                  const int32_t *__restrict__ pt_off_new, int32_t *__restrict__ pt_new,
                                  const int32_t *__restrict__ need_off, int32_t pool_top, int32_t nlist) {
    // one warp per list
    const int32_t l = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (l >= nlist) return;
    const int32_t ob = pt_off_old[l], had = pt_off_old[l + 1] - ob;
    const int32_t nb = pt_off_new[l], now = pt_off_new[l + 1] - nb;
    const int32_t first_new = pool_top + need_off[l];
    for (int32_t j = lane; j < now; j += 32) pt_new[nb + j] = j < had ? pt_old[ob + j] : first_new + (j - had);
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

cudaError_t launch_count_positions(const int32_t *assign, int64_t n, int32_t nlist, int32_t *list_len, int32_t *pos,
                                   int32_t *bad, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    count_positions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(assign, n, nlist, list_len, pos, bad);
    return cudaGetLastError();
}

cudaError_t launch_page_need(const int32_t *len_old, const int32_t *len_new, int32_t nlist, int32_t *need,
                             int32_t *npg_new, cudaStream_t st) {
    page_need_kernel<<<(nlist + 255) / 256, 256, 0, st>>>(len_old, len_new, nlist, need, npg_new);
    return cudaGetLastError();
}

cudaError_t launch_rebuild_pt(const int32_t *pt_off_old, const int32_t *pt_old, const int32_t *pt_off_new,
                              int32_t *pt_new, const int32_t *need_off, int32_t pool_top, int32_t nlist,
                              cudaStream_t st) {
    const int64_t threads = (int64_t)nlist * 32;
    rebuild_pt_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(pt_off_old, pt_old, pt_off_new, pt_new,
                                                                          need_off, pool_top, nlist);
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
