// K3: k-means centroid update (FAISS Clustering::train -> compute_centroids + split_clusters,
// which Milvus runs server-side when it builds the IVF_FLAT index declared at reference
// src/semcode/storage/milvus_store.py:76-83).
//
// The assign step is K2 (gemm + argmax).  The update reads X once (HBM-bound): one warp per row
// adds the row into its centroid's fp64 accumulator with 64-bit global reductions (red.add.f64),
// which keeps the sums order-insensitive to ~1e-16 so the fp32 means are reproducible; counts are
// integer atomics.  The objective (sum of best similarity / squared distance) is reduced per CTA.
#include "common.cuh"

namespace sc {

namespace {

__global__ void __launch_bounds__(256)
kmeans_accumulate_kernel(const float *__restrict__ x, int64_t n, int ds, const int32_t *__restrict__ assign,
                         const float *__restrict__ best, int metric, double *__restrict__ sums,
                         int32_t *__restrict__ counts, double *__restrict__ objective) {
    __shared__ double obj_s[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 8 + warp;
    double obj = 0.0;
    if (i < n) {
        const int32_t l = assign[i];
        const float4 *src = reinterpret_cast<const float4 *>(x + i * (int64_t)ds);
        double *dst = sums + (int64_t)l * ds;
        float nrm = 0.f;
        for (int c = lane; c < (ds >> 2); c += 32) {
            const float4 v = __ldg(src + c);
            atomicAdd(dst + 4 * c + 0, (double)v.x);
            atomicAdd(dst + 4 * c + 1, (double)v.y);
            atomicAdd(dst + 4 * c + 2, (double)v.z);
            atomicAdd(dst + 4 * c + 3, (double)v.w);
            nrm += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        if (metric == 1) nrm = warp_sum(nrm);
        if (lane == 0) {
            atomicAdd(counts + l, 1);
            // best = similarity to maximise: IP -> x.c ; L2 -> 2 x.c - |c|^2, so |x-c|^2 = |x|^2 - best
            obj = metric == 0 ? (double)best[i] : (double)nrm - (double)best[i];
        }
    }
    if (lane == 0) obj_s[warp] = obj;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += obj_s[w];
        atomicAdd(objective, t);
    }
}

__global__ void kmeans_finalize_kernel(const double *__restrict__ sums, const int32_t *__restrict__ counts,
                                       int32_t nlist, int ds, float *__restrict__ centroids) {
    const int64_t total = (int64_t)nlist * ds;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t l = (int32_t)(i / ds);
        const int32_t c = counts[l];
        if (c > 0) centroids[i] = (float)(sums[i] / (double)c);
    }
}

// FAISS split_clusters: ci <- cj*(1 +- eps), cj <- cj*(1 -+ eps), eps = 1/1024, sign alternating per dim
__global__ void split_centroid_kernel(float *__restrict__ centroids, int ds, int32_t ci, int32_t cj) {
    const float eps = 1.0f / 1024.0f;
    for (int t = threadIdx.x; t < ds; t += blockDim.x) {
        const float b = centroids[(int64_t)cj * ds + t];
        const float up = (t & 1) == 0 ? 1.f + eps : 1.f - eps;
        const float dn = (t & 1) == 0 ? 1.f - eps : 1.f + eps;
        centroids[(int64_t)ci * ds + t] = b * up;
        centroids[(int64_t)cj * ds + t] = b * dn;
    }
}

__global__ void gather_rows_kernel(const float *__restrict__ x, const int64_t *__restrict__ rows, int64_t n, int ds,
                                   float *__restrict__ out) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const float4 *src = reinterpret_cast<const float4 *>(x + rows[i] * (int64_t)ds);
    float4 *dst = reinterpret_cast<float4 *>(out + i * (int64_t)ds);
    for (int c = lane; c < (ds >> 2); c += 32) dst[c] = __ldg(src + c);
}

}  // namespace

cudaError_t launch_kmeans_accumulate(const float *x, int64_t n, int ds, const int32_t *assign, const float *best,
                                     int metric, double *sums, int32_t *counts, double *objective, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    kmeans_accumulate_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, n, ds, assign, best, metric, sums, counts,
                                                                       objective);
    return cudaGetLastError();
}

cudaError_t launch_kmeans_finalize(const double *sums, const int32_t *counts, int32_t nlist, int ds, float *centroids,
                                   cudaStream_t st) {
    const int64_t total = (int64_t)nlist * ds;
    const int64_t want = (total + 255) / 256;
    kmeans_finalize_kernel<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, st>>>(sums, counts, nlist, ds,
                                                                                            centroids);
    return cudaGetLastError();
}

cudaError_t launch_split_centroid(float *centroids, int ds, int32_t ci, int32_t cj, cudaStream_t st) {
    split_centroid_kernel<<<1, 256, 0, st>>>(centroids, ds, ci, cj);
    return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float *x, const int64_t *rows, int64_t n, int ds, float *out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t threads = n * 32;
    gather_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, rows, n, ds, out);
    return cudaGetLastError();
}

}  // namespace sc
