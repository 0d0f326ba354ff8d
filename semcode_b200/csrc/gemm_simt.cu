// K1/K2 (exact-fp32 version): coarse-centroid similarity as a tiled NT contraction on the FP32
// pipe.  S[m,n] = sum_k A[m,k] * B[n,k]; optional L2 epilogue S = 2*S - |b_n|^2.
//
// This is the parity-exact contraction used for the coarse quantizer (FAISS IndexFlat::search on
// the centroids, reached from reference src/semcode/storage/milvus_store.py:141-147) and for
// k-means/list assignment (quantizer->assign, milvus_store.py:128-130).  Roofline: FP32 FMA pipe
// (148 SMs x 128 lanes x 2 flop x clock); operands stream from L2.
#include <algorithm>

#include "common.cuh"

namespace sc {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int LDS = BM + 4;  // padded leading dim of the k-major smem tiles (keeps float4 alignment)

__device__ __forceinline__ float4 ldg_tile(const float *base, int64_t row, int64_t rows, int k, int K) {
    if (row < rows && k < K) return __ldg(reinterpret_cast<const float4 *>(base + row * (int64_t)K + k));
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256, 2)
gemm_nt_kernel(const float *__restrict__ A, int64_t M, const float *__restrict__ B, int N, int K,
               const float *__restrict__ bnorm, float *__restrict__ C) {
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int ty = tid >> 4, tx = tid & 15;

    // global->smem mapping: 2 float4 of A and of B per thread per k-tile
    const int lrow0 = tid >> 2, lkq = (tid & 3) * 4;  // rows lrow0 and lrow0+64, k offset lkq

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    const int nk = (K + BK - 1) / BK;

    auto load_tile = [&](int kt) {
        const int k = kt * BK + lkq;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ra[i] = ldg_tile(A, m0 + lrow0 + 64 * i, M, k, K);
            rb[i] = ldg_tile(B, (int64_t)n0 + lrow0 + 64 * i, N, k, K);
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow0 + 64 * i;
            As[buf][lkq + 0][r] = ra[i].x;
            As[buf][lkq + 1][r] = ra[i].y;
            As[buf][lkq + 2][r] = ra[i].z;
            As[buf][lkq + 3][r] = ra[i].w;
            Bs[buf][lkq + 0][r] = rb[i].x;
            Bs[buf][lkq + 1][r] = rb[i].y;
            Bs[buf][lkq + 2][r] = rb[i].z;
            Bs[buf][lkq + 3][r] = rb[i].w;
        }
    };

    load_tile(0);
    store_tile(0);
    __syncthreads();

    int cur = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tile(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }

    const bool vec_ok = (N & 3) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = n0 + h * 64 + tx * 4;
            if (col >= N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = acc[i][h * 4 + j];
                if (bnorm != nullptr && col + j < N) v[j] = 2.f * v[j] - __ldg(bnorm + col + j);
            }
            float *dst = C + row * (int64_t)N + col;
            if (vec_ok && col + 3 < N) {
                *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < N) dst[j] = v[j];
            }
        }
    }
}

// ---- coarse scores for tiny batches (nq <= 16): one warp per centroid, queries staged in shared memory --------
// At batch 1 the 128-row MMA tile is >99 % padding and the TMA/MMA pipeline costs ~45 us of latency; streaming the
// centroid table once (nlist * 4 * dim bytes, 50 MB at C2: ~8 us from HBM, less from L2) is the roofline here.
// Exact fp32 (FFMA), same epilogue as the contraction kernels: S = alpha * q.c - bias[c].
template <int NQ>
__global__ void __launch_bounds__(256) coarse_small_kernel(const float *__restrict__ q, int nq, const float *__restrict__ cent,
                                                           int nlist, int ds, float alpha, const float *__restrict__ bias,
                                                           float *__restrict__ scores) {
    extern __shared__ __align__(16) float4 qs4[];  // [nq][ds4]
    const int ds4 = ds >> 2;
    pdl_launch_dependents();  // the probe selection may take its place on an SM while the table streams
    for (int i = threadIdx.x; i < nq * ds4; i += blockDim.x) qs4[i] = __ldg(reinterpret_cast<const float4 *>(q) + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = blockIdx.x * 8 + warp; c < nlist; c += gridDim.x * 8) {
        const float4 *row = reinterpret_cast<const float4 *>(cent + (int64_t)c * ds);
        float acc[NQ];
#pragma unroll
        for (int j = 0; j < NQ; ++j) acc[j] = 0.f;
        for (int k = lane; k < ds4; k += 32) {
            const float4 x = ld_stream_f4(row + k);
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                if (j < nq) {
                    const float4 v = qs4[j * ds4 + k];
                    acc[j] = fmaf(x.x, v.x, acc[j]);
                    acc[j] = fmaf(x.y, v.y, acc[j]);
                    acc[j] = fmaf(x.z, v.z, acc[j]);
                    acc[j] = fmaf(x.w, v.w, acc[j]);
                }
            }
        }
        const float b = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            if (j < nq) {
                const float sum = warp_sum(acc[j]);
                if (lane == 0) scores[(int64_t)j * nlist + c] = alpha * sum - b;
            }
        }
    }
}

__global__ void row_norms_kernel(const float *__restrict__ x, int64_t rows, int ds, float *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float4 *p = reinterpret_cast<const float4 *>(x + r * (int64_t)ds);
    float s = 0.f;
    for (int c = lane; c < (ds >> 2); c += 32) {
        const float4 v = __ldg(p + c);
        s = fmaf(v.x, v.x, s);
        s = fmaf(v.y, v.y, s);
        s = fmaf(v.z, v.z, s);
        s = fmaf(v.w, v.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
}

__global__ void argmax_rows_kernel(const float *__restrict__ scores, int64_t M, int N, int32_t *__restrict__ out_idx,
                                   float *__restrict__ out_val) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    const float *p = scores + r * (int64_t)N;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < N; c += 32) {
        const float v = __ldg(p + c);
        if (v > bv || (v == bv && c < bi) || bi == 0x7fffffff) {
            bv = v;
            bi = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) {
            bv = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        out_idx[r] = bi == 0x7fffffff ? 0 : bi;
        if (out_val) out_val[r] = bv;
    }
}

__global__ void pad_rows_kernel(const float *__restrict__ in, int64_t n, int d, int ds, float *__restrict__ out) {
    const int64_t total = n * (int64_t)ds;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ds;
        const int c = (int)(i - r * ds);
        out[i] = c < d ? in[r * (int64_t)d + c] : 0.f;
    }
}

}  // namespace

cudaError_t launch_gemm_nt(const float *A, int64_t M, const float *B, int N, int K, const float *bnorm, float *C,
                           cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    gemm_nt_kernel<<<grid, 256, 0, st>>>(A, M, B, N, K, bnorm, C);
    return cudaGetLastError();
}

// returns cudaErrorNotSupported when the batch does not fit (caller uses the contraction kernels)
cudaError_t launch_coarse_small(const float *q, int64_t nq, const float *cent, int nlist, int ds, float alpha, const float *bias,
                                float *scores, int num_sms, cudaStream_t st) {
    if (nq < 1 || nq > 16) return cudaErrorNotSupported;
    const size_t smem = (size_t)nq * ds * 4;
    if (smem > 96 * 1024) return cudaErrorNotSupported;
    const int grid = std::min((nlist + 7) / 8, num_sms * 16);
    cudaError_t e;
#define SC_LAUNCH_SMALL(NQ)                                                                                            \
    {                                                                                                                  \
        auto kern = coarse_small_kernel<NQ>;                                                                           \
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e; \
        kern<<<grid, 256, smem, st>>>(q, (int)nq, cent, nlist, ds, alpha, bias, scores);                               \
    }
    if (nq == 1) SC_LAUNCH_SMALL(1) else if (nq <= 4) SC_LAUNCH_SMALL(4) else SC_LAUNCH_SMALL(16)
#undef SC_LAUNCH_SMALL
    return cudaGetLastError();
}

cudaError_t launch_row_norms(const float *x, int64_t rows, int ds, float *out, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    const int wpb = 8;
    row_norms_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(x, rows, ds, out);
    return cudaGetLastError();
}

cudaError_t launch_argmax_rows(const float *scores, int64_t M, int N, int32_t *out_idx, float *out_val,
                               cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    const int wpb = 8;
    argmax_rows_kernel<<<(unsigned)((M + wpb - 1) / wpb), wpb * 32, 0, st>>>(scores, M, N, out_idx, out_val);
    return cudaGetLastError();
}

cudaError_t launch_pad_rows(const float *in, int64_t n, int d, int ds, float *out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * (int64_t)ds;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    pad_rows_kernel<<<blocks, 256, 0, st>>>(in, n, d, ds, out);
    return cudaGetLastError();
}

}  // namespace sc
