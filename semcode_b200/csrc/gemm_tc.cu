// K1/K2 on the 5th-generation tensor cores: S[m,n] = alpha * sum_k A[m,k] B[n,k] - bias[n] as a
// TMA-fed, warp-specialised tcgen05 contraction with 3xTF32 splitting, fp32 accumulators in TMEM.
//
// Replaces the sgemm FAISS runs for the coarse quantizer (IndexFlat::search on the centroids,
// reached from reference src/semcode/storage/milvus_store.py:141-147) and for k-means / list
// assignment (quantizer->assign behind Collection.upsert, milvus_store.py:128-130).
//
// 3xTF32: every fp32 operand x is split once (split_tf32_kernel) into hi = tf32(x) and
// lo = tf32(x - hi); the product is accumulated as hi*hi + hi*lo + lo*hi (the lo*lo term is below
// 2^-22 relative), which keeps fp32-class accuracy on kind::tf32 MMAs.
//
// CTA = 6 warps, one CTA per SM (persistent):
//   warp 0  TMA producer: per k-block (32 floats = one 128-byte swizzled row segment) four bulk tensor
//           loads -- A_hi, A_lo [128 x 32], B_hi, B_lo [256 x 32] -- into a 2-stage ring (96 KB / stage)
//   warp 1  TMEM allocator + MMA issuer: 4 k-steps x 3 tcgen05.mma (128x256x8, kind::tf32) per stage;
//           tcgen05.commit releases the stage and, after the last k-block, publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld 32 columns at a time from one of two 256-column accumulators
//           (double buffered, so the epilogue of tile i overlaps the MMAs of tile i+1)
// Epilogues: SCORES writes S to global memory (coarse probe, small M); ARGMAX keeps a per-row running
// best over the CTA's whole sweep of N tiles (one thread owns one row: no cross-thread reduction) and
// never materialises the M x N matrix (k-means assignment, bulk add).
// Roofline: tensor pipe (kind::tf32 dense, 3 MMAs per logical product), operands stream from L2.
#include <cuda.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace sc {

namespace {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 2;
constexpr int A_BYTES = BM * BK * 4;                     // 16 KB
constexpr int B_BYTES = BN * BK * 4;                     // 32 KB
constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);     // hi + lo of both operands = 96 KB
constexpr int TMEM_COLS = 512;                           // two 256-column fp32 accumulators
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NTHREADS = 192;

struct TcArgs {
    int64_t M;
    int N, K;
    float alpha;
    const float *bias;  // [N] or nullptr
    float *C;           // SCORES: [M, N]
    float *best_val;    // ARGMAX: [M]
    int32_t *best_idx;  // ARGMAX: [M]
    int n_base;         // ARGMAX: added to the column index (slab offset)
    int merge;          // ARGMAX: 1 = compare with the values already in best_val/best_idx
    unsigned long long *packed;  // ARGMAX over 256x256 tiles: [M] (key << 32 | ~index), merged with atomicMax
    int m_tiles, n_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of the (converged) warp.  Unlike `lane == 0`, ptxas knows that exactly one thread runs the guarded region,
// so every tcgen05.mma / TMA instruction in it takes its operands with plain R2UR moves; under `lane == 0` each one is
// wrapped in an ELECT / R2UR.BROADCAST / branch loop (~100 cycles per MMA -- as long as a 128x256x8 tf32 MMA runs).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row atoms 1024 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

// kind::tf32, fp32 accumulate, A and B K-major, M x N instruction shape
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// iteration i of this CTA -> (m_tile, n_tile); false when the CTA has no more work
template <bool ARGMAX>
__device__ __forceinline__ bool tile_of(const TcArgs &a, int i, int &mt, int &nt) {
    if (ARGMAX) {  // m tiles are dealt to CTAs, each sweeps every n tile (running best stays in registers)
        mt = blockIdx.x + (i / a.n_tiles) * gridDim.x;
        nt = i % a.n_tiles;
        return mt < a.m_tiles;
    }
    const int64_t t = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;  // flat (m, n) order, n fastest
    if (t >= (int64_t)a.m_tiles * a.n_tiles) return false;
    mt = (int)(t / a.n_tiles);
    nt = (int)(t % a.n_tiles);
    return true;
}

// BN_: centroid rows per tile.  256 is the default; 128 (three stages of 64 KB) is used when 256-wide tiles would leave SMs
// idle -- a 128-query coarse pass (one rank's share of a 1024-query batch on 8 GPUs) is ONE row tile: 64 tiles of 256 on
// 148 SMs, each CTA pulling 2.3 MB through its own L2 -> SMEM path (42 us); 128 tiles of 128 pull 1.5 MB each.
template <bool ARGMAX, int BN_>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, const TcArgs a) {
    constexpr int B_BYTES_ = BN_ * BK * 4;
    constexpr int STAGE_BYTES_ = 2 * (A_BYTES + B_BYTES_);
    constexpr int STAGES_ = BN_ == 256 ? STAGES : 3;
    static_assert(BN_ == 256 || BN_ == 128, "tile widths: 256 or 128 centroid rows");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES_ * STAGE_BYTES_);
    // bars: [0,S) full  [S,2S) empty  [2S,2S+2) accumulator full  [2S+2,2S+4) accumulator empty
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES_ + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES_ + s); };
    auto tfull_bar = [&](int acc) { return bar0 + 8u * (2 * STAGES_ + acc); };
    auto tempty_bar = [&](int acc) { return bar0 + 8u * (2 * STAGES_ + 2 + acc); };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES_; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int acc = 0; acc < 2; ++acc) {
            mbar_init(tfull_bar(acc), 1);
            mbar_init(tempty_bar(acc), 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int kblocks = (a.K + BK - 1) / BK;

    if (warp == 0) {
        if (elect_one()) {  // ---- TMA producer ----
            int stage = 0;
            uint32_t phase = 0;
            int mt, nt;
            for (int i = 0; tile_of<ARGMAX>(a, i, mt, nt); ++i) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES_);
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES_);
                    tma_load_2d(sa, &map_ahi, full_bar(stage), kb * BK, mt * BM);
                    tma_load_2d(sa + A_BYTES, &map_alo, full_bar(stage), kb * BK, mt * BM);
                    tma_load_2d(sa + 2 * A_BYTES, &map_bhi, full_bar(stage), kb * BK, nt * BN_);
                    tma_load_2d(sa + 2 * A_BYTES + B_BYTES_, &map_blo, full_bar(stage), kb * BK, nt * BN_);
                    if (++stage == STAGES_) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {  // ---- MMA issuer ----
            constexpr uint32_t idesc = umma_idesc_tf32(BM, BN_);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            int mt, nt;
            for (int i = 0; tile_of<ARGMAX>(a, i, mt, nt); ++i) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN_);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES_);
                    const uint64_t d_ahi = umma_desc_sw128(sa), d_alo = umma_desc_sw128(sa + A_BYTES);
                    const uint64_t d_bhi = umma_desc_sw128(sa + 2 * A_BYTES), d_blo = umma_desc_sw128(sa + 2 * A_BYTES + B_BYTES_);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t off = (uint64_t)((ks * 8 * 4) >> 4);  // 32 bytes per k-step inside the swizzled row
                        umma_tf32(tmem_d, d_ahi + off, d_bhi + off, idesc, (kb | ks) != 0 ? 1u : 0u);
                        umma_tf32(tmem_d, d_ahi + off, d_blo + off, idesc, 1u);
                        umma_tf32(tmem_d, d_alo + off, d_bhi + off, idesc, 1u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
                    if (++stage == STAGES_) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(acc));  // accumulator complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter (warp % 4), one row per thread ----
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        float best = -INFINITY;
        int best_i = 0;
        int mt, nt;
        for (int i = 0; tile_of<ARGMAX>(a, i, mt, nt); ++i) {
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int64_t m = (int64_t)mt * BM + row;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN_);
            if (ARGMAX && nt == 0) {
                best = -INFINITY;
                best_i = 0;
                if (a.merge && m < a.M) {
                    best = a.best_val[m];
                    best_i = a.best_idx[m];
                }
            }
#pragma unroll 1
            for (int c = 0; c < BN_; c += 32) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)c, v);
                const int n0 = nt * BN_ + c;
                if (n0 >= a.N) continue;  // whole chunk is padding (loads above stay warp-uniform)
                if (ARGMAX) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = n0 + j;
                        if (n < a.N) {
                            float s = a.alpha * v[j];
                            if (a.bias) s -= __ldg(a.bias + n);
                            if (s > best) {  // strict: ties keep the lowest index (columns ascend)
                                best = s;
                                best_i = a.n_base + n;
                            }
                        }
                    }
                } else if (m < a.M) {
                    float *dst = a.C + m * (int64_t)a.N + n0;
                    const bool vec = ((a.N & 3) == 0) && (n0 + 32 <= a.N);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float s[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            s[t] = a.alpha * v[j + t];
                            if (a.bias && n0 + j + t < a.N) s[t] -= __ldg(a.bias + n0 + j + t);
                        }
                        if (vec) {
                            *reinterpret_cast<float4 *>(dst + j) = make_float4(s[0], s[1], s[2], s[3]);
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                if (n0 + j + t < a.N) dst[j + t] = s[t];
                        }
                    }
                }
            }
            if (ARGMAX && nt == a.n_tiles - 1 && m < a.M) {
                a.best_val[m] = best;
                a.best_idx[m] = best_i;
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

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- 256 x 256 tile variant of the fused argmax (bulk assignment / k-means) -------------------------------
// The 128 x 256 kernel above is L2->SMEM bound: hi+lo of A (128 rows) and B (256 rows) = 96 KB per 1536
// MMA cycles = 62 B/cycle/SM, above what L2 delivers, so the tensor pipe idles ~40 % of the time.  Here one
// CTA owns TWO 128-row halves that share every B tile: 64 KB per 1536 MMA cycles (-33 % traffic per MAC).
// To fit three stages the k-block shrinks to 16 floats (64-byte rows, 64-byte swizzle); the two 128x256
// accumulators fill all 512 TMEM columns, so the epilogue is not overlapped with the next tile's MMAs
// (~3 % of a tile at K = 768).
namespace big {
constexpr int BM2 = 256, BN2 = 256, BK2 = 16, STAGES2 = 3;
constexpr int A2_BYTES = BM2 * BK2 * 4;                    // 16 KB (both halves)
constexpr int B2_BYTES = BN2 * BK2 * 4;                    // 16 KB
constexpr int STAGE2_BYTES = 2 * (A2_BYTES + B2_BYTES);    // 64 KB
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;
}  // namespace big

// K-major operand tile, 64-byte swizzle: rows of 64 B, 8-row atoms 512 B apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(512u >> 4) << 32;  // stride byte offset
    d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;            // SWIZZLE_64B
    return d;
}

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc256_argmax_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                         const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, const TcArgs a) {
    using namespace big;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES2 * STAGE2_BYTES);
    // bars: [0,S) full  [S,2S) empty  [2S] accumulators full  [2S+1] accumulators empty
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES2 + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES2 + s); };
    const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES2), tempty_bar = bar0 + 8u * (2 * STAGES2 + 1);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);  // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int kblocks = (a.K + BK2 - 1) / BK2;
    // flat (row tile, centroid tile) order with the centroid tile fastest: the CTAs that run at the same time
    // share a handful of A row tiles (and the whole B slab), so both operands are served from L2 -- a CTA that
    // sweeps all centroid tiles of its own row tile re-streams 1.5 MB of A per tile from DRAM (measured:
    // 51.8 GB of DRAM reads per 400k x 8192 launch).  The per-row best is therefore merged across CTAs with
    // one 64-bit atomicMax per row per tile on a packed (order-preserving key, ~index) word.
    auto tile = [&](int i, int &mt, int &nt) {
        const int64_t t = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
        if (t >= (int64_t)a.m_tiles * a.n_tiles) return false;
        mt = (int)(t / a.n_tiles);
        nt = (int)(t % a.n_tiles);
        return true;
    };

    if (warp == 0) {
        if (elect_one()) {  // ---- TMA producer ----
            int stage = 0;
            uint32_t phase = 0;
            int mt, nt;
            for (int i = 0; tile(i, mt, nt); ++i) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
                    mbar_expect_tx(full_bar(stage), STAGE2_BYTES);
                    tma_load_2d(sa, &map_ahi, full_bar(stage), kb * BK2, mt * BM2);
                    tma_load_2d(sa + A2_BYTES, &map_alo, full_bar(stage), kb * BK2, mt * BM2);
                    tma_load_2d(sa + 2 * A2_BYTES, &map_bhi, full_bar(stage), kb * BK2, nt * BN2);
                    tma_load_2d(sa + 2 * A2_BYTES + B2_BYTES, &map_blo, full_bar(stage), kb * BK2, nt * BN2);
                    if (++stage == STAGES2) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {  // ---- MMA issuer ----
            constexpr uint32_t idesc = umma_idesc_tf32(128, BN2);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            int mt, nt;
            for (int i = 0; tile(i, mt, nt); ++i) {
                mbar_wait(tempty_bar, acc_phase ^ 1u);  // epilogue has drained both accumulators
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE2_BYTES);
                    const uint64_t d_bhi = umma_desc_sw64(sa + 2 * A2_BYTES), d_blo = umma_desc_sw64(sa + 2 * A2_BYTES + B2_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {  // the two 128-row halves share the B tile
                        const uint64_t d_ahi = umma_desc_sw64(sa + h * (A2_BYTES / 2)), d_alo = umma_desc_sw64(sa + A2_BYTES + h * (A2_BYTES / 2));
                        const uint32_t tmem_d = tmem_base + (uint32_t)(h * BN2);
#pragma unroll
                        for (int ks = 0; ks < BK2 / 8; ++ks) {
                            const uint64_t off = (uint64_t)((ks * 8 * 4) >> 4);
                            umma_tf32(tmem_d, d_ahi + off, d_bhi + off, idesc, (kb | ks) != 0 ? 1u : 0u);
                            umma_tf32(tmem_d, d_ahi + off, d_blo + off, idesc, 1u);
                            umma_tf32(tmem_d, d_alo + off, d_bhi + off, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == STAGES2) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar);
                acc_phase ^= 1u;
            }
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lane quarter (warp % 4); one row per thread in EACH half ----
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        uint32_t acc_phase = 0;
        int mt, nt;
        for (int i = 0; tile(i, mt, nt); ++i) {
            mbar_wait(tfull_bar, acc_phase);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t m = (int64_t)mt * BM2 + h * 128 + row;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * BN2);
                float best = -INFINITY;
                int best_i = 0x7fffffff;
#pragma unroll 1
                for (int c = 0; c < BN2; c += 32) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)c, v);
                    const int n0 = nt * BN2 + c;
                    if (n0 >= a.N) continue;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = n0 + j;
                        if (n < a.N) {
                            float sc = a.alpha * v[j];
                            if (a.bias) sc -= __ldg(a.bias + n);
                            if (sc > best) {  // strict: ties keep the lowest index (columns ascend)
                                best = sc;
                                best_i = a.n_base + n;
                            }
                        }
                    }
                }
                if (m < a.M && best_i != 0x7fffffff) {
                    const unsigned long long w = ((unsigned long long)f2key(best) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)best_i);
                    atomicMax(a.packed + m, w);  // larger key wins; equal keys -> lower index wins
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

__global__ void unpack_argmax_kernel(const unsigned long long *__restrict__ packed, int64_t M, float *__restrict__ best_val,
                                     int32_t *__restrict__ best_idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const unsigned long long w = packed[i];
    best_idx[i] = w == 0ull ? 0 : (int32_t)(0xffffffffu - (uint32_t)w);
    if (best_val) best_val[i] = w == 0ull ? -INFINITY : key2f((uint32_t)(w >> 32));
}

__global__ void split_tf32_kernel(const float4 *__restrict__ x, int64_t n4, float4 *__restrict__ hi, float4 *__restrict__ lo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        float4 h, l;
        const float in[4] = {v.x, v.y, v.z, v.w};
        float ho[4], lw[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t hb, lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[t]));
            ho[t] = __uint_as_float(hb);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(in[t] - ho[t]));
            lw[t] = __uint_as_float(lb);
        }
        h = make_float4(ho[0], ho[1], ho[2], ho[3]);
        l = make_float4(lw[0], lw[1], lw[2], lw[3]);
        hi[i] = h;
        lo[i] = l;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// row-major fp32 [rows, K] (row stride K floats) as a 2-D tiled map with boxes of [box_rows x 32 floats], 128B swizzle
bool make_map(CUtensorMap *map, const float *base, int64_t rows, int K, int box_rows, int box_k = BK) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, box_k * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool ARGMAX, int BN_>
cudaError_t launch_tc_bn(const float *ahi, const float *alo, int64_t M, const float *bhi, const float *blo, int N, int K,
                         TcArgs a, int num_sms, cudaStream_t st) {
    constexpr int stages = BN_ == 256 ? STAGES : 3;
    constexpr int smem_bytes = stages * 2 * (A_BYTES + BN_ * BK * 4) + 1024 /*align*/ + 256 /*barriers*/;
    CUtensorMap mah, mal, mbh, mbl;
    if (!make_map(&mah, ahi, M, K, BM) || !make_map(&mal, alo, M, K, BM) || !make_map(&mbh, bhi, N, K, BN_) ||
        !make_map(&mbl, blo, N, K, BN_))
        return cudaErrorInvalidValue;
    a.M = M;
    a.N = N;
    a.K = K;
    a.m_tiles = (int)((M + BM - 1) / BM);
    a.n_tiles = (N + BN_ - 1) / BN_;
    auto kern = gemm_tc_kernel<ARGMAX, BN_>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    const int64_t work = ARGMAX ? a.m_tiles : (int64_t)a.m_tiles * a.n_tiles;
    const int grid = (int)(work < num_sms ? work : num_sms);
    kern<<<grid, NTHREADS, smem_bytes, st>>>(mah, mal, mbh, mbl, a);
    return cudaGetLastError();
}

template <bool ARGMAX>
cudaError_t launch_tc(const float *ahi, const float *alo, int64_t M, const float *bhi, const float *blo, int N, int K,
                      TcArgs a, int num_sms, cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    // score tiles of 128 centroid rows when tiles of 256 cannot occupy every SM (small batches)
    const int64_t tiles256 = ((M + BM - 1) / BM) * (int64_t)((N + BN - 1) / BN);
    static const bool bn128 = [] {  // SEMCODE_COARSE_BN128=0 keeps the 256-wide tiles (A/B measurements)
        const char *e = getenv("SEMCODE_COARSE_BN128");
        return !(e && e[0] == '0');
    }();
    if (!ARGMAX && tiles256 < num_sms && bn128)
        return launch_tc_bn<ARGMAX, 128>(ahi, alo, M, bhi, blo, N, K, a, num_sms, st);
    return launch_tc_bn<ARGMAX, 256>(ahi, alo, M, bhi, blo, N, K, a, num_sms, st);
}

}  // namespace

cudaError_t launch_split_tf32(const float *x, int64_t n, float *hi, float *lo, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t n4 = n / 4;  // callers pass ds-padded rows: n is a multiple of 4
    const int64_t want = (n4 + 255) / 256;
    split_tf32_kernel<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const float4 *>(x), n4, reinterpret_cast<float4 *>(hi), reinterpret_cast<float4 *>(lo));
    return cudaGetLastError();
}

cudaError_t launch_gemm_tc_scores(const float *ahi, const float *alo, int64_t M, const float *bhi, const float *blo, int N,
                                  int K, float alpha, const float *bias, float *C, int num_sms, cudaStream_t st) {
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.alpha = alpha;
    a.bias = bias;
    a.C = C;
    return launch_tc<false>(ahi, alo, M, bhi, blo, N, K, a, num_sms, st);
}

cudaError_t launch_gemm_tc_argmax(const float *ahi, const float *alo, int64_t M, const float *bhi, const float *blo, int N,
                                  int K, float alpha, const float *bias, float *best_val, int32_t *best_idx, int n_base,
                                  int merge, int num_sms, unsigned long long *packed, cudaStream_t st) {
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.alpha = alpha;
    a.bias = bias;
    a.best_val = best_val;
    a.best_idx = best_idx;
    a.n_base = n_base;
    a.merge = merge;
    a.packed = packed;
    if (packed == nullptr || M <= 0 || N <= 0) return launch_tc<true>(ahi, alo, M, bhi, blo, N, K, a, num_sms, st);
    // 256 x 256 tiles, 64-byte swizzle
    using namespace big;
    CUtensorMap mah, mal, mbh, mbl;
    if (!make_map(&mah, ahi, M, K, BM2, BK2) || !make_map(&mal, alo, M, K, BM2, BK2) || !make_map(&mbh, bhi, N, K, BN2, BK2) ||
        !make_map(&mbl, blo, N, K, BN2, BK2))
        return cudaErrorInvalidValue;
    a.M = M;
    a.N = N;
    a.K = K;
    a.m_tiles = (int)((M + BM2 - 1) / BM2);
    a.n_tiles = (N + BN2 - 1) / BN2;
    cudaError_t e = cudaFuncSetAttribute(gemm_tc256_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e != cudaSuccess) return e;
    const int64_t work = (int64_t)a.m_tiles * a.n_tiles;
    const int grid = (int)(work < num_sms ? work : num_sms);
    gemm_tc256_argmax_kernel<<<grid, NTHREADS, SMEM2_BYTES, st>>>(mah, mal, mbh, mbl, a);
    return cudaGetLastError();
}

cudaError_t launch_unpack_argmax(const unsigned long long *packed, int64_t M, float *best_val, int32_t *best_idx, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    unpack_argmax_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(packed, M, best_val, best_idx);
    return cudaGetLastError();
}

}  // namespace sc
