/*
 * C restatement of FAISS IndexIVFFlat search / assign (CPU).  TEST INFRASTRUCTURE ONLY.
 * PARITY UNPINNED -- the reference holds no golden vectors for this path and its engine
 * (Milvus v2.4.4 / knowhere / FAISS, reached from src/semcode/storage/milvus_store.py:128-147)
 * is not in /root/reference; see oracle/__init__.py.
 *
 * What follows which published FAISS routine [EXT]:
 *   orc_dot / orc_l2sqr ........ fvec_inner_product / fvec_L2sqr (exact fp32, direct form)
 *   orc_top_probes ............. IndexFlat::search on the centroids, keep the nprobe best
 *   orc_scan_search ............ IndexIVFFlat::search_preassigned -> IVFFlatScanner::scan_codes
 *                                with a per-query binary heap (HeapResultHandler); rows whose
 *                                skip[] byte is set are dropped before ranking (IDSelector /
 *                                knowhere BitsetView); IP descending, squared-L2 ascending;
 *                                missing results id -1
 *   orc_gemm_nt ................ the sgemm FAISS uses for the coarse pass (blocked, OpenMP)
 *   orc_assign ................. quantizer->assign = argbest over centroids
 * OpenMP runs over queries, as FAISS parallel_mode 0 does.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp, x86-64-v3 baseline with AVX-512 clones).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__x86_64__) && defined(__GNUC__)
#define ORC_CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define ORC_CLONES
#endif

#define ORC_IP 0
#define ORC_L2 1

ORC_CLONES
float orc_dot(const float *a, const float *b, int d) {
    float acc[16] = {0};
    int i = 0;
    for (; i + 16 <= d; i += 16)
        for (int j = 0; j < 16; ++j) acc[j] += a[i + j] * b[i + j];
    float s = 0.f;
    for (int j = 0; j < 16; ++j) s += acc[j];
    for (; i < d; ++i) s += a[i] * b[i];
    return s;
}

ORC_CLONES
float orc_l2sqr(const float *a, const float *b, int d) {
    float acc[16] = {0};
    int i = 0;
    for (; i + 16 <= d; i += 16)
        for (int j = 0; j < 16; ++j) {
            float t = a[i + j] - b[i + j];
            acc[j] += t * t;
        }
    float s = 0.f;
    for (int j = 0; j < 16; ++j) s += acc[j];
    for (; i < d; ++i) {
        float t = a[i] - b[i];
        s += t * t;
    }
    return s;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU arm asks for the cores it may run on instead */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * heap keyed on "worse-ness": root = current worst kept result.  better(a,b): a ranks before b.
 * score is the similarity to MAXIMISE (IP: dot, L2: -dist^2 is NOT used; we keep the raw value
 * and flip the comparison instead so the returned distances are the raw ones).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float v;
    int64_t id;
} orc_hit;

static inline int hit_before(int metric, orc_hit a, orc_hit b) { /* a ranks strictly before b */
    if (a.v != b.v) return metric == ORC_IP ? (a.v > b.v) : (a.v < b.v);
    return a.id < b.id;
}

static void heap_sift_down(orc_hit *h, int n, int i, int metric) {
    for (;;) { /* root = worst: parent must rank AFTER children */
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && hit_before(metric, h[w], h[l])) w = l;
        if (r < n && hit_before(metric, h[w], h[r])) w = r;
        if (w == i) return;
        orc_hit t = h[i];
        h[i] = h[w];
        h[w] = t;
        i = w;
    }
}

static void heap_push(orc_hit *h, int *n, int k, orc_hit x, int metric) {
    if (*n < k) {
        int i = (*n)++;
        h[i] = x;
        while (i > 0) {
            int p = (i - 1) / 2;
            if (hit_before(metric, h[p], h[i])) { /* parent better than child -> swap up */
                orc_hit t = h[p];
                h[p] = h[i];
                h[i] = t;
                i = p;
            } else
                break;
        }
    } else if (hit_before(metric, x, h[0])) {
        h[0] = x;
        heap_sift_down(h, k, 0, metric);
    }
}

static int cmp_metric_g;
#pragma omp threadprivate(cmp_metric_g)
static int hit_cmp(const void *a, const void *b) {
    orc_hit x = *(const orc_hit *)a, y = *(const orc_hit *)b;
    if (hit_before(cmp_metric_g, x, y)) return -1;
    if (hit_before(cmp_metric_g, y, x)) return 1;
    return 0;
}

/* scores [nq, nlist] (similarity to maximise) -> probes [nq, nprobe] best first, ties -> lower id */
void orc_top_probes(const float *scores, int64_t nq, int nlist, int nprobe, int32_t *probes) {
    if (nprobe > nlist) nprobe = nlist;
#pragma omp parallel
    {
        orc_hit *h = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)nprobe);
#pragma omp for schedule(dynamic, 4)
        for (int64_t q = 0; q < nq; ++q) {
            int n = 0;
            const float *s = scores + q * (int64_t)nlist;
            for (int j = 0; j < nlist; ++j) {
                orc_hit x = {s[j], j};
                heap_push(h, &n, nprobe, x, ORC_IP);
            }
            cmp_metric_g = ORC_IP;
            qsort(h, (size_t)n, sizeof(orc_hit), hit_cmp);
            for (int j = 0; j < nprobe; ++j) probes[q * nprobe + j] = (int32_t)h[j].id;
        }
        free(h);
    }
}

/* C[m,n] = A[m,k] . B[n,k]^T   (both row-major, k contiguous), optional bias: C = alpha*C - bias[n] */
ORC_CLONES
static void gemm_block(const float *A, const float *B, float *C, int64_t m0, int64_t m1, int n0, int n1,
                       int k, int64_t ldc) {
    for (int64_t i = m0; i < m1; ++i) {
        const float *a = A + i * (int64_t)k;
        int j = n0;
        for (; j + 4 <= n1; j += 4) {
            const float *b0 = B + (int64_t)j * k, *b1 = b0 + k, *b2 = b1 + k, *b3 = b2 + k;
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int t = 0; t < k; ++t) {
                float av = a[t];
                s0 += av * b0[t];
                s1 += av * b1[t];
                s2 += av * b2[t];
                s3 += av * b3[t];
            }
            C[i * ldc + j] = s0;
            C[i * ldc + j + 1] = s1;
            C[i * ldc + j + 2] = s2;
            C[i * ldc + j + 3] = s3;
        }
        for (; j < n1; ++j) C[i * ldc + j] = orc_dot(a, B + (int64_t)j * k, k);
    }
}

void orc_gemm_nt(const float *A, int64_t m, const float *B, int n, int k, float *C) {
    const int NB = 64;
    const int64_t MB = 32;
    int64_t mblocks = (m + MB - 1) / MB;
    int nblocks = (n + NB - 1) / NB;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int64_t bi = 0; bi < mblocks; ++bi)
        for (int bj = 0; bj < nblocks; ++bj) {
            int64_t m0 = bi * MB, m1 = m0 + MB < m ? m0 + MB : m;
            int n0 = bj * NB, n1 = n0 + NB < n ? n0 + NB : n;
            gemm_block(A, B, C, m0, m1, n0, n1, k, n);
        }
}

/* similarity to maximise: IP -> q.c ; L2 -> 2 q.c - |c|^2 */
void orc_coarse_scores(const float *q, int64_t nq, const float *c, int nlist, int d, int metric, float *scores) {
    orc_gemm_nt(q, nq, c, nlist, d, scores);
    if (metric == ORC_L2) {
        float *cn = (float *)malloc(sizeof(float) * (size_t)nlist);
        for (int j = 0; j < nlist; ++j) cn[j] = orc_dot(c + (int64_t)j * d, c + (int64_t)j * d, d);
#pragma omp parallel for
        for (int64_t i = 0; i < nq; ++i)
            for (int j = 0; j < nlist; ++j) scores[i * nlist + j] = 2.0f * scores[i * nlist + j] - cn[j];
        free(cn);
    }
}

void orc_assign(const float *x, int64_t n, const float *c, int nlist, int d, int metric, int32_t *out) {
    const int64_t CH = 4096;
    float *scores = (float *)malloc(sizeof(float) * (size_t)CH * (size_t)nlist);
    for (int64_t s = 0; s < n; s += CH) {
        int64_t m = n - s < CH ? n - s : CH;
        orc_coarse_scores(x + s * d, m, c, nlist, d, metric, scores);
#pragma omp parallel for
        for (int64_t i = 0; i < m; ++i) {
            const float *r = scores + i * nlist;
            int best = 0;
            for (int j = 1; j < nlist; ++j)
                if (r[j] > r[best]) best = j;
            out[s + i] = best;
        }
    }
    free(scores);
}

/*
 * IVF_FLAT search with pre-computed probes.
 *   list_off [nlist+1], vecs [n,d] list-contiguous, ids [n], skip [n] or NULL (1 = skip row)
 *   out_dist/out_ids [nq,k]; raw IP or squared L2; pad: id -1, dist -FLT_MAX (IP) / FLT_MAX (L2)
 */
void orc_scan_search(const float *q, int64_t nq, int d, int metric, const int32_t *probes, int nprobe,
                     const int64_t *list_off, const float *vecs, const int64_t *ids, const uint8_t *skip,
                     int k, float *out_dist, int64_t *out_ids) {
#pragma omp parallel
    {
        orc_hit *h = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)k);
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; ++qi) {
            int n = 0;
            const float *qv = q + qi * (int64_t)d;
            for (int p = 0; p < nprobe; ++p) {
                int l = probes[qi * nprobe + p];
                if (l < 0) continue;
                for (int64_t r = list_off[l]; r < list_off[l + 1]; ++r) {
                    if (skip && skip[r]) continue;
                    orc_hit x;
                    x.v = metric == ORC_IP ? orc_dot(qv, vecs + r * d, d) : orc_l2sqr(qv, vecs + r * d, d);
                    x.id = ids[r];
                    heap_push(h, &n, k, x, metric);
                }
            }
            cmp_metric_g = metric;
            qsort(h, (size_t)n, sizeof(orc_hit), hit_cmp);
            for (int j = 0; j < k; ++j) {
                if (j < n) {
                    out_dist[qi * k + j] = h[j].v;
                    out_ids[qi * k + j] = h[j].id;
                } else {
                    out_dist[qi * k + j] = metric == ORC_IP ? -FLT_MAX : FLT_MAX;
                    out_ids[qi * k + j] = -1;
                }
            }
        }
        free(h);
    }
}

/* full search = coarse + probes + scan, for the timed CPU baseline */
void orc_search(const float *q, int64_t nq, int d, int metric, const float *centroids, int nlist, int nprobe,
                const int64_t *list_off, const float *vecs, const int64_t *ids, const uint8_t *skip, int k,
                float *out_dist, int64_t *out_ids) {
    if (nprobe > nlist) nprobe = nlist;
    float *scores = (float *)malloc(sizeof(float) * (size_t)nq * (size_t)nlist);
    int32_t *probes = (int32_t *)malloc(sizeof(int32_t) * (size_t)nq * (size_t)nprobe);
    orc_coarse_scores(q, nq, centroids, nlist, d, metric, scores);
    orc_top_probes(scores, nq, nlist, nprobe, probes);
    orc_scan_search(q, nq, d, metric, probes, nprobe, list_off, vecs, ids, skip, k, out_dist, out_ids);
    free(scores);
    free(probes);
}
