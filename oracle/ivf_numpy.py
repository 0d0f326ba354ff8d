"""NumPy restatement of FAISS ``IndexIVFFlat`` semantics (train / add / search).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/__init__.py).

What each function follows:

* index spec (metric IP, IVF_FLAT, nlist) ........ reference src/semcode/storage/milvus_store.py:76-83
* search params (nprobe=16, limit=top_k) ......... reference src/semcode/storage/milvus_store.py:141-147
* k-means train, add, search ..................... FAISS IndexIVFFlat as summarised in SURVEY.md
  section 8a rows a9-a11 [EXT]: coarse quantizer = exact flat search over the centroids with the
  index metric; assignment = best centroid under that metric (max inner product for IP, min
  squared L2 for L2); centroids = plain means; scan = exact fp32 inner product / squared L2
  (no sqrt) against every stored vector of the probed lists, rows whose mask bit is set are
  skipped *before* ranking; IP results descending, L2 ascending; short results padded with
  id -1.

Deviations that are ours (the FAISS RNG cannot be reproduced bit-for-bit and is [EXT]):
k-means initial centroids are the rows ``default_rng(seed).permutation(n)[:nlist]`` and an empty
cluster is re-seeded from the currently largest cluster (FAISS picks the donor at random,
weighted by size) with FAISS's +-1/1024 symmetric perturbation.

Tie rule of the oracle: equal scores are ordered by ascending id.  The product may order exact
ties differently; parity tests accept that and nothing else.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

METRIC_IP = 0
METRIC_L2 = 1

EPS_SPLIT = 1.0 / 1024.0  # FAISS Clustering.cpp split_clusters EPS [EXT]


def metric_code(metric) -> int:
    if metric in (METRIC_IP, "IP", "ip"):
        return METRIC_IP
    if metric in (METRIC_L2, "L2", "l2"):
        return METRIC_L2
    raise ValueError(f"unknown metric {metric!r}")


# --------------------------------------------------------------------------------------
# coarse quantizer
# --------------------------------------------------------------------------------------
def coarse_similarity(q: np.ndarray, centroids: np.ndarray, metric, dtype=np.float32) -> np.ndarray:
    """Similarity to *maximise*: IP -> q.c ; L2 -> 2 q.c - |c|^2 (= |q|^2 - |q-c|^2)."""
    q = np.ascontiguousarray(q, dtype=dtype)
    c = np.ascontiguousarray(centroids, dtype=dtype)
    s = q @ c.T
    if metric_code(metric) == METRIC_L2:
        s = 2.0 * s - np.einsum("ij,ij->i", c, c)[None, :]
    return s.astype(dtype, copy=False)


def top_desc(scores: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k largest per row, best first, ties -> lower index."""
    n = scores.shape[1]
    k = min(k, n)
    # stable argsort of -scores keeps lower index first on ties
    order = np.argsort(-scores, axis=1, kind="stable")
    return order[:, :k].astype(np.int32)


def coarse_probe(q, centroids, metric, nprobe: int, dtype=np.float32) -> np.ndarray:
    nprobe = min(int(nprobe), centroids.shape[0])  # FAISS clamps nprobe to nlist [EXT]
    return top_desc(coarse_similarity(q, centroids, metric, dtype), nprobe)


def assign(x, centroids, metric, dtype=np.float32, chunk: int = 65536) -> np.ndarray:
    """list(i) = argbest_j dist(x_i, c_j)  (FAISS IndexIVF::add_core -> quantizer->assign)."""
    out = np.empty(x.shape[0], dtype=np.int32)
    for s in range(0, x.shape[0], chunk):
        sim = coarse_similarity(x[s : s + chunk], centroids, metric, dtype)
        out[s : s + chunk] = np.argmax(sim, axis=1)
    return out


# --------------------------------------------------------------------------------------
# k-means (FAISS Clustering::train restated; see module docstring for the deviations)
# --------------------------------------------------------------------------------------
def kmeans_init_rows(n: int, nlist: int, seed: int) -> np.ndarray:
    return np.sort(np.random.default_rng(seed).permutation(n)[:nlist]).astype(np.int64)


def kmeans_subsample_rows(n: int, nlist: int, max_points_per_centroid: int, seed: int) -> Optional[np.ndarray]:
    """FAISS subsamples training to max_points_per_centroid*nlist rows [EXT]. None = use all."""
    if max_points_per_centroid <= 0 or n <= max_points_per_centroid * nlist:
        return None
    rows = np.random.default_rng(seed + 1).permutation(n)[: max_points_per_centroid * nlist]
    return np.sort(rows).astype(np.int64)


def split_empty_clusters(centroids: np.ndarray, counts: np.ndarray) -> int:
    """Re-seed every empty cluster from the largest one (in place). Returns #splits."""
    counts = counts.astype(np.int64, copy=True)
    d = centroids.shape[1]
    sign = np.where(np.arange(d) % 2 == 0, 1.0 + EPS_SPLIT, 1.0 - EPS_SPLIT).astype(centroids.dtype)
    sign_r = np.where(np.arange(d) % 2 == 0, 1.0 - EPS_SPLIT, 1.0 + EPS_SPLIT).astype(centroids.dtype)
    nsplit = 0
    for ci in np.flatnonzero(counts == 0):
        cj = int(np.argmax(counts))  # ties -> lowest index
        if counts[cj] < 2:
            break
        base = centroids[cj].copy()
        centroids[ci] = base * sign
        centroids[cj] = base * sign_r
        counts[ci] = counts[cj] // 2
        counts[cj] -= counts[ci]
        nsplit += 1
    return nsplit


def kmeans_train(
    x: np.ndarray,
    nlist: int,
    metric,
    niter: int = 10,
    seed: int = 1234,
    max_points_per_centroid: int = 256,
    init_centroids: Optional[np.ndarray] = None,
) -> Tuple[np.ndarray, list]:
    """Lloyd iterations; returns (centroids fp32 [nlist,d], objective per iteration).

    Objective = sum over training rows of the best similarity (IP) or best squared distance (L2),
    evaluated with the centroids *entering* the iteration (as FAISS logs it).
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    rows = kmeans_subsample_rows(n, nlist, max_points_per_centroid, seed)
    if rows is not None:
        x = x[rows]
        n = x.shape[0]
    if init_centroids is None:
        c = x[kmeans_init_rows(n, nlist, seed)].copy()
    else:
        c = np.array(init_centroids, dtype=np.float32, copy=True)
    mcode = metric_code(metric)
    objective = []
    xn = np.einsum("ij,ij->i", x.astype(np.float64), x.astype(np.float64))
    for _ in range(niter):
        a = np.empty(n, dtype=np.int32)
        best = np.empty(n, dtype=np.float64)
        for s in range(0, n, 65536):
            sim = coarse_similarity(x[s : s + 65536], c, mcode)
            a[s : s + 65536] = np.argmax(sim, axis=1)
            best[s : s + 65536] = sim[np.arange(sim.shape[0]), a[s : s + 65536]]
        objective.append(float(best.sum()) if mcode == METRIC_IP else float((xn - best).sum()))
        counts = np.bincount(a, minlength=nlist)
        sums = np.zeros((nlist, d), dtype=np.float64)
        np.add.at(sums, a, x.astype(np.float64))
        nz = counts > 0
        c[nz] = (sums[nz] / counts[nz, None]).astype(np.float32)
        split_empty_clusters(c, counts)
    return c, objective


# --------------------------------------------------------------------------------------
# index object (CSR inverted lists) + search
# --------------------------------------------------------------------------------------
@dataclass
class OracleIndex:
    metric: int
    centroids: np.ndarray  # [nlist, d] fp32
    list_off: np.ndarray  # [nlist+1] int64
    vecs: np.ndarray  # [n, d] fp32, list-contiguous
    ids: np.ndarray  # [n] int64
    repo_tags: np.ndarray  # [n] uint32
    lang_tags: np.ndarray  # [n] uint8

    @property
    def nlist(self) -> int:
        return self.centroids.shape[0]

    @property
    def ntotal(self) -> int:
        return self.vecs.shape[0]


def build_index(x, ids, centroids, metric, repo_tags=None, lang_tags=None, assignment=None) -> OracleIndex:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    ids = np.asarray(ids, dtype=np.int64)
    repo_tags = np.zeros(n, np.uint32) if repo_tags is None else np.asarray(repo_tags, np.uint32)
    lang_tags = np.zeros(n, np.uint8) if lang_tags is None else np.asarray(lang_tags, np.uint8)
    a = assign(x, centroids, metric) if assignment is None else np.asarray(assignment)
    order = np.argsort(a, kind="stable")
    nlist = centroids.shape[0]
    off = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(np.bincount(a, minlength=nlist), out=off[1:])
    return OracleIndex(
        metric_code(metric),
        np.ascontiguousarray(centroids, dtype=np.float32),
        off,
        x[order],
        ids[order],
        repo_tags[order],
        lang_tags[order],
    )


def row_mask(index: OracleIndex, repos=None, langs=None, removed_ids=None) -> Optional[np.ndarray]:
    """Boolean [n] -- True = row is *skipped* (knowhere BitsetView convention [EXT])."""
    if repos is None and langs is None and removed_ids is None:
        return None
    m = np.zeros(index.ntotal, dtype=bool)
    if repos is not None:
        m |= ~np.isin(index.repo_tags, np.asarray(list(repos), dtype=np.uint32))
    if langs is not None:
        m |= ~np.isin(index.lang_tags, np.asarray(list(langs), dtype=np.uint8))
    if removed_ids is not None:
        m |= np.isin(index.ids, np.asarray(list(removed_ids), dtype=np.int64))
    return m


def scan_scores(index: OracleIndex, q: np.ndarray, lo: int, hi: int, dtype=np.float32) -> np.ndarray:
    """Exact distance of one query against rows [lo,hi): IP or squared L2 (direct form)."""
    xs = index.vecs[lo:hi].astype(dtype, copy=False)
    qq = q.astype(dtype, copy=False)
    if index.metric == METRIC_IP:
        return xs @ qq
    diff = xs - qq[None, :]
    return np.einsum("ij,ij->i", diff, diff)


def search(
    index: OracleIndex,
    q: np.ndarray,
    k: int,
    nprobe: int,
    mask: Optional[np.ndarray] = None,
    probes: Optional[np.ndarray] = None,
    dtype=np.float32,
) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (dist [nq,k] fp32/fp64, ids [nq,k] int64); missing results: id -1, dist -/+ inf-like."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    nq = q.shape[0]
    if probes is None:
        probes = coarse_probe(q, index.centroids, index.metric, nprobe)
    pad = -np.finfo(np.float32).max if index.metric == METRIC_IP else np.finfo(np.float32).max
    out_d = np.full((nq, k), pad, dtype=dtype)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for qi in range(nq):
        ds, iz = [], []
        for l in probes[qi]:
            lo, hi = int(index.list_off[l]), int(index.list_off[l + 1])
            if hi == lo:
                continue
            s = scan_scores(index, q[qi], lo, hi, dtype)
            ii = index.ids[lo:hi]
            if mask is not None:
                keep = ~mask[lo:hi]
                s, ii = s[keep], ii[keep]
            ds.append(s)
            iz.append(ii)
        if not ds:
            continue
        s = np.concatenate(ds)
        ii = np.concatenate(iz)
        key = -s if index.metric == METRIC_IP else s
        order = np.lexsort((ii, key))[:k]
        out_d[qi, : order.size] = s[order]
        out_i[qi, : order.size] = ii[order]
    return out_d, out_i


def brute_force(x, ids, q, k, metric, dtype=np.float64) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k over all rows in `dtype` (ground truth for recall and for nprobe=nlist)."""
    x = np.asarray(x, dtype=dtype)
    q = np.asarray(q, dtype=dtype)
    ids = np.asarray(ids, dtype=np.int64)
    if metric_code(metric) == METRIC_IP:
        s = q @ x.T
        key = -s
    else:
        s = (q * q).sum(1)[:, None] - 2.0 * (q @ x.T) + (x * x).sum(1)[None, :]
        key = s
    k = min(k, x.shape[0])
    out_d = np.empty((q.shape[0], k), dtype=dtype)
    out_i = np.empty((q.shape[0], k), dtype=np.int64)
    for r in range(q.shape[0]):
        order = np.lexsort((ids, key[r]))[:k]
        out_d[r] = s[r, order]
        out_i[r] = ids[order]
    return out_d, out_i


def merge_topk(part_d: np.ndarray, part_i: np.ndarray, k: int, metric) -> Tuple[np.ndarray, np.ndarray]:
    """Merge [parts, nq, k'] partial results into [nq, k] (Milvus proxy reduce [EXT])."""
    parts, nq, kk = part_d.shape
    d = np.transpose(part_d, (1, 0, 2)).reshape(nq, parts * kk)
    i = np.transpose(part_i, (1, 0, 2)).reshape(nq, parts * kk)
    ip = metric_code(metric) == METRIC_IP
    pad = -np.finfo(np.float32).max if ip else np.finfo(np.float32).max
    out_d = np.full((nq, k), pad, dtype=part_d.dtype)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for r in range(nq):
        valid = i[r] >= 0
        dv, iv = d[r][valid], i[r][valid]
        order = np.lexsort((iv, -dv if ip else dv))[:k]
        out_d[r, : order.size] = dv[order]
        out_i[r, : order.size] = iv[order]
    return out_d, out_i


def recall_at_k(found_ids: np.ndarray, truth_ids: np.ndarray) -> float:
    k = truth_ids.shape[1]
    hit = 0
    for r in range(truth_ids.shape[0]):
        hit += np.intersect1d(found_ids[r][found_ids[r] >= 0], truth_ids[r]).size
    return hit / float(truth_ids.shape[0] * k)
