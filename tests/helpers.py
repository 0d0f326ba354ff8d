"""Shared checkers for the parity tests (the oracle is imported here and in tests only)."""

from __future__ import annotations

import numpy as np

# north_star tolerance: distances within 1e-5 RELATIVE.  A purely relative bound is ill-defined only for results near
# zero, which both metrics produce by cancellation: an inner product of near-orthogonal unit rows, and the coarse L2
# similarity 2 x.c - |c|^2 of a row next to its centroid (two terms of size ~1).  Any fp32 evaluation -- FAISS's AVX lanes,
# numpy's pairwise sums, a GPU's tree -- carries an ABSOLUTE rounding error of a few ulp of the largest intermediate
# there, so results that small are held to 4 ulp of the largest intermediate (|x|^2 + |c|^2 = 2 for the unit-norm rows
# used here): ATOL = 4 * 2^-24 * 2 ~ 5e-7.  (Round 1 used 1e-6; 2e-7 was tried and fails exactly on the L2 coarse
# similarity -0.0022 = 0.9978 - 1.0000 at d = 30, by 2.4e-7.)
RTOL = 1e-5
ATOL = 5e-7


def unit_rows(rng, n, d, dtype=np.float32):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(dtype)


def close(a, b):
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) <= RTOL * np.abs(b.astype(np.float64)) + ATOL


def assert_topk_parity(dist, ids, ref_dist, ref_ids, what=""):
    """ids identical except inside exact-distance ties; distances within tolerance, position by
    position; padding (id -1) identical."""
    dist, ids, ref_dist, ref_ids = map(np.asarray, (dist, ids, ref_dist, ref_ids))
    assert dist.shape == ref_dist.shape and ids.shape == ref_ids.shape, (what, dist.shape, ref_dist.shape)
    nq, k = ids.shape
    bad = []
    for q in range(nq):
        pad_g, pad_r = ids[q] < 0, ref_ids[q] < 0
        if not np.array_equal(pad_g, pad_r):
            bad.append((q, "padding", ids[q].tolist(), ref_ids[q].tolist()))
            continue
        v = ~pad_r
        if not close(dist[q][v], ref_dist[q][v]).all():
            j = int(np.flatnonzero(~close(dist[q][v], ref_dist[q][v]))[0])
            bad.append((q, "distance", j, float(dist[q][j]), float(ref_dist[q][j])))
            continue
        diff = np.flatnonzero(ids[q] != ref_ids[q])
        for j in diff:
            # a differing id is legitimate only inside a run of tied distances (or at the cut-off,
            # where the tied partner fell outside the top-k)
            d = ref_dist[q]
            tied_prev = j > 0 and close(d[j - 1 : j], d[j : j + 1])[0]
            tied_next = j + 1 < k and ref_ids[q][j + 1] >= 0 and close(d[j + 1 : j + 2], d[j : j + 1])[0]
            last = j == int(v.sum()) - 1
            if not (tied_prev or tied_next or last):
                bad.append((q, "id", int(j), int(ids[q][j]), int(ref_ids[q][j])))
                break
            if last and not (tied_prev or tied_next):
                # cut-off tie: the other id must not appear elsewhere in the reference row
                if ids[q][j] in ref_ids[q]:
                    bad.append((q, "id-dup", int(j), int(ids[q][j]), int(ref_ids[q][j])))
                    break
    assert not bad, f"{what}: {len(bad)} / {nq} queries differ, first: {bad[:3]}"


def assert_sorted(dist, ids, metric_ip: bool):
    dist, ids = np.asarray(dist), np.asarray(ids)
    for q in range(dist.shape[0]):
        v = ids[q] >= 0
        d = dist[q][v]
        if metric_ip:
            assert np.all(d[:-1] >= d[1:]), f"query {q} not descending"
        else:
            assert np.all(d[:-1] <= d[1:]), f"query {q} not ascending"
        assert len(set(ids[q][v].tolist())) == int(v.sum()), f"query {q} has duplicate ids"
        assert not v[int(v.sum()):].any(), f"query {q}: padding is not at the tail"
