"""Shared checkers for the parity tests (the oracle is imported here and in tests only)."""

from __future__ import annotations

import numpy as np

# north_star tolerance: distances within 1e-5 RELATIVE.  A purely relative bound is ill-defined only for results near
# zero: an fp32 dot product of d terms carries an ABSOLUTE rounding error of about sqrt(d) * 2^-24 * |q||x| whatever its
# summation order (FAISS's AVX lanes, numpy's pairwise sums and a GPU's tree all differ), so a result smaller than 2 % of
# |q||x| is held to the absolute precision of a result of that size: ATOL = RTOL * 0.02 for the unit-norm rows used here
# (5x tighter than round 1's 1e-6; measured worst case on the GPU box is 4e-7 at d = 3072 on scores >= 0.04).
RTOL = 1e-5
ATOL = 2e-7


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
