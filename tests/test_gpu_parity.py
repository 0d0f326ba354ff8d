"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (north_star): identical ids except exact-distance ties, distances within 1e-5 relative
(tests/helpers.py states the tolerance).  Everything here runs on cuda:0 of the GPU box.
"""

import numpy as np
import pytest

from helpers import assert_sorted, assert_topk_parity, close, unit_rows

pytestmark = pytest.mark.gpu

FMAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def sb(native_lib):
    import semcode_b200

    return semcode_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import ivf_numpy

    return ivf_numpy


@pytest.fixture(scope="module")
def orc_c():
    from oracle import ivf_c

    ivf_c.build()
    return ivf_c


def make_case(orc, n, d, nlist, nq, metric, seed=0, normalise=True):
    rng = np.random.default_rng(seed)
    x = unit_rows(rng, n, d) if normalise else rng.standard_normal((n, d)).astype(np.float32)
    q = unit_rows(rng, nq, d) if normalise else rng.standard_normal((nq, d)).astype(np.float32)
    cent = x[orc.kmeans_init_rows(n, nlist, seed)].copy()
    ids = (np.arange(n, dtype=np.int64) * 7 + 3)  # ids are not row numbers
    return x, q, cent, ids


def build_pair(sb, orc, x, ids, cent, metric, repo=None, lang=None):
    """Same rows, same centroids, same assignment in the oracle index and the GPU index."""
    assign = orc.assign(x, cent, metric)
    oidx = orc.build_index(x, ids, cent, metric, repo, lang, assignment=assign)
    g = sb.IVFFlatIndex(x.shape[1], nlist=cent.shape[0], metric=metric)
    g.set_centroids(cent)
    g.add(x, ids, repo, lang, lists=assign)
    return g, oidx, assign


# ---- KAT-1: tiny, hand-checkable ---------------------------------------------------------------
def test_tiny_known_answer(sb):
    # 8 points on the axes of R^4, two lists; query = e0 + 0.5 e1
    x = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1],
                  [2, 0, 0, 0], [0, 2, 0, 0], [0, 0, 2, 0], [0, 0, 0, 2]], dtype=np.float32)
    ids = np.arange(10, 18, dtype=np.int64)
    cent = np.array([[1, 1, 0, 0], [0, 0, 1, 1]], dtype=np.float32)
    q = np.array([[1, 0.5, 0, 0]], dtype=np.float32)
    lists = np.array([0, 0, 1, 1, 0, 0, 1, 1], dtype=np.int32)
    # inner product: list 0 holds ids 10,11,14,15 with q.x = 1, .5, 2, 1
    g = sb.IVFFlatIndex(4, nlist=2, metric="IP")
    g.set_centroids(cent)
    g.add(x, ids, lists=lists)
    assert g.probe(q, 1).tolist() == [[0]]
    d, i = g.search(q, 3, nprobe=1)
    assert i.tolist()[0][0] == 14 and d.tolist()[0] == [2.0, 1.0, 1.0]
    assert sorted(i.tolist()[0][1:]) == [10, 15]
    d, i = g.search(q, 8, nprobe=2)
    assert i.tolist()[0][0] == 14 and set(i.tolist()[0]) == set(range(10, 18))
    # k larger than the probed list: padded with -1 / -FLT_MAX
    d, i = g.search(q, 6, nprobe=1)
    assert i.tolist()[0][4:] == [-1, -1] and d[0, 4] == -FMAX
    # squared L2 (no sqrt): list 0 distances to q: 10 -> .25, 11 -> 1.25, 14 -> 1.25, 15 -> 3.25
    g2 = sb.IVFFlatIndex(4, nlist=2, metric="L2")
    g2.set_centroids(cent)
    g2.add(x, ids, lists=lists)
    d, i = g2.search(q, 4, nprobe=1)
    assert i.tolist()[0][0] == 10 and i.tolist()[0][3] == 15
    np.testing.assert_allclose(d[0], [0.25, 1.25, 1.25, 3.25], rtol=1e-6)
    d, i = g2.search(q, 5, nprobe=1)
    assert i[0, 4] == -1 and d[0, 4] == FMAX


# ---- coarse quantizer ---------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
@pytest.mark.parametrize("d", [768, 100, 30])
def test_probe_and_assign_match_oracle(sb, orc, metric, d):
    x, q, cent, ids = make_case(orc, 3000, d, 200, 150, metric, seed=d)
    g = sb.IVFFlatIndex(d, nlist=200, metric=metric)
    g.set_centroids(cent)
    np.testing.assert_array_equal(g.get_centroids(), cent)
    sim = orc.coarse_similarity(q, cent, metric, dtype=np.float64)
    want = orc.top_desc(sim, 17)
    got, sc = g.probe(q, 17, with_scores=True)
    for r in range(q.shape[0]):
        if not np.array_equal(got[r], want[r]):
            # only near-ties (fp32 rounding of the contraction) may reorder
            assert sorted(got[r]) == sorted(want[r]) or np.allclose(
                np.sort(sim[r][got[r]]), np.sort(sim[r][want[r]]), rtol=1e-5, atol=5e-7
            ), f"query {r}: probes differ beyond rounding"
        # coarse similarities only rank the lists (they are not returned distances): the 3xTF32 contraction drops the
        # lo x lo term, 2^-22 of |x||c|, which the L2 form 2 x.c - |c|^2 doubles and then cancels against (0.0011 +- 5.5e-7
        # seen at d = 30) -- so they get 1e-6 absolute on top of the relative bound, the final distances do not
        ref = sim[r][got[r]].astype(np.float32)
        assert (np.abs(sc[r].astype(np.float64) - ref) <= 1e-5 * np.abs(ref) + 1e-6).all()
    a = g.assign(x)
    a_ref = np.argmax(orc.coarse_similarity(x, cent, metric, dtype=np.float64), axis=1)
    diff = np.flatnonzero(a != a_ref)
    simx = orc.coarse_similarity(x[diff], cent, metric, dtype=np.float64)
    for j, r in enumerate(diff):
        assert abs(simx[j, a[r]] - simx[j, a_ref[r]]) <= 1e-5 * abs(simx[j, a_ref[r]]) + 1e-6


# ---- the list scan + top-k against the oracle, same probes --------------------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
@pytest.mark.parametrize("d,k", [(768, 10), (768, 50), (100, 10), (30, 7), (2048, 50), (3072, 10), (64, 200)])
def test_scan_parity_same_probes(sb, orc, orc_c, metric, d, k):
    n, nlist, nq, nprobe = 6000, 64, 37, 9
    x, q, cent, ids = make_case(orc, n, d, nlist, nq, metric, seed=d + k)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, metric)
    probes = orc.coarse_probe(q, cent, metric, nprobe)
    rd, ri = orc.search(oidx, q, k, nprobe, probes=probes)
    gd, gi = g.search(q, k, lists=probes)
    assert_topk_parity(gd, gi, rd, ri, f"numpy oracle {metric} d={d} k={k}")
    assert_sorted(gd, gi, metric == "IP")
    # second, independent checker: the C restatement
    cd, ci = orc_c.scan_search(q, oidx.metric, probes, oidx.list_off, oidx.vecs, oidx.ids, k)
    assert_topk_parity(gd, gi, cd, ci, f"C oracle {metric} d={d} k={k}")


@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_full_search_matches_oracle(sb, orc, metric):
    """coarse + scan + top-k end to end; the GPU must choose the oracle's probes."""
    x, q, cent, ids = make_case(orc, 20000, 128, 256, 200, metric, seed=5)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, metric)
    rd, ri = orc.search(oidx, q, 10, 16)
    gd, gi = g.search(q, 10, nprobe=16)
    same_probes = np.array([sorted(a) == sorted(b) for a, b in zip(g.probe(q, 16), orc.coarse_probe(q, cent, metric, 16))])
    assert same_probes.mean() > 0.97  # a near-tied 16th/17th centroid may flip
    assert_topk_parity(gd[same_probes], gi[same_probes], rd[same_probes], ri[same_probes], f"full {metric}")


@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_nprobe_equals_nlist_is_brute_force(sb, orc, metric):
    x, q, cent, ids = make_case(orc, 5000, 96, 32, 64, metric, seed=9, normalise=False)
    g, _, _ = build_pair(sb, orc, x, ids, cent, metric)
    bd, bi = orc.brute_force(x, ids, q, 10, metric)  # fp64 ground truth
    for nprobe in (32, 1000):  # nprobe is clamped to nlist
        gd, gi = g.search(q, 10, nprobe=nprobe)
        assert_topk_parity(gd, gi, bd.astype(np.float32), bi, f"brute {metric} nprobe={nprobe}")


def test_device_tensors_in_and_out(sb, orc):
    import torch

    x, q, cent, ids = make_case(orc, 4000, 768, 32, 50, "IP", seed=2)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, "IP")
    rd, ri = orc.search(oidx, q, 10, 8)
    qd = torch.from_numpy(q).cuda()
    gd, gi = g.search(qd, 10, nprobe=8)
    assert gd.is_cuda and gi.is_cuda and gi.dtype == torch.int64
    torch.cuda.synchronize()
    hd, hi = g.search(q, 10, nprobe=8)
    np.testing.assert_array_equal(gi.cpu().numpy(), hi)
    np.testing.assert_array_equal(gd.cpu().numpy(), hd)
    # device-side add: same index when the rows come from a CUDA tensor
    g2 = sb.IVFFlatIndex(768, nlist=32, metric="IP")
    g2.set_centroids(torch.from_numpy(cent).cuda())
    g2.add(torch.from_numpy(x).cuda(), torch.from_numpy(ids).cuda())
    d2, i2 = g2.search(q, 10, nprobe=8)
    assert_topk_parity(d2, i2, hd, hi, "device add")
    # non-default stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sd, si = g.search(qd, 10, nprobe=8)
    s.synchronize()
    np.testing.assert_array_equal(si.cpu().numpy(), hi)


# ---- filters, tombstones, upsert ----------------------------------------------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_filtered_search_parity(sb, orc, orc_c, metric):
    n, d, nlist = 30000, 256, 64
    x, q, cent, ids = make_case(orc, n, d, nlist, 60, metric, seed=11)
    rng = np.random.default_rng(3)
    repo = rng.zipf(1.3, n).clip(max=199).astype(np.uint32)
    lang = rng.integers(0, 2, n).astype(np.uint8)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, metric, repo, lang)
    probes = orc.coarse_probe(q, cent, metric, 24)
    # ~5 % selectivity: a handful of mid-frequency repos, one language
    repos = [5, 6, 7, 8, 9, 10, 11]
    mask = orc.row_mask(oidx, repos=repos, langs=[1])
    sel = 1.0 - mask.mean()
    assert 0.005 < sel < 0.2
    rd, ri = orc.search(oidx, q, 50, 24, mask=mask, probes=probes)
    gd, gi = g.search(q, 50, repos=repos, langs=[1], lists=probes)
    assert_topk_parity(gd, gi, rd, ri, f"filter {metric}")
    cd, ci = orc_c.scan_search(q, oidx.metric, probes, oidx.list_off, oidx.vecs, oidx.ids, 50, skip=mask)
    assert_topk_parity(gd, gi, cd, ci, f"filter C {metric}")
    # language only / repo only / unknown repo tag -> nothing
    m2 = orc.row_mask(oidx, langs=[0])
    rd, ri = orc.search(oidx, q, 10, 24, mask=m2, probes=probes)
    gd, gi = g.search(q, 10, langs=[0], lists=probes)
    assert_topk_parity(gd, gi, rd, ri, "lang only")
    gd, gi = g.search(q, 10, repos=[100000], lists=probes)
    assert (gi == -1).all()


def test_remove_and_reinsert(sb, orc):
    x, q, cent, ids = make_case(orc, 8000, 64, 16, 40, "IP", seed=21)
    g, oidx, assign = build_pair(sb, orc, x, ids, cent, "IP")
    probes = orc.coarse_probe(q, cent, "IP", 16)
    rd, ri = orc.search(oidx, q, 10, 16, probes=probes)
    victims = np.unique(ri[:, :3].ravel())
    victims = victims[victims >= 0]
    assert g.remove_ids(victims) == victims.size
    assert g.remove_ids(victims) == 0  # already gone
    assert g.ntotal == 8000 - victims.size
    mask = orc.row_mask(oidx, removed_ids=victims)
    rd2, ri2 = orc.search(oidx, q, 10, 16, mask=mask, probes=probes)
    gd, gi = g.search(q, 10, lists=probes)
    assert_topk_parity(gd, gi, rd2, ri2, "after remove")
    assert not np.isin(gi, victims).any()
    # upsert = remove + add under the same id with a new vector
    row = int(np.flatnonzero(ids == victims[0])[0])
    newv = q[:1].copy()
    g.add(newv, victims[:1], lists=assign[row : row + 1])
    gd, gi = g.search(q[:1], 1, nprobe=16)
    assert gi[0, 0] == victims[0] and abs(gd[0, 0] - float(newv[0] @ q[0])) < 1e-5


def test_incremental_add_equals_bulk(sb, orc):
    x, q, cent, ids = make_case(orc, 9000, 48, 40, 64, "L2", seed=31)
    g, oidx, assign = build_pair(sb, orc, x, ids, cent, "L2")
    g2 = sb.IVFFlatIndex(48, nlist=40, metric="L2")
    g2.set_centroids(cent)
    cuts = [0, 1, 33, 34, 2000, 2001, 7777, 9000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        g2.add(x[a:b], ids[a:b])  # GPU picks the lists itself
    np.testing.assert_array_equal(np.sort(g2.list_sizes()), np.sort(np.bincount(assign, minlength=40)))
    d1, i1 = g.search(q, 10, nprobe=40)
    d2, i2 = g2.search(q, 10, nprobe=40)
    assert_topk_parity(d2, i2, d1, i1, "incremental")
    off, vecs, eids, tags = g2.export_csr()
    assert off[-1] == 9000 and sorted(eids.tolist()) == sorted(ids.tolist())
    pos = {int(v): j for j, v in enumerate(ids)}
    np.testing.assert_array_equal(vecs[:50], x[[pos[int(v)] for v in eids[:50]]])


def test_empty_ragged_and_edge_cases(sb, orc):
    g = sb.IVFFlatIndex(20, nlist=8, metric="IP")
    with pytest.raises(sb.NativeError):
        g.search(np.zeros((1, 20), np.float32), 5)  # no centroids yet
    cent = np.eye(8, 20, dtype=np.float32)
    g.set_centroids(cent)
    d, i = g.search(np.ones((3, 20), np.float32), 5, nprobe=4)  # empty index
    assert (i == -1).all() and (d == -FMAX).all()
    d, i = g.search(np.zeros((0, 20), np.float32), 5)
    assert d.shape == (0, 5)
    # all rows in one list, the others empty; k above the row count
    x = np.tile(cent[2], (5, 1)) * np.arange(1, 6, dtype=np.float32)[:, None]
    g.add(x, np.arange(5, dtype=np.int64))
    assert g.list_sizes().tolist() == [0, 0, 5, 0, 0, 0, 0, 0]
    d, i = g.search(cent[2:3], 8, nprobe=8)
    assert i[0].tolist() == [4, 3, 2, 1, 0, -1, -1, -1]
    d, i = g.search(cent[5:6], 3, nprobe=1)  # probes an empty list only
    assert (i == -1).all()
    with pytest.raises(sb.NativeError):
        g.search(cent[:1], 0)
    with pytest.raises(sb.NativeError):
        g.search(cent[:1], 4096)
    with pytest.raises(sb.NativeError):
        g.add(x, np.arange(5, dtype=np.int64), lists=np.full(5, 8, np.int32))  # list id out of range
    assert g.ntotal == 5
    with pytest.raises(ValueError):
        g.search(np.zeros((1, 21), np.float32), 5)
    g.reset()
    assert g.ntotal == 0 and g.list_sizes().sum() == 0


def test_exact_ties_keep_both(sb):
    d = 32
    rng = np.random.default_rng(0)
    base = unit_rows(rng, 100, d)
    x = np.concatenate([base, base[:10]])  # ten exact duplicates under other ids
    ids = np.arange(110, dtype=np.int64)
    g = sb.IVFFlatIndex(d, nlist=1, metric="IP")
    g.set_centroids(np.zeros((1, d), np.float32))
    g.add(x, ids)
    dd, ii = g.search(base[:10], 2, nprobe=1)
    for r in range(10):
        assert sorted(ii[r].tolist()) == [r, 100 + r] and dd[r, 0] == dd[r, 1]


@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_one_pass_selection_and_its_fallbacks(sb, orc, metric):
    # >= 8192 candidates and k <= 128: the top-k kernel keeps 4 pairs per thread in ONE pass over the candidates;
    # k = 129 and 300 take the radix select.  One list, so candidate index == insertion order.
    n, d = 50_000, 32
    rng = np.random.default_rng(5)
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, 24, d)
    # queries 0..3: their ten best rows sit 512 slots apart -> one thread of the selection CTA meets them all,
    # keeps four, and the CTA must notice and fall back
    for qi in range(4):
        for j in range(10):
            x[(7 + 31 * qi) + 512 * j] = (q[qi] + 0.01 * (j + 1) * unit_rows(rng, 1, d)[0])
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    # query 4: sixty exact duplicates of its best row -> ties across the k-th position
    x[1000:1060] = x[999]
    q[4] = x[999]
    ids = np.arange(n, dtype=np.int64) + 11
    g = sb.IVFFlatIndex(d, nlist=1, metric=metric)
    g.set_centroids(np.zeros((1, d), np.float32))
    g.add(x, ids)
    for k in (1, 10, 50, 128, 129, 300):
        bd, bi = orc.brute_force(x, ids, q, k, metric)
        gd, gi = g.search(q, k, nprobe=1)
        assert_topk_parity(gd, gi, bd.astype(np.float32), bi, f"one list, k={k} {metric}")
    gd, gi = g.search(q[4:5], 10, nprobe=1)  # all ten out of the 61 tied rows (which ten: slot order inside the list)
    assert set(gi[0].tolist()) <= set((np.arange(999, 1060) + 11).tolist()) and len(set(gi[0].tolist())) == 10
    assert (gd[0] == gd[0, 0]).all()


def test_one_pass_selection_of_probes(sb, orc):
    # the same kernel ranks the centroids: 12000 lists >= 8192 -> one pass for nprobe <= 128
    d, nlist = 64, 12_000
    rng = np.random.default_rng(6)
    cent = unit_rows(rng, nlist, d)
    q = unit_rows(rng, 200, d)
    g = sb.IVFFlatIndex(d, nlist=nlist, metric="IP")
    g.set_centroids(cent)
    sim = orc.coarse_similarity(q, cent, "IP", dtype=np.float64)
    for nprobe in (1, 8, 32, 128, 200):
        want = orc.top_desc(sim, nprobe)
        got, sc = g.probe(q, nprobe, with_scores=True)
        for r in range(q.shape[0]):
            assert len(set(got[r].tolist())) == nprobe
            assert close(sc[r], sim[r][got[r]].astype(np.float32)).all()
            assert (np.diff(sc[r]) <= 0).all()
            if not np.array_equal(got[r], want[r]):  # only near-ties (fp32 rounding of the contraction) may reorder
                assert np.allclose(np.sort(sim[r][got[r]]), np.sort(sim[r][want[r]]), rtol=1e-5, atol=5e-7), (nprobe, r)


def test_chunked_search_equals_single_pass(sb, orc):
    x, q, cent, ids = make_case(orc, 20000, 64, 128, 700, "IP", seed=41)
    g, _, _ = build_pair(sb, orc, x, ids, cent, "IP")
    d1, i1 = g.search(q, 10, nprobe=12)
    g.set_param("scratch_bytes", 1 << 20)  # forces many query chunks
    d2, i2 = g.search(q, 10, nprobe=12)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(d1, d2)
    for v in (1, 2, 3, 4):
        g.set_param("scan_variant", v)
        d3, i3 = g.search(q, 10, nprobe=12)
        assert_topk_parity(d3, i3, d1, i1, f"scan variant {v}")


def test_merge_topk(sb, orc):
    import torch

    rng = np.random.default_rng(1)
    for metric in ("IP", "L2"):
        pd = rng.standard_normal((4, 33, 10)).astype(np.float32)
        pd = -np.sort(-pd, axis=2) if metric == "IP" else np.sort(pd, axis=2)
        pi = rng.permutation(4 * 33 * 10).reshape(4, 33, 10).astype(np.int64)
        pi[1, :, 7:] = -1  # a shard with short results
        rd, ri = orc.merge_topk(pd, pi, 10, metric)
        gd, gi = sb.merge_topk(torch.from_numpy(pd).cuda(), torch.from_numpy(pi).cuda(), 10, metric)
        assert_topk_parity(gd.cpu().numpy(), gi.cpu().numpy(), rd, ri, f"merge {metric}")


# ---- k-means --------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_kmeans_tracks_oracle(sb, orc, metric):
    rng = np.random.default_rng(17)
    centres = rng.standard_normal((24, 40)).astype(np.float32) * 3
    x = (centres[rng.integers(0, 24, 6000)] + rng.standard_normal((6000, 40)).astype(np.float32)).astype(np.float32)
    if metric == "IP":
        x /= np.linalg.norm(x, axis=1, keepdims=True)  # embeddings are ~unit norm; keeps IP clusters non-empty
    c_ref, obj_ref = orc.kmeans_train(x, 24, metric, niter=8, seed=5, max_points_per_centroid=0)
    g = sb.IVFFlatIndex(40, nlist=24, metric=metric)
    obj = g.train(x, niter=8, seed=5, max_points_per_centroid=0)
    # a point whose two best centroids tie within fp32 rounding may be assigned differently, which
    # moves two centroids by ~|x-c|/count: tolerances are set for a handful of such flips
    np.testing.assert_allclose(obj, obj_ref, rtol=1e-3)
    if metric == "L2":
        assert all(b <= a * (1 + 1e-6) for a, b in zip(obj[:-1], obj[1:])), "L2 objective must not increase"
    np.testing.assert_allclose(g.get_centroids(), c_ref, rtol=1e-2, atol=2e-2)
    # Python-driven Lloyd (the building blocks the sharded trainer uses) gives the same result
    g2 = sb.IVFFlatIndex(40, nlist=24, metric=metric)
    obj2 = g2.train(x, niter=8, init_centroids=x[orc.kmeans_init_rows(6000, 24, 5)])
    np.testing.assert_allclose(obj2, obj, rtol=1e-6)
    np.testing.assert_allclose(g2.get_centroids(), g.get_centroids(), rtol=1e-5, atol=5e-7)


def test_kmeans_empty_cluster_split(sb, orc):
    # two tight blobs but four centroids, two of them initialised far away -> empty -> split
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.standard_normal((500, 8)) * 0.01 + 5, rng.standard_normal((300, 8)) * 0.01 - 5]).astype(np.float32)
    init = np.stack([x[0], x[600], np.full(8, 100, np.float32), np.full(8, -100, np.float32)])
    c_ref, obj_ref = orc.kmeans_train(x, 4, "L2", niter=3, init_centroids=init, max_points_per_centroid=0)
    g = sb.IVFFlatIndex(8, nlist=4, metric="L2")
    obj = g.train(x, niter=3, init_centroids=init)
    np.testing.assert_allclose(obj, obj_ref, rtol=5e-2)  # |x|^2 - best cancels: fp32 noise dominates tiny objectives
    np.testing.assert_allclose(g.get_centroids(), c_ref, rtol=1e-3, atol=1e-3)
    g.add(x, np.arange(800, dtype=np.int64))
    assert (g.list_sizes() > 0).all()


# ---- size-independent properties at a larger size -----------------------------------------------
def test_properties_at_scale(sb, orc):
    import torch

    n, d, nlist, nq = 400_000, 768, 1024, 512
    gen = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn((n, d), generator=gen, device="cuda")
    x = torch.nn.functional.normalize(x, dim=1)
    g = sb.IVFFlatIndex(d, nlist=nlist, metric="IP")
    g.train(x, niter=2, max_points_per_centroid=64)
    ids = torch.arange(n, device="cuda", dtype=torch.int64)
    g.add(x, ids)
    assert g.ntotal == n and int(g.list_sizes().sum()) == n
    q = x[:nq].clone()
    prev_hit = None
    for nprobe in (1, 8, 64):
        dd, ii = g.search(q, 10, nprobe=nprobe)
        dd, ii = dd.cpu().numpy(), ii.cpu().numpy()
        assert_sorted(dd, ii, True)
        # a stored vector queried with itself comes back first, with inner product 1
        assert (ii[:, 0] == np.arange(nq)).all()
        np.testing.assert_allclose(dd[:, 0], 1.0, rtol=1e-5)
        if prev_hit is not None:
            # probing more lists can only improve the k-th best similarity
            assert (dd[:, 9] >= prev_hit - 1e-6).all()
        prev_hit = dd[:, 9]
    # exhaustive probing == exact search (fp64 check on a sample)
    dd, ii = g.search(q[:32], 10, nprobe=nlist)
    exact = (q[:32].double() @ x.double().T).topk(10, dim=1)
    np.testing.assert_array_equal(ii.cpu().numpy(), exact.indices.cpu().numpy())
    np.testing.assert_allclose(dd.cpu().numpy(), exact.values.cpu().numpy(), rtol=1e-5)
    t = g.stats()
    assert t.npages >= n // 32 and t.bytes_lists >= n * d * 4


def test_exhaustive_probe_beyond_select_limit(sb, orc):
    """nprobe >= nlist needs no ranking, so it is not bound by the 2048-wide selection kernel."""
    x, q, cent, ids = make_case(orc, 9000, 32, 3000, 20, "IP", seed=51)
    g, _, _ = build_pair(sb, orc, x, ids, cent, "IP")
    bd, bi = orc.brute_force(x, ids, q, 10, "IP")
    gd, gi = g.search(q, 10, nprobe=3000)
    assert_topk_parity(gd, gi, bd.astype(np.float32), bi, "exhaustive nlist=3000")
    gd, gi = g.search(q, 10, nprobe=100000)
    assert_topk_parity(gd, gi, bd.astype(np.float32), bi, "exhaustive clamp")
    with pytest.raises(sb.NativeError):
        g.search(q, 10, nprobe=2500)
    with pytest.raises(sb.NativeError):
        g.probe(q, 2500)


# ---- tcgen05 3xTF32 contraction (gemm_tc.cu) vs exact arithmetic -----------------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
@pytest.mark.parametrize("d,nlist,nq", [(768, 1000, 300), (96, 257, 129), (3072, 512, 5), (100, 64, 1)])
def test_tensor_core_coarse_scores_have_fp32_accuracy(sb, orc, metric, d, nlist, nq):
    rng = np.random.default_rng(d + nlist)
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    g = sb.IVFFlatIndex(d, nlist=nlist, metric=metric)
    g.set_centroids(cent)
    g.set_param("small_coarse", 0)  # keep tiny batches on the contraction kernels under test here
    exact = orc.coarse_similarity(q, cent, metric, dtype=np.float64)
    scale = np.abs(q.astype(np.float64)) @ np.abs(cent.astype(np.float64)).T * (2 if metric == "L2" else 1)
    if metric == "L2":
        scale = scale + (cent.astype(np.float64) ** 2).sum(1)[None, :]
    errs = {}
    for impl in (0, 1):  # 0 = tcgen05 3xTF32, 1 = fp32 SIMT
        g.set_param("coarse_impl", impl)
        lists, sc = g.probe(q, min(nlist, 64), with_scores=True)
        ref = np.take_along_axis(exact, lists.astype(np.int64), axis=1)
        errs[impl] = float(np.max(np.abs(sc - ref) / np.take_along_axis(scale, lists.astype(np.int64), axis=1)))
        # ranking is the exact ranking up to rounding-level ties
        want = orc.top_desc(exact, min(nlist, 64))
        for r in range(nq):
            if not np.array_equal(lists[r], want[r]):
                a, b = np.sort(exact[r][lists[r]]), np.sort(exact[r][want[r]])
                assert np.allclose(a, b, rtol=1e-5, atol=1e-5 * scale[r].max()), (impl, r)
    # error relative to sum |a||b|: fp32 FMA chains give ~1.5e-7; 3xTF32 on the tensor pipe (truncating
    # accumulation over K/8 x 3 MMAs) measures ~2e-6 at K=3072; plain TF32 would be ~5e-4
    assert errs[1] < 1e-6 and errs[0] < 5e-6, errs


@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_tensor_core_fused_argmax_matches_simt(sb, orc, metric):
    # n >= 4096 takes the fused contraction+argmax kernel; nlist spans several slabs and a ragged tail
    n, d, nlist = 9000, 2048, 6500
    rng = np.random.default_rng(3)
    x = unit_rows(rng, n, d)
    cent = unit_rows(rng, nlist, d) * rng.uniform(0.5, 1.5, (nlist, 1)).astype(np.float32)
    g = sb.IVFFlatIndex(d, nlist=nlist, metric=metric)
    g.set_centroids(cent)
    a_tc = g.assign(x)  # 256 x 256 tiles (64-byte swizzle)
    g.set_param("tc_variant", 1)
    a_tc1 = g.assign(x)  # 128 x 256 tiles (128-byte swizzle)
    g.set_param("coarse_impl", 1)
    a_simt = g.assign(x)
    exact = orc.coarse_similarity(x, cent, metric, dtype=np.float64)
    a_ref = np.argmax(exact, axis=1)
    for a in (a_tc, a_tc1, a_simt):
        diff = np.flatnonzero(a != a_ref)
        assert diff.size < 10
        for r in diff:  # only rounding-level ties may differ
            assert abs(exact[r, a[r]] - exact[r, a_ref[r]]) <= 1e-5 * abs(exact[r, a_ref[r]]) + 1e-6


# ---- list-major scan (scan_lists.cu) == query-major scan (scan.cu) == oracle ---------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
@pytest.mark.parametrize("d,nlist,nq,nprobe", [(768, 16, 100, 9), (200, 7, 41, 7), (128, 40, 300, 5), (3072, 8, 70, 3)])
def test_list_major_scan_matches_query_major_and_oracle(sb, orc, metric, d, nlist, nq, nprobe):
    # few lists + many queries -> every list is probed by 1..nq queries: exercises the 32- and 8-query
    # tiles, ragged chunks (counts like 33, 40, 9), ragged row tiles and a partial last k-stage (d=200)
    n = 5000
    x, q, cent, ids = make_case(orc, n, d, nlist, nq, metric, seed=d + nq)
    rng = np.random.default_rng(1)
    repo = rng.integers(0, 5, n).astype(np.uint32)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, metric, repo=repo)
    g.remove_ids(ids[::17])
    probes = orc.coarse_probe(q, cent, metric, nprobe)
    probes[::7, -1] = -1  # skipped probe slots
    mask = orc.row_mask(oidx, removed_ids=ids[::17])
    out = {}
    for mode in (1, 2):
        g.set_param("scan_mode", mode)
        out[mode] = g.search(q, 10, lists=probes)
        out[mode, "f"] = g.search(q, 10, lists=probes, repos=[1, 3])
    rd, ri = orc.search(oidx, q, 10, nprobe, mask=mask, probes=probes_valid(orc, probes))
    for key in (1, 2):
        assert_topk_parity(out[key][0], out[key][1], rd, ri, f"mode {key} {metric} d={d}")
    fmask = mask | orc.row_mask(oidx, repos=[1, 3])
    fd, fi = orc.search(oidx, q, 10, nprobe, mask=fmask, probes=probes_valid(orc, probes))
    for key in ((1, "f"), (2, "f")):
        assert_topk_parity(out[key][0], out[key][1], fd, fi, f"filtered mode {key} {metric} d={d}")
    np.testing.assert_array_equal(out[1][1], out[2][1])


@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_plan_scans_across_many_ctas(sb, metric):
    # The pair plan (1024 pairs per CTA) and the list plan (1024 lists per CTA) are single-pass scans with a
    # look-back over the preceding CTAs: 140800 pairs -> 138 CTAs (look-back windows of 32), 40000 lists -> 40 CTAs.
    # Query-major and list-major must agree, and both must match a brute-force scan of the probed lists.
    n, d, nlist, nq, nprobe, k = 120_000, 128, 40_000, 2200, 64, 10
    rng = np.random.default_rng(99)
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    ids = np.arange(n, dtype=np.int64) * 3 + 1
    assign = rng.integers(0, nlist, n).astype(np.int32)
    g = sb.IVFFlatIndex(d, nlist=nlist, metric=metric)
    g.set_centroids(unit_rows(rng, nlist, d))
    g.add(x, ids, lists=assign)
    start = rng.integers(0, nlist, nq)
    probes = ((start[:, None] + np.arange(nprobe)[None, :] * 617) % nlist).astype(np.int32)  # distinct per query
    out = {}
    for mode in (1, 2):
        g.set_param("scan_mode", mode)
        for rep in range(3):  # the look-back words are reused from launch to launch (epochs)
            out[mode] = g.search(q, k, lists=probes)
    g.set_param("plan_epoch", (1 << 22) - 3)  # ... and zeroed when the 22-bit epoch wraps
    for rep in range(4):
        wd, wi = g.search(q, k, lists=probes)
        np.testing.assert_array_equal(wi, out[2][1])
    np.testing.assert_array_equal(out[1][1], out[2][1])
    assert_topk_parity(out[2][0], out[2][1], out[1][0], out[1][1], f"list-major vs query-major {metric}")
    order = np.argsort(assign, kind="stable")
    bounds = np.searchsorted(assign[order], np.arange(nlist + 1))
    for qi in list(range(0, nq, 97)) + [nq - 1]:
        rows = np.concatenate([order[bounds[l] : bounds[l + 1]] for l in probes[qi]])
        xr = x[rows].astype(np.float64)
        if metric == "IP":
            sc = xr @ q[qi].astype(np.float64)
            best = np.argsort(-sc, kind="stable")[:k]
        else:
            sc = ((xr - q[qi].astype(np.float64)) ** 2).sum(axis=1)
            best = np.argsort(sc, kind="stable")[:k]
        rd = np.full(k, FMAX if metric == "L2" else -FMAX, dtype=np.float64)
        ri = np.full(k, -1, dtype=np.int64)
        rd[: best.size] = sc[best]
        ri[: best.size] = ids[rows[best]]
        assert_topk_parity(out[1][0][qi : qi + 1], out[1][1][qi : qi + 1], rd[None, :].astype(np.float32), ri[None, :],
                           f"query {qi} vs brute force {metric}")


def probes_valid(orc, probes):
    """The NumPy oracle indexes list_off with every probe: map the skipped (-1) slots to an empty
    trailing list by handing it a ragged Python structure instead."""

    class _Rows:
        def __init__(self, p):
            self.p = p

        def __getitem__(self, qi):
            r = self.p[qi]
            return r[r >= 0]

    return _Rows(probes)


def test_list_major_auto_mode_full_search(sb, orc):
    # nq * nprobe >= 2 * nlist switches the scan to list-major automatically
    x, q, cent, ids = make_case(orc, 30000, 256, 64, 600, "IP", seed=77)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, "IP")
    d0, i0 = g.search(q, 10, nprobe=12)  # 7200 pairs >= 1.5 * 64 lists of ~470 rows: auto -> list-major
    g.set_param("scan_mode", 1)
    d1, i1 = g.search(q, 10, nprobe=12)
    np.testing.assert_array_equal(i0, i1)
    assert_topk_parity(d0, i0, d1, i1, "auto vs query-major")
    g.set_profiling(True)
    g.set_param("scan_mode", 0)
    g.search(q, 10, nprobe=12)
    assert g.last_search_times().scan_launches in (6, 7)  # count, plan, fill + tile items (+ query split for the tcgen05 tiles) + the two page scans
    g.set_param("scan_mode", 1)
    g.search(q, 10, nprobe=12)
    assert g.last_search_times().scan_launches == 1


@pytest.mark.parametrize("metric", ["IP", "L2"])
@pytest.mark.parametrize("per_list", [3, 7, 13])
def test_multi_query_page_scan_buckets_and_slices(sb, orc, metric, per_list):
    """The two multi-query page scans (scan_mq.cu): remainders of 1..4 queries (bucket 0), 5..16 queries in one or
    two passes of 8 (bucket 1), over single-slice, multi-slice and ragged-slice dimensions (lists_cfg = 2 only changes
    the stage width of the 32-query tile)."""
    for d in (128, 200, 768, 1024, 2048, 3072):
        n = 3000 if d <= 1024 else 1500
        x, q, cent, ids = make_case(orc, n, d, 24, 12 * per_list, metric, seed=d + per_list)
        g, oidx, _ = build_pair(sb, orc, x, ids, cent, metric)
        g.remove_ids(ids[::13])
        probes = orc.coarse_probe(q, cent, metric, 2)  # ~per_list queries per list
        g.set_param("scan_mode", 1)
        d1, i1 = g.search(q, 10, lists=probes)
        rd, ri = orc.search(oidx, q, 10, 2, mask=orc.row_mask(oidx, removed_ids=ids[::13]), probes=probes)
        for cfg in (0, 2, 4):  # 4: the 8-query bucket on mma.sync (inner product, dim % 16 == 0)
            g.set_param("scan_mode", 2)
            g.set_param("lists_cfg", cfg)
            d3, i3 = g.search(q, 10, lists=probes)
            assert_topk_parity(d3, i3, d1, i1, f"mq cfg={cfg} {metric} d={d}")
            assert_topk_parity(d3, i3, rd, ri, f"mq cfg={cfg} vs oracle {metric} d={d}")
        # both buckets in one launch (an option; same arithmetic, so the very same bits as the two launches)
        g.set_param("lists_cfg", 0)
        d3, i3 = g.search(q, 10, lists=probes)
        g.set_param("mq_fused", 1)
        d4, i4 = g.search(q, 10, lists=probes)
        np.testing.assert_array_equal(i4, i3)
        np.testing.assert_array_equal(d4, d3)


@pytest.mark.parametrize("d,nlist,nq,nprobe,k", [(768, 6, 500, 3, 10), (128, 3, 200, 2, 50), (3072, 2, 150, 1, 10), (1024, 12, 900, 5, 7),
                                                 (64, 5, 333, 2, 10)])
def test_tensor_core_tiles_match_ffma_tiles_and_oracle(sb, orc, d, nlist, nq, nprobe, k):
    """Lists probed by many queries (inner product): tcgen05 tile items (scan_lists_tc.cu, 64-query chunks, split
    tf32 operands, four accumulators per tile) against the exact-fp32 FFMA tiles (lists_cfg = 1), the query-major
    scan and the oracle -- with tombstones, a language filter, ragged last tiles and ragged query chunks."""
    n = 9000 if d <= 1024 else 2500
    rng = np.random.default_rng(d + nq)
    x, q, cent, ids = make_case(orc, n, d, nlist, nq, "IP", seed=d)
    lang = rng.integers(0, 3, size=n).astype(np.uint8)
    assign = orc.assign(x, cent, "IP")
    oidx = orc.build_index(x, ids, cent, "IP", None, lang, assignment=assign)
    g = sb.IVFFlatIndex(d, nlist=nlist, metric="IP")
    g.set_centroids(cent)
    g.add(x, ids, None, lang, lists=assign)
    g.remove_ids(ids[::11])
    probes = orc.coarse_probe(q, cent, "IP", nprobe)
    for langs in (None, [0, 2]):
        mask = orc.row_mask(oidx, langs=langs, removed_ids=ids[::11])
        rd, ri = orc.search(oidx, q, k, nprobe, mask=mask, probes=probes)
        g.set_param("scan_mode", 1)
        d1, i1 = g.search(q, k, lists=probes, langs=langs)
        assert_topk_parity(d1, i1, rd, ri, f"query-major d={d} langs={langs}")
        for cfg in (0, 1, 5):  # 0: list rows as a tensor-memory operand (scan_lists_ts.cu), 5: both operands in shared memory
            g.set_param("scan_mode", 2)
            g.set_param("lists_cfg", cfg)
            d2, i2 = g.search(q, k, lists=probes, langs=langs)
            assert_topk_parity(d2, i2, rd, ri, f"list-major cfg={cfg} vs oracle d={d} langs={langs}")
            assert_topk_parity(d2, i2, d1, i1, f"list-major cfg={cfg} vs query-major d={d} langs={langs}")


def test_sharded_step_with_a_one_rank_exchange(sb, orc):
    """sc_index_search_sharded with world = 1 (the exchange buffer is this GPU's own memory): the peer stores,
    the flag publication and the waiting merge run on one GPU and must reproduce the plain search."""
    import ctypes as C

    import torch

    from semcode_b200 import _capi

    x, q, cent, ids = make_case(orc, 5000, 128, 16, 70, "IP", seed=5)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, "IP")
    nbytes = 8 << 20
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    arr = (C.c_void_p * 1)(buf.data_ptr())
    h = C.c_void_p()
    L = _capi.lib()
    _capi.check(L.sc_exchange_create(0, 1, arr, nbytes, 0, C.byref(h)))

    class Ex:
        handle = h

    try:
        for rep, (m, k, nprobe) in enumerate(((70, 10, 4), (1, 5, 16), (33, 50, 7), (70, 10, 4))):
            d0, i0 = g.search(q[:m], k, nprobe=nprobe)
            d1, i1 = g.search(torch.from_numpy(q[:m]).cuda(), k, nprobe=nprobe, exchange=Ex)
            assert_topk_parity(d1.cpu().numpy(), i1.cpu().numpy(), d0, i0, f"one-rank exchange step {rep}")
            probes = orc.coarse_probe(q[:m], cent, "IP", nprobe)
            d2, i2 = g.search(torch.from_numpy(q[:m]).cuda(), k, lists=probes, exchange=Ex)
            assert_topk_parity(d2.cpu().numpy(), i2.cpu().numpy(), d0, i0, f"one-rank exchange, given lists, step {rep}")
        t, e = C.c_int32(0), C.c_int64(0)
        _capi.check(L.sc_exchange_status(h, C.byref(t), C.byref(e)))
        assert t.value == 0 and e.value == 8
        with pytest.raises(_capi.NativeError, match="too small"):
            g.search(torch.from_numpy(np.repeat(q, 40, axis=0)).cuda(), 2048, nprobe=16, exchange=Ex)
    finally:
        L.sc_exchange_destroy(h)


# ---- randomized shapes: every scan route against the oracle ---------------------------------------------
def test_randomized_shapes_all_scan_routes(sb, orc):
    rng = np.random.default_rng(20261018)
    for trial in range(14):
        d = int(rng.choice([4, 20, 64, 100, 128, 200, 256, 384, 768, 1024, 1536]))
        n = int(rng.integers(200, 6000))
        nlist = int(rng.integers(1, 48))
        nq = int(rng.integers(1, 260))
        nprobe = int(rng.integers(1, nlist + 1))
        k = int(rng.choice([1, 5, 10, 64, 300]))
        metric = "IP" if trial % 2 == 0 else "L2"
        x = rng.standard_normal((n, d)).astype(np.float32)
        q = rng.standard_normal((nq, d)).astype(np.float32)
        # skewed list sizes: a few centroids attract most rows
        cent = x[rng.integers(0, n, nlist)] * rng.uniform(0.2, 2.0, (nlist, 1)).astype(np.float32)
        ids = rng.permutation(10 * n)[:n].astype(np.int64)
        repo = rng.integers(0, 6, n).astype(np.uint32)
        lang = rng.integers(0, 3, n).astype(np.uint8)
        g, oidx, _ = build_pair(sb, orc, x, ids, cent, metric, repo, lang)
        gone = ids[rng.random(n) < 0.05]
        if gone.size:
            g.remove_ids(gone)
        use_filter = trial % 3 == 0
        mask = orc.row_mask(oidx, repos=[1, 2, 4] if use_filter else None, langs=[0, 2] if use_filter else None,
                            removed_ids=gone if gone.size else None)
        probes = orc.coarse_probe(q, cent, metric, nprobe)
        rd, ri = orc.search(oidx, q, k, nprobe, mask=mask, probes=probes)
        kw = dict(repos=[1, 2, 4], langs=[0, 2]) if use_filter else {}
        for mode in (1, 2):
            g.set_param("scan_mode", mode)
            gd, gi = g.search(q, k, lists=probes, **kw)
            assert_topk_parity(gd, gi, rd, ri, f"trial {trial} mode {mode} {metric} d={d} n={n} nlist={nlist} nq={nq} "
                                               f"nprobe={nprobe} k={k} filter={use_filter}")
            assert_sorted(gd, gi, metric == "IP")
        g.close()


# ---- small batches: streamed coarse pass + sub-page split of the query-major scan -------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
@pytest.mark.parametrize("d,nq", [(768, 1), (768, 3), (100, 16), (3072, 2), (30, 5)])
def test_small_batch_path(sb, orc, metric, d, nq):
    x, q, cent, ids = make_case(orc, 7000, d, 50, nq, metric, seed=d * 31 + nq)
    g, oidx, _ = build_pair(sb, orc, x, ids, cent, metric)
    exact = orc.coarse_similarity(q, cent, metric, dtype=np.float64)
    lists, sc = g.probe(q, 11, with_scores=True)
    want = orc.top_desc(exact, 11)
    for r in range(nq):
        if not np.array_equal(lists[r], want[r]):
            assert np.allclose(np.sort(exact[r][lists[r]]), np.sort(exact[r][want[r]]), rtol=1e-5, atol=5e-7)
        assert close(sc[r], exact[r][lists[r]].astype(np.float32)).all()
    g.set_param("small_coarse", 0)
    lists_tc = g.probe(q, 11)
    assert np.mean([sorted(a) == sorted(b) for a, b in zip(lists, lists_tc)]) >= 0.5  # same ranking up to near-ties
    g.set_param("small_coarse", 1)
    for nprobe in (1, 4, 50):  # 3 .. 150 pages for one query: 4-, 2- and 1-way page splits
        probes = orc.coarse_probe(q, cent, metric, nprobe)
        rd, ri = orc.search(oidx, q, 10, nprobe, probes=probes)
        g.set_param("scan_mode", 1)
        gd, gi = g.search(q, 10, lists=probes)
        assert_topk_parity(gd, gi, rd, ri, f"small batch {metric} d={d} nq={nq} nprobe={nprobe}")
        gd2, gi2 = g.search(q, 10, lists=probes, repos=[0])
        assert_topk_parity(gd2, gi2, rd, ri, "all rows carry repo tag 0")
        gd3, gi3 = g.search(q, 10, lists=probes, repos=[7])
        assert (gi3 == -1).all()


# ---- failure atomicity of inserts, tag validation ---------------------------------------------------
def test_failed_add_chunk_leaves_a_consistent_index(sb, orc):
    """A multi-chunk add whose third chunk fails AFTER its slots were claimed: chunks 1-2 stay committed, the list
    lengths of the failed chunk are rolled back on the device and the host mirror follows, so the next search sizes its
    scratch from the truth; later inserts work.  Same for a bad list id and an oversize repo tag in a later chunk."""
    n, d, nlist = 6000, 64, 8
    x, q, cent, ids = make_case(orc, n, d, nlist, 40, "IP", seed=3)
    assign = orc.assign(x, cent, "IP")
    g = sb.IVFFlatIndex(d, nlist=nlist, metric="IP")
    g.set_centroids(cent)
    g.set_param("add_chunk_rows", 1500)
    g.set_param("fail_add_after", 3)
    with pytest.raises(sb.NativeError, match="injected"):
        g.add(x, ids, lists=assign)
    assert g.ntotal == 3000
    np.testing.assert_array_equal(g.list_sizes(), np.bincount(assign[:3000], minlength=nlist))
    oidx = orc.build_index(x[:3000], ids[:3000], cent, "IP", assignment=assign[:3000])
    rd, ri = orc.search(oidx, q, 10, 3)
    for mode in (1, 2):
        g.set_param("scan_mode", mode)
        gd, gi = g.search(q, 10, nprobe=3)
        assert_topk_parity(gd, gi, rd, ri, f"after a failed chunk, mode {mode}")
    bad = assign[3000:].copy()
    bad[1600] = 99  # second chunk of this call
    with pytest.raises(sb.NativeError, match="list id outside"):
        g.add(x[3000:], ids[3000:], lists=bad)
    assert g.ntotal == 4500
    np.testing.assert_array_equal(g.list_sizes(), np.bincount(assign[:4500], minlength=nlist))
    tags = np.zeros(1500, dtype=np.uint32)
    tags[7] = 1 << 23  # does not fit the 23-bit repo field: must be refused, not aliased to repo 0
    with pytest.raises(sb.NativeError, match="repo tag"):
        g.add(x[4500:], ids[4500:], repo_tags=tags, lists=assign[4500:])
    assert g.ntotal == 4500
    g.add(x[4500:], ids[4500:], lists=assign[4500:])
    oidx = orc.build_index(x, ids, cent, "IP", assignment=assign)
    rd, ri = orc.search(oidx, q, 10, 3)
    for mode in (1, 2):
        g.set_param("scan_mode", mode)
        gd, gi = g.search(q, 10, nprobe=3)
        assert_topk_parity(gd, gi, rd, ri, f"after the retries, mode {mode}")
    sums, counts, obj = g.kmeans_buffers()
    with pytest.raises(sb.NativeError, match="non-empty"):
        g.kmeans_update(sums, counts)  # would re-centre lists that were assigned under the old centroids


# ---- compaction + bulk export ----------------------------------------------------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_compaction_keeps_results_and_recycles_pages(sb, orc, metric):
    rng = np.random.default_rng(17)
    n, d, nlist, nq = 20000, 96, 12, 60
    x, q, cent, ids = make_case(orc, n, d, nlist, nq, metric, seed=11)
    repo = rng.integers(0, 5, n).astype(np.uint32)
    lang = rng.integers(0, 3, n).astype(np.uint8)
    g, oidx, assign = build_pair(sb, orc, x, ids, cent, metric, repo, lang)
    gone = rng.random(n) < 0.4
    gone[assign == 3] = True  # one list loses every row
    assert g.remove_ids(ids[gone]) == int(gone.sum())
    before = {}
    for mode in (1, 2):
        g.set_param("scan_mode", mode)
        before[mode] = (g.search(q, 10, nprobe=5), g.search(q, 10, nprobe=5, repos=[1, 3], langs=[0, 2]))
    pages0 = g.stats().npages
    freed = g.compact()
    st = g.stats()
    assert freed > 0 and st.nfree_pages == freed and st.nremoved == 0 and st.ntotal == n - int(gone.sum()) and st.npages == pages0
    np.testing.assert_array_equal(g.list_sizes(), np.bincount(assign[~gone], minlength=nlist))
    live = orc.build_index(x[~gone], ids[~gone], cent, metric, repo[~gone], lang[~gone], assignment=assign[~gone])
    rd, ri = orc.search(live, q, 10, 5)
    fd, fi = orc.search(live, q, 10, 5, mask=orc.row_mask(live, repos=[1, 3], langs=[0, 2]))
    for mode in (1, 2):
        g.set_param("scan_mode", mode)
        (d0, i0), (d1, i1) = g.search(q, 10, nprobe=5), g.search(q, 10, nprobe=5, repos=[1, 3], langs=[0, 2])
        np.testing.assert_array_equal(i0, before[mode][0][1])  # slot order inside a list is kept: even ties stay put
        np.testing.assert_array_equal(d0, before[mode][0][0])
        np.testing.assert_array_equal(i1, before[mode][1][1])
        assert_topk_parity(d0, i0, rd, ri, f"compacted vs oracle rebuilt from the live rows, mode {mode} {metric}")
        assert_topk_parity(d1, i1, fd, fi, f"compacted + filter vs oracle, mode {mode} {metric}")
    assert g.compact() == 0  # nothing left to drop
    # new rows take the freed pages before the pool grows
    m = 3000
    xn = unit_rows(rng, m, d)
    idn = np.arange(m, dtype=np.int64) + 10**7
    an = orc.assign(xn, cent, metric)
    g.add(xn, idn, repo[:m], lang[:m], lists=an)
    st2 = g.stats()
    assert st2.npages == pages0 and st2.nfree_pages < freed and st2.ntotal == st.ntotal + m
    both = orc.build_index(np.concatenate([x[~gone], xn]), np.concatenate([ids[~gone], idn]), cent, metric,
                           np.concatenate([repo[~gone], repo[:m]]), np.concatenate([lang[~gone], lang[:m]]),
                           assignment=np.concatenate([assign[~gone], an]))
    rd, ri = orc.search(both, q, 10, 5)
    for mode in (1, 2):
        g.set_param("scan_mode", mode)
        d0, i0 = g.search(q, 10, nprobe=5)
        assert_topk_parity(d0, i0, rd, ri, f"after re-filling the freed pages, mode {mode} {metric}")
    # bulk export == per-list export, host and device
    off, vecs, eids, tags = g.export_lists(2, 9)
    _, dv, di, dt = g.export_lists(2, 9, device=True)
    np.testing.assert_array_equal(dv.cpu().numpy(), vecs)
    np.testing.assert_array_equal(di.cpu().numpy(), eids)
    np.testing.assert_array_equal(dt.cpu().numpy().view(np.uint32), tags)
    for l in range(2, 9):
        v1, i1, t1 = g.export_list(l)
        a, b = int(off[l - 2]), int(off[l - 1])
        np.testing.assert_array_equal(vecs[a:b], v1)
        np.testing.assert_array_equal(eids[a:b], i1)
        np.testing.assert_array_equal(tags[a:b], t1)
    coff, cv, ci, ct = g.export_csr()
    assert coff[-1] == st2.ntotal and set(ci.tolist()) == set(ids[~gone].tolist()) | set(idn.tolist())


def test_exchange_timeout_is_surfaced(sb, orc):
    """a peer that never arrives must not yield a silently wrong result.  Rank 0 of a 2-rank exchange whose peer
    never calls: the waiting kernels give up after the (shortened) timeout and set the pinned status word; poll() sees it
    without a sync, status() after one, and the next step is refused."""
    import ctypes as C

    import torch

    from semcode_b200 import _capi

    x, q, cent, ids = make_case(orc, 3000, 64, 8, 20, "IP", seed=9)
    g, _, _ = build_pair(sb, orc, x, ids, cent, "IP")
    nbytes = 4 << 20
    mine = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    peer = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")  # stands in for the absent rank 1
    arr = (C.c_void_p * 2)(mine.data_ptr(), peer.data_ptr())
    h = C.c_void_p()
    L = _capi.lib()
    _capi.check(L.sc_exchange_create(0, 2, arr, nbytes, 0, C.byref(h)))

    class Ex:
        handle = h

    try:
        _capi.check(L.sc_exchange_set_timeout_ms(h, 30))
        t = C.c_int32(0)
        _capi.check(L.sc_exchange_poll(h, C.byref(t)))
        assert t.value == 0
        g.search(torch.from_numpy(q).cuda(), 5, nprobe=4, exchange=Ex)  # issues; the waits time out on the device
        torch.cuda.synchronize()
        _capi.check(L.sc_exchange_poll(h, C.byref(t)))
        assert t.value == 1
        e = C.c_int64(0)
        _capi.check(L.sc_exchange_status(h, C.byref(t), C.byref(e)))
        assert t.value == 1 and e.value == 1
        with pytest.raises(_capi.NativeError, match="unusable"):
            g.search(torch.from_numpy(q).cuda(), 5, nprobe=4, exchange=Ex)
        d0, i0 = g.search(q, 5, nprobe=4)  # the index itself is fine
        rd, ri = orc.search(orc.build_index(x, ids, cent, "IP"), q, 5, 4)
        assert_topk_parity(d0, i0, rd, ri, "plain search after a timed-out exchange step")
    finally:
        L.sc_exchange_destroy(h)


# ---- the headline regime: nlist 16384, thousands of queries, every list-major consumer at once -------------------
@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_headline_regime_list_major_against_oracle(sb, orc_c, metric):
    """BASELINE.json configs[1]'s shape at a quarter of its rows: 2.5M x 768, nlist 16384, nq 2048, nprobe 32, top-10 on a
    Zipf-clustered set, so one batch holds lists probed once, a few times (scan_mq<4>, scan_mq<8>) and hundreds of times
    (tile items), and the one-pass top-k reads candidates from all of them.  The whole batch is compared with the
    query-major kernel (ids identical), a random slice of it with the C oracle on the same lists."""
    import torch

    n, d, nlist, nq, nprobe, k = 2_500_000, 768, 16384, 2048, 32, 10
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(2024)
    centres = torch.randn((2048, d), generator=gen, device=dev)
    w = 1.0 / torch.arange(1, 2049, device=dev, dtype=torch.float32) ** 1.1

    def rows(m):
        c = torch.multinomial(w / w.sum(), m, replacement=True, generator=gen)
        return torch.nn.functional.normalize(centres[c] + 0.3 * torch.randn((m, d), generator=gen, device=dev), dim=1)

    g = sb.IVFFlatIndex(d, nlist=nlist, metric=metric)
    g.train(rows(300_000), niter=2, max_points_per_centroid=0)
    for s in range(0, n, 500_000):
        g.add(rows(500_000), torch.arange(s, s + 500_000, device=dev, dtype=torch.int64) * 3 + 1)
    assert g.ntotal == n
    q = rows(nq)

    g.set_profiling(True)
    d0, i0 = g.search(q, k, nprobe=nprobe)  # automatic switch: nq * nprobe = 4 * nlist -> list-major
    torch.cuda.synchronize()
    t = g.last_search_times()
    assert t.unique_rows > 0 and t.scanned_rows > 2 * t.unique_rows, "the batch must have gone list-major with real sharing"
    assert t.scan_launches >= 6, t.scan_launches  # count, plan, fill, tile items (+ split) and both multi-query page scans
    g.set_profiling(False)
    g.set_param("scan_mode", 1)
    d1, i1 = g.search(q, k, nprobe=nprobe)
    g.set_param("scan_mode", 0)
    d0, i0, d1, i1 = (a.cpu().numpy() for a in (d0, i0, d1, i1))
    assert_sorted(d0, i0, metric == "IP")
    assert_topk_parity(d0, i0, d1, i1, "list-major vs query-major, whole batch")
    # the clustered set is dense near the top: a handful of 1e-5-relative near-ties order differently under the tile
    # kernel's 3xTF32 sums and the page scan's FFMA sums (assert_topk_parity has just checked that is all they are)
    assert (i0 == i1).mean() > 0.998

    # C oracle on a random slice of the SAME output, over the lists those queries probe
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(nq, 128, replace=False))
    qs = q[torch.from_numpy(pick).to(dev)].cpu().numpy()
    probes = g.probe(qs, nprobe)
    used = np.unique(probes)
    remap = -np.ones(nlist, dtype=np.int32)
    remap[used] = np.arange(used.size, dtype=np.int32)
    parts = [g.export_list(int(l)) for l in used]
    off = np.zeros(used.size + 1, dtype=np.int64)
    np.cumsum([p[0].shape[0] for p in parts], out=off[1:])
    vecs = np.concatenate([p[0] for p in parts])
    ids = np.concatenate([p[1] for p in parts])
    od, oi = orc_c.scan_search(qs, g.metric, remap[probes], off, vecs, ids, k)
    assert_topk_parity(d0[pick], i0[pick], od, oi, "list-major vs C oracle")
    if metric == "IP":
        # the option that sends remainders of 5..16 queries to the tensor-core tiles as well: same results
        g.set_param("tile_rem", 4)
        d2, i2 = g.search(q, k, nprobe=nprobe)
        g.set_param("tile_rem", 0)
        assert_topk_parity(d2.cpu().numpy(), i2.cpu().numpy(), d1, i1, "tile_rem = 4 vs query-major")
    g.close()


def test_two_searches_in_flight_on_one_handle(sb, orc):
    """Two host threads, each on its own CUDA stream, search the same handle at once (the handle keeps two scratch slots;
    searches share the lock, writers take it alone) while a third thread inserts rows.  Every result equals the result of
    the same search made alone, and the inserted rows are all there afterwards."""
    import threading

    import torch

    x, q, cent, ids = make_case(orc, 60000, 256, 64, 600, "IP", seed=21)
    g, _, _ = build_pair(sb, orc, x[:50000], ids[:50000], cent, "IP")
    dev = torch.device("cuda", 0)
    qd = torch.from_numpy(q).to(dev)
    batches = [qd[i * 100:(i + 1) * 100].contiguous() for i in range(6)]
    # (list-major for the 100-query batches at nprobe 12, query-major at nprobe 2: both routes run concurrently)
    want = {(b, npb): tuple(t.cpu().numpy() for t in g.search(batches[b], 10, nprobe=npb)) for b in range(6) for npb in (2, 12)}
    errors, results = [], {}

    def searcher(tid):
        try:
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                for rep in range(8):
                    for b in range(tid, 6, 2):
                        for npb in (2, 12):
                            d, i = g.search(batches[b], 10, nprobe=npb)
                            results[(tid, rep, b, npb)] = (d, i)
            st.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def writer():
        try:
            far = np.zeros((1, 256), dtype=np.float32)
            far[0, 0] = -1.0  # rows nobody retrieves: the searches' expected results do not change
            for j in range(20):
                g.add(np.repeat(far, 50, axis=0), np.arange(10**9 + 50 * j, 10**9 + 50 * (j + 1), dtype=np.int64))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=searcher, args=(0,)), threading.Thread(target=searcher, args=(1,)), threading.Thread(target=writer)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    torch.cuda.synchronize()
    assert g.ntotal == 50000 + 1000
    for (tid, rep, b, npb), (d, i) in results.items():
        wd, wi = want[(b, npb)]
        np.testing.assert_array_equal(i.cpu().numpy(), wi, err_msg=f"thread {tid} rep {rep} batch {b} nprobe {npb}")
        np.testing.assert_array_equal(d.cpu().numpy(), wd)
    g.close()


def test_scratch_guards_stay_intact(sb, orc):
    """Stand-in for compute-sanitizer memcheck (closed on the GPU pool, see profiles/): with `debug_canary` every scratch buffer
    sits between two 256-byte guards and is allocated without slack; after a tour through every scan route, the pair plans, both
    selection paths, filters, removal and insertion no guard byte may have changed."""
    rng = np.random.default_rng(8)
    x, q, cent, ids = make_case(orc, 40000, 768, 48, 700, "IP", seed=31)
    repo = rng.integers(0, 20, x.shape[0]).astype(np.uint32)
    lang = rng.integers(0, 3, x.shape[0]).astype(np.uint8)
    g = sb.IVFFlatIndex(768, nlist=48, metric="IP")
    try:
        g.set_param("debug_canary", 1)
        g.set_centroids(cent)
        g.add(x[:30000], ids[:30000], repo[:30000], lang[:30000])
        want = g.search(q[:64], 10, nprobe=6)
        for nq, k, nprobe, kw in ((1, 10, 6, {}), (5, 10, 6, {}), (64, 10, 6, {}), (700, 10, 6, {}), (700, 10, 30, {}), (300, 200, 48, {}),
                                  (700, 50, 12, {"repos": [1, 2, 3], "langs": [1]}), (16, 7, 3, {"langs": [0, 2]}), (129, 1, 1, {})):
            g.search(q[:nq], k, nprobe=nprobe, **kw)
            g.set_param("check_canaries", 0)
        for mode in (1, 2):
            g.set_param("scan_mode", mode)
            g.search(q[:300], 10, nprobe=9)
        g.set_param("scan_mode", 0)
        for cfg in (1, 3, 5, 0):
            g.set_param("lists_cfg", cfg)
            g.search(q, 10, nprobe=24)
        g.remove_ids(ids[:5000])
        g.add(x[30000:], ids[30000:], repo[30000:], lang[30000:])
        g.compact()
        g.assign(x[:3000])
        g.probe(q[:100], 9)
        g.search(q, 33, nprobe=17)
        g.set_param("check_canaries", 8)  # at least 8 guarded buffers were in play
        # and the guarded run computes what the unguarded one does
        g.set_param("lists_cfg", 0)
        g2 = sb.IVFFlatIndex(768, nlist=48, metric="IP")
        g.set_param("debug_canary", 0)
        g2.set_centroids(cent)
        g2.add(x[:30000], ids[:30000], repo[:30000], lang[:30000])
        d2, i2 = g2.search(q[:64], 10, nprobe=6)
        np.testing.assert_array_equal(want[1], i2)
        np.testing.assert_array_equal(want[0], d2)
        g2.close()
    finally:
        g.set_param("debug_canary", 0)
        g.close()
