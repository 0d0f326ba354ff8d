"""GPU tests of the drop-in wrapper: the call sequences of IndexerService (services/indexer.py:103-120)
and SemanticSearchPipeline (rag/pipeline.py:93-169) against MilvusVectorStore, plus golden fixtures."""

import os
from dataclasses import dataclass
from typing import List

import numpy as np
import pytest

from helpers import assert_topk_parity, unit_rows

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@dataclass
class EmbeddingPayload:  # reference src/semcode/embeddings/providers.py:21-28
    id: str
    text: str
    vector: List[float]
    metadata: dict


def payloads(x, repo_of, lang_of, prefix="c"):
    out = []
    for i, v in enumerate(x):
        meta = {"repo": repo_of(i), "path": f"src/f{i}.py", "language": lang_of(i), "start_line": 1, "end_line": 9}
        out.append(EmbeddingPayload(id=f"{prefix}{i:06d}", text=f"chunk {i}", vector=[float(t) for t in v], metadata=meta))
    return out


@pytest.fixture()
def store_mod(native_lib):
    from semcode_b200 import storage

    yield storage
    for name in ("t_small", "t_seal", "t_share"):
        storage.drop_collection(name)


def test_upsert_and_search_like_the_callers(store_mod):
    rng = np.random.default_rng(0)
    x = unit_rows(rng, 300, 64)
    store = store_mod.MilvusVectorStore("t_small", dim=64)
    store.connect()
    store.connect()  # idempotent, attaches to the existing collection
    calls = []
    store.upsert_embeddings(payloads(x, lambda i: "demo", lambda i: "python"), progress=lambda a, b: calls.append((a, b)))
    assert calls[0] == (0, 300) and calls[-1] == (300, 300) and [c[0] for c in calls[1:]] == [128, 256, 300]
    calls.clear()
    store.upsert_embeddings([], progress=lambda a, b: calls.append((a, b)))
    assert calls == [(0, 0)]
    # the RAG retriever's unwrap (rag/pipeline.py:113-169)
    results = store.search([float(t) for t in x[17]], top_k=5)
    assert results
    hits = next(iter(results))
    assert len(hits) == 5
    h = hits[0]
    assert h.id == "c000017" and abs(h.distance - 1.0) < 1e-5 and h.score == h.distance
    assert h.entity.get("repo") == "demo" and h.entity.get("path") == "src/f17.py"
    assert h.entity.get("language") == "python" and h.entity.get("text") == "chunk 17"
    assert h.entity.get("metadata")["end_line"] == 9
    ds = [t.distance for t in hits]
    assert ds == sorted(ds, reverse=True)
    # small collection -> exact (growing segment): equals brute force
    want = np.argsort(-(x @ x[17]))[:5]
    assert [t.id for t in hits] == [f"c{j:06d}" for j in want]
    # a second store object in the same process sees the same collection (indexer + pipeline singletons)
    other = store_mod.MilvusVectorStore("t_small", dim=64)
    other.connect()
    assert next(iter(other.search([float(t) for t in x[3]], top_k=1)))[0].id == "c000003"
    with pytest.raises(ValueError):
        store_mod.MilvusVectorStore("t_small", dim=32).connect()
    with pytest.raises(ValueError):
        store.search([0.0] * 63)


def test_upsert_replaces_by_primary_key(store_mod):
    rng = np.random.default_rng(1)
    x = unit_rows(rng, 50, 32)
    store = store_mod.MilvusVectorStore("t_small", dim=32)
    store.connect()
    store.upsert_embeddings(payloads(x, lambda i: "a", lambda i: "cpp"))
    newv = unit_rows(rng, 1, 32)[0]
    p = EmbeddingPayload(id="c000007", text="new text", vector=newv.tolist(), metadata={"repo": "b"})
    store.upsert_embeddings([p, p])  # duplicate in one batch: one row
    assert store._collection.num_entities == 50
    hits = store.search(newv.tolist(), top_k=3)[0]
    assert hits[0].id == "c000007" and hits[0].entity.get("text") == "new text" and hits[0].entity.get("repo") == "b"
    assert hits[0].entity.get("language") == ""  # missing metadata keys -> "" (milvus_store.py:121-123)
    old = store.search(x[7].tolist(), top_k=50)[0]
    assert [h.id for h in old].count("c000007") == 1 and len(old) == 50


def test_seal_trains_ivf_and_filters_push_down(store_mod, monkeypatch):
    from oracle import ivf_numpy as orc

    rng = np.random.default_rng(2)
    n, d = 3000, 48
    x = unit_rows(rng, n, d)
    monkeypatch.setattr(store_mod.milvus_store.settings, "ivf_nlist", 16, raising=False)
    monkeypatch.setattr(store_mod.milvus_store.settings, "ivf_train_niter", 4, raising=False)
    store = store_mod.MilvusVectorStore("t_seal", dim=d)
    store.connect()
    col = store._collection
    assert col.nlist == 16 and col.seal_rows == 39 * 16
    ids = [f"r{i}" for i in range(n)]
    repos = [f"repo{i % 7}" for i in range(n)]
    langs = ["python" if i % 3 else "cpp" for i in range(n)]
    store.upsert_arrays(ids[:500], x[:500], repos=repos[:500], languages=langs[:500])
    assert col.index is None  # still growing: exact
    store.upsert_arrays(ids[500:], x[500:], repos=repos[500:], languages=langs[500:], texts=[f"t{i}" for i in range(500, n)])
    ivf = col.index
    assert ivf is not None and ivf.nlist == 16 and ivf.ntotal == n and col._growing_rows == 0
    # same centroids / lists in the oracle -> identical results at nprobe 4
    off, vecs, rid, tags = ivf.export_csr()
    oidx = orc.OracleIndex(0, ivf.get_centroids(), off, vecs, rid, (tags >> 8).astype(np.uint32), (tags & 0xFF).astype(np.uint8))
    q = unit_rows(rng, 20, d)
    probes = ivf.probe(q, 4)
    rd, ri = orc.search(oidx, q, 10, 4, probes=probes)
    gd, gi = store.search_arrays(q, top_k=10, nprobe=4)
    assert_topk_parity(gd, gi, rd, ri, "store vs oracle")
    res = store.search_batch(q, top_k=10, nprobe=4)
    assert [[h.id for h in hits] for hits in res] == [[f"r{j}" for j in row if j >= 0] for row in ri]
    # filter pushdown: every hit satisfies the predicate and k survivors come back (filter-then-rank)
    res = store.search(q[0].tolist(), top_k=10, nprobe=16, repos=["repo3"], languages=["python"])
    hits = res[0]
    assert len(hits) == 10 and all(h.entity.get("repo") == "repo3" and h.entity.get("language") == "python" for h in hits)
    mask = orc.row_mask(oidx, repos=[col._repo_vocab["repo3"]], langs=[col._lang_vocab["python"]])
    rd, ri = orc.search(oidx, q[:1], 10, 16, mask=mask)
    assert [h.id for h in hits] == [f"r{j}" for j in ri[0]]
    assert store.search(q[0].tolist(), top_k=5, repos=["nope"]) == [[]]
    # rows added after the seal go to the inverted lists and are found; replaced rows disappear
    v = unit_rows(rng, 1, d)[0]
    store.upsert_arrays(["r5", "fresh"], np.stack([v, -v]), repos=["repo0", "repo0"], languages=["go", "go"])
    assert ivf.ntotal == n + 1 and col.num_entities == n + 1
    top = store.search(v.tolist(), top_k=1, nprobe=16)[0][0]
    assert top.id == "r5" and top.entity.get("language") == "go"


def test_golden_fixtures_through_the_c_abi(native_lib):
    import semcode_b200 as sb
    from oracle import ivf_numpy as orc

    z = np.load(os.path.join(GOLD, "kat_small.npz"))
    for metric in ("IP", "L2"):
        cent = z[f"cent_{metric}"]
        g = sb.IVFFlatIndex(24, nlist=12, metric=metric)
        g.set_centroids(cent)
        g.add(z["x"], z["ids"], z["repo"], z["lang"], lists=orc.assign(z["x"], cent, metric))
        np.testing.assert_array_equal(g.probe(z["q"], 4), z[f"probes_{metric}"])
        d, i = g.search(z["q"], 10, lists=z[f"probes_{metric}"])
        assert_topk_parity(d, i, z[f"dist_{metric}"], z[f"ids_{metric}"], f"golden {metric}")
        d, i = g.search(z["q"], 10, lists=z[f"probes_{metric}"], repos=[1, 2, 3], langs=[1])
        assert_topk_parity(d, i, z[f"fdist_{metric}"], z[f"fids_{metric}"], f"golden filtered {metric}")
    # BASELINE.json configs[0] shape: 100k x 768, nlist 1024, nprobe 16, top-10
    c1 = np.load(os.path.join(GOLD, "kat_c1.npz"))
    n, dim, nlist = 100_000, 768, 1024
    x = unit_rows(np.random.default_rng(1234), n, dim)
    q = unit_rows(np.random.default_rng(4321), 64, dim)
    cent = x[orc.kmeans_init_rows(n, nlist, 1234)].copy()
    g = sb.IVFFlatIndex(dim, nlist=nlist, metric="IP")
    g.set_centroids(cent)
    g.add(x, np.arange(n, dtype=np.int64))  # GPU assigns the lists
    sizes = g.list_sizes()
    assert int(np.abs(sizes - c1["list_sizes"]).sum()) <= 20  # fp32 near-ties in the assignment only
    got = g.probe(q, 16)
    same = np.array([sorted(a) == sorted(b) for a, b in zip(got, c1["probes"])])
    assert same.mean() >= 0.95
    d, i = g.search(q, 10, lists=c1["probes"])
    moved = np.abs(sizes - c1["list_sizes"]).sum() > 0
    if not moved:
        assert_topk_parity(d, i, c1["dist"], c1["ids"], "golden C1")
    else:  # a row that changed list may enter/leave a result; everything else is identical
        agree = np.mean([len(np.intersect1d(a, b)) / 10 for a, b in zip(i, c1["ids"])])
        assert agree > 0.995
    d2, i2 = g.search(q, 10, nprobe=16)
    assert_topk_parity(d2[same], i2[same], d[same], i[same], "full path vs preassigned")


def test_snapshot_round_trip(store_mod, monkeypatch, tmp_path):
    """connect() on an existing collection loads it (milvus_store.py:51-54) -- here from a snapshot dir."""
    rng = np.random.default_rng(5)
    n, d = 1500, 40
    x = unit_rows(rng, n, d)
    monkeypatch.setattr(store_mod.milvus_store.settings, "ivf_nlist", 8, raising=False)
    monkeypatch.setattr(store_mod.milvus_store.settings, "ivf_train_niter", 3, raising=False)
    monkeypatch.setattr(store_mod.milvus_store.settings, "ivf_persist_dir", str(tmp_path), raising=False)
    store = store_mod.MilvusVectorStore("t_share", dim=d)
    store.connect()
    ids = [f"k{i}" for i in range(n)]
    store.upsert_arrays(ids[:1000], x[:1000], repos=["a"] * 1000, languages=["python"] * 1000,
                        texts=[f"t{i}" for i in range(1000)], metadata=[{"i": i} for i in range(1000)])
    assert store._collection.index is not None  # sealed at 39 * 8 rows
    store.upsert_arrays(ids[1000:], x[1000:], repos=["b"] * 500, languages=["cpp"] * 500)
    store.upsert_arrays(["k3"], x[3:4] * -1.0, repos=["a"], languages=["python"])  # a replaced row
    q = unit_rows(rng, 12, d)
    before = store.search_batch(q, top_k=7, nprobe=8)
    fb = store.search(q[0].tolist(), top_k=5, nprobe=8, repos=["b"])
    store.flush()
    store_mod.drop_collection("t_share")
    again = store_mod.MilvusVectorStore("t_share", dim=d)
    again.connect()
    col = again._collection
    assert col.num_entities == n and col.index is not None and col.index.ntotal == n
    after = again.search_batch(q, top_k=7, nprobe=8)
    assert [[h.id for h in hits] for hits in after] == [[h.id for h in hits] for hits in before]
    np.testing.assert_allclose([[h.distance for h in hits] for hits in after], [[h.distance for h in hits] for hits in before], rtol=1e-6)
    assert after[0][0].entity.get("metadata") == before[0][0].entity.get("metadata")
    fa = again.search(q[0].tolist(), top_k=5, nprobe=8, repos=["b"])
    assert [h.id for h in fa[0]] == [h.id for h in fb[0]] and all(h.entity.get("repo") == "b" for h in fa[0])
    # upserts keep working on the reloaded collection (primary-key map restored)
    again.upsert_arrays(["k5"], x[5:6], repos=["a"], languages=["python"])
    assert col.num_entities == n


# ---- the host-logic scenarios of tests/test_store_host.py, with the real engine ------------------------------------
@pytest.fixture()
def real_store(native_lib):
    import semcode_b200.storage.milvus_store as ms

    yield ms
    for name in list(ms._REGISTRY):
        ms.drop_collection(name)


def test_persist_without_flush_on_the_gpu(real_store, monkeypatch, tmp_path):
    from test_store_host import scenario_persist_without_flush

    scenario_persist_without_flush(real_store, monkeypatch, tmp_path)


def test_journal_on_the_gpu(real_store, monkeypatch, tmp_path):
    from test_store_host import scenario_journal

    scenario_journal(real_store, monkeypatch, tmp_path)


def test_compaction_and_retrain_on_the_gpu(real_store, monkeypatch):
    from test_store_host import scenario_compaction_and_retrain

    scenario_compaction_and_retrain(real_store, monkeypatch)


def test_concurrent_searches_on_the_gpu(real_store, monkeypatch):
    from test_store_host import scenario_concurrent_searches

    scenario_concurrent_searches(real_store, monkeypatch)


def test_reference_caller_traffic_replayed_on_the_gpu(real_store):
    """tests/golden/callers.json was recorded in the build container by running the reference's OWN IndexerService and
    SemanticSearchPipeline (unmodified, /root/reference) against the drop-in store with the oracle-backed engine double
    (tests/golden/make_callers_golden.py).  Here the same payloads go through the same wrapper with the CUDA engine; the
    documents the reference's `_hit_to_document` built there must be the hits returned here."""
    import json

    with open(os.path.join(GOLD, "callers.json")) as f:
        gold = json.load(f)
    store = real_store.MilvusVectorStore("callers_replay", dim=gold["dim"])
    store.connect()
    store.upsert_embeddings([EmbeddingPayload(p["id"], p["text"], p["vector"], p["metadata"]) for p in gold["payloads"]])
    assert store._collection.num_entities == len(gold["payloads"])
    for case in gold["queries"]:
        hits = next(iter(store.search(case["vector"], top_k=case["top_k"])))
        docs = case["documents"]
        assert len(hits) == len(docs)
        for h, d in zip(hits, docs):
            assert abs(h.distance - d["score"]) <= 1e-5 * abs(d["score"]) + 1e-6
        # ids identical except inside exact-score ties
        got, want = [h.entity.get("text") for h in hits], [d["snippet"] for d in docs]
        for j, (g, w) in enumerate(zip(got, want)):
            if g != w:
                assert abs(docs[j]["score"] - hits[j].distance) <= 1e-6 and g in want, (j, g[:40], w[:40])
        assert hits[0].entity.get("repo") == docs[0]["repo"] and hits[0].entity.get("metadata") == docs[0]["metadata"]


# ---- one collection over several devices, one process (ivf_devices) ----------------------------------------------
def test_multi_device_collection_behind_the_wrapper(real_store, monkeypatch, tmp_path):
    """`ivf_devices` gives the collection a row-sharded backend (semcode_b200/multidevice.py): upsert deals the rows, seal
    trains data-parallel, search reduces the shards' partial top-k on the first device, flush / connect persist and reload
    the shards.  With one visible GPU both shards live on cuda:0 -- same code path, same checks."""
    import torch

    ms = real_store
    ngpu = torch.cuda.device_count()
    devs = "0,1" if ngpu >= 2 else "0,0"
    print(f"multi-device collection on devices {devs} ({ngpu} visible GPU(s))")
    rng = np.random.default_rng(5)
    n, d = 6000, 96
    x = unit_rows(rng, n, d)
    repos = ["alpha" if i % 4 else "beta" for i in range(n)]
    langs = ["python" if i % 3 else "cpp" for i in range(n)]
    ids = [f"k{i:06d}" for i in range(n)]
    monkeypatch.setenv("SEMCODE_IVF_NLIST", "16")
    monkeypatch.setenv("SEMCODE_IVF_SEAL_ROWS", "100000")  # sealed explicitly below, with given centroids
    single = ms.MilvusVectorStore("md_single", dim=d)
    single.connect()
    monkeypatch.setenv("SEMCODE_IVF_DEVICES", devs)
    monkeypatch.setenv("SEMCODE_IVF_PERSIST_DIR", str(tmp_path / "p"))
    multi = ms.MilvusVectorStore("md_multi", dim=d)
    multi.connect()
    assert multi._collection.devices == [int(v) for v in devs.split(",")]
    cent = x[rng.choice(n, 16, replace=False)].copy()
    for st in (single, multi):
        for a in range(0, n, 1700):  # uneven batches exercise the global round-robin cursor
            st.upsert_arrays(ids[a:a + 1700], x[a:a + 1700], repos=repos[a:a + 1700], languages=langs[a:a + 1700])
        st.build_index(centroids=cent)
    from semcode_b200.multidevice import MultiDeviceIVFFlat

    mi = multi._collection.index
    assert isinstance(mi, MultiDeviceIVFFlat) and mi.ntotal == n and abs(mi.shards[0].ntotal - mi.shards[1].ntotal) <= 1
    q = unit_rows(rng, 70, d)
    for kw in ({}, {"repos": ["beta"]}, {"languages": ["cpp"], "repos": ["alpha"]}):
        d1, r1 = single.search_arrays(q, 10, nprobe=5, **kw)
        d2, r2 = multi.search_arrays(q, 10, nprobe=5, **kw)
        assert_topk_parity(d2, r2, d1, r1, f"two shards vs one index {kw}")
    # later rows go straight to the shards; replaced keys are tombstoned on whichever shard holds them
    extra = unit_rows(rng, 501, d)
    for st in (single, multi):
        st.upsert_arrays(ids[:501], extra, repos=repos[:501], languages=langs[:501])
    d1, r1 = single.search_arrays(extra[:40], 5, nprobe=16)
    d2, r2 = multi.search_arrays(extra[:40], 5, nprobe=16)
    assert_topk_parity(d2, r2, d1, r1, "after an upsert-replace")
    # device tensors in -> device tensors out
    dq = torch.from_numpy(q).cuda()
    d3, r3 = multi.search_arrays(dq, 10, nprobe=5)
    d1, r1 = single.search_arrays(q, 10, nprobe=5)
    assert d3.is_cuda and r3.is_cuda
    assert_topk_parity(d3.cpu().numpy(), r3.cpu().numpy(), d1, r1, "device tensors")
    # persisted by the upserts above (no flush() call); a new 'process' reloads both shards
    hits_before = multi.search(q[3].tolist(), top_k=4, nprobe=5)[0]
    multi.flush()
    ms._REGISTRY.pop("md_multi").close()
    again = ms.MilvusVectorStore("md_multi", dim=d)
    again.connect()
    assert isinstance(again._collection.index, MultiDeviceIVFFlat) and again._collection.num_entities == n
    hits_after = again.search(q[3].tolist(), top_k=4, nprobe=5)[0]
    assert [h.id for h in hits_after] == [h.id for h in hits_before]
    # data-parallel k-means over the shards tracks the single-device Lloyd iterations
    one = ms.IVFFlatIndex(d, nlist=16, metric="IP")
    two = MultiDeviceIVFFlat(d, nlist=16, metric="IP", devices=[int(v) for v in devs.split(",")])
    o1 = one.train(x, niter=4, seed=9, max_points_per_centroid=0)
    o2 = two.train(x, niter=4, seed=9, max_points_per_centroid=0)
    np.testing.assert_allclose(o2, o1, rtol=1e-6)
    np.testing.assert_allclose(two.get_centroids(), one.get_centroids(), rtol=1e-4, atol=1e-6)


def test_c3_shaped_collection_on_every_visible_gpu(real_store, monkeypatch, tmp_path):
    """BASELINE.json configs[2] (3072-d rows that do not fit one card) through the reference-facing wrapper: the collection is
    row-sharded over EVERY visible GPU (`ivf_devices`), filled by upserts, sealed (data-parallel k-means), persisted as a
    snapshot, re-opened as a new process would, and searched.  Rows scale with the box -- 125k x 3072 per GPU: 1M rows /
    12.3 GB on 8 GPUs (the full 10M-row C3 runs in bench.py's `c3` key at N = 8); on one GPU two shards share the card.
    Checked: exhaustive probing equals exact search computed by torch in fp64 on the regenerated rows, the nprobe-32 result
    has the recall the coarse quantizer allows, and the reloaded collection answers exactly like the one that was saved."""
    import time

    import torch

    ms = real_store
    ngpu = torch.cuda.device_count()
    devices = list(range(ngpu)) if ngpu >= 2 else [0, 0]
    d, per_gpu, chunk = 3072, 125_000, 25_000
    n = per_gpu * max(ngpu, 2) if ngpu >= 2 else 200_000
    nlist = 16384 if n >= 39 * 16384 else 2048
    monkeypatch.setenv("SEMCODE_IVF_NLIST", str(nlist))
    monkeypatch.setenv("SEMCODE_IVF_SEAL_ROWS", str(10 * n))  # sealed explicitly below
    monkeypatch.setenv("SEMCODE_IVF_DEVICES", ",".join(str(v) for v in devices))
    monkeypatch.setenv("SEMCODE_IVF_PERSIST_DIR", str(tmp_path / "c3"))
    dev0 = torch.device("cuda", devices[0])

    def rows(a, b):  # seeded per chunk: regenerated for the exact check instead of kept
        g = torch.Generator(device=dev0).manual_seed(1000 + a)
        centres = torch.randn((512, d), generator=torch.Generator(device=dev0).manual_seed(7), device=dev0)
        c = torch.randint(0, 512, (b - a,), generator=g, device=dev0)
        return torch.nn.functional.normalize(centres[c] + 0.35 * torch.randn((b - a, d), generator=g, device=dev0), dim=1)

    st = ms.MilvusVectorStore("c3_shape", dim=d)
    st.connect()
    t0 = time.time()
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        st.upsert_arrays([f"c{i:08d}" for i in range(a, b)], rows(a, b), repos=["r%d" % (i % 7) for i in range(a, b)],
                         languages=["python" if i % 3 else "cpp" for i in range(a, b)])
    t_insert = time.time() - t0
    t0 = time.time()
    st.build_index(niter=3)  # seal: k-means over the shards' devices, rows dealt to the shards, snapshot written
    t_seal = time.time() - t0
    col = st._collection
    from semcode_b200.multidevice import MultiDeviceIVFFlat

    assert isinstance(col.index, MultiDeviceIVFFlat) and len(col.index.shards) == len(devices) and col.index.ntotal == n
    sizes = [sh.ntotal for sh in col.index.shards]
    assert max(sizes) - min(sizes) <= 1
    if ngpu >= 2:
        assert sorted({sh.device for sh in col.index.shards}) == devices

    q = rows(n, n + 64)
    # exact search on the regenerated rows (fp64 on the first device)
    best_d = torch.full((64, 10), -2.0, dtype=torch.float64, device=dev0)
    best_i = torch.full((64, 10), -1, dtype=torch.int64, device=dev0)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        s = q.double() @ rows(a, b).double().T
        dd, ii = torch.cat([best_d, s], 1).topk(10, dim=1)
        cand = torch.cat([best_i, torch.arange(a, b, device=dev0).expand(64, -1)], 1)
        best_d, best_i = dd, torch.gather(cand, 1, ii)
    exact_keys = [[f"c{int(v):08d}" for v in r] for r in best_i.cpu().numpy()]
    de, re_ = st.search_arrays(q, 10, nprobe=nlist)
    got_keys = [[col._pk[int(v)] for v in r] for r in (re_.cpu().numpy() if torch.is_tensor(re_) else re_)]
    assert got_keys == exact_keys
    np.testing.assert_allclose(de.cpu().numpy() if torch.is_tensor(de) else de, best_d.cpu().numpy(), rtol=1e-5)
    t0 = time.time()
    d32, r32 = st.search_arrays(q, 10, nprobe=32)
    torch.cuda.synchronize()
    t_search = time.time() - t0
    r32 = r32.cpu().numpy() if torch.is_tensor(r32) else r32
    recall = np.mean([len(set(r32[i].tolist()) & {col._row_of[key] for key in exact_keys[i]}) / 10 for i in range(64)])
    assert recall > 0.5, recall
    # what a user sees: Hits with ids and scores
    hits = st.search(q[0].cpu().tolist(), top_k=5, nprobe=32)[0]
    assert [h.id for h in hits] == [col._pk[int(v)] for v in r32[0][:5]]

    # a new process: the snapshot build_index() published is all there is
    snap = col.snapshot_dir(col.persist_dir)
    assert snap and os.path.isdir(snap)
    ms._REGISTRY.pop("c3_shape").close()
    t0 = time.time()
    again = ms.MilvusVectorStore("c3_shape", dim=d)
    again.connect()
    t_load = time.time() - t0
    assert again._collection.num_entities == n and isinstance(again._collection.index, MultiDeviceIVFFlat)
    d2, r2 = again.search_arrays(q, 10, nprobe=32)
    np.testing.assert_array_equal(r2.cpu().numpy() if torch.is_tensor(r2) else r2, r32)
    print(f"C3-shaped collection: {n} x {d} on devices {devices}: upserts {t_insert:.1f} s, seal {t_seal:.1f} s, "
          f"search(64 x nprobe 32) {t_search * 1e3:.1f} ms, recall@10 {recall:.3f}, reload {t_load:.1f} s")


def test_search_stats_for_telemetry(real_store, monkeypatch):
    """`ivf_profile`: the wrapper reports what a recent search on the sealed index cost (per-phase CUDA events of the engine),
    which the optional caller patch records in the API's telemetry."""
    ms = real_store
    rng = np.random.default_rng(3)
    n, d = 5000, 64
    x = unit_rows(rng, n, d)
    monkeypatch.setenv("SEMCODE_IVF_NLIST", "16")
    monkeypatch.setenv("SEMCODE_IVF_SEAL_ROWS", "100000")
    monkeypatch.setenv("SEMCODE_IVF_PROFILE", "1")
    st = ms.MilvusVectorStore("stats_col", dim=d)
    st.connect()
    st.upsert_arrays([f"k{i}" for i in range(n)], x)
    assert st.last_search_stats() is None
    st.build_index(niter=2)
    st.search(x[0].tolist(), top_k=5, nprobe=4)  # switches the events on
    hits = st.search(x[1].tolist(), top_k=5, nprobe=4)[0]
    assert hits[0].id == "k1"
    s = st.last_search_stats()
    assert s is not None and s["nq"] == 1 and s["nprobe"] == 4 and s["scan"] == "query-major"
    assert s["scanned_rows"] > 0 and s["scanned_bytes"] == s["scanned_rows"] * 4 * d and s["scan_ms"] > 0 and s["scan_GBps"] > 0
