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
