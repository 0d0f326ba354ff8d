"""Host logic of the drop-in store that needs no particular engine: persistence without flush(), crash-safe snapshot
generations, automatic compaction and re-training, concurrent searches next to an upsert.

Run twice: here with the oracle-backed engine double (CPU suite), and on the B200 with the real engine
(tests/test_gpu_store.py imports the same scenario functions)."""

import os
import threading

import numpy as np
import pytest

from helpers import unit_rows


class Payload:  # field-identical to semcode.embeddings.providers.EmbeddingPayload (providers.py:21-28)
    def __init__(self, id, text, vector, metadata):
        self.id, self.text, self.vector, self.metadata = id, text, vector, metadata


def make_payloads(rng, n, dim, repo="r", start=0):
    x = unit_rows(rng, n, dim)
    return [Payload(f"{repo}-{start + i:06d}", f"text {repo} {start + i}", x[i].tolist(),
                    {"repo": repo, "path": f"f{(start + i) % 7}.py", "language": "python" if (start + i) % 3 else "cpp"})
            for i in range(n)], x


@pytest.fixture
def double_engine(monkeypatch):
    import semcode_b200.storage.milvus_store as ms
    from engine_double import OracleIVFFlat, merge_parts

    monkeypatch.setattr(ms, "IVFFlatIndex", OracleIVFFlat)
    monkeypatch.setattr(ms, "_merge_parts", merge_parts)
    yield ms
    for name in list(ms._REGISTRY):
        ms.drop_collection(name)


def scenario_persist_without_flush(ms, monkeypatch, tmp_path):
    """The reference never calls flush() (its milvus_store.py:128-133): rows ingested by one process must be there for the
    next one.  'Next process' = the registry forgets the collection and connect() finds the snapshot."""
    monkeypatch.setenv("SEMCODE_IVF_PERSIST_DIR", str(tmp_path / "persist"))
    monkeypatch.setenv("SEMCODE_IVF_NLIST", "4")
    monkeypatch.setenv("SEMCODE_IVF_SEAL_ROWS", "200")
    rng = np.random.default_rng(1)
    store = ms.MilvusVectorStore("persisted", dim=16)
    store.connect()
    pay, x = make_payloads(rng, 300, 16)
    store.upsert_embeddings(pay)  # seals at 200 rows; snapshot written on return
    before = store.search(x[5].tolist(), top_k=3)[0]
    assert before[0].id == pay[5].id
    snap = tmp_path / "persist" / "persisted"
    cur = (snap / "CURRENT").read_text()
    assert (snap / cur / "collection.json").exists() and [d for d in os.listdir(snap) if d.startswith("gen-")] == [cur]
    ms._REGISTRY.pop("persisted").close()  # "process exit"
    again = ms.MilvusVectorStore("persisted", dim=16)
    again.connect()
    assert again._collection.num_entities == 300 and again._collection.index is not None
    after = again.search(x[5].tolist(), top_k=3)[0]
    assert [h.id for h in after] == [h.id for h in before]
    assert [round(h.distance, 5) for h in after] == [round(h.distance, 5) for h in before]
    assert after[0].entity.get("text") == pay[5].text and after[0].entity.get("metadata")["path"] == pay[5].metadata["path"]
    # a small ingest is journalled next to the published generation, not a rewrite of the collection ...
    again.upsert_embeddings(make_payloads(rng, 20, 16, repo="s")[0])
    assert (snap / "CURRENT").read_text() == cur and sorted(f for f in os.listdir(snap / cur) if f.startswith("journal-")) == ["journal-000001.npz"]
    ms._REGISTRY.pop("persisted").close()
    again = ms.MilvusVectorStore("persisted", dim=16)
    again.connect()  # ... and replayed by the next process
    assert again._collection.num_entities == 320
    assert [h.id for h in again.search(x[5].tolist(), top_k=3)[0]] == [h.id for h in before]
    # an explicit flush writes a new generation; a crash in the middle of a snapshot leaves the published one loadable
    again.flush()
    cur2 = (snap / "CURRENT").read_text()
    assert cur2 != cur and not (snap / cur).exists() and not [f for f in os.listdir(snap / cur2) if f.startswith("journal-")]
    os.makedirs(snap / "gen-99999999.tmp" / "ivf")  # debris of an interrupted save
    ms._REGISTRY.pop("persisted").close()
    third = ms.MilvusVectorStore("persisted", dim=16)
    third.connect()
    assert third._collection.num_entities == 320
    # a snapshot whose files do not belong together is refused, not mis-served
    ms._REGISTRY.pop("persisted").close()
    lines = (snap / cur2 / "columns.jsonl").read_text().splitlines()
    (snap / cur2 / "columns.jsonl").write_text("\n".join(lines[:-5]) + "\n")
    with pytest.raises(ValueError, match="inconsistent"):
        ms.MilvusVectorStore("persisted", dim=16).connect()


def scenario_journal(ms, monkeypatch, tmp_path):
    """Incremental persistence: upserts, replacements and deletes after a snapshot are appended to it as journal files and
    replayed on connect(); the journal turns into a new generation once it has grown to `ivf_journal_ratio` of the snapshot;
    a half-written journal file (crash) is ignored."""
    monkeypatch.setenv("SEMCODE_IVF_PERSIST_DIR", str(tmp_path / "persist"))
    monkeypatch.setenv("SEMCODE_IVF_NLIST", "4")
    monkeypatch.setenv("SEMCODE_IVF_SEAL_ROWS", "200")
    monkeypatch.setenv("SEMCODE_IVF_JOURNAL_RATIO", "0.5")
    rng = np.random.default_rng(9)
    name = "journalled"
    snap = tmp_path / "persist" / name

    def reopen():
        ms._REGISTRY.pop(name).close()
        st = ms.MilvusVectorStore(name, dim=16)
        st.connect()
        return st

    st = ms.MilvusVectorStore(name, dim=16)
    st.connect()
    pay, x = make_payloads(rng, 400, 16)
    st.upsert_embeddings(pay)  # seals at 200 rows -> a generation
    gen = (snap / "CURRENT").read_text()
    journals = lambda: sorted(f for f in os.listdir(snap / (snap / "CURRENT").read_text()) if f.startswith("journal-"))  # noqa: E731
    assert journals() == []
    # three small changes: new rows, replacements of existing keys (new vectors), a delete
    new, xn = make_payloads(rng, 30, 16, repo="late")
    st.upsert_embeddings(new)
    repl, xr = make_payloads(rng, 10, 16)  # same ids as pay[:10]
    st.upsert_embeddings(repl)
    st._collection.delete([pay[20].id, new[3].id])
    st._collection.flush()
    assert (snap / "CURRENT").read_text() == gen and journals() == ["journal-000001.npz", "journal-000002.npz", "journal-000003.npz"]
    want = {kk: [h.id for h in st.search(v.tolist(), top_k=4, nprobe=4)[0]] for kk, v in (("old", x[77]), ("new", xn[5]), ("repl", xr[2]))}
    (snap / gen / "journal-000004.npz.tmp").write_bytes(b"half a file")  # crash while journalling
    st = reopen()
    col = st._collection
    assert col.num_entities == 400 + 30 - 2 and col._journal_files == 3
    for kk, v in (("old", x[77]), ("new", xn[5]), ("repl", xr[2])):
        assert [h.id for h in st.search(v.tolist(), top_k=4, nprobe=4)[0]] == want[kk]
    top = st.search(xr[2].tolist(), top_k=1, nprobe=4)[0][0]
    assert top.id == repl[2].id and abs(top.distance - 1.0) < 1e-5  # the replacement's vector answers for that key
    assert pay[20].id not in col._row_of and new[3].id not in col._row_of
    # the journal continues where it stopped ...
    st.upsert_embeddings(make_payloads(rng, 5, 16, repo="later")[0])
    assert journals()[-1] == "journal-000004.npz" and (snap / "CURRENT").read_text() == gen
    # ... until it has grown to half the snapshot: then the next implicit flush writes a generation and the journal is gone
    big, _ = make_payloads(rng, 600, 16, repo="bulk")
    st.upsert_embeddings(big)
    gen2 = (snap / "CURRENT").read_text()
    col = st._collection
    assert gen2 != gen and journals() == [] and not (snap / gen).exists(), (col._snapshot_bytes, col._journal_bytes, col._pending_bytes)
    st = reopen()
    assert st._collection.num_entities == 400 + 30 - 2 + 5 + 600 and st._collection._journal_files == 0


def scenario_compaction_and_retrain(ms, monkeypatch):
    monkeypatch.setenv("SEMCODE_IVF_NLIST", "8")
    monkeypatch.setenv("SEMCODE_IVF_SEAL_ROWS", "400")
    monkeypatch.setenv("SEMCODE_IVF_COMPACT_RATIO", "0.25")
    monkeypatch.setenv("SEMCODE_IVF_RETRAIN_FACTOR", "3")
    rng = np.random.default_rng(2)
    store = ms.MilvusVectorStore("maintained", dim=24)
    store.connect()
    col = store._collection
    pay, x = make_payloads(rng, 400, 24)
    store.upsert_embeddings(pay)
    assert col.index is not None and col._trained_rows == 400
    # re-upserting a third of the keys tombstones the old rows: past 25 % of the slots the index compacts itself
    pay2, x2 = make_payloads(rng, 140, 24)
    store.upsert_embeddings(pay2)
    assert col.maintenance["compactions"] >= 1 and col.index.stats().nremoved == 0 and col.num_entities == 400
    hit = store.search(x2[7].tolist(), top_k=1)[0][0]
    assert hit.id == pay2[7].id and abs(hit.distance - 1.0) < 1e-5  # the NEW vector answers for that key
    assert store.search(x[7].tolist(), top_k=400, nprobe=8)[0].ids.count(pay[7].id) == 1
    # growth to 3x the trained size re-clusters the index on its current rows
    more, xm = make_payloads(rng, 900, 24, repo="grown")
    store.upsert_embeddings(more)
    assert col.maintenance["retrains"] == 1 and col._trained_rows >= 1200 and col.num_entities == 1300
    for probe, p in ((xm[3], more[3]), (x2[9], pay2[9]), (x[399], pay[399])):
        h = store.search(probe.tolist(), top_k=1, nprobe=8)[0][0]
        assert h.id == p.id and abs(h.distance - 1.0) < 1e-5
    got = store.search(xm[3].tolist(), top_k=5, nprobe=8, repos=["grown"], languages=["cpp"])[0]
    assert got and all(h.entity.get("repo") == "grown" and h.entity.get("language") == "cpp" for h in got)


def scenario_concurrent_searches(ms, monkeypatch):
    """FastAPI answers /query from a thread pool while BackgroundTasks ingests (api/main.py:160,202): searches share the read
    side of the collection lock, an upsert takes the write side; every reader sees a consistent collection."""
    monkeypatch.setenv("SEMCODE_IVF_NLIST", "4")
    monkeypatch.setenv("SEMCODE_IVF_SEAL_ROWS", "300")
    rng = np.random.default_rng(3)
    store = ms.MilvusVectorStore("concurrent", dim=16)
    store.connect()
    pay, x = make_payloads(rng, 300, 16)
    store.upsert_embeddings(pay)
    errors, done = [], threading.Event()

    def reader(i):
        try:
            while not done.is_set():
                j = (i * 37) % 300
                hits = store.search(x[j].tolist(), top_k=2)[0]
                assert hits[0].id == pay[j].id, (hits[0].id, pay[j].id)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=reader, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for r in range(4):
        store.upsert_embeddings(make_payloads(rng, 60, 16, repo=f"w{r}")[0])
    done.set()
    for t in threads:
        t.join(timeout=60)
    assert not errors and store._collection.num_entities == 540
    # readers really overlap: two of them inside the read section at once
    lock = store._collection._lock
    inside, peak = [0], [0]
    gate = threading.Barrier(2, timeout=20)

    def overlap():
        with lock.read():
            inside[0] += 1
            peak[0] = max(peak[0], inside[0])
            gate.wait()
            inside[0] -= 1

    ts = [threading.Thread(target=overlap) for _ in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=30)
    assert peak[0] == 2


def test_persist_without_flush(double_engine, monkeypatch, tmp_path):
    scenario_persist_without_flush(double_engine, monkeypatch, tmp_path)


def test_journal(double_engine, monkeypatch, tmp_path):
    scenario_journal(double_engine, monkeypatch, tmp_path)


def test_compaction_and_retrain(double_engine, monkeypatch):
    scenario_compaction_and_retrain(double_engine, monkeypatch)


def test_concurrent_searches(double_engine, monkeypatch):
    scenario_concurrent_searches(double_engine, monkeypatch)
