"""The reference's OWN callers, unmodified, against the drop-in store.

`/root/reference/src/semcode/services/indexer.py` (IndexerService), `rag/pipeline.py` (SemanticSearchPipeline) and the
reference's integration test `tests/integration/test_indexer_service.py` are imported from the read-only reference tree
and executed with `semcode.storage.MilvusVectorStore` resolved to `semcode_b200.storage.MilvusVectorStore` -- the
one-line swap of INTEGRATION.md section 1, made here through `sys.modules` so that no reference file is touched.

Third-party imports the image lacks (structlog, langchain, tree_sitter) are satisfied by the stand-ins under
tests/shims/.  The engine is the real one when a B200 is present; in the GPU-less build container it is an
oracle-backed TEST DOUBLE (tests/engine_double.py) -- what is under test here is the wrapper's contract with the callers.
The reference tree does not travel to the GPU box: there these tests skip (and say so) and
tests/test_gpu_store.py::test_reference_caller_traffic_replayed_on_the_gpu replays the traffic recorded here.
"""

import hashlib
import importlib
import importlib.util
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference"
REF_SRC = os.path.join(REF_ROOT, "src")
SHIMS = os.path.join(ROOT, "tests", "shims")

pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF_SRC, "semcode")),
    reason="the reference tree (/root/reference) exists only in the build container, not on the GPU box",
)

DIM = 32


class HashEmbedding:
    """Deterministic stand-in for the embedding provider: unit vector seeded by the text."""

    def __init__(self, dim=DIM):
        self.dim = dim

    def _one(self, text):
        seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little")
        v = np.random.default_rng(seed).standard_normal(self.dim)
        return (v / np.linalg.norm(v)).astype(np.float32).tolist()

    def embed_documents(self, texts):
        return [self._one(t) for t in texts]

    def embed_query(self, text):
        return self._one(text)


def _purge_semcode():
    for name in [m for m in sys.modules if m == "semcode" or m.startswith("semcode.")]:
        del sys.modules[name]


def _install(monkeypatch, src_root):
    """Shims + reference on sys.path, the engine (double without a GPU), and the documented storage swap."""
    import torch

    import semcode_b200.storage.milvus_store as ms

    monkeypatch.syspath_prepend(os.path.join(ROOT, "tests"))
    monkeypatch.syspath_prepend(SHIMS)
    monkeypatch.syspath_prepend(src_root)
    if not torch.cuda.is_available():
        from engine_double import OracleIVFFlat, merge_parts

        monkeypatch.setattr(ms, "IVFFlatIndex", OracleIVFFlat)
        monkeypatch.setattr(ms, "_merge_parts", merge_parts)
    _purge_semcode()
    importlib.import_module("semcode")  # the reference package itself
    # `from .milvus_store import MilvusVectorStore` in the reference's storage/__init__.py now
    # binds the drop-in
    monkeypatch.setitem(sys.modules, "semcode.storage.milvus_store", ms)
    storage = importlib.import_module("semcode.storage")
    assert storage.MilvusVectorStore is ms.MilvusVectorStore
    # inside semcode the drop-in reads the reference's own settings object (milvus_store.py: `from semcode.settings import
    # settings` at import time); here it was imported before semcode was importable, so hand it over now
    monkeypatch.setattr(ms, "settings", importlib.import_module("semcode.settings").settings)
    return ms


@pytest.fixture
def ref(monkeypatch, tmp_path):
    ms = _install(monkeypatch, REF_SRC)
    yield ms
    for name in list(ms._REGISTRY):
        ms.drop_collection(name)
    _purge_semcode()


def _make_repo(root):
    root.mkdir()
    for i in range(3):
        body = "\n".join(f"def fn_{i}_{j}(x):\n    return x * {j} + {i}\n" for j in range(140))  # > 200 lines: several chunks
        (root / f"mod_{i}.py").write_text(body)
    (root / "native.cpp").write_text("int add(int a, int b) {\n  return a + b;\n}\n")
    (root / "notes.txt").write_text("not a source file\n")


def _index_demo(ref, tmp_path, monkeypatch, collection):
    from semcode.ingestion import RepositoryIngestionManager
    from semcode.services import IndexerService, IndexingCallbacks
    from semcode.settings import settings
    from semcode.storage import MilvusVectorStore, RepositoryRegistry

    workspace = tmp_path / "workspace"
    monkeypatch.setattr(settings, "workspace_root", workspace)
    monkeypatch.setattr(settings, "embedding_dimension", DIM, raising=False)
    monkeypatch.setattr("semcode.services.indexer.EmbeddingProviderFactory.create", lambda provider=None, model=None: HashEmbedding())
    src = tmp_path / "demo_src"
    if not src.exists():
        _make_repo(src)
    seen = {"stages": [], "upsert": [], "embed": []}
    cb = IndexingCallbacks(stage=seen["stages"].append, upsert_progress=lambda a, b: seen["upsert"].append((a, b)),
                           embed_progress=lambda a, b: seen["embed"].append((a, b)))
    service = IndexerService(
        ingestion_manager=RepositoryIngestionManager(workspace=workspace),
        registry=RepositoryRegistry(registry_path=workspace / "registry.json"),
        vector_store=MilvusVectorStore(collection_name=collection, dim=DIM),
    )  # auto_connect=True: IndexerService.__init__ calls store.connect() (indexer.py:54-63)
    assert service._connected
    return service, src, seen, cb


def test_indexer_service_runs_unchanged_against_the_drop_in(ref, tmp_path, monkeypatch):
    service, src, seen, cb = _index_demo(ref, tmp_path, monkeypatch, "callers_idx")
    result = service.index_repository(paths=[src], name="demo", callbacks=cb)
    n = result.chunk_count
    assert n >= 9 and result.embeddings_indexed == n and result.milvus_collection == "callers_idx"
    assert "upsert_completed" in seen["stages"] and "upsert_failed" not in seen["stages"]
    assert seen["upsert"][0] == (0, n) and seen["upsert"][-1] == (n, n)  # progress protocol of milvus_store.py:101-133
    assert [a for a, _ in seen["upsert"]] == sorted(a for a, _ in seen["upsert"])
    col = service.vector_store._collection
    assert col.num_entities == n
    langs = {col._language[r] for r in col._row_of.values()}
    assert langs == {"python", "cpp"}
    assert any(rec.name == "demo" for rec in service.registry.list())
    # the same repository again: ids are md5(repo:path:start:end) (indexer.py:186-188) -> upsert replaces by primary key
    service.index_repository(paths=[src], name="demo", force=True, callbacks=cb)
    assert col.num_entities == n


def test_rag_retrieval_runs_unchanged_against_the_drop_in(ref, tmp_path, monkeypatch):
    from semcode.rag import SemanticSearchPipeline
    from semcode.settings import settings

    service, src, _, cb = _index_demo(ref, tmp_path, monkeypatch, "semcode_chunks")
    service.index_repository(paths=[src], name="demo", callbacks=cb)
    col = service.vector_store._collection
    monkeypatch.setattr(settings, "rag_max_context_sources", 4, raising=False)
    pipeline = SemanticSearchPipeline()  # collection "semcode_chunks", its own MilvusVectorStore (pipeline.py:42)
    assert type(pipeline.vector_store) is ref.MilvusVectorStore
    pipeline._embedding = HashEmbedding()
    target = next(r for r in col._row_of.values() if col._language[r] == "python")
    docs = pipeline._retrieve_documents(col._text[target])  # pipeline.py:93-131 -> _hit_to_document :133-169
    assert len(docs) == 4 and pipeline._last_retrieval_error is None
    top = docs[0]
    assert set(top) == {"repo", "path", "language", "snippet", "score", "metadata"}
    assert top["snippet"] == col._text[target] and top["repo"] == "demo" and top["language"] == "python"
    assert abs(top["score"] - 1.0) < 1e-5 and top["metadata"]["start_line"] >= 1
    assert [d["score"] for d in docs] == sorted((d["score"] for d in docs), reverse=True)

    class _LLM:
        def invoke(self, messages):
            return type("R", (), {"content": f"{len(messages)} messages"})()

    monkeypatch.setattr(pipeline, "_create_llm", lambda: _LLM())
    out = pipeline.query(col._text[target])  # pipeline.py:49-88
    assert out["answer"] == "2 messages" and len(out["sources"]) == 4 and out["meta"] == {"fallback_used": False}
    # an empty collection under another name: one empty hit list (as pymilvus answers [EXT]) -> no documents, and query()
    # takes the caller's own "no_documents" branch (pipeline.py:54-62)
    empty = SemanticSearchPipeline(collection_name="callers_empty", fallback_enabled=False)
    empty._embedding = HashEmbedding()
    assert empty._retrieve_documents("anything") == [] and empty._last_retrieval_error is None
    assert empty.query("anything")["meta"] == {"fallback_used": False, "reason": "no_documents"}


def test_reference_integration_test_with_the_store_injected(ref, tmp_path, monkeypatch):
    """reference tests/integration/test_indexer_service.py:32-68, body unchanged; only its DummyVectorStore is the drop-in."""
    path = os.path.join(REF_ROOT, "tests", "integration", "test_indexer_service.py")
    spec = importlib.util.spec_from_file_location("ref_test_indexer_service", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class StoreUnderTest(ref.MilvusVectorStore):
        def __init__(self):
            super().__init__(collection_name="test_semcode_chunks", dim=1)  # the reference's DummyEmbedding is 1-d
            self.connect()  # the reference test sets service._connected = True itself (test_indexer_service.py:58)

        @property
        def payloads(self):  # what the reference test asserts on (:66)
            return list(self._collection._row_of)

    monkeypatch.setattr(mod, "DummyVectorStore", StoreUnderTest)
    mod.test_indexer_service_integration(tmp_path, monkeypatch)
    assert ref.has_collection("test_semcode_chunks") and ref._REGISTRY["test_semcode_chunks"].num_entities > 0


def test_reference_api_app_ingests_and_answers_from_the_drop_in(ref, tmp_path, monkeypatch):
    """The reference's FastAPI application (api/main.py), unmodified, over HTTP (fastapi.testclient): POST /ingest runs the
    real IndexerService into the drop-in store, POST /query runs the real SemanticSearchPipeline out of it -- the module
    constructs both at import time (api/main.py:24-29) with whatever `semcode.storage.MilvusVectorStore` is.  Embeddings and
    the LLM are the only stand-ins (they are network services)."""
    from fastapi.testclient import TestClient
    from semcode.settings import settings

    workspace = tmp_path / "workspace"
    monkeypatch.setattr(settings, "workspace_root", workspace)
    monkeypatch.setattr(settings, "embedding_dimension", DIM, raising=False)
    monkeypatch.setattr(settings, "api_key", "secret", raising=False)
    monkeypatch.setattr(settings, "telemetry_enabled", True, raising=False)
    monkeypatch.setattr(settings, "rag_max_context_sources", 3, raising=False)
    monkeypatch.setattr("semcode.services.indexer.EmbeddingProviderFactory.create", lambda provider=None, model=None: HashEmbedding())
    api = importlib.import_module("semcode.api.main")
    assert type(api.indexer.vector_store) is ref.MilvusVectorStore and type(api.pipeline.vector_store) is ref.MilvusVectorStore
    api.pipeline._embedding = HashEmbedding()

    class _LLM:
        def invoke(self, messages):
            return type("R", (), {"content": "an answer"})()

    monkeypatch.setattr(api.pipeline, "_create_llm", lambda: _LLM())
    root = tmp_path / "checkout"
    root.mkdir()
    _make_repo(root / "src")
    client = TestClient(api.app)
    hdr = {"X-API-Key": "secret"}
    assert client.post("/query", json={"question": "x"}).status_code == 401  # the app's own auth still guards the route
    r = client.post("/ingest", headers=hdr, json={"name": "demo", "root": str(root), "include": ["src"]})
    assert r.status_code == 200, r.text
    body = r.json()
    assert body["name"] == "demo" and body["chunk_count"] > 3 and set(body["languages"]) == {"python", "cpp"}
    col = ref._REGISTRY["semcode_chunks"]  # ONE collection behind both the indexer's and the pipeline's store objects
    assert col.num_entities == body["chunk_count"]
    assert [rp["name"] for rp in client.get("/repos", headers=hdr).json()] == ["demo"]
    text = col._text[next(rr for rr in col._row_of.values() if col._language[rr] == "cpp")]
    r = client.post("/query", headers=hdr, json={"question": text})
    assert r.status_code == 200, r.text
    out = r.json()
    assert out["answer"] == "an answer" and len(out["sources"]) == 3 and out["meta"] == {"fallback_used": False}
    assert out["sources"][0]["snippet"] == text and out["sources"][0]["language"] == "cpp" and abs(out["sources"][0]["score"] - 1.0) < 1e-5
    tele = client.get("/telemetry", headers=hdr).json()
    assert tele["ingest"]["count"] == 1 and tele["query"]["count"] == 1 and tele["query"]["failures"] == 0


def test_reference_api_endpoint_test_runs_with_the_store_swapped(ref, tmp_path, monkeypatch):
    """reference tests/integration/test_api_endpoints.py, body unchanged: importing semcode.api.main constructs the indexer and
    the pipeline on top of the drop-in store."""
    path = os.path.join(REF_ROOT, "tests", "integration", "test_api_endpoints.py")
    spec = importlib.util.spec_from_file_location("ref_test_api_endpoints", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert type(mod.api_main.indexer.vector_store) is ref.MilvusVectorStore
    keep = mod.api_main.pipeline
    try:
        mod.test_api_endpoints_with_stubs(tmp_path, monkeypatch)
    finally:
        mod.api_main.pipeline = keep  # the reference test assigns its stub without monkeypatch


def test_reference_cli_ingests_and_a_later_process_answers(ref, tmp_path, monkeypatch):
    """`semcode ingest` (cli.py:118-301, unmodified, through typer's CliRunner) fills the collection in one process; with
    `ivf_persist_dir` set the rows are there for the NEXT process -- here the reference's API app -- although the reference
    never calls flush() (its Milvus server persisted for it): the one-process limitation ADVICE round 1 pointed at."""
    from fastapi.testclient import TestClient
    from semcode.settings import settings
    from typer.testing import CliRunner

    workspace = tmp_path / "workspace"
    monkeypatch.setattr(settings, "workspace_root", workspace)
    monkeypatch.setattr(settings, "embedding_dimension", DIM, raising=False)
    monkeypatch.setattr(settings, "ivf_persist_dir", str(tmp_path / "persist"), raising=False)
    monkeypatch.setattr(settings, "api_key", "", raising=False)
    monkeypatch.setattr(settings, "rag_max_context_sources", 2, raising=False)
    monkeypatch.setattr("semcode.services.indexer.EmbeddingProviderFactory.create", lambda provider=None, model=None: HashEmbedding())
    root = tmp_path / "checkout"
    root.mkdir()
    _make_repo(root / "src")
    cli = importlib.import_module("semcode.cli")
    res = CliRunner().invoke(cli.app, ["ingest", "--name", "demo", "--include", "src", "--root", str(root), "--yes"])
    assert res.exit_code == 0, res.output
    assert "Ingested demo" in res.output and "chunks=" in res.output
    n = ref._REGISTRY["semcode_chunks"].num_entities
    assert n > 3 and os.path.exists(tmp_path / "persist" / "semcode_chunks" / "CURRENT")
    ref._REGISTRY.pop("semcode_chunks").close()  # the CLI process is gone

    api = importlib.import_module("semcode.api.main")  # a new process: the API server
    api.pipeline._embedding = HashEmbedding()
    monkeypatch.setattr(api.pipeline, "_create_llm", lambda: type("L", (), {"invoke": lambda self, m: type("R", (), {"content": "ok"})()})())
    client = TestClient(api.app)
    r = client.post("/query", json={"question": "int add(int a, int b) {\n  return a + b;\n}"})
    assert r.status_code == 200, r.text
    out = r.json()
    assert ref._REGISTRY["semcode_chunks"].num_entities == n  # loaded from the snapshot the CLI's upsert left behind
    assert out["answer"] == "ok" and len(out["sources"]) == 2 and out["sources"][0]["repo"] == "demo"


def test_optional_caller_patch_pushes_filters_down_and_batches(monkeypatch, tmp_path):
    """patches/semcode_filter_pushdown_and_batch.patch: QueryRequest.repos / languages reach the
    index scan through SemanticSearchPipeline.query, retrieve_batch issues ONE store call for many questions, and the API's
    telemetry records the scan's bytes and GB/s."""
    src = tmp_path / "patched" / "src"
    shutil.copytree(REF_SRC, src)
    patch = os.path.join(ROOT, "patches", "semcode_filter_pushdown_and_batch.patch")
    r = subprocess.run(["patch", "-p1", "-i", patch], cwd=src.parent, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ms = _install(monkeypatch, str(src))
    try:
        from semcode.rag import SemanticSearchPipeline
        from semcode.settings import settings

        service, repo_src, _, cb = _index_demo(ms, tmp_path, monkeypatch, "semcode_chunks")
        service.index_repository(paths=[repo_src], name="demo", callbacks=cb)
        other = tmp_path / "other_src"
        other.mkdir()
        (other / "solo.py").write_text("def solo():\n    return 42\n")
        service.index_repository(paths=[other], name="other", callbacks=cb)
        col = service.vector_store._collection
        monkeypatch.setattr(settings, "rag_max_context_sources", 5, raising=False)
        pipeline = SemanticSearchPipeline()
        pipeline._embedding = HashEmbedding()
        q = col._text[next(r for r in col._row_of.values() if col._repo[r] == "demo" and col._language[r] == "python")]
        assert {d["repo"] for d in pipeline._retrieve_documents(q)} == {"demo"}
        only_other = pipeline._retrieve_documents(q, repos=["other"])
        assert [d["repo"] for d in only_other] == ["other"]  # filter-then-rank: the one row of that repo, not an empty top-k
        cpp = pipeline._retrieve_documents(q, languages=["cpp"])
        assert cpp and {d["language"] for d in cpp} == {"cpp"}
        calls = []
        real = pipeline.vector_store.search_batch
        monkeypatch.setattr(pipeline.vector_store, "search_batch", lambda *a, **k: calls.append(k) or real(*a, **k))
        texts = [col._text[r] for r in list(col._row_of.values())[:6]]
        batches = pipeline.retrieve_batch(texts, repos=["demo"])
        assert len(calls) == 1 and calls[0]["repos"] == ["demo"] and len(batches) == 6
        assert all(b[0]["snippet"] == t for b, t in zip(batches, texts))
        # the API model carries the new fields and the handler threads them through
        api = importlib.import_module("semcode.api.main")
        req = api.QueryRequest(question="q", repos=["other"], languages=None)
        seen = {}
        monkeypatch.setattr(api.pipeline, "query", lambda question, **kw: seen.update(kw) or {"answer": "a", "sources": []})
        assert api.query(req).answer == "a" and seen == {"repos": ["other"], "languages": None}
        # telemetry: with `ivf_profile` the store reports what a search on the sealed index cost and the
        # patched API records it next to the query's duration (api/telemetry.py)
        monkeypatch.setattr(settings, "ivf_profile", True, raising=False)
        monkeypatch.setattr(settings, "telemetry_enabled", True, raising=False)
        store = api.pipeline.vector_store
        store.connect()
        store.build_index(niter=2)  # seal: searches now run on the IVF lists
        vec = HashEmbedding().embed_query(q)
        store.search(vec, top_k=3)  # switches the per-phase events on
        store.search(vec, top_k=3)
        stats = store.last_search_stats()
        assert stats and stats["nq"] == 1 and stats["scanned_bytes"] > 0 and stats["scan_GBps"] > 0 and stats["scan"] in ("query-major", "list-major")
        api.query(api.QueryRequest(question="q"))
        event = api.telemetry.snapshot()["recent_events"][0]
        assert event["kind"] == "query" and event["metadata"]["retrieval"]["scanned_bytes"] == stats["scanned_bytes"]
        # the Streamlit front end sends a narrowed repo / language selection along instead of only post-filtering the answer
        # (streamlit is not in the image: a bare stand-in module lets the patched file import)
        monkeypatch.setitem(sys.modules, "streamlit", type(sys)("streamlit"))
        fe = importlib.import_module("semcode.frontend.app")
        sent = {}
        monkeypatch.setattr(fe, "_request", lambda method, url, api_key=None, **kw: sent.update(kw["json"]) or type("R", (), {"json": lambda self: {}})())
        fe._run_query("http://x", None, "q", repos=["other"], languages=None)
        assert sent == {"question": "q", "repos": ["other"]}
    finally:
        for name in list(ms._REGISTRY):
            ms.drop_collection(name)
        _purge_semcode()
