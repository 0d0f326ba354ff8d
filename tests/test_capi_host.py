"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol of
include/semcode_ivf.h; the product has no CPU fallback and never touches oracle/; the drop-in
wrapper keeps the reference's surface (reference src/semcode/storage/milvus_store.py:29-148)."""

import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(native_lib):
    from semcode_b200 import _capi

    hdr = open(os.path.join(ROOT, "include", "semcode_ivf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(native_lib, name), f"{name} is declared in the header but not exported"
    assert native_lib.sc_abi_version() == 3
    assert re.search(r"#define SC_ABI_VERSION 3\b", hdr)


def test_struct_layouts_match_header():
    import ctypes as C

    from semcode_b200 import _capi

    assert C.sizeof(_capi.ScFilter) == 32
    assert C.sizeof(_capi.ScStats) == 6 * 4 + 6 * 8 + 2 * 4
    assert C.sizeof(_capi.ScSearchTimes) == 6 * 4 + 2 * 8 + 2 * 4


def test_no_cpu_fallback_without_gpu(native_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import semcode_b200 as sb

    assert native_lib.sc_device_count() == 0
    with pytest.raises(sb.NativeError, match="no CUDA device|no CPU fallback"):
        sb.IVFFlatIndex(16, nlist=4)
    from semcode_b200.storage import MilvusVectorStore

    store = MilvusVectorStore("t_nogpu", dim=16)
    with pytest.raises(sb.NativeError):
        store.connect()
    assert store._collection is None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semcode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liborc" not in src, f


def test_wrapper_surface_matches_reference():
    from semcode_b200.storage import MilvusVectorStore

    sig = inspect.signature(MilvusVectorStore.__init__)
    assert list(sig.parameters) == ["self", "collection_name", "dim"]
    assert sig.parameters["collection_name"].default == "semcode_chunks" and sig.parameters["dim"].default is None
    up = inspect.signature(MilvusVectorStore.upsert_embeddings)
    assert list(up.parameters) == ["self", "payloads", "progress"] and up.parameters["progress"].default is None
    se = inspect.signature(MilvusVectorStore.search)
    assert list(se.parameters)[:3] == ["self", "vector", "top_k"] and se.parameters["top_k"].default == 10
    for extra in list(se.parameters)[3:]:
        assert se.parameters[extra].kind is inspect.Parameter.KEYWORD_ONLY
    store = MilvusVectorStore(dim=8)
    assert store.collection_name == "semcode_chunks" and store.dim == 8 and store._collection is None
    msg = "Milvus collection is not initialized. Call connect() first."
    with pytest.raises(RuntimeError, match=re.escape(msg)):
        store.search([0.0] * 8)
    with pytest.raises(RuntimeError, match=re.escape(msg)):
        store.upsert_embeddings([])


def test_default_dim_comes_from_settings(monkeypatch):
    from semcode_b200.storage import milvus_store as ms

    assert ms.MilvusVectorStore().dim == ms.settings.embedding_dimension


def test_hit_shape_is_what_the_rag_pipeline_reads():
    # rag/pipeline.py:133-169: hit.entity.get(name) for repo/path/language/text/metadata; score
    from semcode_b200.storage import Hit, Hits, SearchResult

    h = Hit("abc", 0.75, {"repo": "r", "path": "p", "language": "python", "text": "t", "metadata": {"a": 1}})
    assert h.entity.get("repo") == "r" and h.entity.get("missing") is None and h.entity.get("metadata") == {"a": 1}
    assert h.score == h.distance == 0.75 and h.id == "abc"
    res = SearchResult([Hits([h])])
    assert res and next(iter(res))[0] is h and res[0].ids == ["abc"] and res[0].distances == [0.75]
    assert not SearchResult()


def test_kmeans_row_selection_matches_oracle():
    from oracle import ivf_numpy as orc
    from semcode_b200 import index

    np.testing.assert_array_equal(index.kmeans_init_rows(5000, 64, 9), orc.kmeans_init_rows(5000, 64, 9))
    np.testing.assert_array_equal(index.kmeans_subsample_rows(50000, 16, 256, 9), orc.kmeans_subsample_rows(50000, 16, 256, 9))
    assert index.kmeans_subsample_rows(100, 16, 256, 9) is None


def test_every_tuning_knob_is_documented_in_the_header():
    """sc_index_set_param takes names, not enums: every name index.cu accepts must be described in include/semcode_ivf.h."""
    import re

    src = open(os.path.join(ROOT, "semcode_b200", "csrc", "index.cu")).read()
    names = sorted(set(re.findall(r'strcmp\(name, "([a-z_0-9]+)"\)', src)))
    header = open(os.path.join(ROOT, "include", "semcode_ivf.h")).read()
    assert len(names) >= 15
    assert [n for n in names if f'"{n}"' not in header] == []
