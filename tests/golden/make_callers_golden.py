"""Record the traffic of the reference's own callers for tests/test_gpu_store.py::test_reference_caller_traffic_replayed_on_the_gpu.

Runs in the build container only (needs /root/reference): the UNMODIFIED reference IndexerService indexes a small
repository into the drop-in store (engine = oracle-backed double, tests/engine_double.py) and the UNMODIFIED
SemanticSearchPipeline retrieves documents for a few questions; payloads, query vectors and the documents that the
reference's `_hit_to_document` produced are written to tests/golden/callers.json.

    python tests/golden/make_callers_golden.py
"""
import json
import os
import sys
import tempfile
from pathlib import Path

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import pytest  # noqa: E402

import test_reference_callers as trc  # noqa: E402


def main():
    mp = pytest.MonkeyPatch()
    try:
        with tempfile.TemporaryDirectory() as tmp:
            tmp = Path(tmp)
            ms = trc._install(mp, trc.REF_SRC)
            from semcode.rag import SemanticSearchPipeline
            from semcode.settings import settings

            recorded = []
            real_upsert = ms.MilvusVectorStore.upsert_embeddings

            def spy(self, payloads, progress=None):
                payloads = list(payloads)
                recorded.extend(payloads)
                return real_upsert(self, payloads, progress)

            mp.setattr(ms.MilvusVectorStore, "upsert_embeddings", spy)
            service, src, _, cb = trc._index_demo(ms, tmp, mp, "semcode_chunks")
            service.index_repository(paths=[src], name="demo", callbacks=cb)
            mp.setattr(settings, "rag_max_context_sources", 5, raising=False)
            pipeline = SemanticSearchPipeline()
            emb = trc.HashEmbedding()
            pipeline._embedding = emb
            queries = []
            for i, p in enumerate(recorded[:: max(1, len(recorded) // 6)][:6]):
                question = p.text if i % 2 == 0 else f"where is fn_{i}_3 defined?"
                docs = pipeline._retrieve_documents(question)
                queries.append({"question": question, "vector": emb.embed_query(question), "top_k": 5, "documents": docs})
            out = {"dim": trc.DIM, "payloads": [{"id": p.id, "text": p.text, "vector": p.vector, "metadata": p.metadata} for p in recorded],
                   "queries": queries}
            with open(os.path.join(HERE, "callers.json"), "w") as f:
                json.dump(out, f)
            print(f"{len(recorded)} payloads, {len(queries)} queries -> tests/golden/callers.json")
    finally:
        mp.undo()


if __name__ == "__main__":
    main()
