"""What a caller of the reference-facing wrapper sees for ONE query (rag/pipeline.py:96 sends one per request):
MilvusVectorStore.search(vector: list[float], top_k) -> [Hits] on a C1-shaped collection (100k x 768, nlist 1024, nprobe 16)
and on the raw engine with host arrays.  Wall-clock per call, median of 300."""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SEMCODE_IVF_NLIST", "1024")
os.environ.setdefault("SEMCODE_IVF_SEAL_ROWS", "1000000")
import numpy as np
import torch

import semcode_b200.storage.milvus_store as ms

n, d = 100_000, 768
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.nn.functional.normalize(torch.randn((n, d), generator=g, device="cuda"), dim=1)
st = ms.MilvusVectorStore("lat", dim=d)
st.connect()
st.upsert_arrays([f"k{i}" for i in range(n)], x, repos=["r"] * n, languages=["python"] * n, texts=["t"] * n)
st.build_index(niter=4)
q = x[:400].cpu().numpy()
qlist = [v.tolist() for v in q]


def med(fn, reps=300):
    for i in range(20):
        fn(i)
    t = []
    for i in range(reps):
        t0 = time.perf_counter()
        fn(i)
        t.append(time.perf_counter() - t0)
    return statistics.median(t) * 1e6


out = {
    "wrapper_search_list_of_floats_us": med(lambda i: st.search(qlist[i % 400], top_k=10, nprobe=16)),
    "wrapper_search_arrays_numpy_us": med(lambda i: st.search_arrays(q[i % 400:i % 400 + 1], 10, nprobe=16)),
    "engine_numpy_in_numpy_out_us": med(lambda i: st._collection.index.search(q[i % 400:i % 400 + 1], 10, nprobe=16)),
}
qd = x[:400].contiguous()
od, oi = torch.empty((1, 10), device="cuda"), torch.empty((1, 10), dtype=torch.int64, device="cuda")


def dev(i):
    st._collection.index.search(qd[i % 400:i % 400 + 1], 10, nprobe=16, out=(od, oi))
    torch.cuda.synchronize()


out["engine_device_in_out_plus_sync_us"] = med(dev)
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
