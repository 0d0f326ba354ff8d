"""Regenerates the golden fixtures of tests/golden/ from the NumPy oracle.

PARITY UNPINNED at the reference boundary: rmontanana/semcode holds no golden vector for its
IVF_FLAT path and its engine (Milvus/knowhere/FAISS) cannot be run here, so these fixtures pin
the *oracle* (and through it the CUDA path), not the reference.  Run from the repo root:

    python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ivf_numpy as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def kat_small():
    """KAT-2/4: small seeded set, inputs stored (N=600, d=24, nlist=12), both metrics, with and
    without a 3-repo/1-language filter."""
    rng = np.random.default_rng(20261018)
    x, q = unit(rng, 600, 24), unit(rng, 16, 24)
    ids = rng.permutation(10_000)[:600].astype(np.int64)
    repo = rng.integers(0, 12, 600).astype(np.uint32)
    lang = rng.integers(0, 3, 600).astype(np.uint8)
    out = dict(x=x, q=q, ids=ids, repo=repo, lang=lang)
    for metric in ("IP", "L2"):
        cent = x[orc.kmeans_init_rows(600, 12, 7)].copy()
        idx = orc.build_index(x, ids, cent, metric, repo, lang)
        probes = orc.coarse_probe(q, cent, metric, 4)
        d, i = orc.search(idx, q, 10, 4, probes=probes)
        mask = orc.row_mask(idx, repos=[1, 2, 3], langs=[1])
        fd, fi = orc.search(idx, q, 10, 4, mask=mask, probes=probes)
        out.update({f"cent_{metric}": cent, f"probes_{metric}": probes, f"dist_{metric}": d, f"ids_{metric}": i,
                    f"fdist_{metric}": fd, f"fids_{metric}": fi})
    np.savez_compressed(os.path.join(HERE, "kat_small.npz"), **out)


def kat_c1():
    """KAT-3: BASELINE.json configs[0] shape -- 100k x 768, nlist 1024, nprobe 16, top-10, IP.
    Inputs are regenerated from the seeds (default_rng(1234) DB, (4321) queries); centroids are the
    rows kmeans_init_rows(n, nlist, 1234); only the expected results of 64 queries are stored."""
    n, d, nlist, nq = 100_000, 768, 1024, 64
    x = unit(np.random.default_rng(1234), n, d)
    q = unit(np.random.default_rng(4321), nq, d)
    ids = np.arange(n, dtype=np.int64)
    cent = x[orc.kmeans_init_rows(n, nlist, 1234)].copy()
    idx = orc.build_index(x, ids, cent, "IP")
    probes = orc.coarse_probe(q, cent, "IP", 16)
    dist, out_ids = orc.search(idx, q, 10, 16, probes=probes)
    np.savez_compressed(os.path.join(HERE, "kat_c1.npz"), probes=probes, dist=dist, ids=out_ids,
                        list_sizes=np.diff(idx.list_off).astype(np.int32))


if __name__ == "__main__":
    kat_small()
    kat_c1()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
