"""CPU stand-in for semcode_b200.IVFFlatIndex, backed by the oracle -- TESTS ONLY.

The product has no CPU engine.  This double lets the host-side logic of the drop-in store (growing segment, seal,
primary-key replace, tags, hit objects, persistence hooks) and the REFERENCE's own callers run in the GPU-less container;
on the GPU box the same tests use the real engine.  It implements exactly the engine surface GpuCollection drives."""

from __future__ import annotations

import numpy as np

from oracle import ivf_numpy as orc


class _Stats:
    def __init__(self, e):
        self.dim = self.dim_padded = e.dim
        self.nlist, self.metric, self.trained = e.nlist, e.metric, int(e.c is not None)
        self.ntotal, self.nremoved, self.npages, self.nfree_pages = e.ntotal, e._removed, 0, 0


class OracleIVFFlat:
    def __init__(self, dim, nlist=128, metric="IP", device=0):
        self.dim, self.nlist, self.metric, self.device = int(dim), int(nlist), orc.metric_code(metric), int(device)
        self.c = None
        self._x = np.zeros((0, self.dim), np.float32)
        self._ids = np.zeros(0, np.int64)
        self._repo = np.zeros(0, np.uint32)
        self._lang = np.zeros(0, np.uint8)
        self._list = np.zeros(0, np.int32)
        self._dead = np.zeros(0, bool)
        self._removed = 0

    # -- lifecycle / quantizer
    def close(self):
        pass

    def reset(self):
        c = self.c
        self.__init__(self.dim, self.nlist, self.metric, self.device)
        self.c = c

    def set_centroids(self, c):
        self.c = np.asarray(c, np.float32).reshape(self.nlist, self.dim).copy()

    def get_centroids(self):
        return self.c.copy()

    @property
    def is_trained(self):
        return self.c is not None

    def train(self, x, niter=25, seed=1234, max_points_per_centroid=256, init_centroids=None):
        self.c, obj = orc.kmeans_train(np.asarray(x, np.float32), self.nlist, self.metric, niter=niter, seed=seed,
                                       max_points_per_centroid=max_points_per_centroid, init_centroids=init_centroids)
        return obj

    def assign(self, x):
        return orc.assign(np.asarray(x, np.float32), self.c, self.metric).astype(np.int32)

    # -- rows
    def add(self, x, ids, repo_tags=None, lang_tags=None, lists=None):
        x = np.asarray(x, np.float32).reshape(-1, self.dim)
        n = x.shape[0]
        if repo_tags is not None and np.asarray(repo_tags).max(initial=0) > (1 << 23) - 1:
            raise ValueError("repo tag above the 23-bit field")
        self._x = np.concatenate([self._x, x])
        self._ids = np.concatenate([self._ids, np.asarray(ids, np.int64)])
        self._repo = np.concatenate([self._repo, np.zeros(n, np.uint32) if repo_tags is None else np.asarray(repo_tags, np.uint32)])
        self._lang = np.concatenate([self._lang, np.zeros(n, np.uint8) if lang_tags is None else np.asarray(lang_tags, np.uint8)])
        self._list = np.concatenate([self._list, self.assign(x) if lists is None else np.asarray(lists, np.int32)])
        self._dead = np.concatenate([self._dead, np.zeros(n, bool)])

    def remove_ids(self, ids):
        hit = np.isin(self._ids, np.asarray(ids, np.int64)) & ~self._dead
        self._dead |= hit
        self._removed += int(hit.sum())
        return int(hit.sum())

    def compact(self):
        keep = ~self._dead
        self._x, self._ids, self._repo, self._lang, self._list = (a[keep] for a in (self._x, self._ids, self._repo, self._lang, self._list))
        self._dead = np.zeros(self._ids.size, bool)
        self._removed = 0
        return 0

    @property
    def ntotal(self):
        return int((~self._dead).sum())

    def stats(self):
        return _Stats(self)

    def list_sizes(self):
        return np.bincount(self._list, minlength=self.nlist).astype(np.int32)

    def export_list(self, l):
        m = self._list == l
        tags = (self._repo[m] << np.uint32(8)) | self._lang[m].astype(np.uint32) | (self._dead[m].astype(np.uint32) << np.uint32(31))
        return self._x[m].copy(), self._ids[m].copy(), tags.astype(np.uint32)

    def export_lists(self, lo, hi, device=False):
        parts = [self.export_list(l) for l in range(lo, hi)]
        off = np.zeros(hi - lo + 1, np.int64)
        np.cumsum([p[1].size for p in parts], out=off[1:])
        return (off, np.concatenate([p[0] for p in parts]) if parts else self._x[:0], np.concatenate([p[1] for p in parts]),
                np.concatenate([p[2] for p in parts]))

    def list_ranges(self, max_bytes=512 << 20):
        return [(0, self.nlist)]

    # -- profiling hooks (what IVFFlatIndex.set_profiling / last_search_times report; here: row counts only)
    def set_profiling(self, enabled):
        self._prof = bool(enabled)

    def last_search_times(self):
        from types import SimpleNamespace

        if not getattr(self, "_prof", False) or getattr(self, "_last", None) is None:
            raise RuntimeError("profiling is off or no search has run")
        nq, nprobe = self._last
        rows = int(nq * nprobe * max(1, len(self._ids)) / self.nlist)
        return SimpleNamespace(coarse_ms=0.01, probe_select_ms=0.01, plan_ms=0.01, scan_ms=0.1, topk_ms=0.01, total_ms=0.14,
                               scanned_rows=rows, unique_rows=0, scan_launches=1, total_launches=5)

    # -- search
    def search(self, q, k, nprobe=16, repos=None, langs=None, lists=None, out=None, exchange=None):
        q = np.asarray(q, np.float32).reshape(-1, self.dim)
        self._last = (q.shape[0], min(int(nprobe), self.nlist))
        idx = orc.build_index(self._x, self._ids, self.c, self.metric, self._repo, self._lang, assignment=self._list)
        mask = orc.row_mask(idx, repos=repos, langs=langs, removed_ids=self._ids[self._dead] if self._dead.any() else None)
        return orc.search(idx, q, int(k), min(int(nprobe), self.nlist), mask=mask, probes=lists)

    # -- persistence hooks used by GpuCollection.save / from_snapshot
    def save(self, path):
        import os

        os.makedirs(path, exist_ok=True)
        np.savez(os.path.join(path, "double.npz"), c=self.c if self.c is not None else np.zeros((0, self.dim), np.float32),
                 x=self._x, ids=self._ids, repo=self._repo, lang=self._lang, lst=self._list, dead=self._dead,
                 meta=np.array([self.dim, self.nlist, self.metric]))

    @classmethod
    def load(cls, path, device=0):
        import os

        z = np.load(os.path.join(path, "double.npz"))
        dim, nlist, metric = (int(v) for v in z["meta"])
        e = cls(dim, nlist, metric, device)
        if z["c"].shape[0]:
            e.c = z["c"]
        e._x, e._ids, e._repo, e._lang, e._list, e._dead = z["x"], z["ids"], z["repo"], z["lang"], z["lst"], z["dead"]
        return e


def merge_parts(parts, k, metric, device):
    """Stand-in for the device merge of (growing, sealed) partial results."""
    pd = np.stack([np.asarray(p[0]) for p in parts])
    pi = np.stack([np.asarray(p[1]) for p in parts])
    return orc.merge_topk(pd, pi, k, metric)
