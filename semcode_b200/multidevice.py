"""MultiDeviceIVFFlat -- one IVF_FLAT index row-sharded over several GPUs of the box, driven by ONE process.

This is the backend the drop-in store uses when `ivf_devices` names more than one device:
the FastAPI / CLI processes of the reference are single processes, so the collection behind
`MilvusVectorStore` cannot rely on a torchrun launch (that is `ShardedIVFFlat`, one process per GPU, with the fused
peer-memory exchange).  Same partitioning as there: centroids replicated, every inverted list's rows dealt round-robin
to the shards with global int64 ids, so the union of the shards IS the single index and the merged result equals the
1-GPU result (exact-tie order aside).

  search  every shard runs coarse pass + scan + top-k on its own device and stream (asynchronously: device tensors in
          and out, no host sync in between), the [nq, k] partials are copied to the first device over NVLink and reduced
          there by sc_merge_topk -- the Milvus proxy's reduce over query nodes [EXT]
  train   data-parallel Lloyd: kmeans_step per shard, the fp64 sums / counts / objective added on the first device,
          the identical kmeans_update on every shard
  save    one sub-directory per shard (bulk sc_index_export_lists underneath), `shards.json` names the layout

It implements the engine surface GpuCollection drives (the same calls as IVFFlatIndex).
"""

from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence

import numpy as np

from ._capi import torch
from .index import KMEANS_MAX_POINTS_PER_CENTROID, KMEANS_NITER, KMEANS_SEED, IVFFlatIndex, kmeans_init_rows, \
    kmeans_subsample_rows, merge_topk, metric_code


class _Stats:
    def __init__(self, parts):
        first = parts[0]
        self.dim, self.dim_padded, self.metric, self.nlist, self.trained = (first.dim, first.dim_padded, first.metric, first.nlist,
                                                                            first.trained)
        for name in ("ntotal", "nremoved", "npages", "nfree_pages", "bytes_lists", "bytes_scratch"):
            setattr(self, name, sum(int(getattr(p, name)) for p in parts))
        self.max_list_len = max(p.max_list_len for p in parts)
        self.min_list_len = min(p.min_list_len for p in parts)


class MultiDeviceIVFFlat:
    def __init__(self, dim: int, nlist: int = 128, metric="IP", devices: Sequence[int] = (0,)):
        if not devices:
            raise ValueError("devices must name at least one CUDA ordinal")
        self.dim, self.nlist, self.metric = int(dim), int(nlist), metric_code(metric)
        self.devices = [int(d) for d in devices]
        self.shards: List[IVFFlatIndex] = [IVFFlatIndex(dim, nlist=nlist, metric=metric, device=d) for d in self.devices]
        self._next_row = 0  # global round-robin cursor of the deal

    @property
    def device(self) -> int:
        return self.devices[0]

    def _dev(self, s: int):
        return torch.device("cuda", self.devices[s])

    def close(self) -> None:
        for sh in self.shards:
            sh.close()

    def reset(self) -> None:
        for sh in self.shards:
            sh.reset()
        self._next_row = 0

    # -- coarse quantizer -------------------------------------------------------------------------
    def set_centroids(self, centroids) -> None:
        c = centroids.detach().cpu().numpy() if torch.is_tensor(centroids) else np.asarray(centroids, dtype=np.float32)
        for sh in self.shards:
            sh.set_centroids(c)

    def get_centroids(self) -> np.ndarray:
        return self.shards[0].get_centroids()

    @property
    def is_trained(self) -> bool:
        return self.shards[0].is_trained

    def train(self, x, niter: int = KMEANS_NITER, seed: int = KMEANS_SEED,
              max_points_per_centroid: int = KMEANS_MAX_POINTS_PER_CENTROID, init_centroids=None) -> List[float]:
        """Data-parallel Lloyd over the shards' devices; same subsample / init rows as IVFFlatIndex.train."""
        if torch.is_tensor(x):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        if n < self.nlist:
            raise ValueError(f"need at least nlist={self.nlist} training rows, got {n}")
        rows = kmeans_subsample_rows(n, self.nlist, max_points_per_centroid, seed)
        if rows is not None:
            x = np.ascontiguousarray(x[rows])
            n = x.shape[0]
        self.set_centroids(x[kmeans_init_rows(n, self.nlist, seed)] if init_centroids is None else init_centroids)
        g = len(self.shards)
        parts = [torch.from_numpy(np.ascontiguousarray(x[s::g])).to(self._dev(s)) for s in range(g)]
        bufs = [sh.kmeans_buffers() for sh in self.shards]
        dev0 = self._dev(0)
        out = []
        for _ in range(niter):
            for s, sh in enumerate(self.shards):
                for b in bufs[s]:
                    b.zero_()
                if parts[s].shape[0]:
                    sh.kmeans_step(parts[s], *bufs[s])
            sums = bufs[0][0].clone()
            counts = bufs[0][1].clone()
            obj = bufs[0][2].clone()
            for s in range(1, g):  # the one exchange of an iteration: nlist * (8 ds + 4) + 8 bytes per shard over NVLink
                sums += bufs[s][0].to(dev0)
                counts += bufs[s][1].to(dev0)
                obj += bufs[s][2].to(dev0)
            out.append(float(obj.item()))
            for s, sh in enumerate(self.shards):
                sh.kmeans_update(sums.to(self._dev(s)), counts.to(self._dev(s)))
        return out

    # -- rows ---------------------------------------------------------------------------------------
    def add(self, x, ids, repo_tags=None, lang_tags=None, lists=None) -> None:
        """Deal the batch round-robin by global arrival number: shard s keeps the rows congruent to s."""
        n = int(x.shape[0])
        g = len(self.shards)
        for s, sh in enumerate(self.shards):
            first = (s - self._next_row) % g
            if first >= n:
                continue
            sl = slice(first, n, g)

            def take(a, s=s, sl=sl):
                if a is None:
                    return None
                a = a[sl]
                if torch.is_tensor(a):
                    return a.to(self._dev(s)).contiguous() if a.is_cuda else a.contiguous()
                return np.ascontiguousarray(a)

            sh.add(take(x), take(ids), take(repo_tags), take(lang_tags), lists=take(lists))
        self._next_row = (self._next_row + n) % g

    def remove_ids(self, ids) -> int:
        return sum(sh.remove_ids(ids) for sh in self.shards)

    def compact(self) -> int:
        return sum(sh.compact() for sh in self.shards)

    @property
    def ntotal(self) -> int:
        return sum(sh.ntotal for sh in self.shards)

    def stats(self) -> _Stats:
        return _Stats([sh.stats() for sh in self.shards])

    def list_sizes(self) -> np.ndarray:
        return np.sum([sh.list_sizes() for sh in self.shards], axis=0).astype(np.int32)

    def list_ranges(self, max_bytes: int = 512 << 20):
        sizes = self.list_sizes().astype(np.int64)
        per_row = self.dim * 4 + 12
        out, lo, acc = [], 0, 0
        for l in range(self.nlist):
            b = int(sizes[l]) * per_row
            if acc and acc + b > max_bytes:
                out.append((lo, l))
                lo, acc = l, 0
            acc += b
        out.append((lo, self.nlist))
        return out

    def export_lists(self, list_begin: int, list_end: int, device: bool = False):
        """Lists [list_begin, list_end) of the WHOLE index (each list: shard 0's slots, then shard 1's, ...), host arrays."""
        parts = [sh.export_lists(list_begin, list_end) for sh in self.shards]
        n = list_end - list_begin
        off = np.zeros(n + 1, dtype=np.int64)
        for p in parts:
            off[1:] += np.diff(p[0])
        off = np.concatenate([[0], np.cumsum(off[1:])]).astype(np.int64)
        vecs = np.empty((int(off[-1]), self.dim), dtype=np.float32)
        ids = np.empty(int(off[-1]), dtype=np.int64)
        tags = np.empty(int(off[-1]), dtype=np.uint32)
        cur = off[:-1].copy()
        for po, pv, pi, pt in parts:
            for l in range(n):
                a, b = int(po[l]), int(po[l + 1])
                c = int(cur[l])
                vecs[c:c + b - a], ids[c:c + b - a], tags[c:c + b - a] = pv[a:b], pi[a:b], pt[a:b]
                cur[l] += b - a
        return off, vecs, ids, tags

    # -- search ---------------------------------------------------------------------------------------
    def search(self, q, k: int, nprobe: int = 16, repos=None, langs=None, lists=None, out=None, exchange=None):
        on_dev = torch.is_tensor(q) and q.is_cuda
        if torch.is_tensor(q):
            qt = q.to(torch.float32)
        else:
            qt = torch.from_numpy(np.ascontiguousarray(np.asarray(q, dtype=np.float32)))
        if qt.dim() == 1:
            qt = qt.unsqueeze(0)
        dev0 = self._dev(0)
        pd, pi = [], []
        for s, sh in enumerate(self.shards):  # asynchronous launches, one device after the other
            qs = qt.to(self._dev(s), non_blocking=True).contiguous()
            ls = None if lists is None else torch.as_tensor(lists).to(self._dev(s))
            with torch.cuda.device(self._dev(s)):
                d, i = sh.search(qs, k, nprobe=nprobe, repos=repos, langs=langs, lists=ls)
            pd.append(d)
            pi.append(i)
        with torch.cuda.device(dev0):
            gd = torch.stack([d.to(dev0) for d in pd])
            gi = torch.stack([i.to(dev0) for i in pi])
            md, mi = merge_topk(gd, gi, k, self.metric, self.devices[0])
        if on_dev:
            return md.to(q.device), mi.to(q.device)
        return md.cpu().numpy(), mi.cpu().numpy()

    # -- persistence ----------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        for s, sh in enumerate(self.shards):
            sh.save(os.path.join(path, f"shard-{s:02d}"))
        tmp = os.path.join(path, "shards.json.tmp")
        with open(tmp, "w") as f:
            json.dump({"format": 1, "shards": len(self.shards), "dim": self.dim, "nlist": self.nlist, "metric": self.metric,
                       "next_row": self._next_row}, f)
        os.replace(tmp, os.path.join(path, "shards.json"))

    @classmethod
    def load(cls, path: str, devices: Optional[Sequence[int]] = None) -> "MultiDeviceIVFFlat":
        with open(os.path.join(path, "shards.json")) as f:
            meta = json.load(f)
        g = int(meta["shards"])
        if not devices:
            devices = list(range(g))
        if len(devices) != g:
            raise ValueError(f"snapshot {path!r} holds {g} shards, {len(devices)} devices were configured")
        self = cls.__new__(cls)
        self.dim, self.nlist, self.metric = int(meta["dim"]), int(meta["nlist"]), int(meta["metric"])
        self.devices = [int(d) for d in devices]
        self.shards = [IVFFlatIndex.load(os.path.join(path, f"shard-{s:02d}"), device=d) for s, d in enumerate(self.devices)]
        self._next_row = int(meta.get("next_row", 0))
        return self
