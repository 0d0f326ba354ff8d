"""IVFFlatIndex -- Python host mirror of one sc_index (one B200).

This is the object the drop-in ``MilvusVectorStore`` (storage/milvus_store.py) drives in place of
the Milvus collection index declared at reference src/semcode/storage/milvus_store.py:76-83
(``IVF_FLAT``, ``metric_type`` IP, ``nlist``) and searched at :141-147 (``nprobe``).  All arithmetic
happens in libsemcode_ivf.so; torch is used for device buffers and streams only.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _capi
from ._capi import METRIC_IP, METRIC_L2, torch

Array = Union[np.ndarray, "torch.Tensor"]

# FAISS ClusteringParameters defaults the reference engine would use [EXT]
KMEANS_NITER = 25
KMEANS_SEED = 1234
KMEANS_MAX_POINTS_PER_CENTROID = 256
KMEANS_MIN_POINTS_PER_CENTROID = 39


def metric_code(metric) -> int:
    if metric in (METRIC_IP, "IP", "ip"):
        return METRIC_IP
    if metric in (METRIC_L2, "L2", "l2"):
        return METRIC_L2
    raise ValueError(f"unknown metric {metric!r} (expected 'IP' or 'L2')")


def kmeans_init_rows(n: int, nlist: int, seed: int) -> np.ndarray:
    """Rows that seed the centroids: a seeded random sample without replacement, sorted."""
    return np.sort(np.random.default_rng(seed).permutation(n)[:nlist]).astype(np.int64)


def kmeans_subsample_rows(n: int, nlist: int, max_points_per_centroid: int, seed: int) -> Optional[np.ndarray]:
    """FAISS trains on at most max_points_per_centroid*nlist rows [EXT]; None = use every row."""
    if max_points_per_centroid <= 0 or n <= max_points_per_centroid * nlist:
        return None
    rows = np.random.default_rng(seed + 1).permutation(n)[: max_points_per_centroid * nlist]
    return np.sort(rows).astype(np.int64)


def _is_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _as_rows(x: Array, dim: int, name: str) -> Array:
    """float32, C-contiguous [n, dim] without changing where the data lives."""
    if _is_tensor(x):
        if x.dim() == 1:
            x = x.unsqueeze(0)
        if x.dim() != 2 or x.shape[1] != dim:
            raise ValueError(f"{name}: expected shape [n, {dim}], got {tuple(x.shape)}")
        return x.to(torch.float32).contiguous()
    a = np.asarray(x, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2 or a.shape[1] != dim:
        raise ValueError(f"{name}: expected shape [n, {dim}], got {a.shape}")
    return np.ascontiguousarray(a)


def _as_vec(x, kind: str, n: int, name: str) -> Optional[Array]:
    if x is None:
        return None
    if _is_tensor(x):
        want = _capi._torch_dtype(kind)
        x = x.to(want).contiguous().reshape(-1)
    else:
        x = np.ascontiguousarray(np.asarray(x, dtype=_capi._NP[kind]).reshape(-1))
    if x.shape[0] != n:
        raise ValueError(f"{name}: expected {n} entries, got {x.shape[0]}")
    return x


@dataclass
class SearchTimes:
    coarse_ms: float
    probe_select_ms: float
    plan_ms: float
    scan_ms: float
    topk_ms: float
    total_ms: float
    scanned_rows: int
    unique_rows: int
    scan_launches: int
    total_launches: int


class IVFFlatIndex:
    """One IVF_FLAT index on one GPU (device ordinal ``device``)."""

    def __init__(self, dim: int, nlist: int = 128, metric="IP", device: int = 0):
        self._h = None
        L = _capi.lib()
        self.dim = int(dim)
        self.nlist = int(nlist)
        self.metric = metric_code(metric)
        self.device = int(device)
        h = C.c_void_p()
        _capi.check(L.sc_index_create(self.dim, self.metric, self.nlist, self.device, C.byref(h)))
        self._h = h
        self._L = L

    # -- lifecycle ------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._L.sc_index_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        _capi.check(self._L.sc_index_reset(self._h))

    def _stream(self) -> int:
        return _capi.current_stream(self.device)

    def _dev(self):
        return torch.device("cuda", self.device)

    # -- coarse quantizer -------------------------------------------------------------------------
    @property
    def is_trained(self) -> bool:
        return bool(self.stats().trained)

    def set_centroids(self, centroids: Array) -> None:
        c = _as_rows(centroids, self.dim, "centroids")
        if c.shape[0] != self.nlist:
            raise ValueError(f"centroids: expected {self.nlist} rows, got {c.shape[0]}")
        _capi.check(self._L.sc_index_set_centroids(self._h, _capi.ptr(c, "f32"), self.nlist, self._stream()))

    def get_centroids(self) -> np.ndarray:
        out = np.empty((self.nlist, self.dim), dtype=np.float32)
        _capi.check(self._L.sc_index_get_centroids(self._h, _capi.ptr(out, "f32"), self._stream()))
        return out

    def train(
        self,
        x: Array,
        niter: int = KMEANS_NITER,
        seed: int = KMEANS_SEED,
        max_points_per_centroid: int = KMEANS_MAX_POINTS_PER_CENTROID,
        init_centroids: Optional[Array] = None,
    ) -> List[float]:
        """Lloyd k-means on x (FAISS Clustering::train restated); returns the objective entering
        each iteration.  Subsampling and init rows follow oracle/ivf_numpy.py exactly."""
        x = _as_rows(x, self.dim, "x")
        n = x.shape[0]
        if n < self.nlist:
            raise ValueError(f"need at least nlist={self.nlist} training rows, got {n}")
        rows = kmeans_subsample_rows(n, self.nlist, max_points_per_centroid, seed)
        if rows is not None:
            if _is_tensor(x):
                x = x.index_select(0, torch.from_numpy(rows).to(x.device))
            else:
                x = np.ascontiguousarray(x[rows])
            n = x.shape[0]
        obj = np.zeros(max(niter, 1), dtype=np.float64)
        if init_centroids is not None:
            self.set_centroids(init_centroids)
            return self._lloyd(x, niter)
        init = kmeans_init_rows(n, self.nlist, seed)
        _capi.check(
            self._L.sc_index_train(
                self._h, _capi.ptr(x, "f32"), n, int(niter), _capi.ptr(init, "i64"), _capi.ptr(obj, "f64"),
                self._stream(),
            )
        )
        return [float(v) for v in obj[:niter]]

    def _lloyd(self, x: Array, niter: int) -> List[float]:
        """Lloyd iterations from the current centroids, driven from Python through the
        kmeans_step / kmeans_update building blocks (the same ones the sharded trainer uses)."""
        sums, counts, obj = self.kmeans_buffers()
        out = []
        for _ in range(niter):
            sums.zero_()
            counts.zero_()
            obj.zero_()
            self.kmeans_step(x, sums, counts, obj)
            out.append(float(obj.item()))
            self.kmeans_update(sums, counts)
        return out

    def tensor_device(self):
        """Where this engine wants its tensors (ShardedIVFFlat asks the engine, not torch.cuda)."""
        return self._dev()

    def kmeans_buffers(self):
        """Zeroed accumulators for kmeans_step: sums [nlist*dim_padded] fp64, counts [nlist] int32, objective [1]."""
        dev = self._dev()
        ds = self.stats().dim_padded
        return (torch.zeros(self.nlist * ds, dtype=torch.float64, device=dev),
                torch.zeros(self.nlist, dtype=torch.int32, device=dev),
                torch.zeros(1, dtype=torch.float64, device=dev))

    def kmeans_step(self, x: Array, sums, counts, obj) -> None:
        x = _as_rows(x, self.dim, "x")
        _capi.check(
            self._L.sc_index_kmeans_step(
                self._h, _capi.ptr(x, "f32"), x.shape[0], _capi.ptr(sums, "f64"), _capi.ptr(counts, "i32"),
                _capi.ptr(obj, "f64"), self._stream(),
            )
        )

    def kmeans_update(self, sums, counts) -> int:
        ns = C.c_int32(0)
        _capi.check(
            self._L.sc_index_kmeans_update(
                self._h, _capi.ptr(sums, "f64"), _capi.ptr(counts, "i32"), C.addressof(ns), self._stream()
            )
        )
        return int(ns.value)

    def assign(self, x: Array) -> Array:
        x = _as_rows(x, self.dim, "x")
        n = x.shape[0]
        if _is_tensor(x) and x.is_cuda:
            out = torch.empty(n, dtype=torch.int32, device=x.device)
        else:
            out = np.empty(n, dtype=np.int32)
        _capi.check(self._L.sc_index_assign(self._h, _capi.ptr(x, "f32"), n, _capi.ptr(out, "i32"), self._stream()))
        return out

    def probe(self, q: Array, nprobe: int, with_scores: bool = False):
        q = _as_rows(q, self.dim, "q")
        nq = q.shape[0]
        nprobe = min(int(nprobe), self.nlist)
        if _is_tensor(q) and q.is_cuda:
            lists = torch.empty((nq, nprobe), dtype=torch.int32, device=q.device)
            scores = torch.empty((nq, nprobe), dtype=torch.float32, device=q.device) if with_scores else None
        else:
            lists = np.empty((nq, nprobe), dtype=np.int32)
            scores = np.empty((nq, nprobe), dtype=np.float32) if with_scores else None
        _capi.check(
            self._L.sc_index_probe(
                self._h, _capi.ptr(q, "f32"), nq, nprobe, _capi.ptr(lists, "i32"), _capi.ptr(scores, "f32"),
                self._stream(),
            )
        )
        return (lists, scores) if with_scores else lists

    # -- insert / remove ----------------------------------------------------------------------------
    def add(
        self,
        x: Array,
        ids: Array,
        repo_tags: Optional[Array] = None,
        lang_tags: Optional[Array] = None,
        lists: Optional[Array] = None,
    ) -> None:
        """Append rows (FAISS add_with_ids).  ``lists`` pre-assigns the inverted list of every row."""
        x = _as_rows(x, self.dim, "x")
        n = x.shape[0]
        ids = _as_vec(ids, "i64", n, "ids")
        repo_tags = _as_vec(repo_tags, "u32", n, "repo_tags")
        lang_tags = _as_vec(lang_tags, "u8", n, "lang_tags")
        lists = _as_vec(lists, "i32", n, "lists")
        args = [self._h, _capi.ptr(x, "f32"), _capi.ptr(ids, "i64"), _capi.ptr(repo_tags, "u32"),
                _capi.ptr(lang_tags, "u8")]
        if lists is None:
            _capi.check(self._L.sc_index_add(*args, n, self._stream()))
        else:
            _capi.check(self._L.sc_index_add_preassigned(*args, _capi.ptr(lists, "i32"), n, self._stream()))

    def remove_ids(self, ids: Array) -> int:
        ids = _as_vec(ids, "i64", len(ids), "ids")
        out = C.c_int64(0)
        _capi.check(
            self._L.sc_index_remove_ids(self._h, _capi.ptr(ids, "i64"), ids.shape[0], C.addressof(out), self._stream())
        )
        return int(out.value)

    # -- search ---------------------------------------------------------------------------------------
    @staticmethod
    def _filter(repos: Optional[Iterable[int]], langs: Optional[Iterable[int]]):
        if repos is None and langs is None:
            return None, ()
        f = _capi.ScFilter()
        keep = []
        if repos is not None:
            r = np.ascontiguousarray(np.asarray(list(repos), dtype=np.uint32))
            if r.size == 0:
                raise ValueError("repos filter is empty: pass None for 'any repo'")
            f.repo_tags = r.ctypes.data_as(C.POINTER(C.c_uint32))
            f.n_repos = r.size
            keep.append(r)
        if langs is not None:
            l = np.ascontiguousarray(np.asarray(list(langs), dtype=np.uint8))
            if l.size == 0:
                raise ValueError("langs filter is empty: pass None for 'any language'")
            f.lang_tags = l.ctypes.data_as(C.POINTER(C.c_uint8))
            f.n_langs = l.size
            keep.append(l)
        return f, keep

    def search(
        self,
        q: Array,
        k: int,
        nprobe: int = 16,
        repos: Optional[Iterable[int]] = None,
        langs: Optional[Iterable[int]] = None,
        lists: Optional[Array] = None,
        out: Optional[Tuple[Array, Array]] = None,
        exchange: Optional["PeerExchange"] = None,
    ) -> Tuple[Array, Array]:
        """Top-k per query: (dist [nq,k] fp32, ids [nq,k] int64), best first, id -1 = no result.

        CUDA-tensor queries give CUDA-tensor results without synchronising the stream; numpy / CPU
        queries give numpy results (synchronous).  With `exchange` this index is one shard of a row-sharded
        index and the call is one fused step (sc_index_search_sharded): every rank passes the same queries
        and receives the same merged result."""
        q = _as_rows(q, self.dim, "q")
        nq = q.shape[0]
        k = int(k)
        on_dev = _is_tensor(q) and q.is_cuda
        if out is not None:
            dist, ids = out
        elif on_dev:
            dist = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            ids = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        else:
            dist = np.empty((nq, k), dtype=np.float32)
            ids = np.empty((nq, k), dtype=np.int64)
        f, keep = self._filter(repos, langs)
        fp = C.byref(f) if f is not None else None
        if lists is not None:
            if _is_tensor(lists):
                lists = lists.to(torch.int32).contiguous()
            else:
                lists = np.ascontiguousarray(np.asarray(lists, dtype=np.int32))
            if lists.shape[0] != nq:
                raise ValueError("lists: one row of probes per query expected")
        if exchange is not None:
            _capi.check(
                self._L.sc_index_search_sharded(
                    self._h, exchange.handle, _capi.ptr(q, "f32"), nq, k,
                    int(nprobe) if lists is None else int(lists.shape[1]), _capi.ptr(lists, "i32") if lists is not None else None,
                    fp, _capi.ptr(dist, "f32"), _capi.ptr(ids, "i64"), self._stream(),
                )
            )
        elif lists is None:
            _capi.check(
                self._L.sc_index_search(
                    self._h, _capi.ptr(q, "f32"), nq, k, int(nprobe), fp, _capi.ptr(dist, "f32"),
                    _capi.ptr(ids, "i64"), self._stream(),
                )
            )
        else:
            _capi.check(
                self._L.sc_index_search_preassigned(
                    self._h, _capi.ptr(q, "f32"), nq, k, int(lists.shape[1]), _capi.ptr(lists, "i32"), fp,
                    _capi.ptr(dist, "f32"), _capi.ptr(ids, "i64"), self._stream(),
                )
            )
        del keep
        return dist, ids

    # -- introspection ------------------------------------------------------------------------------
    def stats(self) -> _capi.ScStats:
        s = _capi.ScStats()
        _capi.check(self._L.sc_index_stats(self._h, C.byref(s)))
        return s

    @property
    def ntotal(self) -> int:
        return int(self.stats().ntotal)

    def list_sizes(self) -> np.ndarray:
        out = np.empty(self.nlist, dtype=np.int32)
        _capi.check(self._L.sc_index_list_sizes(self._h, _capi.ptr(out, "i32")))
        return out

    def export_list(self, l: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(vectors [len,dim], ids [len], tags [len]) of one inverted list, slot order."""
        n = int(self.list_sizes()[l])
        vec = np.empty((n, self.dim), dtype=np.float32)
        ids = np.empty(n, dtype=np.int64)
        tags = np.empty(n, dtype=np.uint32)
        ln = C.c_int64(0)
        _capi.check(
            self._L.sc_index_export_list(
                self._h, int(l), n, _capi.ptr(vec, "f32"), _capi.ptr(ids, "i64"), _capi.ptr(tags, "u32"),
                C.addressof(ln), self._stream(),
            )
        )
        return vec, ids, tags

    def export_lists(self, list_begin: int, list_end: int, device: bool = False):
        """Lists [list_begin, list_end) back to back in slot order, ONE call (sc_index_export_lists):
        (off [n + 1] int64, vecs [rows, dim], ids [rows], tags [rows]) as numpy arrays, or CUDA tensors with
        ``device=True`` (tags as int32 bits).  Tombstoned slots are included (tag bit 31)."""
        n = int(list_end) - int(list_begin)
        off = np.zeros(n + 1, dtype=np.int64)
        _capi.check(self._L.sc_index_export_lists(self._h, int(list_begin), int(list_end), 0, None, None, None,
                                                  _capi.ptr(off, "i64"), self._stream()))
        rows = int(off[-1])
        if device:
            dev = self._dev()
            vecs = torch.empty((rows, self.dim), dtype=torch.float32, device=dev)
            ids = torch.empty(rows, dtype=torch.int64, device=dev)
            tags = torch.empty(rows, dtype=torch.int32, device=dev)
        else:
            vecs = np.empty((rows, self.dim), dtype=np.float32)
            ids = np.empty(rows, dtype=np.int64)
            tags = np.empty(rows, dtype=np.uint32)
        if rows:
            _capi.check(self._L.sc_index_export_lists(self._h, int(list_begin), int(list_end), rows, _capi.ptr(vecs, "f32"),
                                                      _capi.ptr(ids, "i64"), _capi.ptr(tags, "u32"), None, self._stream()))
        return off, vecs, ids, tags

    def list_ranges(self, max_bytes: int = 512 << 20):
        """Consecutive list ranges of at most ~max_bytes of vectors each (bulk export / re-insert in bounded steps)."""
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

    def export_csr(self, live_only: bool = True):
        """Whole index as CSR on the host: (list_off [nlist+1], vecs [n,dim], ids [n], tags [n]).
        Used by persistence and to hand the same lists to the CPU baseline.  A few bulk calls
        (sc_index_export_lists over ~512 MB ranges), not one call per list."""
        sizes = self.list_sizes().astype(np.int64)
        total = int(sizes.sum())
        vecs = np.empty((total, self.dim), dtype=np.float32)
        ids = np.empty(total, dtype=np.int64)
        tags = np.empty(total, dtype=np.uint32)
        off = np.zeros(self.nlist + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        st = self._stream()
        for lo, hi in self.list_ranges():
            a, b = int(off[lo]), int(off[hi])
            if b == a:
                continue
            _capi.check(
                self._L.sc_index_export_lists(
                    self._h, lo, hi, b - a, vecs[a:b].ctypes.data, ids[a:b].ctypes.data, tags[a:b].ctypes.data, None, st,
                )
            )
        if live_only and total:
            live = (tags & np.uint32(0x80000000)) == 0
            if not live.all():
                row_list = np.repeat(np.arange(self.nlist), sizes)
                new_sizes = np.bincount(row_list[live], minlength=self.nlist).astype(np.int64)
                vecs, ids, tags = vecs[live], ids[live], tags[live]
                off = np.zeros(self.nlist + 1, dtype=np.int64)
                np.cumsum(new_sizes, out=off[1:])
        return off, vecs, ids, tags

    def compact(self) -> int:
        """Drop the tombstoned slots of every list in place (sc_index_compact); returns the pages freed."""
        out = C.c_int64(0)
        _capi.check(self._L.sc_index_compact(self._h, C.addressof(out), self._stream()))
        return int(out.value)

    # -- persistence ----------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        """Write the index as memory-mappable .npy files under directory `path`
        (centroids, CSR lists with live rows only, ids, tags) plus meta.json."""
        import json
        import os

        os.makedirs(path, exist_ok=True)
        off, vecs, ids, tags = self.export_csr(live_only=True)
        trained = self.is_trained
        np.save(os.path.join(path, "centroids.npy"), self.get_centroids() if trained else np.zeros((0, self.dim), np.float32))
        np.save(os.path.join(path, "list_off.npy"), off)
        np.save(os.path.join(path, "vecs.npy"), vecs)
        np.save(os.path.join(path, "ids.npy"), ids)
        np.save(os.path.join(path, "tags.npy"), tags)
        meta = {"format": 1, "dim": self.dim, "nlist": self.nlist, "metric": self.metric, "ntotal": int(off[-1]),
                "trained": bool(trained)}
        tmp = os.path.join(path, "meta.json.tmp")
        with open(tmp, "w") as f:
            json.dump(meta, f)
        os.replace(tmp, os.path.join(path, "meta.json"))  # meta.json last: its presence marks a complete snapshot

    @classmethod
    def load(cls, path: str, device: int = 0, chunk_rows: int = 1 << 18) -> "IVFFlatIndex":
        import json
        import os

        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") != 1:
            raise ValueError(f"unknown snapshot format {meta.get('format')!r}")
        idx = cls(meta["dim"], nlist=meta["nlist"], metric=meta["metric"], device=device)
        if not meta["trained"]:
            return idx
        idx.set_centroids(np.load(os.path.join(path, "centroids.npy")))
        off = np.load(os.path.join(path, "list_off.npy"))
        vecs = np.load(os.path.join(path, "vecs.npy"), mmap_mode="r")
        ids = np.load(os.path.join(path, "ids.npy"), mmap_mode="r")
        tags = np.load(os.path.join(path, "tags.npy"), mmap_mode="r")
        if not (int(off[-1]) == vecs.shape[0] == ids.shape[0] == tags.shape[0] == int(meta["ntotal"])) or off.shape[0] != meta["nlist"] + 1:
            raise ValueError(f"snapshot {path!r} is inconsistent: list_off ends at {int(off[-1])}, vecs {vecs.shape[0]}, "
                             f"ids {ids.shape[0]}, tags {tags.shape[0]}, meta {meta['ntotal']}")
        lists = np.repeat(np.arange(meta["nlist"], dtype=np.int32), np.diff(off))
        for s in range(0, int(off[-1]), chunk_rows):
            e = min(int(off[-1]), s + chunk_rows)
            t = np.asarray(tags[s:e])
            idx.add(np.ascontiguousarray(vecs[s:e]), np.ascontiguousarray(ids[s:e]),
                    (t >> np.uint32(8)) & np.uint32((1 << 23) - 1), (t & np.uint32(0xFF)).astype(np.uint8), lists=lists[s:e])
        return idx

    # -- profiling / tuning -------------------------------------------------------------------------
    def set_profiling(self, enabled: bool) -> None:
        _capi.check(self._L.sc_index_set_profiling(self._h, 1 if enabled else 0))

    def last_search_times(self) -> SearchTimes:
        t = _capi.ScSearchTimes()
        _capi.check(self._L.sc_index_last_search_times(self._h, C.byref(t)))
        return SearchTimes(t.coarse_ms, t.probe_select_ms, t.plan_ms, t.scan_ms, t.topk_ms, t.total_ms,
                           int(t.scanned_rows), int(t.unique_rows), int(t.scan_launches), int(t.total_launches))

    def set_param(self, name: str, value: int) -> None:
        _capi.check(self._L.sc_index_set_param(self._h, name.encode(), int(value)))


class PeerExchange:
    """Peer-mapped exchange buffers of one rank (sc_exchange_t): the fused scatter/gather of a sharded search.

    torch is the plumbing: `torch.distributed._symmetric_memory` allocates one buffer per rank and maps every
    peer's buffer into this process over NVLink; the kernels of libsemcode_ivf store into / spin on those
    mappings directly (no NCCL call on the data path).  Raises when symmetric memory cannot be set up; the
    caller (ShardedIVFFlat) then keeps the NCCL all-gather + merge route."""

    def __init__(self, device: int, group=None, nbytes: int = 64 << 20):
        import torch.distributed as dist

        self.device = int(device)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("PeerExchange covers the GPUs of one box (world <= 8)")
        dev = torch.device("cuda", self.device)
        # every rank must take the same route: agree on whether symmetric memory came up everywhere
        ptrs, err = None, None
        try:
            ptrs = self._symmetric(dev, group, int(nbytes))
        except Exception as e:  # noqa: BLE001
            err = f"{type(e).__name__}: {e}"
        oks = [None] * self.world
        dist.all_gather_object(oks, ptrs is not None, group=group)
        if not all(oks):
            # CUDA IPC: each rank allocates its buffer and opens the peers' (cudaIpcOpenMemHandle underneath torch's tensor
            # sharing).  This is what works when two ranks share one device -- torch's symmetric memory refuses that -- and on
            # boxes without fabric handles; the C ABI only ever sees plain pointers.
            self._hdl = None
            try:
                ptrs = self._ipc(dev, group, int(nbytes))
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(f"no peer-mapped buffers: symmetric memory: {err}; CUDA IPC: {type(e).__name__}: {e}") from e
            self.transport = "cuda-ipc"
        else:
            self.transport = "symmetric-memory"
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)  # every rank's flags are zero before anybody stores into them
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("rendezvous returned no peer pointers")
        arr = (C.c_void_p * self.world)(*ptrs)
        h = C.c_void_p()
        _capi.check(_capi.lib().sc_exchange_create(self.rank, self.world, arr, int(nbytes), self.device, C.byref(h)))
        self.handle = h
        self.nbytes = int(nbytes)
        self._group = group

    def _symmetric(self, dev, group, nbytes: int):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        g = group if group is not None else dist.group.WORLD
        self._buf = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        self._hdl = symm.rendezvous(self._buf, g)
        self._buf.zero_()
        return [int(p) for p in self._hdl.buffer_ptrs]

    def _ipc(self, dev, group, nbytes: int):
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor

        self._buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        handles = [None] * self.world
        dist.all_gather_object(handles, reduce_tensor(self._buf), group=group)
        self._peers = []
        ptrs = []
        for r, (rebuild, args) in enumerate(handles):
            t = self._buf if r == self.rank else rebuild(*args)
            self._peers.append(t)  # keeps the mapping alive
            ptrs.append(int(t.data_ptr()))
        return ptrs

    def status(self) -> Tuple[bool, int]:
        """(timed_out, steps issued); synchronises the device."""
        t, e = C.c_int32(0), C.c_int64(0)
        _capi.check(_capi.lib().sc_exchange_status(self.handle, C.byref(t), C.byref(e)))
        return bool(t.value), int(e.value)

    def poll(self) -> bool:
        """True once a finished step gave up waiting for a peer (or failed midway); does NOT synchronise."""
        t = C.c_int32(0)
        _capi.check(_capi.lib().sc_exchange_poll(self.handle, C.byref(t)))
        return bool(t.value)

    def set_timeout_ms(self, ms: int) -> None:
        _capi.check(_capi.lib().sc_exchange_set_timeout_ms(self.handle, int(ms)))

    def close(self) -> None:
        if getattr(self, "handle", None):
            _capi.lib().sc_exchange_destroy(self.handle)
            self.handle = None
        self._peers = None  # drop the mappings of the peers' buffers before this rank's own buffer goes away

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def merge_topk(part_dist, part_ids, k: int, metric, device: Optional[int] = None):
    """Device merge of partial results [parts, nq, kin] -> ([nq,k], [nq,k]) (cross-shard reduce)."""
    if not (_is_tensor(part_dist) and part_dist.is_cuda):
        raise TypeError("merge_topk works on CUDA tensors")
    parts, nq, kin = part_dist.shape
    part_dist = part_dist.to(torch.float32).contiguous()
    part_ids = part_ids.to(torch.int64).contiguous()
    dev = part_dist.device.index if device is None else device
    out_d = torch.empty((nq, k), dtype=torch.float32, device=part_dist.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=part_dist.device)
    _capi.check(
        _capi.lib().sc_merge_topk(
            _capi.ptr(part_dist, "f32"), _capi.ptr(part_ids, "i64"), parts, nq, kin, int(k), metric_code(metric),
            _capi.ptr(out_d, "f32"), _capi.ptr(out_i, "i64"), dev, _capi.current_stream(dev),
        )
    )
    return out_d, out_i
