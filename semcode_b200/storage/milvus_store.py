"""Drop-in replacement for reference src/semcode/storage/milvus_store.py.

Same class name, constructor, methods, error strings and log events as the reference wrapper
(milvus_store.py:29-148), but the Milvus server behind it is replaced by an in-process collection
whose IVF_FLAT index lives on a B200 (semcode_b200.IVFFlatIndex -> libsemcode_ivf.so):

  reference call (milvus_store.py)                         here
  ------------------------------------------------------   -----------------------------------------
  connections.connect + utility.has_collection (:42-54)    process-wide collection registry (+ optional
                                                           on-disk snapshot under SEMCODE_IVF_PERSIST_DIR)
  CollectionSchema(7 fields) + create_index(IVF_FLAT, IP,  GpuCollection(dim, nlist=128, metric="IP")
    nlist=128) + load()  (:59-84)
  Collection.upsert([7 columns])  (:128-130)               GpuCollection.upsert -> remove_ids + add
  Collection.search(data=[vector], nprobe=16, limit)       GpuCollection.search -> sc_index_search
    (:141-147)

Rows not yet covered by a trained index sit in a *growing segment* that is searched exactly (a
one-list index), as Milvus does for unsealed data [EXT]; once `seal_rows` rows exist (default
39 x nlist, FAISS' min_points_per_centroid) k-means runs and later rows are added straight to
the inverted lists (FAISS add).  Scalar columns (id, repo, path, language, text, metadata) stay on
the host; repo and language are additionally encoded as integer tags next to each vector so that
`repos=` / `languages=` filters are evaluated inside the list scan.

Keyword-only extensions (absent from the reference, defaults reproduce it): nprobe, repos,
languages on search(); search_batch(); search_arrays(); upsert_arrays(); build_index().
"""

from __future__ import annotations

import atexit
import json
import os
import shutil
import threading
import time
import weakref
from contextlib import contextmanager
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np

from ..index import KMEANS_MIN_POINTS_PER_CENTROID, IVFFlatIndex, merge_topk, metric_code
from .._capi import METRIC_IP, torch

try:  # the reference's own settings / logger when the drop-in runs inside semcode
    from semcode.settings import settings  # type: ignore
except Exception:  # pragma: no cover - exercised when semcode is not importable

    class _EnvSettings:
        """Minimal stand-in for semcode.settings.AppSettings (settings.py:30-82): SEMCODE_* env."""

        def __getattr__(self, name: str):
            env = os.environ.get("SEMCODE_" + name.upper())
            defaults = {
                "embedding_dimension": 3072,  # settings.py:47
                "milvus_uri": "http://localhost:19530",  # settings.py:40
                "milvus_upsert_batch_size": 128,  # settings.py:76
            }
            if env is None:
                if name in defaults:
                    return defaults[name]
                raise AttributeError(name)
            d = defaults.get(name)
            return type(d)(env) if d is not None else env

    settings = _EnvSettings()

try:
    from semcode.logger import get_logger  # type: ignore
except Exception:  # pragma: no cover
    import logging

    class _KVLogger:
        def __init__(self, name: str):
            self._l = logging.getLogger(name)

        def _fmt(self, event: str, kw: Dict[str, Any]) -> str:
            return event + "".join(f" {k}={v}" for k, v in kw.items())

        def info(self, event: str, **kw):
            self._l.info(self._fmt(event, kw))

        def warning(self, event: str, **kw):
            self._l.warning(self._fmt(event, kw))

        def error(self, event: str, **kw):
            self._l.error(self._fmt(event, kw))

    def get_logger(name: str):
        return _KVLogger(name)


log = get_logger(__name__)

OUTPUT_FIELDS = ("repo", "path", "language", "text", "metadata")  # milvus_store.py:146


def _setting(name: str, default):
    """Optional knob: settings.<name> if present (AppSettings has extra='allow', settings.py:36)."""
    try:
        v = getattr(settings, name)
    except Exception:
        return default
    if v is None:
        return default
    try:
        return type(default)(v) if default is not None else v
    except Exception:
        return default


# --------------------------------------------------------------------------------------------------
# result objects: the shape SemanticSearchPipeline._hit_to_document reads (rag/pipeline.py:133-169)
# --------------------------------------------------------------------------------------------------
class Entity:
    __slots__ = ("_fields",)

    def __init__(self, fields: Dict[str, Any]):
        self._fields = fields

    def get(self, name: str, default=None):
        return self._fields.get(name, default)

    def __getitem__(self, name: str):
        return self._fields[name]

    def to_dict(self) -> Dict[str, Any]:
        return dict(self._fields)

    def __repr__(self) -> str:
        return f"Entity({self._fields!r})"


class Hit:
    __slots__ = ("id", "distance", "entity")

    def __init__(self, pk: str, distance: float, fields: Dict[str, Any]):
        self.id = pk
        self.distance = float(distance)
        self.entity = Entity(fields)

    @property
    def score(self) -> float:
        return self.distance

    def get(self, name: str, default=None):
        return self.entity.get(name, default)

    def to_dict(self) -> Dict[str, Any]:
        return {"id": self.id, "distance": self.distance, "entity": self.entity.to_dict()}

    def __repr__(self) -> str:
        return f"Hit(id={self.id!r}, distance={self.distance:.6g})"


class Hits(list):
    @property
    def ids(self) -> List[str]:
        return [h.id for h in self]

    @property
    def distances(self) -> List[float]:
        return [h.distance for h in self]


class SearchResult(list):
    """list of Hits, one per query vector (pymilvus SearchResult shape)."""


def _merge_parts(parts, top_k: int, metric: int, device: int):
    """Reduce the partial results of the sealed index and the growing segment ([(dist, ids), ...], each [nq, k]) on the
    device (sc_merge_topk): what the Milvus proxy does across segments [EXT].  numpy in -> numpy out."""
    (d0, i0), (d1, i1) = parts
    if isinstance(d0, np.ndarray):
        dev = torch.device("cuda", device)
        pd = torch.stack([torch.from_numpy(d0), torch.from_numpy(d1)]).to(dev)
        pi = torch.stack([torch.from_numpy(i0), torch.from_numpy(i1)]).to(dev)
        d, i = merge_topk(pd, pi, top_k, metric, device)
        return d.cpu().numpy(), i.cpu().numpy()
    return merge_topk(torch.stack([d0, d1]), torch.stack([i0, i1]), top_k, metric, device)


class _RWLock:
    """Many searches or one writer (FastAPI serves /query from a thread pool, api/main.py:202, next to
    BackgroundTasks ingestion, api/main.py:160).  Writers are preferred so that a stream of queries cannot starve an
    upsert; re-entrant for the thread that holds the write side (seal -> build_index -> add)."""

    def __init__(self):
        self._cv = threading.Condition(threading.Lock())
        self._readers = 0
        self._writer = None
        self._depth = 0
        self._waiting_writers = 0

    @contextmanager
    def read(self):
        me = threading.get_ident()
        with self._cv:
            if self._writer == me:  # the writer may read its own state
                self._depth += 1
                nested = True
            else:
                nested = False
                while self._writer is not None or self._waiting_writers:
                    self._cv.wait()
                self._readers += 1
        try:
            yield
        finally:
            with self._cv:
                if nested:
                    self._depth -= 1
                else:
                    self._readers -= 1
                    if self._readers == 0:
                        self._cv.notify_all()

    @contextmanager
    def write(self):
        me = threading.get_ident()
        with self._cv:
            if self._writer == me:
                self._depth += 1
            else:
                self._waiting_writers += 1
                while self._writer is not None or self._readers:
                    self._cv.wait()
                self._waiting_writers -= 1
                self._writer = me
                self._depth = 1
        try:
            yield
        finally:
            with self._cv:
                self._depth -= 1
                if self._depth == 0:
                    self._writer = None
                    self._cv.notify_all()


# --------------------------------------------------------------------------------------------------
# the in-process collection
# --------------------------------------------------------------------------------------------------
class GpuCollection:
    """What `Collection(name)` is to the reference wrapper: schema + index + rows."""

    def __init__(self, name: str, dim: int, nlist: int = 128, metric: str = "IP", device: int = 0,
                 seal_rows: Optional[int] = None, train_niter: int = 25, devices: Optional[Sequence[int]] = None,
                 compact_ratio: float = 0.2, retrain_factor: float = 0.0):
        self.name = name
        self.dim = int(dim)
        self.nlist = int(nlist)
        self.metric = metric_code(metric)
        self.devices = [int(d) for d in devices] if devices else [int(device)]
        self.device = self.devices[0]
        self.seal_rows = int(seal_rows) if seal_rows else KMEANS_MIN_POINTS_PER_CENTROID * self.nlist
        self.train_niter = int(train_niter)
        # maintenance Milvus does in the background [EXT]: drop tombstoned slots once they are this fraction of a sealed
        # index's slots; re-cluster once the index holds retrain_factor x the rows it was trained on (0 = never)
        self.compact_ratio = float(compact_ratio)
        self.retrain_factor = float(retrain_factor)
        self.persist_dir: Optional[str] = None
        self._generation = 0
        self._dirty = False
        # Incremental persistence: between full snapshots ("generations") the logical operations -- upserted rows, deleted
        # keys -- are appended to the published generation as journal files and replayed by from_snapshot(), so an ingest of a
        # few hundred rows into a large collection does not rewrite the collection.  A full snapshot follows once the journal
        # has grown to journal_ratio x the snapshot's size, after a structural change (seal, re-train), or on request.
        self.journal_ratio = 0.5
        self._pending: List[tuple] = []
        self._pending_bytes = 0
        self._needs_full = True  # no generation yet
        self._journal_bytes = 0
        self._journal_files = 0
        self._snapshot_bytes = 0
        self._replaying = False
        self._lock = _RWLock()
        # host scalar columns, indexed by row number (== the int64 id stored next to the vector)
        self._pk: List[Optional[str]] = []
        self._repo: List[str] = []
        self._path: List[str] = []
        self._language: List[str] = []
        self._text: List[str] = []
        self._metadata: List[Any] = []
        self._row_of: Dict[str, int] = {}
        self._repo_vocab: Dict[str, int] = {}
        self._lang_vocab: Dict[str, int] = {}
        # growing segment: a one-list index == exact search
        self._growing = IVFFlatIndex(self.dim, nlist=1, metric=self.metric, device=self.device)
        self._growing.set_centroids(np.zeros((1, self.dim), dtype=np.float32))
        self._growing_rows = 0
        self._ivf = None  # IVFFlatIndex, or MultiDeviceIVFFlat when several devices are configured
        # telemetry (`ivf_profile`): per-phase CUDA events of the sealed index's searches; last_search_stats() reports them
        self.profile = False
        self._last_stats: Optional[Dict[str, Any]] = None
        self._trained_rows = 0
        self.maintenance = {"compactions": 0, "retrains": 0}

    def _new_index(self, nlist: int):
        if len(self.devices) > 1:
            from ..multidevice import MultiDeviceIVFFlat

            return MultiDeviceIVFFlat(self.dim, nlist=nlist, metric=self.metric, devices=self.devices)
        return IVFFlatIndex(self.dim, nlist=nlist, metric=self.metric, device=self.device)

    def _load_index(self, path: str):
        if os.path.exists(os.path.join(path, "shards.json")):
            from ..multidevice import MultiDeviceIVFFlat

            return MultiDeviceIVFFlat.load(path, devices=self.devices)
        return IVFFlatIndex.load(path, device=self.device)

    # -- pymilvus-compatible no-ops --------------------------------------------------------------
    def load(self) -> None:
        return None

    def flush(self, full: bool = False) -> None:
        """Collection.flush(): with a persist directory configured, make everything ingested so far durable -- as a journal
        file next to the published snapshot while the changes are small, as a new snapshot generation otherwise (`full`)."""
        if not self.persist_dir:
            return
        with self._lock.write():
            grown = self._journal_bytes + self._pending_bytes > self.journal_ratio * max(self._snapshot_bytes, 1)
            if full or self._needs_full or self._generation == 0 or self.journal_ratio <= 0 or grown:
                self.save(self.persist_dir)
            elif self._pending:
                self._write_journal()
            self._dirty = False

    _PENDING_MAX = 256 << 20  # bytes of un-flushed vectors kept for the journal; beyond that the next flush is a full snapshot

    def _record(self, op: tuple, nbytes: int) -> None:
        if not self.persist_dir or self._replaying or self._needs_full:
            return
        if self._pending_bytes + nbytes > self._PENDING_MAX:
            self._pending, self._pending_bytes, self._needs_full = [], 0, True
            return
        self._pending.append(op)
        self._pending_bytes += nbytes

    def _write_journal(self) -> None:
        gen_dir = os.path.join(self.persist_dir, f"gen-{self._generation:08d}")
        ops, arrays = [], {}
        for i, op in enumerate(self._pending):
            if op[0] == "upsert":
                _, ids, vec, repos, paths, languages, texts, metadata = op
                ops.append({"op": "upsert", "ids": ids, "repos": repos, "paths": paths, "languages": languages, "texts": texts,
                            "metadata": metadata, "vec": f"v{i}"})
                arrays[f"v{i}"] = vec
            else:
                ops.append({"op": "delete", "ids": op[1]})
        arrays["ops"] = np.frombuffer(json.dumps(ops).encode("utf-8"), dtype=np.uint8)
        name = os.path.join(gen_dir, f"journal-{self._journal_files + 1:06d}.npz")
        with open(name + ".tmp", "wb") as f:
            np.savez(f, **arrays)
            f.flush()
            os.fsync(f.fileno())
        os.replace(name + ".tmp", name)  # the publication point of this batch of operations
        self._journal_files += 1
        self._journal_bytes += os.path.getsize(name)
        self._pending, self._pending_bytes = [], 0

    def _replay_journal(self, snap: str) -> None:
        files = sorted(f for f in os.listdir(snap) if f.startswith("journal-") and f.endswith(".npz"))
        self._replaying = True
        try:
            for fname in files:
                with np.load(os.path.join(snap, fname)) as z:
                    for op in json.loads(bytes(z["ops"]).decode("utf-8")):
                        if op["op"] == "upsert":
                            self.upsert_columns(op["ids"], z[op["vec"]], op["repos"], op["paths"], op["languages"], op["texts"],
                                                op["metadata"])
                        else:
                            self.delete(op["ids"])
                self._journal_bytes += os.path.getsize(os.path.join(snap, fname))
            self._journal_files = len(files)
        finally:
            self._replaying = False

    # -- persistence (connect() on an existing collection just loads,
    #    milvus_store.py:51-54; Milvus keeps its data in a volume, docker-compose.yml:13-14) -----------
    def save(self, path: str) -> None:
        """Write a complete snapshot into a fresh generation directory and publish it by replacing `CURRENT`:
        a crash at any point leaves the previous generation loadable, never a mix of old and new files."""
        with self._lock.write():
            os.makedirs(path, exist_ok=True)
            gen = self._generation + 1
            name = f"gen-{gen:08d}"
            tmp = os.path.join(path, name + ".tmp")
            shutil.rmtree(tmp, ignore_errors=True)
            os.makedirs(tmp)
            self._growing.save(os.path.join(tmp, "growing"))
            if self._ivf is not None:
                self._ivf.save(os.path.join(tmp, "ivf"))
            with open(os.path.join(tmp, "columns.jsonl"), "w") as f:
                for r, pk in enumerate(self._pk):
                    if pk is None:
                        f.write("null\n")
                    else:
                        f.write(json.dumps([pk, self._repo[r], self._path[r], self._language[r], self._text[r],
                                            self._metadata[r]]) + "\n")
            meta = {"format": 2, "generation": gen, "name": self.name, "dim": self.dim, "nlist": self.nlist, "metric": self.metric,
                    "seal_rows": self.seal_rows, "train_niter": self.train_niter, "has_ivf": self._ivf is not None,
                    "growing_rows": self._growing_rows, "rows": len(self._pk), "trained_rows": self._trained_rows,
                    "compact_ratio": self.compact_ratio, "retrain_factor": self.retrain_factor,
                    "repo_vocab": self._repo_vocab, "lang_vocab": self._lang_vocab}
            with open(os.path.join(tmp, "collection.json"), "w") as f:
                json.dump(meta, f)
            final = os.path.join(path, name)
            shutil.rmtree(final, ignore_errors=True)
            os.replace(tmp, final)
            with open(os.path.join(path, "CURRENT.tmp"), "w") as f:
                f.write(name)
            os.replace(os.path.join(path, "CURRENT.tmp"), os.path.join(path, "CURRENT"))  # the publication point
            self._generation = gen
            self._dirty = False
            self._pending, self._pending_bytes, self._needs_full = [], 0, False
            self._journal_bytes, self._journal_files = 0, 0
            self._snapshot_bytes = sum(os.path.getsize(os.path.join(dp, fn)) for dp, _, fns in os.walk(final) for fn in fns)
            for old in os.listdir(path):
                if old.startswith("gen-") and old != name:
                    shutil.rmtree(os.path.join(path, old), ignore_errors=True)

    @staticmethod
    def snapshot_dir(path: str) -> Optional[str]:
        """Directory of the published generation under `path` (or `path` itself for a round-1 flat snapshot)."""
        cur = os.path.join(path, "CURRENT")
        if os.path.exists(cur):
            with open(cur) as f:
                d = os.path.join(path, f.read().strip())
            return d if os.path.exists(os.path.join(d, "collection.json")) else None
        return path if os.path.exists(os.path.join(path, "collection.json")) else None

    @classmethod
    def from_snapshot(cls, path: str, device: int = 0, devices: Optional[Sequence[int]] = None) -> "GpuCollection":
        snap = cls.snapshot_dir(path)
        if snap is None:
            raise FileNotFoundError(f"no published snapshot under {path!r}")
        with open(os.path.join(snap, "collection.json")) as f:
            meta = json.load(f)
        col = cls(meta["name"], meta["dim"], nlist=meta["nlist"], metric=meta["metric"], device=device, devices=devices,
                  seal_rows=meta["seal_rows"], train_niter=meta["train_niter"],
                  compact_ratio=meta.get("compact_ratio", 0.2), retrain_factor=meta.get("retrain_factor", 0.0))
        col._generation = int(meta.get("generation", 0))
        col._growing.close()
        col._growing = IVFFlatIndex.load(os.path.join(snap, "growing"), device=col.device)
        col._growing_rows = int(meta["growing_rows"])
        col._trained_rows = int(meta.get("trained_rows", 0))
        if meta["has_ivf"]:
            col._ivf = col._load_index(os.path.join(snap, "ivf"))
        col._repo_vocab = {k: int(v) for k, v in meta["repo_vocab"].items()}
        col._lang_vocab = {k: int(v) for k, v in meta["lang_vocab"].items()}
        with open(os.path.join(snap, "columns.jsonl")) as f:
            for r, line in enumerate(f):
                row = json.loads(line)
                if row is None:
                    row = [None, "", "", "", "", None]
                pk, repo, pth, lang, text, md = row
                col._pk.append(pk)
                col._repo.append(repo)
                col._path.append(pth)
                col._language.append(lang)
                col._text.append(text)
                col._metadata.append(md)
                if pk is not None:
                    col._row_of[pk] = r
        live = col._growing.ntotal + (col._ivf.ntotal if col._ivf is not None else 0)
        if "rows" in meta and len(col._pk) != int(meta["rows"]) or live != len(col._row_of):
            raise ValueError(f"snapshot {snap!r} is inconsistent: {len(col._pk)} column rows, {len(col._row_of)} live keys, "
                             f"{live} live vectors")
        col._needs_full = False
        col._snapshot_bytes = sum(os.path.getsize(os.path.join(dp, fn)) for dp, _, fns in os.walk(snap) for fn in fns
                                  if not fn.startswith("journal-"))
        col._replay_journal(snap)  # what was ingested after that generation was written
        col._dirty = False
        return col

    @property
    def num_entities(self) -> int:
        with self._lock.read():
            return len(self._row_of)

    @property
    def index(self):
        return self._ivf

    def close(self) -> None:
        with self._lock.write():
            self._growing.close()
            if self._ivf is not None:
                self._ivf.close()
                self._ivf = None

    # -- tags ---------------------------------------------------------------------------------------
    def _tag(self, vocab: Dict[str, int], value: str, limit: int) -> int:
        t = vocab.get(value)
        if t is None:
            t = len(vocab)
            if t > limit:
                raise ValueError(f"too many distinct values for a tag column (limit {limit + 1})")
            vocab[value] = t
        return t

    def repo_tags(self, repos: Iterable[str]) -> List[int]:
        return [self._repo_vocab[r] for r in repos if r in self._repo_vocab]

    def language_tags(self, languages: Iterable[str]) -> List[int]:
        return [self._lang_vocab[l] for l in languages if l in self._lang_vocab]

    # -- insert -------------------------------------------------------------------------------------
    def upsert(self, data: Sequence[Sequence[Any]]) -> int:
        """Collection.upsert([ids, repos, paths, languages, texts, vectors, metadata])
        (milvus_store.py:128-130): replace by primary key, then insert."""
        ids, repos, paths, languages, texts, vectors, metadata = data
        return self.upsert_columns(ids, vectors, repos, paths, languages, texts, metadata)

    def upsert_columns(self, ids: Sequence[str], vectors, repos=None, paths=None, languages=None, texts=None,
                       metadata=None) -> int:
        n = len(ids)
        if n == 0:
            return 0
        if torch is not None and isinstance(vectors, torch.Tensor):
            vec = vectors.to(torch.float32)
            if vec.dim() != 2 or vec.shape[0] != n or vec.shape[1] != self.dim:
                raise ValueError(f"vectors: expected shape [{n}, {self.dim}], got {tuple(vec.shape)}")
        else:
            vec = np.asarray(vectors, dtype=np.float32)
            if vec.ndim != 2 or vec.shape[0] != n or vec.shape[1] != self.dim:
                raise ValueError(f"vectors: expected shape [{n}, {self.dim}], got {vec.shape}")

        def col(c, default):
            return list(c) if c is not None else [default] * n

        repos, paths, languages, texts = col(repos, ""), col(paths, ""), col(languages, ""), col(texts, "")
        metadata = col(metadata, None)
        with self._lock.write():
            # last occurrence of a primary key inside the batch wins
            last: Dict[str, int] = {}
            for i, pk in enumerate(ids):
                last[str(pk)] = i
            keep = sorted(last.values())
            stale = [self._row_of[pk] for pk in last if pk in self._row_of]
            if stale:
                self._remove_rows(stale)
            base = len(self._pk)
            row_ids = np.arange(base, base + len(keep), dtype=np.int64)
            rtags = np.empty(len(keep), dtype=np.uint32)
            ltags = np.empty(len(keep), dtype=np.uint8)
            for j, i in enumerate(keep):
                pk = str(ids[i])
                r, l = str(repos[i] or ""), str(languages[i] or "")
                self._pk.append(pk)
                self._repo.append(r)
                self._path.append(str(paths[i] or ""))
                self._language.append(l)
                self._text.append(texts[i])
                self._metadata.append(metadata[i])
                self._row_of[pk] = base + j
                rtags[j] = self._tag(self._repo_vocab, r, (1 << 23) - 1)
                ltags[j] = self._tag(self._lang_vocab, l, 255)
            if len(keep) != n:
                vec = vec[torch.as_tensor(keep, device=vec.device)] if not isinstance(vec, np.ndarray) else vec[keep]
            self._dirty = True
            if self.persist_dir and not self._replaying and not self._needs_full:
                host = vec if isinstance(vec, np.ndarray) else vec.detach().cpu().numpy()
                self._record(("upsert", [str(ids[i]) for i in keep], np.array(host, dtype=np.float32, copy=True),
                              [str(repos[i] or "") for i in keep], [str(paths[i] or "") for i in keep],
                              [str(languages[i] or "") for i in keep], [texts[i] for i in keep], [metadata[i] for i in keep]),
                             int(host.nbytes) + sum(len(str(texts[i])) for i in keep) + 96 * len(keep))
            if self._ivf is not None:
                self._ivf.add(vec, row_ids, rtags, ltags)
                self._maintain()
            else:
                self._growing.add(vec, row_ids, rtags, ltags, lists=np.zeros(len(keep), dtype=np.int32))
                self._growing_rows += len(keep)
                if self._growing_rows >= self.seal_rows:
                    self.build_index()
            return len(keep)

    def _remove_rows(self, rows: List[int]) -> None:
        arr = np.asarray(rows, dtype=np.int64)
        removed = self._growing.remove_ids(arr)
        self._growing_rows -= removed
        if self._ivf is not None and removed < len(rows):
            self._ivf.remove_ids(arr)
        for r in rows:
            pk = self._pk[r]
            if pk is not None:
                self._row_of.pop(pk, None)
            self._pk[r] = None
            self._text[r] = ""
            self._metadata[r] = None
        self._dirty = True

    def delete(self, ids: Sequence[str]) -> int:
        with self._lock.write():
            rows = [self._row_of[str(pk)] for pk in ids if str(pk) in self._row_of]
            if rows:
                self._remove_rows(rows)
                self._record(("delete", [str(pk) for pk in ids]), 64 * len(rows))
                self._maintain()
            return len(rows)

    # -- maintenance: what Milvus' compaction / index rebuild do behind the reference's back [EXT] -------------
    def _maintain(self) -> None:
        if self._ivf is None:
            return
        st = self._ivf.stats()
        slots = st.ntotal + st.nremoved
        if self.compact_ratio > 0 and st.nremoved > 0 and st.nremoved >= self.compact_ratio * slots:
            self.compact()
        if self.retrain_factor > 0 and self._trained_rows > 0 and st.ntotal >= self.retrain_factor * self._trained_rows:
            self.retrain()

    def compact(self) -> int:
        """Drop the tombstoned slots of the sealed index in place (sc_index_compact); returns the pages freed."""
        with self._lock.write():
            if self._ivf is None:
                return 0
            freed = self._ivf.compact()
            self.maintenance["compactions"] += 1
            log.info("index_compacted", collection=self.name, pages_freed=freed)
            return freed

    def retrain(self, niter: Optional[int] = None) -> None:
        """Re-cluster the sealed index on the rows it holds NOW and move every live row to its new list: the centroids of
        the first seal stop describing a collection that has grown several-fold (FAISS itself never re-trains; Milvus
        builds a fresh index per sealed segment [EXT])."""
        with self._lock.write():
            old = self._ivf
            if old is None or old.ntotal == 0:
                return
            n = old.ntotal
            nlist = self.nlist if self.nlist * KMEANS_MIN_POINTS_PER_CENTROID <= n else max(1, n // KMEANS_MIN_POINTS_PER_CENTROID)
            new = self._new_index(nlist)
            ranges = old.list_ranges()
            # train on a sample spread over all lists (FAISS caps the training set at 256 rows per centroid [EXT])
            cap = 256 * nlist
            stride = max(1, n // cap)
            sample = []
            for lo, hi in ranges:
                _, v, _, t = old.export_lists(lo, hi)
                livev = v[(t & np.uint32(0x80000000)) == 0]
                sample.append(livev[::stride])
            new.train(np.concatenate(sample), niter=self.train_niter if niter is None else niter, max_points_per_centroid=0)
            for lo, hi in ranges:
                _, v, i, t = old.export_lists(lo, hi)
                live = (t & np.uint32(0x80000000)) == 0
                if live.any():
                    new.add(v[live], i[live], (t[live] >> np.uint32(8)) & np.uint32((1 << 23) - 1), (t[live] & np.uint32(0xFF)).astype(np.uint8))
            self._ivf = new
            old.close()
            self._trained_rows = n
            self._dirty = True
            self._needs_full = True  # the lists were rebuilt: replaying a journal on the old snapshot would not reproduce them
            self.maintenance["retrains"] += 1
            log.info("index_retrained", collection=self.name, nlist=nlist, rows=n)

    # -- index build ----------------------------------------------------------------------------------
    def build_index(self, niter: Optional[int] = None, centroids=None):
        """Seal the growing segment: k-means (or the supplied centroids), then move its rows into
        the inverted lists.  With fewer than 39 rows per list nlist shrinks as knowhere does [EXT]."""
        with self._lock.write():
            if self._growing_rows == 0:
                return self._ivf
            vec, rid, tags = self._growing.export_list(0)
            live = (tags & np.uint32(0x80000000)) == 0
            vec, rid, tags = vec[live], rid[live], tags[live]
            n = vec.shape[0]
            if self._ivf is None:
                if centroids is not None:
                    nlist = int(np.asarray(centroids).shape[0])
                else:
                    nlist = self.nlist
                    if nlist * KMEANS_MIN_POINTS_PER_CENTROID > n:
                        nlist = max(1, n // KMEANS_MIN_POINTS_PER_CENTROID)
                ivf = self._new_index(nlist)
                if centroids is not None:
                    ivf.set_centroids(centroids)
                else:
                    ivf.train(vec, niter=self.train_niter if niter is None else niter)
                log.info("index_trained", collection=self.name, nlist=nlist, rows=n)
                self._ivf = ivf
                self._trained_rows = n
            self._ivf.add(vec, rid, (tags >> np.uint32(8)) & np.uint32((1 << 23) - 1), (tags & np.uint32(0xFF)).astype(np.uint8))
            self._growing.reset()
            self._growing_rows = 0
            self._dirty = True
            if not self._replaying:
                self._needs_full = True  # sealed: the next flush writes a generation instead of a journal file
            return self._ivf

    # -- search ---------------------------------------------------------------------------------------
    def search_arrays(self, vectors, top_k: int, nprobe: int = 16, repos: Optional[Iterable[str]] = None,
                      languages: Optional[Iterable[str]] = None):
        """Raw tensor path: (dist [nq,k] fp32, row ids [nq,k] int64, -1 = none).  CUDA in -> CUDA out.
        Searches hold the READ side of the collection lock: they run concurrently with each other."""
        with self._lock.read():
            rt = lt = None
            if repos is not None:
                rt = self.repo_tags(repos)
                if not rt:
                    return self._empty(vectors, top_k)
            if languages is not None:
                lt = self.language_tags(languages)
                if not lt:
                    return self._empty(vectors, top_k)
            parts = []
            if self._ivf is not None and self._ivf.ntotal > 0:
                parts.append(self._ivf.search(vectors, top_k, nprobe=nprobe, repos=rt, langs=lt))
            if self._growing_rows > 0 or not parts:
                parts.append(self._growing.search(vectors, top_k, nprobe=1, repos=rt, langs=lt))
            if len(parts) == 1:
                return parts[0]
            return _merge_parts(parts, top_k, self.metric, self.device)

    def _empty(self, vectors, top_k: int):
        nq = len(vectors) if not hasattr(vectors, "shape") else (1 if len(vectors.shape) == 1 else vectors.shape[0])
        pad = -np.finfo(np.float32).max if self.metric == METRIC_IP else np.finfo(np.float32).max
        return np.full((nq, top_k), pad, dtype=np.float32), np.full((nq, top_k), -1, dtype=np.int64)

    def search(self, data, anns_field: str = "embedding", param: Optional[dict] = None, limit: int = 10,
               expr: Optional[str] = None, output_fields: Optional[Sequence[str]] = None, *,
               repos: Optional[Iterable[str]] = None, languages: Optional[Iterable[str]] = None, **_) -> SearchResult:
        """Collection.search(...) with the keyword set the reference passes (milvus_store.py:141-147)."""
        if anns_field != "embedding":
            raise ValueError(f"unknown vector field {anns_field!r}")
        if expr:
            raise NotImplementedError("boolean expressions are not parsed; use repos= / languages=")
        param = param or {}
        want = param.get("metric_type")
        if want is not None and metric_code(want) != self.metric:
            raise ValueError(f"metric_type {want!r} does not match the index metric")
        nprobe = int((param.get("params") or {}).get("nprobe", 16))
        q = np.asarray(data, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"vector dimension mismatch: expected {self.dim}, got {q.shape[-1]}")
        fields = tuple(output_fields) if output_fields is not None else ()
        with self._lock.read():
            dist, rows = self.search_arrays(q, int(limit), nprobe=nprobe, repos=repos, languages=languages)
            if not isinstance(dist, np.ndarray):
                dist, rows = dist.cpu().numpy(), rows.cpu().numpy()
            out = SearchResult()
            for qi in range(q.shape[0]):
                hits = Hits()
                for d, r in zip(dist[qi], rows[qi]):
                    if r < 0:
                        continue  # Milvus returns short result lists rather than padding
                    r = int(r)
                    f = {}
                    for name in fields:
                        if name == "repo":
                            f[name] = self._repo[r]
                        elif name == "path":
                            f[name] = self._path[r]
                        elif name == "language":
                            f[name] = self._language[r]
                        elif name == "text":
                            f[name] = self._text[r]
                        elif name == "metadata":
                            f[name] = self._metadata[r]
                        elif name == "id":
                            f[name] = self._pk[r]
                        else:
                            raise ValueError(f"unknown output field {name!r}")
                    hits.append(Hit(self._pk[r], float(d), f))
                out.append(hits)
            if self.profile:
                self._record_stats(q.shape[0], nprobe)
            return out

    def _record_stats(self, nq: int, nprobe: int) -> None:
        """What the last search on the sealed index cost (telemetry; under concurrency it is SOME recent search)."""
        ivf = self._ivf
        if ivf is None or not hasattr(ivf, "last_search_times"):
            return
        try:
            if not getattr(ivf, "_profiling_on", False):
                ivf.set_profiling(True)  # takes effect from the next search on
                ivf._profiling_on = True
                return
            t = ivf.last_search_times()
        except Exception:  # no search on the sealed index yet
            return
        rows = t.unique_rows if t.unique_rows > 0 else t.scanned_rows
        self._last_stats = {
            "nq": int(nq), "nprobe": int(nprobe), "search_ms": float(t.total_ms), "scan_ms": float(t.scan_ms),
            "scanned_rows": int(t.scanned_rows), "scanned_bytes": int(rows) * 4 * self.dim,
            "scan_GBps": (int(rows) * 4 * self.dim / (t.scan_ms * 1e6)) if t.scan_ms > 0 else 0.0,
            "scan": "list-major" if t.unique_rows > 0 else "query-major",
        }

    def last_search_stats(self) -> Optional[Dict[str, Any]]:
        return dict(self._last_stats) if self._last_stats else None


_REGISTRY: Dict[str, GpuCollection] = {}
_REGISTRY_LOCK = threading.Lock()


def has_collection(name: str) -> bool:
    """utility.has_collection (milvus_store.py:51)."""
    with _REGISTRY_LOCK:
        return name in _REGISTRY


def drop_collection(name: str) -> None:
    with _REGISTRY_LOCK:
        c = _REGISTRY.pop(name, None)
    if c is not None:
        c.close()


def _parse_devices(value) -> Optional[List[int]]:
    """`ivf_devices`: "0,1,2,3" / [0, 1] / "all" -> CUDA ordinals of the row-sharded collection; empty = single device."""
    if value in (None, "", [], ()):
        return None
    if isinstance(value, str):
        if value.strip().lower() == "all":
            return list(range(torch.cuda.device_count())) if torch is not None else None
        return [int(v) for v in value.replace(";", ",").split(",") if v.strip() != ""]
    return [int(v) for v in value]


def _flush_all_at_exit() -> None:
    # Milvus persists server-side; the reference never calls flush() (its milvus_store.py:128-133).  Collections with a
    # persist directory are therefore written after every upsert_embeddings call and, as a last resort, here.
    for c in list(_REGISTRY.values()):
        try:
            if c.persist_dir and c._dirty:
                c.flush()
        except Exception:  # pragma: no cover - interpreter teardown
            pass


atexit.register(_flush_all_at_exit)


# --------------------------------------------------------------------------------------------------
# the drop-in wrapper
# --------------------------------------------------------------------------------------------------
class MilvusVectorStore:
    """Same surface as the reference wrapper (milvus_store.py:29-148), GPU-backed."""

    def __init__(self, collection_name: str = "semcode_chunks", dim: Optional[int] = None) -> None:
        self.collection_name = collection_name
        self.dim = dim or settings.embedding_dimension
        self._collection: Optional[GpuCollection] = None

    def connect(self) -> None:
        """Attach to the process-wide collection (creating it on first use)."""
        log.info("connecting_milvus", uri=_setting("milvus_uri", "gpu://in-process"))
        self._collection = self._ensure_collection()

    def _ensure_collection(self) -> GpuCollection:
        with _REGISTRY_LOCK:
            existing = _REGISTRY.get(self.collection_name)
            if existing is not None:
                if existing.dim != self.dim:
                    raise ValueError(
                        f"collection {self.collection_name!r} exists with dim {existing.dim}, requested {self.dim}"
                    )
                existing.load()
                return existing
            persist = _setting("ivf_persist_dir", "")
            snap = os.path.join(persist, self.collection_name) if persist else ""
            devices = _parse_devices(_setting("ivf_devices", ""))
            if snap and GpuCollection.snapshot_dir(snap) is not None:
                collection = GpuCollection.from_snapshot(snap, device=_setting("ivf_device", 0), devices=devices)
                if collection.dim != self.dim:
                    collection.close()
                    raise ValueError(f"snapshot {snap!r} has dim {collection.dim}, requested {self.dim}")
                collection.persist_dir = snap
                collection.journal_ratio = float(_setting("ivf_journal_ratio", 0.5))
                _REGISTRY[self.collection_name] = collection
                return collection
            log.info("creating_milvus_collection", collection=self.collection_name, dim=self.dim)
            if not snap:
                # one-process limitation: without a persist directory the rows live and die with this process (a Milvus
                # server would keep them for the next CLI / API process)
                log.warning("collection_not_persisted", collection=self.collection_name,
                            hint="set SEMCODE_IVF_PERSIST_DIR to keep the collection across processes")
            collection = GpuCollection(
                self.collection_name,
                self.dim,
                nlist=_setting("ivf_nlist", 128),  # milvus_store.py:81
                metric=_setting("ivf_metric", "IP"),  # milvus_store.py:79
                device=_setting("ivf_device", 0),
                devices=devices,
                seal_rows=_setting("ivf_seal_rows", 0) or None,
                train_niter=_setting("ivf_train_niter", 25),
                compact_ratio=_setting("ivf_compact_ratio", 0.2),
                retrain_factor=_setting("ivf_retrain_factor", 0.0),
            )
            collection.persist_dir = snap or None
            collection.journal_ratio = float(_setting("ivf_journal_ratio", 0.5))
            collection.load()
            _REGISTRY[self.collection_name] = collection
            return collection

    def _require(self) -> GpuCollection:
        if self._collection is None:
            raise RuntimeError("Milvus collection is not initialized. Call connect() first.")
        return self._collection

    def upsert_embeddings(self, payloads: Sequence[Any], progress: Optional[Callable[[int, int], None]] = None) -> None:
        """Insert or update embeddings (same batching and progress protocol as milvus_store.py:87-133)."""
        collection = self._require()
        payload_list = list(payloads)
        total = len(payload_list)
        log.info("upserting_embeddings", count=total)
        if progress:
            progress(0, total)
        if total == 0:
            return
        batch_size = max(1, _setting("milvus_upsert_batch_size", 128))
        inserted = 0
        for start in range(0, total, batch_size):
            batch = payload_list[start : start + batch_size]
            ids, repos, paths, languages, texts, vectors, metadata = [], [], [], [], [], [], []
            for payload in batch:
                ids.append(payload.id)
                repos.append(payload.metadata.get("repo", ""))
                paths.append(payload.metadata.get("path", ""))
                languages.append(payload.metadata.get("language", ""))
                texts.append(payload.text)
                vectors.append(payload.vector)
                metadata.append(payload.metadata)
            collection.upsert([ids, repos, paths, languages, texts, vectors, metadata])
            inserted += len(batch)
            if progress:
                progress(inserted, total)
        if collection.persist_dir:  # what the Milvus server does without being asked: the rows survive this process
            collection.flush()

    def search(self, vector: "list[float]", top_k: int = 10, *, nprobe: int = 16,
               repos: Optional[Iterable[str]] = None, languages: Optional[Iterable[str]] = None) -> list:
        """Run a raw vector search (milvus_store.py:135-148); returns [Hits] for the one query."""
        collection = self._require()
        collection.profile = bool(_setting("ivf_profile", False))
        return collection.search(
            data=[vector],
            anns_field="embedding",
            param={"metric_type": "IP" if collection.metric == METRIC_IP else "L2", "params": {"nprobe": nprobe}},
            limit=top_k,
            output_fields=list(OUTPUT_FIELDS),
            repos=repos,
            languages=languages,
        )

    # ---- extensions (BASELINE.json configs need batches, raw arrays and bulk insert) ---------------
    def search_batch(self, vectors, top_k: int = 10, *, nprobe: int = 16, repos: Optional[Iterable[str]] = None,
                     languages: Optional[Iterable[str]] = None) -> list:
        collection = self._require()
        return collection.search(
            data=vectors, param={"params": {"nprobe": nprobe}}, limit=top_k, output_fields=list(OUTPUT_FIELDS),
            repos=repos, languages=languages,
        )

    def search_arrays(self, vectors, top_k: int = 10, *, nprobe: int = 16, repos: Optional[Iterable[str]] = None,
                      languages: Optional[Iterable[str]] = None):
        """(dist [nq,k], row ids [nq,k]) without materialising Python hit objects."""
        return self._require().search_arrays(vectors, top_k, nprobe=nprobe, repos=repos, languages=languages)

    def upsert_arrays(self, ids: Sequence[str], vectors, repos=None, paths=None, languages=None, texts=None,
                      metadata=None) -> int:
        """Bulk insert: `vectors` is one [n, dim] numpy array or torch tensor (CPU or CUDA)."""
        return self._require().upsert_columns(ids, vectors, repos, paths, languages, texts, metadata)

    def build_index(self, niter: Optional[int] = None, centroids=None):
        collection = self._require()
        out = collection.build_index(niter=niter, centroids=centroids)
        if collection.persist_dir:
            collection.flush()
        return out

    def compact(self) -> int:
        """Drop tombstoned slots now (done automatically once they pass `ivf_compact_ratio` of the sealed index)."""
        return self._require().compact()

    def retrain(self, niter: Optional[int] = None) -> None:
        """Re-cluster the sealed index on its current rows (automatic at `ivf_retrain_factor` x the trained size)."""
        self._require().retrain(niter=niter)

    def flush(self) -> None:
        """Persist the collection when `ivf_persist_dir` (SEMCODE_IVF_PERSIST_DIR) is configured: an explicit flush writes a
        complete snapshot generation (the implicit ones after upsert_embeddings journal small changes instead)."""
        self._require().flush(full=True)

    def last_search_stats(self) -> Optional[dict]:
        """With `ivf_profile` set: what a recent search on the sealed index cost -- search / scan milliseconds, bytes of list
        vectors it had to read and the scan's GB/s -- for the API's telemetry (api/telemetry.py:89-104 records per-query
        metadata); None otherwise."""
        collection = self._collection
        return None if collection is None else collection.last_search_stats()
