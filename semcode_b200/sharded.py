"""ShardedIVFFlat -- one IVF_FLAT index row-sharded over the GPUs of a box, one process per GPU.

SURVEY.md section 8e: the centroids are replicated, every inverted list's rows are dealt
round-robin to the ranks (global int64 ids), every rank runs the identical coarse pass and scans
its 1/G slice, and the partial top-k are exchanged with ONE small all-gather (12*nq*k bytes per
rank over NVLink) and merged on the device.  Because the union of the slices equals the single
index's lists, the merged result equals the 1-GPU result (tie order aside) -- unlike Milvus'
per-segment indexes [EXT].  k-means is data-parallel Lloyd: local assignment pass, all-reduce of
the per-centroid sums / counts / objective, identical update everywhere.

torch.distributed is the plumbing (NCCL on GPUs; the host logic is also exercised with gloo and a
CPU test double in tests/test_sharded_gloo.py).  All arithmetic is in the per-rank engine.
"""

from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from .index import KMEANS_NITER, KMEANS_SEED, metric_code


def _default_merge(metric: int, device):
    from .index import merge_topk

    return lambda pd, pi, k: merge_topk(pd, pi, k, metric, device)


class ShardedIVFFlat:
    def __init__(self, dim: int, nlist: int, metric="IP", device: Optional[int] = None, group=None,
                 engine=None, merge: Optional[Callable] = None, shard_by: str = "rows", exchange: str = "auto",
                 exchange_bytes: int = 64 << 20, inflight: int = 1):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dim, self.nlist, self.metric = int(dim), int(nlist), metric_code(metric)
        if engine is None:
            from .index import IVFFlatIndex, merge_topk

            if device is None:
                device = torch.cuda.current_device()
            engine = IVFFlatIndex(dim, nlist=nlist, metric=metric, device=device)
            merge = merge or (lambda pd, pi, k: merge_topk(pd, pi, k, self.metric, device))
        if merge is None:
            raise ValueError("a custom engine needs a merge function")
        if shard_by not in ("rows", "lists"):
            raise ValueError("shard_by must be 'rows' or 'lists'")
        # "rows":  every list's rows are dealt round-robin to the ranks (each rank holds 1/G of every list)
        # "lists": list l lives entirely on rank l % G (whole lists per rank: longer per-rank lists, which the
        #          list-major scan prefers; each query's probes are split over the ranks instead of its rows)
        # Either way the union of the shards is exactly the single index, so merged results are identical.
        self.shard_by = shard_by
        self.local = engine
        self._merge = merge
        self._next_row = 0  # global round-robin cursor, identical on every rank
        # "p2p": scatter of the probe rows and gather + merge of the partial top-k fused into the engine's kernels
        # over peer-mapped memory (PeerExchange); "nccl": all-gather + merge kernel; "auto": p2p when it can be set up
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'nccl'")
        # inflight > 1: that many exchanges ("lanes"), so that many fused steps can be in flight per rank, each on its own
        # stream (search(..., lane=j) / search_batches): while one step waits for the slowest peer's partial top-k the next
        # step's coarse pass and scan run.  The engine hands each lane its own scratch slot (two slots per handle).
        self.exchange = None
        self.exchanges = []
        self._lane_streams = None
        self.exchange_error: Optional[str] = None
        if exchange != "nccl" and self.world > 1 and hasattr(engine, "_h"):
            ok = torch.zeros(1, dtype=torch.int32, device=engine.tensor_device())
            try:
                from .index import PeerExchange

                self.exchanges = [PeerExchange(engine.device, group, exchange_bytes) for _ in range(max(1, int(inflight)))]
                ok += 1
            except Exception as e:  # symmetric memory unavailable on this box / build
                self.exchange_error = f"{type(e).__name__}: {e}"
            dist.all_reduce(ok, group=group)  # all ranks or none
            if int(ok.item()) != self.world:
                self.exchanges = []
                if exchange == "p2p":
                    raise RuntimeError(f"peer-memory exchange unavailable: {self.exchange_error}")
            if self.exchanges:
                self.exchange = self.exchanges[0]

    # -- coarse quantizer -------------------------------------------------------------------------
    def set_centroids(self, centroids=None, src: int = 0) -> None:
        """Broadcast `centroids` (given on rank `src`) and install them on every rank."""
        dev = self.local.tensor_device()
        if self.rank == src:
            c = torch.as_tensor(np.asarray(centroids, dtype=np.float32) if not torch.is_tensor(centroids) else centroids)
            c = c.to(dev, torch.float32).contiguous()
        else:
            c = torch.empty((self.nlist, self.dim), dtype=torch.float32, device=dev)
        dist.broadcast(c, src, group=self.group)
        self.local.set_centroids(c)

    def train(self, x_local, niter: int = KMEANS_NITER, seed: int = KMEANS_SEED, init_centroids=None) -> List[float]:
        """Data-parallel Lloyd over the ranks' local rows.  Initial centroids: `init_centroids` on
        rank 0, else `nlist` seeded random rows of rank 0's shard.  Returns the global objective
        entering each iteration (identical on every rank)."""
        from .index import kmeans_init_rows

        if self.rank == 0 and init_centroids is None:
            n0 = x_local.shape[0]
            if n0 < self.nlist:
                raise ValueError(f"rank 0 holds {n0} rows, fewer than nlist={self.nlist}")
            rows = kmeans_init_rows(n0, self.nlist, seed)
            init_centroids = x_local[torch.from_numpy(rows).to(x_local.device)] if torch.is_tensor(x_local) else x_local[rows]
        self.set_centroids(init_centroids, src=0)
        sums, counts, obj = self.local.kmeans_buffers()
        out = []
        for _ in range(niter):
            sums.zero_()
            counts.zero_()
            obj.zero_()
            self.local.kmeans_step(x_local, sums, counts, obj)
            # the one exchange of an iteration: nlist*(8*ds + 4) + 8 bytes per rank
            dist.all_reduce(sums, group=self.group)
            dist.all_reduce(counts, group=self.group)
            dist.all_reduce(obj, group=self.group)
            out.append(float(obj.item()))
            self.local.kmeans_update(sums, counts)
        return out

    # -- insert -------------------------------------------------------------------------------------
    def add(self, x, ids, repo_tags=None, lang_tags=None) -> None:
        """Every rank passes the SAME global batch; rank r keeps the rows whose global arrival
        number is congruent to r (round-robin deal), so each list is spread evenly."""
        n = x.shape[0]
        if self.shard_by == "lists":
            lists = self.local.assign(x)  # every rank ranks the whole batch; it keeps the rows of ITS lists
            lists_np = lists.cpu().numpy() if torch.is_tensor(lists) else np.asarray(lists)
            keep = np.flatnonzero(lists_np % self.world == self.rank)
            if keep.size == 0:
                return

            def pick(a):
                if a is None:
                    return None
                if torch.is_tensor(a):
                    return a[torch.from_numpy(keep).to(a.device)].contiguous()
                return np.ascontiguousarray(np.asarray(a)[keep])

            self.local.add(pick(x), pick(ids), pick(repo_tags), pick(lang_tags), lists=lists_np[keep].astype(np.int32))
            return
        first = (self.rank - self._next_row) % self.world
        sl = slice(first, n, self.world)
        self._next_row = (self._next_row + n) % self.world
        if len(range(n)[sl]) == 0:
            return

        def take(a):
            if a is None:
                return None
            a = a[sl]
            return a.contiguous() if torch.is_tensor(a) else np.ascontiguousarray(a)

        self.local.add(take(x), take(ids), take(repo_tags), take(lang_tags))

    def add_local(self, x_local, ids_local, repo_tags=None, lang_tags=None) -> None:
        """The caller has already partitioned the rows (ids must be globally unique)."""
        self.local.add(x_local, ids_local, repo_tags, lang_tags)

    def check_exchange(self) -> None:
        """Synchronise and raise if any fused step issued so far timed out waiting for a peer."""
        if any(ex.status()[0] for ex in self.exchanges):
            raise RuntimeError("sharded search: a peer did not arrive within the exchange timeout; the results of that step "
                               "and the ones after it are invalid -- rebuild the exchange after a barrier")

    @property
    def ntotal(self) -> int:
        t = torch.tensor([self.local.ntotal], dtype=torch.int64, device=self.local.tensor_device())
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def close(self) -> None:
        """Collective: every rank drops its mappings of the peers' exchange buffers, then (after a barrier) its own buffer."""
        for ex in self.exchanges:
            ex.close()
        if self.exchanges:
            dist.barrier(group=self.group)
            for ex in self.exchanges:
                ex._buf = None
        self.exchanges, self.exchange = [], None

    # -- persistence ----------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        """Every rank writes its shard under `path/shard-RR/` (the engine's own snapshot: bulk list export underneath), rank 0
        adds `sharded.json`, published last: a directory without it is not a snapshot.  `path` must be visible to every rank."""
        import json
        import os

        os.makedirs(path, exist_ok=True)
        if self.rank == 0 and os.path.exists(os.path.join(path, "sharded.json")):
            os.remove(os.path.join(path, "sharded.json"))  # unpublish before the shards change
        dist.barrier(group=self.group)
        self.local.save(os.path.join(path, f"shard-{self.rank:02d}"))
        dist.barrier(group=self.group)
        if self.rank == 0:
            tmp = os.path.join(path, "sharded.json.tmp")
            with open(tmp, "w") as f:
                json.dump({"format": 1, "world": self.world, "dim": self.dim, "nlist": self.nlist, "metric": self.metric,
                           "shard_by": self.shard_by, "next_row": self._next_row}, f)
            os.replace(tmp, os.path.join(path, "sharded.json"))
        dist.barrier(group=self.group)

    @classmethod
    def load(cls, path: str, device: Optional[int] = None, group=None, exchange: str = "auto", exchange_bytes: int = 64 << 20,
             inflight: int = 1, engine_loader: Optional[Callable] = None, merge: Optional[Callable] = None) -> "ShardedIVFFlat":
        """Re-open a snapshot written by `save` with the SAME number of ranks (the deal of the rows is part of the data)."""
        import json
        import os

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        with open(os.path.join(path, "sharded.json")) as f:
            meta = json.load(f)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if int(meta["world"]) != world:
            raise ValueError(f"snapshot {path!r} was written by {meta['world']} ranks, this job has {world}")
        shard = os.path.join(path, f"shard-{rank:02d}")
        if engine_loader is None:
            from .index import IVFFlatIndex

            if device is None:
                device = torch.cuda.current_device()
            engine = IVFFlatIndex.load(shard, device=device)
        else:
            engine = engine_loader(shard)
        self = cls(int(meta["dim"]), int(meta["nlist"]), int(meta["metric"]), device=device, group=group, engine=engine,
                   merge=merge if merge is not None else (None if engine_loader is not None else _default_merge(int(meta["metric"]), device)),
                   shard_by=meta.get("shard_by", "rows"), exchange=exchange, exchange_bytes=exchange_bytes, inflight=inflight)
        self._next_row = int(meta.get("next_row", 0))
        return self

    # -- search ---------------------------------------------------------------------------------------
    def probe(self, q, nprobe: int):
        """Coarse pass split over the ranks: rank r ranks the centroids for its 1/G of the queries, one
        all-gather of [nq, nprobe] int32 gives every rank the full probe table (centroids are replicated,
        so this equals the unsplit coarse pass)."""
        nq = q.shape[0]
        nprobe = min(int(nprobe), self.nlist)
        per = (nq + self.world - 1) // self.world
        lo, hi = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
        mine = torch.full((per, nprobe), -1, dtype=torch.int32, device=q.device)
        if hi > lo:
            mine[: hi - lo] = torch.as_tensor(self.local.probe(q[lo:hi].contiguous(), nprobe), device=q.device)
        full = torch.empty((self.world * per, nprobe), dtype=torch.int32, device=q.device)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        return full[:nq]

    def search_batches(self, batches, k: int, nprobe: int = 16, repos=None, langs=None):
        """Several query batches, `inflight` of them in flight at a time (one stream and one exchange per lane; every rank
        passes the same batches in the same order).  Returns the list of (dist, ids), complete on the current stream."""
        dev = self.local.tensor_device()
        lanes = max(1, len(self.exchanges))
        if lanes == 1 or self.world == 1:
            return [self.search(q, k, nprobe, repos, langs) for q in batches]
        if self._lane_streams is None:
            self._lane_streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
        cur = torch.cuda.current_stream(dev)
        for st in self._lane_streams:
            st.wait_stream(cur)
        outs = []
        for i, q in enumerate(batches):
            with torch.cuda.stream(self._lane_streams[i % lanes]):
                outs.append(self.search(q, k, nprobe, repos, langs, lane=i % lanes))
        for st in self._lane_streams:
            cur.wait_stream(st)
        return outs

    def search(self, q, k: int, nprobe: int = 16, repos=None, langs=None, shard_coarse: bool = True, lane: int = 0):
        """Every rank passes the same queries and receives the same merged (dist, ids) tensors.  `lane` picks the exchange
        of a fused step (see `inflight`); steps of one lane are ordered by the caller's stream."""
        dev = self.local.tensor_device()
        if not torch.is_tensor(q):
            q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
        q = q.to(dev, torch.float32)
        if self.world == 1:
            return self.local.search(q, k, nprobe=nprobe, repos=repos, langs=langs)
        if self.exchanges and self.shard_by == "rows" and q.shape[0] >= 1:
            ex = self.exchanges[lane]
            # A step whose merge gave up waiting for a peer returned garbage, and the late peer now writes into buffers
            # of later steps: never continue silently.  (The C ABI refuses as well; this gives the Python-level reason.)
            if ex.poll():
                raise RuntimeError("sharded search: a peer did not arrive within the exchange timeout (or a step failed "
                                   "midway); results since then are invalid -- rebuild the exchange after a barrier")
            # fused steps of at most 8192 queries: split coarse pass, probe rows and partial top-k stored into the
            # peers' buffers by the kernels that produce them, merge kernel waiting on the peers' flags
            np_ = min(int(nprobe), self.nlist)
            step = 8192
            while step > 1 and 2 * (step * np_ * 4 + self.world * step * k * 12 + 1024) + 4096 > ex.nbytes:
                step //= 2  # same arithmetic on every rank
            outs = [self.local.search(q[s:s + step], k, nprobe=nprobe, repos=repos, langs=langs, exchange=ex)
                    for s in range(0, q.shape[0], step)]
            if len(outs) == 1:
                return outs[0]
            return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
        if self.shard_by == "lists":
            probes = self.probe(q, nprobe) if q.shape[0] >= 2 * self.world else torch.as_tensor(
                self.local.probe(q, min(int(nprobe), self.nlist)), device=dev)
            mine = torch.where(probes % self.world == self.rank, probes, torch.full_like(probes, -1))  # -1 = skipped
            d, i = self.local.search(q, k, repos=repos, langs=langs, lists=mine.contiguous())
        elif shard_coarse and hasattr(self.local, "probe") and q.shape[0] >= 2 * self.world:
            d, i = self.local.search(q, k, repos=repos, langs=langs, lists=self.probe(q, nprobe))
        else:
            d, i = self.local.search(q, k, nprobe=nprobe, repos=repos, langs=langs)
        nq = d.shape[0]
        # concatenated layout [world*nq, k] (accepted by every backend), viewed as [world, nq, k]
        gd = torch.empty((self.world * nq, k), dtype=d.dtype, device=d.device)
        gi = torch.empty((self.world * nq, k), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(gd, d.contiguous(), group=self.group)
        dist.all_gather_into_tensor(gi, i.contiguous(), group=self.group)
        return self._merge(gd.view(self.world, nq, k), gi.view(self.world, nq, k), k)
