"""Host logic of the row-sharded index (ShardedIVFFlat) with world_size 2 over gloo, on CPU.

The per-rank engine is a TEST DOUBLE backed by the oracle (the product has no CPU engine); what is
under test is the product's sharding / exchange / merge orchestration: the round-robin deal, the
centroid broadcast, the k-means all-reduce and the top-k all-gather must reproduce the single-index
oracle result exactly."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleEngine:
    """CPU stand-in with IVFFlatIndex's engine interface (tests only)."""

    def __init__(self, dim, nlist, metric):
        from oracle import ivf_numpy as orc

        self.orc, self.dim, self.nlist, self.metric = orc, dim, nlist, orc.metric_code(metric)
        self.c = None
        self.x, self.ids = [], []

    def tensor_device(self):
        return torch.device("cpu")

    def set_centroids(self, c):
        self.c = c.numpy().astype(np.float32).copy()

    def get_centroids(self):
        return self.c.copy()

    def kmeans_buffers(self):
        return (torch.zeros(self.nlist * self.dim, dtype=torch.float64), torch.zeros(self.nlist, dtype=torch.int32),
                torch.zeros(1, dtype=torch.float64))

    def kmeans_step(self, x, sums, counts, obj):
        x = np.asarray(x, dtype=np.float32)
        sim = self.orc.coarse_similarity(x, self.c, self.metric)
        a = np.argmax(sim, axis=1)
        best = sim[np.arange(len(a)), a].astype(np.float64)
        s = np.zeros((self.nlist, self.dim))
        np.add.at(s, a, x.astype(np.float64))
        sums += torch.from_numpy(s.reshape(-1))
        counts += torch.from_numpy(np.bincount(a, minlength=self.nlist).astype(np.int32))
        o = best.sum() if self.metric == 0 else ((x.astype(np.float64) ** 2).sum(1) - best).sum()
        obj += float(o)

    def kmeans_update(self, sums, counts):
        s = sums.numpy().reshape(self.nlist, self.dim)
        n = counts.numpy()
        nz = n > 0
        self.c[nz] = (s[nz] / n[nz, None]).astype(np.float32)
        return self.orc.split_empty_clusters(self.c, n)

    def add(self, x, ids, repo_tags=None, lang_tags=None):
        self.x.append(np.asarray(x, dtype=np.float32))
        self.ids.append(np.asarray(ids, dtype=np.int64))

    @property
    def ntotal(self):
        return sum(len(i) for i in self.ids)

    def search(self, q, k, nprobe=16, repos=None, langs=None):
        idx = self.orc.build_index(np.concatenate(self.x), np.concatenate(self.ids), self.c, self.metric)
        d, i = self.orc.search(idx, q.numpy(), k, nprobe)
        return torch.from_numpy(d), torch.from_numpy(i)

    def save(self, path):
        os.makedirs(path, exist_ok=True)
        np.savez(os.path.join(path, "engine.npz"), c=self.c, x=np.concatenate(self.x), ids=np.concatenate(self.ids),
                 meta=np.array([self.dim, self.nlist, self.metric]))

    @classmethod
    def load(cls, path):
        z = np.load(os.path.join(path, "engine.npz"))
        dim, nlist, metric = (int(v) for v in z["meta"])
        e = cls(dim, nlist, metric)
        e.c, e.x, e.ids = z["c"], [z["x"]], [z["ids"]]
        return e


def _worker(rank, world, port, metric, ret, snap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import unit_rows
        from oracle import ivf_numpy as orc
        from semcode_b200.sharded import ShardedIVFFlat

        rng = np.random.default_rng(0)
        n, d, nlist, nq = 3000, 24, 16, 40
        x, q = unit_rows(rng, n, d), unit_rows(rng, nq, d)
        ids = np.arange(n, dtype=np.int64) * 3 + 1

        def merge(pd, pi, k):
            md, mi = orc.merge_topk(pd.numpy(), pi.numpy(), k, metric)
            return torch.from_numpy(md), torch.from_numpy(mi)

        sh = ShardedIVFFlat(d, nlist, metric, engine=OracleEngine(d, nlist, metric), merge=merge)
        # data-parallel k-means on the two halves == single-process k-means on all rows
        init = x[orc.kmeans_init_rows(n, nlist, 5)]
        obj = sh.train(x[rank::world], niter=5, init_centroids=init if rank == 0 else None)
        c_ref, obj_ref = orc.kmeans_train(x, nlist, metric, niter=5, init_centroids=init, max_points_per_centroid=0)
        np.testing.assert_allclose(obj, obj_ref, rtol=1e-9)
        np.testing.assert_allclose(sh.local.get_centroids(), c_ref, rtol=1e-6, atol=1e-7)
        # uneven batches exercise the global round-robin cursor
        for a, b in ((0, 1), (1, 8), (8, 1001), (1001, 3000)):
            sh.add(x[a:b], ids[a:b])
        assert sh.ntotal == n and abs(sh.local.ntotal - n // world) <= 1
        mine = np.concatenate(sh.local.ids)
        np.testing.assert_array_equal(mine, ids[rank::world])
        gd, gi = sh.search(q, 10, nprobe=4)
        full = orc.build_index(x, ids, sh.local.get_centroids(), metric)
        rd, ri = orc.search(full, q, 10, 4)
        np.testing.assert_array_equal(gi.numpy(), ri)
        np.testing.assert_allclose(gd.numpy(), rd, rtol=1e-6)  # BLAS blocking differs with the slice length
        # snapshot: one directory per rank + the published manifest; re-opened, it answers alike and the deal continues where
        # it stopped (the round-robin cursor is part of the snapshot)
        sh.save(snap)
        assert sorted(os.listdir(snap)) == ["shard-00", "shard-01", "sharded.json"]
        sh2 = ShardedIVFFlat.load(snap, engine_loader=OracleEngine.load, merge=merge)
        assert sh2.ntotal == n and sh2._next_row == sh._next_row and sh2.shard_by == "rows"
        gd2, gi2 = sh2.search(q, 10, nprobe=4)
        np.testing.assert_array_equal(gi2.numpy(), gi.numpy())
        extra = unit_rows(rng, 7, d)
        for obj_ in (sh, sh2):
            obj_.add(extra, np.arange(10**6, 10**6 + 7, dtype=np.int64))
        np.testing.assert_array_equal(np.concatenate(sh2.local.ids), np.concatenate(sh.local.ids))
        with pytest.raises(ValueError):
            import json

            meta = json.load(open(os.path.join(snap, "sharded.json")))
            bad = snap + f"-bad{rank}"
            os.makedirs(bad, exist_ok=True)
            json.dump(dict(meta, world=3), open(os.path.join(bad, "sharded.json"), "w"))
            ShardedIVFFlat.load(bad, engine_loader=OracleEngine.load, merge=merge)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback

        ret[rank] = traceback.format_exc()
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric", ["IP", "L2"])
def test_sharded_equals_single_index_world2(metric, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, metric, ret, str(tmp_path / "snap")), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)
