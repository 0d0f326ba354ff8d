#!/usr/bin/env python
"""bench.py -- IVF_FLAT search throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU IVF_FLAT (oracle port) on host cores

Workload (config.workload): BASELINE.json configs[1] -- 10M x 768 fp32 IVF_FLAT, nlist 16384, IP, top-10, synthetic
unit-norm Gaussian embeddings (seed 1234 DB / 4321 queries).  One *step* = one batch of `nq` queries through coarse
quantizer -> nprobe list scan -> top-k.  With N GPUs the index is row-sharded (every list's rows dealt round-robin),
centroids are replicated, the coarse pass is split over the ranks, each rank scans its shard and the partial top-k are
exchanged and merged on the device.  The union of the shards is exactly the single index, so results equal the 1-GPU
results; the total database is fixed, so scaling is "strong".

The line carries (keys beyond the base contract are additive):
  value / ms_per_step  queries per second with the query batch already in HBM (CUDA events, max over ranks)
  e2e                  the same batch through the C ABI with HOST buffers (pinned): H2D of the queries and D2H of
                       (dist, ids) inside the timed region
  parity               the TIMED path checked in this run: the list-major output of step 0 against the query-major kernel
                       on the same batch (ids and distances) and against the C oracle on a slice of that same output;
                       at N > 1 against the per-shard query-major search + NCCL all-gather + merge
  roofline             the list scan of the headline step: compulsory bytes / its share of the unprofiled step, against the
                       measured HBM peak (MEASURED_PEAKS.json); roofline_query_major: logical bytes == DRAM bytes regime
  cpu_baseline         oracle/ (C restatement of FAISS IndexIVFFlat, OpenMP over all host cores, BLAS sgemm coarse pass,
                       built -march=native on this box) on a bounded sample of the same queries against the same lists
  clustered            BASELINE.json's metric as stated -- QPS at recall@10 -- on the Zipf-clustered set B (the iid set has no
                       cluster structure: recall 0.02 by construction): headline config, its roofline, parity, and an
                       nprobe x nq sweep with recall@10 per nprobe
  roofline_tiles       nq 4096 / nprobe 128: every list probed ~32x, the tcgen05 tile kernel (scan_lists_ts.cu)
  c5_filtered          BASELINE.json configs[4]: 10M x 2048, 5 % repo/language filter, top-50
  kmeans               one shard-sized Lloyd iteration of configs[3] (tcgen05 3xTF32 contraction) against a measured TF32 peak
  c3 (N = 8)           BASELINE.json configs[2]: 10M x 3072 row-sharded over the box
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "search QPS (IVF_FLAT 10Mx768 fp32, nlist=16384, IP, top-10)"
UNIT = "queries/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n", type=int, default=10_000_000)
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--nlist", type=int, default=16384)
    p.add_argument("--nprobe", type=int, default=32)
    p.add_argument("--nq", type=int, default=1024)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--metric", default="IP", choices=["IP", "L2"])
    p.add_argument("--train-rows", type=int, default=1_000_000)
    p.add_argument("--train-iters", type=int, default=4)
    p.add_argument("--no-extras", action="store_true", help="headline only: skip clustered / sweep / tiles / c5 / kmeans / c3")
    p.add_argument("--sweep-nq", default="1,256,4096", help="batch sizes of the sweep")
    p.add_argument("--sweep-nprobe", default="8,16,32,64,128")
    p.add_argument("--recall-queries", type=int, default=128)
    p.add_argument("--cpu-queries", type=int, default=96, help="queries in the CPU baseline sample / the oracle parity slice")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work of the cpu_baseline sample (repeated until this long)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--scan-variant", type=int, default=0)
    p.add_argument("--dataset", default="iid", choices=["iid", "clustered"],
                   help="synthetic set of the headline: A (iid Gaussian, SURVEY.md 8d) or B (Zipf-clustered)")
    p.add_argument("--coarse-impl", type=int, default=0, help="0 = tcgen05 3xTF32, 1 = fp32 SIMT")
    p.add_argument("--scan-mode", type=int, default=0, help="0 = auto, 1 = query-major, 2 = list-major")
    p.add_argument("--lists-cfg", type=int, default=0, help="tile configuration of the list-major kernel")
    p.add_argument("--tile-rem", type=int, default=0, help="4 = remainders of 5..16 queries become tcgen05 tile items")
    p.add_argument("--mq-fused", type=int, default=0, help="1 = the two list-major page-scan buckets in one launch (measured slower)")
    p.add_argument("--shard-sim", type=int, default=1,
                   help="single-GPU run over ONE rank's shard of a G-way row-sharded index (rows i with i %% G == 0 of the "
                        "same stream, same centroids): what each rank of --gpus G scans, without the exchange")
    p.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                   help="N > 1: p2p = probe rows and partial top-k stored straight into the peers' buffers by the kernels "
                        "that produce them (NVLink peer memory), merge kernel waits on flags; nccl = all-gathers + merge")
    p.add_argument("--inflight", type=int, default=0,
                   help="N > 1, p2p exchange: fused steps in flight per rank (one stream + one exchange + one scratch slot each); "
                        "0 = automatic: 2 from 4 ranks on (measured at N = 8: 1.36M vs 1.26M QPS -- the second step fills the waits for "
                        "the slowest peer), 1 below (N = 2: 423k vs 438k -- little to hide, and the interleaved scans share L2)")
    p.add_argument("--shard-by", default="rows", choices=["rows", "lists"],
                   help="N > 1: deal every list's rows round-robin (rows) or whole lists (list l on rank l %% N)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic data: unit-norm rows, generated on the device
# ------------------------------------------------------------------------------------------------
_LATENT = {}


def gen_rows(torch, n0, n1, dim, seed, device, dataset="iid"):
    """Rows [n0, n1) of the synthetic matrix.  Each 65536-row block has its own
    seeded generator, so any rank can produce any block without materialising the rest.
      iid        set A: x ~ N(0, I_d), L2-normalised (balanced lists, worst case for recall)
      clustered  set B: 4096 latent centres ~ N(0, I_d), Zipf(1.1) cluster sizes, within-cluster noise
                 sigma = 0.3, L2-normalised (skewed lists, like real code embeddings)"""
    blk = 65536
    out = torch.empty((n1 - n0, dim), dtype=torch.float32, device=device)
    if dataset == "clustered":
        key = (dim, str(device))
        if key not in _LATENT:
            g0 = torch.Generator(device=device).manual_seed(555)
            centres = torch.randn((4096, dim), generator=g0, device=device, dtype=torch.float32)
            w = 1.0 / torch.arange(1, 4097, device=device, dtype=torch.float64) ** 1.1
            _LATENT[key] = (centres, (w / w.sum()).to(torch.float32))
        centres, w = _LATENT[key]
    b = n0 // blk
    while b * blk < n1:
        lo, hi = max(n0, b * blk), min(n1, (b + 1) * blk)
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + b)
        full = torch.randn((blk, dim), generator=g, device=device, dtype=torch.float32)
        if dataset == "clustered":
            c = torch.multinomial(w, blk, replacement=True, generator=g)
            full = centres[c] + 0.3 * full
        out[lo - n0 : hi - n0] = full[lo - b * blk : hi - b * blk]
        b += 1
    # in place (x / max(|x|, 1e-12), what F.normalize computes): a 50M x 768 k-means shard does not fit twice
    return out.div_(out.norm(dim=1, keepdim=True).clamp_min_(1e-12))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.th = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample taken
            for ts, line in self.rows:
                f = [s.strip() for s in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except Exception:
                    pass
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def rel_err(a, b):
    """max |a - b| / max(|b|, tiny) over finite entries (distances of found results)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = np.isfinite(a) & np.isfinite(b) & (np.abs(b) < 1e30)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-30)))


def ids_agreement(gi, ri, gd, rd, rtol=1e-5):
    """Fraction of queries whose id rows are identical, and whether every difference sits inside an exact-distance tie
    (north_star: identical ids except exact-distance ties)."""
    gi, ri = np.asarray(gi), np.asarray(ri)
    same = np.all(gi == ri, axis=1)
    only_ties = True
    for r in np.flatnonzero(~same):
        diff = gi[r] != ri[r]
        if not np.allclose(np.asarray(gd)[r][diff], np.asarray(rd)[r][diff], rtol=rtol, atol=0):
            only_ties = False
            break
    return float(same.mean()), only_ties


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_coarse(ivf_c, q, centroids, metric):
    """Coarse similarities with a BLAS sgemm over all host cores (torch's CPU matmul: MKL / OpenBLAS) -- what FAISS's
    IndexFlat quantizer does on the centroids."""
    import torch

    qt, ct = torch.from_numpy(q), torch.from_numpy(centroids)
    s = torch.mm(qt, ct.t())
    if metric == 1:
        s = 2.0 * s - (ct * ct).sum(1)[None, :]
    return np.ascontiguousarray(s.numpy(), dtype=np.float32)


def cpu_search_sample(ivf_c, q, centroids, metric, nprobe, k, probes_gpu, export_list, budget_s=0.0, max_reps=1000):
    """Time coarse + scan + top-k of `q` on the CPU against the lists those queries probe.
    Only the probed lists are copied to the host (compact CSR, list ids remapped).  Returns the median time, its
    coarse / scan split, the bytes one repetition streams, and the result."""
    used = np.unique(probes_gpu)
    remap = -np.ones(centroids.shape[0], dtype=np.int32)
    remap[used] = np.arange(used.size, dtype=np.int32)
    vec_parts, id_parts, sizes = [], [], []
    for l in used:
        v, i, t = export_list(int(l))
        live = (t & np.uint32(0x80000000)) == 0
        vec_parts.append(v[live])
        id_parts.append(i[live])
        sizes.append(int(live.sum()))
    off = np.zeros(used.size + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    vecs = np.concatenate(vec_parts) if vec_parts else np.zeros((0, q.shape[1]), np.float32)
    ids = np.concatenate(id_parts) if id_parts else np.zeros(0, np.int64)
    t_coarse, t_scan = [], []
    out = None
    probes = None
    started = time.perf_counter()
    while True:  # first repetition = warm-up (page faults of the exported lists); then repeat until `budget_s` of CPU work
        if len(t_scan) >= 2 and (time.perf_counter() - started >= budget_s or len(t_scan) > max_reps):
            break
        t0 = time.perf_counter()
        scores = cpu_coarse(ivf_c, q, centroids, metric)
        probes = ivf_c.top_probes(scores, nprobe)
        t1 = time.perf_counter()
        local = remap[probes]
        t2 = time.perf_counter()
        out = ivf_c.scan_search(q, metric, local, off, vecs, ids, k)
        t3 = time.perf_counter()
        t_coarse.append(t1 - t0)
        t_scan.append(t3 - t2)
    if not np.array_equal(np.sort(probes, 1), np.sort(probes_gpu, 1)):
        # a centroid pair closer than fp32 summation-order noise at the nprobe boundary: compare the SCAN on the probes the
        # device used (the coarse pass has its own parity tests; probes_identical reports the rate)
        out = ivf_c.scan_search(q, metric, remap[probes_gpu], off, vecs, ids, k)
    sizes_arr = np.asarray(sizes, dtype=np.int64)
    scanned_rows = int(sizes_arr[remap[probes]].sum())
    tc, ts = statistics.median(t_coarse[1:]), statistics.median(t_scan[1:])
    return {"seconds": tc + ts, "coarse_s": tc, "scan_s": ts, "scan_bytes": scanned_rows * 4 * q.shape[1], "reps": len(t_scan) - 1,
            "cpu_work_s": time.perf_counter() - started}, out, probes


def run_reference(args):
    """--impl reference: the reference path's CPU implementation.  The reference's engine (Milvus /
    knowhere / FAISS) cannot be installed offline, so this is the oracle port (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import ivf_c

    ivf_c.build(native=True)
    cores = ivf_c.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1; rank 0 runs alone and takes every core it may use
    try:
        torch.set_num_threads(cores)
    except Exception:
        pass
    # A bounded sample of the workload that needs no GPU: the same synthetic rows for the lists a
    # query sample probes.  The CPU arm builds its own (smaller) slice of the index: nlist and
    # nprobe as configured, rows = n, but only the probed lists are materialised.
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    n, d, nlist = args.n, args.dim, args.nlist
    nq = args.cpu_queries
    metric = 0 if args.metric == "IP" else 1
    # centroids: `nlist` synthetic rows (k-means quality does not change the CPU cost per query)
    cent = gen_rows(torch, 0, nlist, d, 99, dev).cpu().numpy()
    per_list = max(1, n // nlist)
    q = gen_rows(torch, 0, nq, d, 4321, dev).cpu().numpy()
    scores = cpu_coarse(ivf_c, q, cent, metric)
    probes = ivf_c.top_probes(scores, args.nprobe)
    used = np.unique(probes)
    remap = -np.ones(nlist, dtype=np.int32)
    remap[used] = np.arange(used.size, dtype=np.int32)
    rng = np.random.default_rng(1234)
    vecs = np.empty((used.size * per_list, d), dtype=np.float32)
    for j, l in enumerate(used):  # rows of list l scatter around its centroid
        blk = cent[l][None, :] + 0.5 * rng.standard_normal((per_list, d)).astype(np.float32)
        vecs[j * per_list : (j + 1) * per_list] = blk / np.linalg.norm(blk, axis=1, keepdims=True)
    ids = np.arange(vecs.shape[0], dtype=np.int64)
    off = np.arange(used.size + 1, dtype=np.int64) * per_list
    times, coarse = [], []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        s = cpu_coarse(ivf_c, q, cent, metric)
        p = ivf_c.top_probes(s, args.nprobe)
        t1 = time.perf_counter()
        ivf_c.scan_search(q, metric, remap[p], off, vecs, ids, args.k)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            coarse.append(t1 - t0)
    total = sum(times)
    qps = nq * len(times) / total
    scan_bytes = nq * args.nprobe * per_list * 4 * d
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = {nq} of the {args.nq} queries of a step x nprobe {args.nprobe} over lists of {per_list} rows "
                                   f"({n}/{nlist}); oracle/ivf_oracle.c built -march=native, OpenMP over queries, BLAS sgemm coarse pass",
                         "coarse_share": sum(coarse) / total,
                         "scan_host_GBps": scan_bytes * len(times) / (total - sum(coarse)) / 1e9},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {
        "workload": f"IVF_FLAT {args.n}x{args.dim} fp32, nlist={args.nlist}, nprobe={args.nprobe}, "
                    f"nq={args.nq}/step, top-{args.k}, metric={args.metric}, {args.dataset} synthetic set "
                    f"(BASELINE.json configs[1])",
        "n": args.n, "dim": args.dim, "nlist": args.nlist, "nprobe": args.nprobe, "nq": args.nq,
        "k": args.k, "metric": args.metric, "dataset": args.dataset, "shard_by": args.shard_by if args.gpus > 1 else None,
        "shard_sim": args.shard_sim if args.shard_sim > 1 else None,
        "l2_policy": "inputs larger than L2: every step streams nq*nprobe lists (>> 126 MB) and rotates query batches",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def build_index(c, n, d, nlist, dataset, metric="IP", tags=None):
    """Train on a prefix (rank 0), broadcast the centroids, add this rank's rows of the seeded stream."""
    torch, dist, sb, args = c.torch, c.dist, c.sb, c.args
    t0 = time.time()
    g = sb.IVFFlatIndex(d, nlist=nlist, metric=metric, device=c.local)
    if args.scan_variant:
        g.set_param("scan_variant", args.scan_variant)
    if args.coarse_impl:
        g.set_param("coarse_impl", args.coarse_impl)
    if args.scan_mode:
        g.set_param("scan_mode", args.scan_mode)
    if args.lists_cfg:
        g.set_param("lists_cfg", args.lists_cfg)
    if args.mq_fused:
        g.set_param("mq_fused", 1)
    if args.tile_rem:
        g.set_param("tile_rem", args.tile_rem)
    cent = torch.empty((nlist, d), dtype=torch.float32, device=c.dev)
    if c.rank == 0:
        tr = gen_rows(torch, 0, min(args.train_rows, n), d, 1234, c.dev, dataset)
        g.train(tr, niter=args.train_iters, max_points_per_centroid=0)
        cent.copy_(torch.from_numpy(g.get_centroids()))
        del tr
    if c.world > 1:
        dist.broadcast(cent, 0)
        if c.rank != 0:
            g.set_centroids(cent)
    t_train = time.time() - t0
    chunk = 1 << 20 if d <= 1024 else 1 << 18
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = gen_rows(torch, s, e, d, 1234, c.dev, dataset)
        ids = torch.arange(s, e, device=c.dev, dtype=torch.int64)
        rt = lt = None
        if tags is not None:
            rt, lt = tags(torch, s, e, c.dev)
        if c.world > 1 and args.shard_by == "rows":  # deal rows round-robin: rank r keeps global rows i with i % world == r
            sl = slice(c.rank, None, c.world)
            g.add(x[sl].contiguous(), ids[sl].contiguous(), None if rt is None else rt[sl].contiguous(),
                  None if lt is None else lt[sl].contiguous())
        elif c.world > 1:  # whole lists: rank r keeps the rows whose list l has l % world == r
            lists = g.assign(x)
            keep = torch.nonzero(lists % c.world == c.rank).squeeze(1)
            g.add(x[keep].contiguous(), ids[keep].contiguous(), lists=lists[keep].contiguous())
            del lists, keep
        elif args.shard_sim > 1 and args.shard_by == "lists":  # rank 0 of a G-way LIST-sharded index: lists l with l % G == 0
            lists = g.assign(x)
            keep = torch.nonzero(lists % args.shard_sim == 0).squeeze(1)
            g.add(x[keep].contiguous(), ids[keep].contiguous(), lists=lists[keep].contiguous())
            del lists, keep
        elif args.shard_sim > 1:
            g.add(x[0::args.shard_sim].contiguous(), ids[0::args.shard_sim].contiguous())
        else:
            g.add(x, ids, rt, lt)
        del x, ids
    torch.cuda.synchronize()
    return g, {"train": t_train, "total": time.time() - t0}


def time_search(c, g, q, k, nprobe, reps, **kw):
    torch = c.torch
    for _ in range(2):
        g.search(q, k, nprobe=nprobe, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = g.search(q, k, nprobe=nprobe, **kw)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def profiled(c, g, q, k, nprobe, reps=3, **kw):
    """Per-phase CUDA events of the library (outside any headline timing): mean over `reps` calls."""
    g.set_profiling(True)
    acc = None
    for _ in range(reps):
        g.search(q, k, nprobe=nprobe, **kw)
        c.torch.cuda.synchronize()
        t = g.last_search_times()
        vals = [t.coarse_ms, t.probe_select_ms, t.plan_ms, t.scan_ms, t.topk_ms, t.total_ms, t.scanned_rows, t.unique_rows,
                t.scan_launches, t.total_launches]
        acc = vals if acc is None else [a + b for a, b in zip(acc, vals)]
    g.set_profiling(False)
    keys = ["coarse_ms", "select_ms", "plan_ms", "scan_ms", "topk_ms", "total_ms", "scanned_rows", "unique_rows", "scan_launches",
            "total_launches"]
    return {kk: v / reps for kk, v in zip(keys, acc)}


def scan_roofline(c, prof, step_ms, d, kernel, peaks, peak_src, traffic=None):
    """Roofline of the list scan inside a step of `step_ms` (unprofiled): the scan's time is its share of the profiled
    pass applied to the unprofiled step (the profiling events add ~5 % to a step)."""
    list_major = prof["unique_rows"] > 0
    bytes_alg = (prof["unique_rows"] if list_major else prof["scanned_rows"]) * 4 * d
    share = prof["scan_ms"] / max(prof["total_ms"], 1e-9)
    kernel_ms = step_ms * share
    achieved = bytes_alg / (kernel_ms / 1e3) / 1e9
    return {
        "bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": achieved / peaks["hbm_gbs"], "frac_of_8TBps_nominal": achieved / 8000.0, "peak_source": peak_src, "traffic": traffic,
        "algorithmic_bytes_per_launch": bytes_alg, "bytes_are": "compulsory (each DISTINCT probed list once)" if list_major else
        "logical (one pass per (query, list) pair)",
        "logical_bytes_per_launch": prof["scanned_rows"] * 4 * d, "kernel_ms": kernel_ms, "kernel_ms_profiled_pass": prof["scan_ms"],
        "kernel_share_of_step": share, "step_ms_unprofiled": step_ms, "step_ms_profiled_pass": prof["total_ms"],
    }


def list_major_name(args, d):
    tiles = ("scan_lists_ts_kernel (tcgen05 tiles, list rows from tensor memory)" if args.metric == "IP" and d % 32 == 0 and args.lists_cfg not in (1, 2)
             else "scan_lists_kernel (FFMA tiles of 32 queries)")
    return ("list-major scan: scan_mq_kernel<4> (remainders of 1..4 queries per list) + scan_mq_kernel<8> (5..16) + "
            f"{tiles} + count / plan / fill")


def check_parity(c, g, q, k, nprobe, out_d, out_i, cq, with_oracle=True, cpu_budget_s=0.0):
    """The batch output (list-major when the automatic switch picked it) against the query-major kernel on the same batch,
    and against the C oracle on the first `cq` queries of that SAME output."""
    torch = c.torch
    res = {"path": "list-major" if c.last_list_major else "query-major", "rtol": 1e-5}
    gd, gi = out_d.cpu().numpy(), out_i.cpu().numpy()
    g.set_param("scan_mode", 1)
    qd, qi = g.search(q, k, nprobe=nprobe)
    g.set_param("scan_mode", c.args.scan_mode)
    torch.cuda.synchronize()
    qd, qi = qd.cpu().numpy(), qi.cpu().numpy()
    same, ties = ids_agreement(gi, qi, gd, qd)
    res["vs_query_major"] = {"queries": int(q.shape[0]), "ids_identical": same, "differences_only_in_ties": ties,
                             "max_rel_err": rel_err(gd, qd)}
    ok = ties and res["vs_query_major"]["max_rel_err"] <= 1e-5
    if with_oracle:
        from oracle import ivf_c

        ivf_c.build(native=True)
        ivf_c.use_all_cores()
        qs = q[:cq].cpu().numpy()
        probes = g.probe(qs, nprobe)
        cpu, (cd, ci), cprobes = cpu_search_sample(ivf_c, qs, g.get_centroids(), g.metric, nprobe, k, probes, g.export_list,
                                                   budget_s=cpu_budget_s)
        same, ties = ids_agreement(gi[:cq], ci, gd[:cq], cd)
        res["vs_oracle"] = {"queries": int(cq), "sliced_from": "the same batch output", "ids_identical": same,
                            "differences_only_in_ties": ties, "max_rel_err": rel_err(gd[:cq], cd),
                            "probes_identical": float(np.mean(np.all(np.sort(probes, 1) == np.sort(cprobes, 1), axis=1)))}
        ok = ok and ties and res["vs_oracle"]["max_rel_err"] <= 1e-5
        res["_cpu"] = cpu
    res["ok"] = bool(ok)
    return res


def recall_at_k(c, g, q, k, ids_ann):
    """recall@k of `ids_ann` against exact search (exhaustive probe of every shard, merged)."""
    torch, dist, sb = c.torch, c.dist, c.sb
    rq = ids_ann.shape[0]
    gd, gi = g.search(q[:rq], k, nprobe=g.nlist)
    if c.world > 1:
        gd_all = torch.empty((c.world, rq, k), dtype=torch.float32, device=c.dev)
        gi_all = torch.empty((c.world, rq, k), dtype=torch.int64, device=c.dev)
        dist.all_gather_into_tensor(gd_all, gd.contiguous())
        dist.all_gather_into_tensor(gi_all, gi.contiguous())
        gd, gi = sb.merge_topk(gd_all, gi_all, k, c.args.metric, c.local)
    torch.cuda.synchronize()
    ia, ie = ids_ann.cpu().numpy(), gi.cpu().numpy()
    return float(np.mean([len(np.intersect1d(ia[r][ia[r] >= 0], ie[r])) / k for r in range(rq)]))


def extra_clustered(c, peaks, peak_src):
    """BASELINE.json's metric on the set where recall means something: headline config + sweep with recall@10."""
    torch, args = c.torch, c.args
    n, d, nlist, k = args.n, args.dim, args.nlist, args.k
    g, build_s = build_index(c, n, d, nlist, "clustered", args.metric)
    out = {"dataset": "clustered (set B: 4096 Zipf(1.1) latent centres, sigma 0.3)", "build_s": build_s}
    q = gen_rows(torch, 0, args.nq, d, 4321, c.dev, "clustered")
    ms, (od, oi) = time_search(c, g, q, k, args.nprobe, max(5, args.steps // 2))
    prof = profiled(c, g, q, k, args.nprobe)
    c.last_list_major = prof["unique_rows"] > 0
    out.update({"nq": args.nq, "nprobe": args.nprobe, "ms_per_step": ms, "qps": args.nq / ms * 1e3,
                "recall_at_10": recall_at_k(c, g, q, k, oi[: args.recall_queries].clone()),
                "roofline": scan_roofline(c, prof, ms, d, list_major_name(args, d) if c.last_list_major else "scan_pages_kernel (query-major)",
                                          peaks, peak_src),
                "parity": {kk: v for kk, v in check_parity(c, g, q, k, args.nprobe, od, oi, min(32, args.cpu_queries)).items() if kk != "_cpu"}})
    # recall@10 depends on nprobe only: measured once per nprobe on a fixed set of queries against exhaustive search
    rq = max(args.recall_queries, 256)
    qr = gen_rows(torch, 0, rq, d, 99991, c.dev, "clustered")
    exact = g.search(qr, k, nprobe=nlist)[1].cpu().numpy()
    recall_by_nprobe = {}
    for np_ in [int(v) for v in args.sweep_nprobe.split(",")]:
        ia = g.search(qr, k, nprobe=np_)[1].cpu().numpy()
        recall_by_nprobe[np_] = float(np.mean([len(np.intersect1d(ia[r][ia[r] >= 0], exact[r])) / k for r in range(rq)]))
    out["recall_at_10_by_nprobe"] = {str(kk): v for kk, v in recall_by_nprobe.items()}
    sweep = []
    for nq_ in [int(v) for v in args.sweep_nq.split(",")]:
        qs = gen_rows(torch, 0, nq_, d, 777 + nq_, c.dev, "clustered")
        for np_ in [int(v) for v in args.sweep_nprobe.split(",")]:
            ms_, _ = time_search(c, g, qs, k, np_, 50 if nq_ <= 16 else (20 if nq_ <= 256 else 3))
            p = profiled(c, g, qs, k, np_, reps=1)
            lm = p["unique_rows"] > 0
            sweep.append({"nq": nq_, "nprobe": np_, "ms": ms_, "qps": nq_ / ms_ * 1e3, "recall_at_10": recall_by_nprobe[np_],
                          "scan": "list-major" if lm else "query-major",
                          "scan_GBps": (p["unique_rows"] if lm else p["scanned_rows"]) * 4 * d / max(p["scan_ms"], 1e-6) / 1e6,
                          "logical_GBps": p["scanned_rows"] * 4 * d / ms_ / 1e6})
    out["sweep"] = sweep
    g.close()
    del g
    torch.cuda.empty_cache()
    return out


def lookup_traffic(args, nq, nprobe, dataset, scan):
    """DRAM bytes per launch of the scan kernels from the committed ncu captures (profiles/scan_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            for ent in json.load(f)["captures"]:
                mm = ent["match"]
                if ent.get("scan", "query-major") == scan and all(mm[key] == val for key, val in (
                        ("n", args.n), ("dim", args.dim), ("nlist", args.nlist), ("nprobe", nprobe), ("nq", nq), ("dataset", dataset))):
                    return ent["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def extra_tiles(c, g, peaks, peak_src):
    """nq 4096 / nprobe 128 on the headline index: every list probed ~32x -> the tcgen05 tile kernel carries the scan."""
    torch, args = c.torch, c.args
    d, k = args.dim, args.k
    q = gen_rows(torch, 0, 4096, d, 9001, c.dev, args.dataset)
    ms, (od, oi) = time_search(c, g, q, k, 128, 5)
    prof = profiled(c, g, q, k, 128)
    c.last_list_major = prof["unique_rows"] > 0
    roof = scan_roofline(c, prof, ms, d, list_major_name(args, d), peaks, peak_src,
                         lookup_traffic(args, 4096, 128, args.dataset, "list-major") if args.shard_sim == 1 else None)
    # parity of this batch: the same 4096 queries through the query-major kernel
    g.set_param("scan_mode", 1)
    qd, qi = g.search(q, k, nprobe=128)
    g.set_param("scan_mode", args.scan_mode)
    torch.cuda.synchronize()
    same, ties = ids_agreement(oi.cpu().numpy(), qi.cpu().numpy(), od.cpu().numpy(), qd.cpu().numpy())
    return {"nq": 4096, "nprobe": 128, "ms_per_step": ms, "qps": 4096 / ms * 1e3, "phases_ms": {kk: prof[kk] for kk in
            ("coarse_ms", "select_ms", "plan_ms", "scan_ms", "topk_ms")}, "roofline": roof,
            "parity": {"vs_query_major": {"queries": 4096, "ids_identical": same, "differences_only_in_ties": ties,
                                          "max_rel_err": rel_err(od.cpu().numpy(), qd.cpu().numpy())}}}


def extra_c5(c, peaks, peak_src):
    """BASELINE.json configs[4]: 10M x 2048 (jina-v4 dims), repo / language masks at ~5 % selectivity, top-50."""
    torch, args = c.torch, c.args
    n, d, nlist, k, nq = args.n, 2048, args.nlist, 50, 256

    def tags(torch, s, e, dev):  # repo ~ Zipf over 200 repos, language in {python, cpp} (tree_sitter_chunker.py:150-156)
        gg = torch.Generator(device=dev).manual_seed(31337 + s)
        w = 1.0 / torch.arange(1, 201, device=dev, dtype=torch.float32)
        repo = torch.multinomial(w / w.sum(), e - s, replacement=True, generator=gg).to(torch.int32)
        lang = (torch.rand(e - s, device=dev, generator=gg) < 0.35).to(torch.uint8)
        return repo, lang

    g, build_s = build_index(c, n, d, nlist, "clustered", "IP", tags=tags)
    # predicate: language cpp (35 %) and repos 1..? chosen to pass ~5 % of the rows
    w = 1.0 / np.arange(1, 201)
    w /= w.sum()
    repos, acc = [], 0.0
    for r in range(199, -1, -1):  # from the tail of the Zipf: many small repos, as a user selecting a handful of projects
        if acc + w[r] > 0.05 / 0.35:
            continue
        repos.append(r)
        acc += w[r]
    q = gen_rows(torch, 0, nq, d, 4321, c.dev, "clustered")
    rows = []
    for np_ in (16, 32, 64):
        ms_f, (fd, fi) = time_search(c, g, q, k, np_, 10, repos=repos, langs=[1])
        ms_u, _ = time_search(c, g, q, k, np_, 5)
        p = profiled(c, g, q, k, np_, reps=1, repos=repos, langs=[1])
        # parity of the filtered list-major result against the filtered query-major kernel
        g.set_param("scan_mode", 1)
        qd, qi = g.search(q, k, nprobe=np_, repos=repos, langs=[1])
        g.set_param("scan_mode", args.scan_mode)
        torch.cuda.synchronize()
        same, ties = ids_agreement(fi.cpu().numpy(), qi.cpu().numpy(), fd.cpu().numpy(), qd.cpu().numpy())
        rows.append({"nprobe": np_, "filtered_ms": ms_f, "filtered_qps": nq / ms_f * 1e3, "unfiltered_ms": ms_u, "unfiltered_qps": nq / ms_u * 1e3,
                     "found_per_query": float((fi >= 0).sum().item()) / nq,
                     "filtered_algorithmic_GBps": (acc * 0.35 * p["scanned_rows"] * 4 * d + 4 * p["scanned_rows"]) / max(p["scan_ms"], 1e-6) / 1e6,
                     "ids_identical_to_query_major": same, "differences_only_in_ties": ties})
    g.close()
    del g
    torch.cuda.empty_cache()
    return {"workload": f"IVF_FLAT {n}x{d}, nlist={nlist}, top-{k}, nq={nq}, clustered set, predicate passes {acc * 0.35:.3f} of the rows "
                        f"({len(repos)} repos x language cpp)", "selectivity": acc * 0.35, "build_s": build_s, "rows": rows}


def extra_kmeans(c, peaks, peak_src):
    """One Lloyd iteration over a slice of a configs[3] shard (768-d, nlist 65536): fused tcgen05 3xTF32 contraction + argmax,
    fp64 accumulation.  Tensor-pipe fraction against a TF32 matmul peak measured here (torch.matmul, library GEMM: measurement only)."""
    torch, sb = c.torch, c.sb
    n, d, nlist = 1_000_000, 768, 65536
    x = gen_rows(torch, 0, n, d, 1234, c.dev, "clustered")
    eng = sb.IVFFlatIndex(d, nlist=nlist, metric="IP", device=c.local)
    from semcode_b200.index import kmeans_init_rows

    eng.set_centroids(x[torch.from_numpy(kmeans_init_rows(n, nlist, 1234)).to(c.dev)])
    sums, counts, obj = eng.kmeans_buffers()
    times = []
    for _ in range(3):
        sums.zero_(); counts.zero_(); obj.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.kmeans_step(x, sums, counts, obj)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
        eng.kmeans_update(sums, counts)
    eng.close()
    del eng, x, sums
    torch.cuda.empty_cache()
    # TF32 peak of this box: 8192^3 torch.matmul with TF32 allowed, best of 10
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    m = 8192
    A = torch.randn((m, m), device=c.dev)
    B = torch.randn((m, m), device=c.dev)
    best = 1e9
    for _ in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        torch.matmul(A, B)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    torch.backends.cuda.matmul.allow_tf32 = prev
    del A, B
    torch.cuda.empty_cache()
    tf32_peak = 2.0 * m ** 3 / (best / 1e3) / 1e12
    it_s = min(times[1:]) / 1e3
    flop = 2.0 * n * nlist * d
    return {"workload": f"k-means assignment pass {n} x {d}, nlist={nlist} (a slice of one BASELINE.json configs[3] shard)",
            "s_per_iteration": it_s, "fp32_equiv_tflops": flop / it_s / 1e12, "mma_tf32_tflops": 3 * flop / it_s / 1e12,
            "tf32_matmul_peak_measured_tflops": tf32_peak, "tensor_pipe_frac_of_measured_tf32": 3 * flop / it_s / 1e12 / tf32_peak,
            "extrapolated_full_shard_6.25M_rows_s": it_s * 6.25}


def run_ours(args):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: semcode_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    import semcode_b200 as sb

    c = Ctx()
    c.torch, c.dist, c.sb, c.args = torch, dist, sb, args
    c.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = rank = int(os.environ.get("RANK", "0"))
    c.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    c.dev = dev = torch.device("cuda", local)
    c.want_cpu_baseline = world == 1 and not args.no_cpu_baseline
    c.last_list_major = False
    if world > 1:
        # NCCL_DEBUG is left as the launcher set it; NCCL logs on fd 1, which main() has pointed at stderr for the run
        dist.init_process_group("nccl", device_id=dev)
    n, d, nlist, k, nprobe, nq = args.n, args.dim, args.nlist, args.k, args.nprobe, args.nq

    g, build_s = build_index(c, n, d, nlist, args.dataset, args.metric)

    # ---- queries: a few rotating batches ----------------------------------------------------------
    nb = 4
    qall = gen_rows(torch, 0, nb * nq, d, 4321, dev, args.dataset)
    qb = [qall[i * nq : (i + 1) * nq].contiguous() for i in range(nb)]
    qhost = [t.cpu().pin_memory() for t in qb]
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    gat_d = torch.empty((world, nq, k), dtype=torch.float32, device=dev) if world > 1 else None
    gat_i = torch.empty((world, nq, k), dtype=torch.int64, device=dev) if world > 1 else None
    host_d = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    host_i = torch.empty((nq, k), dtype=torch.int64).pin_memory()

    per = (nq + world - 1) // world
    my_lo, my_hi = min(nq, rank * per), min(nq, (rank + 1) * per)
    probe_mine = torch.full((per, min(nprobe, nlist)), -1, dtype=torch.int32, device=dev)
    probe_all = torch.empty((world * per, min(nprobe, nlist)), dtype=torch.int32, device=dev)

    ex, ex_err, exs, lanes, lane_out = None, None, [], [], []
    if world > 1 and args.exchange != "nccl" and args.shard_by == "rows":
        ok = torch.zeros(1, dtype=torch.int32, device=dev)
        try:
            from semcode_b200.index import PeerExchange

            # two lanes exist at every N > 1: the end-to-end loop always pipelines two steps (the copies and the host's
            # turn-around of one step hide behind the other); the device-resident loop uses `nlv` of them
            exs = [PeerExchange(local, None, 64 << 20) for _ in range(max(2, args.inflight))]
            ok += 1
        except Exception as e:
            ex_err = f"{type(e).__name__}: {e}"
        dist.all_reduce(ok)
        if int(ok.item()) != world:
            exs = []
            if args.exchange == "p2p":
                raise SystemExit(f"--exchange p2p: peer-memory exchange unavailable ({ex_err})")
        if exs:
            ex = exs[0]
            lanes = [torch.cuda.Stream(device=dev) for _ in exs]
            lane_out = [(torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev))
                        for _ in exs]
    nl = len(lanes)
    nlv = min(nl, args.inflight if args.inflight > 0 else (2 if world >= 4 else 1))  # lanes of the device-resident loop

    def search_sharded(qs):
        if ex is not None:  # ONE C-ABI call: split coarse pass + probe scatter + scan + top-k scatter + waiting merge
            g.search(qs, k, nprobe=nprobe, out=(out_d, out_i), exchange=ex)
            return out_d, out_i
        # the coarse pass is split over the ranks too (centroids are replicated): rank r ranks the centroids
        # for its 1/G of the batch, one small all-gather distributes the probe table
        if my_hi > my_lo:
            probe_mine[: my_hi - my_lo] = g.probe(qs[my_lo:my_hi], nprobe)
        dist.all_gather_into_tensor(probe_all, probe_mine)
        pl = probe_all[:nq]
        if args.shard_by == "lists":  # probes of lists that live elsewhere are skipped (-1)
            pl = torch.where(pl % world == rank, pl, torch.full_like(pl, -1))
        g.search(qs, k, lists=pl, out=(out_d, out_i))
        dist.all_gather_into_tensor(gat_d, out_d)
        dist.all_gather_into_tensor(gat_i, out_i)
        return sb.merge_topk(gat_d, gat_i, k, args.metric, local)

    sim_lists = None
    if world == 1 and args.shard_sim > 1 and args.shard_by == "lists":  # probe tables of the simulated rank, computed up front
        sim_lists = []
        for t in qb:
            pl = g.probe(t, nprobe)
            sim_lists.append(torch.where(pl % args.shard_sim == 0, pl, torch.full_like(pl, -1)).contiguous())

    def step_device(i):
        if nl > 0:  # lane i % nlv: its own stream, exchange and (inside the library) scratch slot
            j = i % nlv
            with torch.cuda.stream(lanes[j]):
                g.search(qb[i % nb], k, nprobe=nprobe, out=lane_out[j], exchange=exs[j])
            return lane_out[j]
        if world > 1:
            return search_sharded(qb[i % nb])
        if sim_lists is not None:
            g.search(qb[i % nb], k, lists=sim_lists[i % nb], out=(out_d, out_i))
            return out_d, out_i
        g.search(qb[i % nb], k, nprobe=nprobe, out=(out_d, out_i))
        return out_d, out_i

    host_lane = [(torch.empty((nq, k), dtype=torch.float32).pin_memory(), torch.empty((nq, k), dtype=torch.int64).pin_memory())
                 for _ in lanes]

    def step_e2e(i):
        if world == 1:
            g.search(qhost[i % nb], k, nprobe=nprobe, out=(host_d, host_i))  # C ABI, host buffers
            return host_d, host_i
        if nl > 0:
            # the same pipeline end to end: lane j's previous step (i - nl) is read back to the host before the lane is reused
            j = i % nl
            lanes[j].synchronize()
            with torch.cuda.stream(lanes[j]):
                qd = qhost[i % nb].to(dev, non_blocking=True)
                g.search(qd, k, nprobe=nprobe, out=lane_out[j], exchange=exs[j])
                host_lane[j][0].copy_(lane_out[j][0], non_blocking=True)
                host_lane[j][1].copy_(lane_out[j][1], non_blocking=True)
            return host_lane[j]
        qd = qhost[i % nb].to(dev, non_blocking=True)
        md, mi = search_sharded(qd)
        host_d.copy_(md, non_blocking=True)
        host_i.copy_(mi, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_d, host_i

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fork():
        cur = torch.cuda.current_stream()
        for st in lanes:
            st.wait_stream(cur)

    def join():
        cur = torch.cuda.current_stream()
        for st in lanes:
            cur.wait_stream(st)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        fork()
        for i in range(steps):
            fn(i)
        join()
        e1.record()
        barrier()
        w1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), w0, w1

    # under `ncu --profile-from-start off` only the search steps are captured, not the index build
    torch.cuda.cudart().cudaProfilerStart()
    fork()
    for i in range(args.warmup):
        step_device(i)
        step_e2e(i)
    join()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # device-resident timing (value): the plain product call, no per-phase events inside the timed region
    ms_total, w0, w1 = timed(step_device, args.steps)
    ms_e2e, _, w1 = timed(step_e2e, args.steps)
    torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    step_ms = ms_total / args.steps

    # ---- the timed path, checked: step 0's output ----------------------------------------------------
    fork()
    d0, i0 = step_device(0)
    join()
    d0, i0 = d0.clone(), i0.clone()
    torch.cuda.synchronize()
    # per-phase CUDA events (recorded by the library on the launching stream): a separate pass over the same inputs
    g.set_profiling(True)
    prof_acc = None
    nprof = min(args.steps, 8)
    for i in range(nprof):
        fork()
        step_device(i)
        join()
        torch.cuda.synchronize()
        t = g.last_search_times()
        vals = [t.coarse_ms, t.probe_select_ms, t.plan_ms, t.scan_ms, t.topk_ms, t.total_ms, t.scanned_rows, t.unique_rows,
                t.scan_launches, t.total_launches]
        prof_acc = vals if prof_acc is None else [a + b for a, b in zip(prof_acc, vals)]
    g.set_profiling(False)
    prof = {kk: v / nprof for kk, v in zip(["coarse_ms", "select_ms", "plan_ms", "scan_ms", "topk_ms", "total_ms", "scanned_rows",
                                            "unique_rows", "scan_launches", "total_launches"], prof_acc)}
    c.last_list_major = list_major = prof["unique_rows"] > 0
    launches_per_step = int(round(prof["total_launches"])) + (4 if world > 1 and ex is None else 0)

    parity = None
    if sim_lists is not None:
        parity = {"path": "shard simulation (lists)", "ok": None}
    elif world == 1:
        parity = check_parity(c, g, qb[0], k, nprobe, d0, i0, min(args.cpu_queries, nq), with_oracle=True,
                              cpu_budget_s=args.cpu_seconds if c.want_cpu_baseline else 0.0)
    else:
        # the exchange route against the plain recipe: every rank scans its shard with the query-major kernel, partials
        # all-gathered by NCCL, merged on the device
        cq = min(256, nq)
        g.set_param("scan_mode", 1)
        ld, li = g.search(qb[0][:cq].contiguous(), k, nprobe=nprobe)
        g.set_param("scan_mode", args.scan_mode)
        pd = torch.empty((world, cq, k), dtype=torch.float32, device=dev)
        pi = torch.empty((world, cq, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(pd, ld.contiguous())
        dist.all_gather_into_tensor(pi, li.contiguous())
        rd, ri = sb.merge_topk(pd, pi, k, args.metric, local)
        torch.cuda.synchronize()
        same, ties = ids_agreement(i0[:cq].cpu().numpy(), ri.cpu().numpy(), d0[:cq].cpu().numpy(), rd.cpu().numpy())
        err = rel_err(d0[:cq].cpu().numpy(), rd.cpu().numpy())
        agree = torch.tensor([1 if (ties and err <= 1e-5) else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        parity = {"path": ("p2p-fused exchange" if ex is not None else "nccl all-gather + merge") + (", list-major" if list_major else ", query-major"),
                  "vs_per_shard_query_major_plus_nccl_merge": {"queries": cq, "ids_identical": same, "differences_only_in_ties": ties,
                                                              "max_rel_err": err},
                  "ok": bool(int(agree.item()) == 1), "rtol": 1e-5}

    # the query-major kernel on the same workload (the regime the HBM-fraction claim of SURVEY.md 8d is made in:
    # it streams every probed list once per (query, list) pair, so logical bytes == DRAM bytes)
    qm = None
    if list_major and not args.scan_mode and world == 1 and sim_lists is None:
        g.set_param("scan_mode", 1)
        qm_ms, _ = time_search(c, g, qb[1], k, nprobe, 4)
        qm_prof = profiled(c, g, qb[1], k, nprobe, reps=2)
        g.set_param("scan_mode", 0)
        qm = (qm_ms, qm_prof)

    recall = recall_at_k(c, g, qb[0], k, i0[: min(args.recall_queries, nq)]) if sim_lists is None else None

    if rank != 0:
        if world > 1 and world == 8 and not args.no_extras:
            g.close()
            del g
            torch.cuda.empty_cache()
            run_c3(c)  # every rank takes part
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    traffic, traffic_lm = None, None
    try:  # DRAM bytes of the scan kernels from the committed ncu --set full captures of this workload (1 GPU)
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            for ent in json.load(f)["captures"]:
                mm = ent["match"]
                if world == 1 and args.shard_sim == 1 and all(mm[key] == val for key, val in (
                        ("n", n), ("dim", d), ("nlist", nlist), ("nprobe", nprobe), ("nq", nq), ("dataset", args.dataset))):
                    if ent.get("scan", "query-major") == "list-major":
                        traffic_lm = ent["dram_bytes_per_launch"]
                    else:
                        traffic = ent["dram_bytes_per_launch"]
    except Exception:
        traffic, traffic_lm = None, None
    qps = nq * args.steps / (ms_total / 1e3)
    e2e_qps = nq * args.steps / (ms_e2e / 1e3)
    line = {
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "recall_at_10": recall,
        "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps,
        "parity": {kk: v for kk, v in parity.items() if kk != "_cpu"},
        "roofline": scan_roofline(c, prof, step_ms, d, list_major_name(args, d) if list_major else "scan_pages_kernel (query-major)",
                                  peaks, peak_src, traffic_lm if list_major else traffic),
        "roofline_query_major": None if qm is None else dict(
            scan_roofline(c, qm[1], qm[0], d, "scan_pages_kernel (query-major, forced with scan_mode=1 on the same workload)", peaks,
                          peak_src, traffic), qps_query_major=nq / qm[0] * 1e3),
        "phases_ms": {kk[:-3]: prof[kk] for kk in ("coarse_ms", "select_ms", "plan_ms", "scan_ms", "topk_ms")},
        "clocks": clocks,
        "build_s": build_s,
    }
    if world > 1:
        line["config"]["exchange"] = (f"p2p-fused (peer-memory stores + flags, no collective call), {nlv} steps in flight per rank "
                                      f"({nl} in the end-to-end loop)"
                                      if ex is not None else f"nccl all-gather + merge ({ex_err or 'requested'})")
        if ex is not None:
            line["exchange_timed_out"] = any(e.status()[0] for e in exs)

    # ---- CPU baseline: timed inside check_parity on the same lists (bounded sample) ----------------------
    if world == 1:
        if c.want_cpu_baseline and "_cpu" in parity:
            from oracle import ivf_c

            cpu = parity["_cpu"]
            cq = min(args.cpu_queries, nq)
            line["cpu_baseline"] = {
                "value": cq / cpu["seconds"], "unit": UNIT, "cores": ivf_c.num_threads(), "kind": "port",
                "sample": f"{cq} of the {nq} step-0 queries, same centroids / lists / nprobe; median of {cpu['reps']} repetitions after one "
                          f"warm-up ({cpu['cpu_work_s']:.1f} s of CPU work); oracle/ivf_oracle.c built -march=native on this box, "
                          f"OpenMP over queries, coarse pass = BLAS sgemm; its result is what parity.vs_oracle compares",
                "seconds": cpu["seconds"], "coarse_share": cpu["coarse_s"] / cpu["seconds"],
                "scan_host_GBps": cpu["scan_bytes"] / cpu["scan_s"] / 1e9,
            }
        else:
            line["cpu_baseline"] = None

    # ---- the rest of BASELINE.json's configs, each guarded: a failure becomes its own key, never a lost headline ---
    if not args.no_extras and args.shard_sim == 1:
        if world == 1:
            for name, fn in (("roofline_tiles", lambda: extra_tiles(c, g, peaks, peak_src)),):
                try:
                    line[name] = fn()
                except Exception as e:  # noqa: BLE001
                    line[name] = {"error": f"{type(e).__name__}: {e}"}
        g.close()
        del g
        torch.cuda.empty_cache()
        if world == 1:
            for name, fn in (("clustered", lambda: extra_clustered(c, peaks, peak_src)), ("c5_filtered", lambda: extra_c5(c, peaks, peak_src)),
                             ("kmeans", lambda: extra_kmeans(c, peaks, peak_src))):
                try:
                    line[name] = fn()
                except Exception as e:  # noqa: BLE001
                    line[name] = {"error": f"{type(e).__name__}: {e}"}
                    torch.cuda.empty_cache()
        elif world == 8:
            try:
                line["c3"] = run_c3(c)
            except Exception as e:  # noqa: BLE001
                line["c3"] = {"error": f"{type(e).__name__}: {e}"}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c3(c):
    """BASELINE.json configs[2]: 10M x 3072 fp32 (123 GB) row-sharded over the 8 GPUs of the box, nq 1024, nprobe 32, top-10."""
    torch, dist, sb, args = c.torch, c.dist, c.sb, c.args
    n, d, nlist, k, nprobe, nq = args.n, 3072, args.nlist, args.k, args.nprobe, args.nq
    g, build_s = build_index(c, n, d, nlist, "clustered", "IP")
    from semcode_b200.index import PeerExchange

    ex = PeerExchange(c.local, None, 64 << 20)
    q = gen_rows(torch, 0, nq, d, 4321, c.dev, "clustered")
    od = torch.empty((nq, k), dtype=torch.float32, device=c.dev)
    oi = torch.empty((nq, k), dtype=torch.int64, device=c.dev)
    for _ in range(3):
        g.search(q, k, nprobe=nprobe, out=(od, oi), exchange=ex)
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 10
    for _ in range(reps):
        g.search(q, k, nprobe=nprobe, out=(od, oi), exchange=ex)
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=c.dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = recall_at_k(c, g, q, k, oi[: args.recall_queries].clone())
    timed_out = ex.status()[0]
    ex.close()
    g.close()
    return {"workload": f"IVF_FLAT {n}x{d} fp32 ({n * d * 4 / 1e9:.0f} GB) row-sharded over {c.world} GPUs, nlist={nlist}, nprobe={nprobe}, nq={nq}, "
                        f"top-{k}, clustered set", "ms_per_step": float(ms.item()), "qps": nq / float(ms.item()) * 1e3, "recall_at_10": rec,
            "exchange_timed_out": timed_out, "build_s": build_s}


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    # stdout carries exactly one JSON line.  Libraries underneath write to fd 1 on their own (NCCL prints its version and, with
    # NCCL_DEBUG=INFO, its whole log there): keep the original stdout for the line and point fd 1 at stderr for the run, so
    # NCCL_DEBUG stays whatever the launcher chose and its output stays visible.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    run_ours(args)


if __name__ == "__main__":
    main()
