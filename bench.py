#!/usr/bin/env python
"""bench.py -- IVF_FLAT search throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU IVF_FLAT (oracle port) on host cores

Workload (config.workload): BASELINE.json configs[1] -- 10M x 768 fp32 IVF_FLAT, nlist 16384, IP,
top-10, synthetic unit-norm Gaussian embeddings (seed 1234 DB / 4321 queries).  One *step* = one
batch of `nq` queries through coarse quantizer -> nprobe list scan -> top-k.  With N GPUs the index
is sharded (--shard-by rows, the default: every list's rows are dealt round-robin; --shard-by lists:
list l lives on rank l % N -- measured 337.6k vs 364.2k QPS at N=2), centroids are replicated, the coarse pass is split over the ranks, each rank
scans its shard and the partial top-k are all-gathered and merged on the device.  The union of
the shards is exactly the single index, so results equal the 1-GPU results; the total database
is fixed, so scaling is "strong".

Reported numbers
  value / ms_per_step  queries per second with the query batch already in HBM (CUDA events, max
                       over ranks)
  e2e                  same batch through the C ABI with HOST buffers (pinned): H2D of the queries
                       and D2H of (dist, ids) inside the timed region
  roofline             the list-scan kernel: algorithmic bytes (sum over probed lists of
                       len * 4 * dim) / its CUDA-event time, against the
                       measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline         oracle/ (C restatement of FAISS IndexIVFFlat, OpenMP over all host cores) on
                       a bounded sample of the same queries against the same lists (N=1, rank 0)
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
    p.add_argument("--sweep", action="store_true", help="also time an nprobe x nq grid (extra key 'sweep')")
    p.add_argument("--sweep-nq", default="1,16,256,4096", help="batch sizes of --sweep")
    p.add_argument("--recall-queries", type=int, default=128)
    p.add_argument("--cpu-queries", type=int, default=96, help="queries in the CPU baseline sample")
    p.add_argument("--cpu-reps", type=int, default=120, help="timed repetitions of the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--scan-variant", type=int, default=0)
    p.add_argument("--dataset", default="iid", choices=["iid", "clustered"],
                   help="synthetic set of SURVEY.md 8d: A (iid Gaussian) or B (Zipf-clustered)")
    p.add_argument("--coarse-impl", type=int, default=0, help="0 = tcgen05 3xTF32, 1 = fp32 SIMT")
    p.add_argument("--scan-mode", type=int, default=0, help="0 = auto, 1 = query-major, 2 = list-major")
    p.add_argument("--lists-cfg", type=int, default=0, help="tile configuration of the list-major kernel")
    p.add_argument("--shard-sim", type=int, default=1,
                   help="single-GPU run over ONE rank's shard of a G-way row-sharded index (rows i with i %% G == 0 of the "
                        "same stream, same centroids): what each rank of --gpus G scans, without the exchange")
    p.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                   help="N > 1: p2p = probe rows and partial top-k stored straight into the peers' buffers by the kernels "
                        "that produce them (NVLink peer memory), merge kernel waits on flags; nccl = all-gathers + merge")
    p.add_argument("--shard-by", default="rows", choices=["rows", "lists"],
                   help="N > 1: deal every list's rows round-robin (rows) or whole lists (list l on rank l %% N)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic data (set A): unit-norm Gaussian rows, generated on the device
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
    return torch.nn.functional.normalize(out, dim=1)


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


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_search_sample(ivf_c, q, centroids, metric, nprobe, k, probes_gpu, export_list, reps=1):
    """Time coarse + scan + top-k of `q` on the CPU against the lists those queries probe.
    Only the probed lists are copied to the host (compact CSR, list ids remapped)."""
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
    times = []
    out = None
    for _ in range(reps + 1):  # first repetition = warm-up (page faults of the exported lists)
        t0 = time.perf_counter()
        scores = ivf_c.coarse_scores(q, centroids, metric)
        probes = ivf_c.top_probes(scores, nprobe)
        t1 = time.perf_counter()
        local = remap[probes]
        t2 = time.perf_counter()
        out = ivf_c.scan_search(q, metric, local, off, vecs, ids, k)
        t3 = time.perf_counter()
        times.append((t1 - t0) + (t3 - t2))
    return statistics.median(times[1:]), out, probes


def run_reference(args):
    """--impl reference: the reference path's CPU implementation.  The reference's engine (Milvus /
    knowhere / FAISS) cannot be installed offline, so this is the oracle port (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import ivf_c

    ivf_c.build()
    cores = ivf_c.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1; rank 0 runs alone and takes every core it may use
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
    scores = ivf_c.coarse_scores(q, cent, metric)
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
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        s = ivf_c.coarse_scores(q, cent, metric)
        p = ivf_c.top_probes(s, args.nprobe)
        ivf_c.scan_search(q, metric, remap[p], off, vecs, ids, args.k)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    qps = nq * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, nq_override=nq),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{nq} queries/step x nprobe {args.nprobe} over lists of {per_list} rows "
                                   f"({n}/{nlist}); oracle/ivf_oracle.c, OpenMP over queries"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, nq_override=None):
    return {
        "workload": f"IVF_FLAT {args.n}x{args.dim} fp32, nlist={args.nlist}, nprobe={args.nprobe}, "
                    f"nq={nq_override or args.nq}/step, top-{args.k}, metric={args.metric}, {args.dataset} synthetic set "
                    f"(BASELINE.json configs[1])",
        "n": args.n, "dim": args.dim, "nlist": args.nlist, "nprobe": args.nprobe, "nq": nq_override or args.nq,
        "k": args.k, "metric": args.metric, "dataset": args.dataset, "shard_by": args.shard_by if args.gpus > 1 else None,
        "shard_sim": args.shard_sim if args.shard_sim > 1 else None,
        "l2_policy": "inputs larger than L2: every step streams nq*nprobe lists (>> 126 MB) and rotates query batches",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: semcode_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    import semcode_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG=VERSION/INFO prints to stdout; stdout must carry exactly one JSON line
        if not os.environ.get("SEMCODE_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner would otherwise land on stdout
        dist.init_process_group("nccl", device_id=dev)
    n, d, nlist, k, nprobe, nq = args.n, args.dim, args.nlist, args.k, args.nprobe, args.nq
    t_build0 = time.time()

    # ---- build: train on a prefix (rank 0), broadcast centroids, add this rank's rows --------------
    g = sb.IVFFlatIndex(d, nlist=nlist, metric=args.metric, device=local)
    if args.scan_variant:
        g.set_param("scan_variant", args.scan_variant)
    if args.coarse_impl:
        g.set_param("coarse_impl", args.coarse_impl)
    if args.scan_mode:
        g.set_param("scan_mode", args.scan_mode)
    if args.lists_cfg:
        g.set_param("lists_cfg", args.lists_cfg)
    cent = torch.empty((nlist, d), dtype=torch.float32, device=dev)
    if rank == 0:
        tr = gen_rows(torch, 0, min(args.train_rows, n), d, 1234, dev, args.dataset)
        g.train(tr, niter=args.train_iters, max_points_per_centroid=0)
        cent.copy_(torch.from_numpy(g.get_centroids()))
        del tr
    if world > 1:
        dist.broadcast(cent, 0)
        if rank != 0:
            g.set_centroids(cent)
    t_train = time.time() - t_build0
    chunk = 1 << 20
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = gen_rows(torch, s, e, d, 1234, dev, args.dataset)
        ids = torch.arange(s, e, device=dev, dtype=torch.int64)
        if world > 1 and args.shard_by == "rows":  # deal rows round-robin: rank r keeps global rows i with i % world == r
            x, ids = x[rank::world].contiguous(), ids[rank::world].contiguous()
            g.add(x, ids)
        elif world > 1:  # whole lists: rank r keeps the rows whose list l has l % world == r
            lists = g.assign(x)
            keep = torch.nonzero(lists % world == rank).squeeze(1)
            g.add(x[keep].contiguous(), ids[keep].contiguous(), lists=lists[keep].contiguous())
            del lists, keep
        elif args.shard_sim > 1:
            g.add(x[0::args.shard_sim].contiguous(), ids[0::args.shard_sim].contiguous())
        else:
            g.add(x, ids)
        del x, ids
    torch.cuda.synchronize()
    t_build = time.time() - t_build0

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

    ex, ex_err = None, None
    if world > 1 and args.exchange != "nccl" and args.shard_by == "rows":
        ok = torch.zeros(1, dtype=torch.int32, device=dev)
        try:
            from semcode_b200.index import PeerExchange

            ex = PeerExchange(local, None, 64 << 20)
            ok += 1
        except Exception as e:
            ex_err = f"{type(e).__name__}: {e}"
        dist.all_reduce(ok)
        if int(ok.item()) != world:
            ex = None
            if args.exchange == "p2p":
                raise SystemExit(f"--exchange p2p: peer-memory exchange unavailable ({ex_err})")

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

    def step_device(i):
        if world > 1:
            return search_sharded(qb[i % nb])
        g.search(qb[i % nb], k, nprobe=nprobe, out=(out_d, out_i))
        return out_d, out_i

    def step_e2e(i):
        if world == 1:
            g.search(qhost[i % nb], k, nprobe=nprobe, out=(host_d, host_i))  # C ABI, host buffers
            return host_d, host_i
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

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        w1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), w0, w1

    # under `ncu --profile-from-start off` only the search steps are captured, not the index build
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(args.warmup):
        step_device(i)
        step_e2e(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # device-resident timing (value): the plain product call, no per-phase events inside the timed region
    barrier()
    scan_ms, scanned_rows, launches, phase = [], [], 0, {"coarse": 0.0, "select": 0.0, "plan": 0.0, "scan": 0.0, "topk": 0.0}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    barrier()
    w1 = time.time()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    # per-phase CUDA events (recorded by the library on the launching stream) for the roofline of the scan kernel:
    # a second pass over the same inputs, outside the headline timing
    g.set_profiling(True)
    step_device(0)
    torch.cuda.synchronize()
    t = g.last_search_times()
    # NCCL route: + coarse split (gemm, select, split) and merge launched from here; the fused route counts its own
    launches_per_step = t.total_launches + (4 if world > 1 and ex is None else 0)
    unique_rows = []
    for i in range(min(args.steps, 8)):
        step_device(i)
        torch.cuda.synchronize()
        t = g.last_search_times()
        scan_ms.append(t.scan_ms)
        scanned_rows.append(t.scanned_rows)
        unique_rows.append(t.unique_rows)
        phase["coarse"] += t.coarse_ms
        phase["select"] += t.probe_select_ms
        phase["plan"] += t.plan_ms
        phase["scan"] += t.scan_ms
        phase["topk"] += t.topk_ms
    nprof = len(scan_ms)
    list_major = statistics.mean(unique_rows) > 0
    # the query-major kernel on the same workload (the regime the HBM-fraction claim of SURVEY.md 8d is made in:
    # it streams every probed list once per (query, list) pair, so logical bytes == DRAM bytes)
    qm = None
    if list_major and not args.scan_mode:
        g.set_param("scan_mode", 1)
        qms, qrows, qtot = [], [], []
        for i in range(min(args.steps, 4) + 1):
            step_device(i)
            torch.cuda.synchronize()
            t = g.last_search_times()
            if i:
                qms.append(t.scan_ms)
                qrows.append(t.scanned_rows)
                qtot.append(t.total_ms)
        g.set_param("scan_mode", 0)
        qm = (statistics.mean(qms), statistics.mean(qrows), statistics.mean(qtot))
    g.set_profiling(False)

    ms_e2e, _, _ = timed(step_e2e, args.steps)
    torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop(w0, w1) if rank == 0 else None

    # ---- recall@10 against exact search (outside the timed region) ---------------------------------
    rq = min(args.recall_queries, nq)
    d_ann, i_ann = step_device(0)
    i_ann = i_ann[:rq].clone()
    gd, gi = g.search(qb[0][:rq], k, nprobe=nlist)  # exhaustive probe of this rank's shard; merged below = exact
    if world > 1:
        gd_all = torch.empty((world, rq, k), dtype=torch.float32, device=dev)
        gi_all = torch.empty((world, rq, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gd_all, gd.contiguous())
        dist.all_gather_into_tensor(gi_all, gi.contiguous())
        gd, gi = sb.merge_topk(gd_all, gi_all, k, args.metric, local)
    torch.cuda.synchronize()
    ia, ie = i_ann.cpu().numpy(), gi.cpu().numpy()
    recall = float(np.mean([len(np.intersect1d(ia[r][ia[r] >= 0], ie[r])) / k for r in range(rq)]))

    if rank != 0:
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
    logical_bytes = statistics.mean(scanned_rows) * 4 * d  # this rank's slice: one pass per (query, list) pair
    scan_s = statistics.mean(scan_ms) / 1e3
    if list_major:  # compulsory bytes: every DISTINCT probed list once
        bytes_per_step = statistics.mean(unique_rows) * 4 * d
        tiles = ("scan_lists_tc_kernel (tcgen05 tiles of 64 queries)" if args.metric == "IP" and d % 32 == 0 and args.lists_cfg not in (1, 2)
                 else "scan_lists_kernel (FFMA tiles of 32 queries)")
        kernel_name = ("list-major scan: scan_mq_kernel<4> (remainders of 1..4 queries per list) + scan_mq_kernel<8> (5..16) + "
                       f"{tiles} + count / plan / fill")
    else:
        bytes_per_step = logical_bytes
        kernel_name = "scan_pages_kernel (query-major)"
    achieved = bytes_per_step / scan_s / 1e9
    qps = nq * args.steps / (ms_total / 1e3)
    e2e_qps = nq * args.steps / (ms_e2e / 1e3)
    line = {
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "recall_at_10": recall,
        "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {
            "bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "frac_of_8TBps_nominal": achieved / 8000.0, "peak_source": peak_src,
            "traffic": traffic if not list_major else traffic_lm, "algorithmic_bytes_per_launch": bytes_per_step,
            "logical_bytes_per_launch": logical_bytes, "logical_GBps": logical_bytes / scan_s / 1e9,
            "kernel_ms": scan_s * 1e3, "kernel_share_of_step": (phase["scan"] / nprof) / (ms_total / args.steps),
        },
        "roofline_query_major": None if qm is None else {
            "bound": "hbm", "kernel": "scan_pages_kernel (query-major, forced with scan_mode=1 on the same workload)",
            "achieved": qm[1] * 4 * d / (qm[0] / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": qm[1] * 4 * d / (qm[0] / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": traffic,
            "algorithmic_bytes_per_launch": qm[1] * 4 * d, "kernel_ms": qm[0], "qps_query_major": nq / qm[2] * 1e3},
        "phases_ms": {kname: v / nprof for kname, v in phase.items()},
        "clocks": clocks,
        "build_s": {"train": t_train, "total": t_build},
    }
    if world > 1:
        line["config"]["exchange"] = "p2p-fused (peer-memory stores + flags, no collective call)" if ex is not None else \
            f"nccl all-gather + merge ({ex_err or 'requested'})"
        if ex is not None:
            line["exchange_timed_out"] = ex.status()[0]

    # ---- optional nprobe x nq sweep --------------------------------------------------------------------
    if args.sweep and world == 1:
        sweep = []
        for nq_ in [int(v) for v in args.sweep_nq.split(",")]:
            qs = gen_rows(torch, 0, nq_, d, 777 + nq_, dev, args.dataset)
            exact_ids = None
            if nq_ == 256:  # recall@10 per nprobe on this batch (exact = exhaustive probe)
                _, ei = g.search(qs, k, nprobe=nlist)
                exact_ids = ei.cpu().numpy()
            for np_ in (8, 16, 32, 64, 128):
                for _ in range(2):
                    g.search(qs, k, nprobe=np_)
                torch.cuda.synchronize()
                reps = 20 if nq_ <= 256 else 3
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    _, ii = g.search(qs, k, nprobe=np_)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                g.set_profiling(True)
                g.search(qs, k, nprobe=np_)
                torch.cuda.synchronize()
                tt = g.last_search_times()
                g.set_profiling(False)
                row = {"nprobe": np_, "nq": nq_, "ms": ms, "qps": nq_ / ms * 1e3,
                       "logical_GBps": tt.scanned_rows * 4 * d / ms / 1e6,
                       "scan_GBps": tt.scanned_rows * 4 * d / max(tt.scan_ms, 1e-6) / 1e6}
                if exact_ids is not None:
                    ia_ = ii.cpu().numpy()
                    row["recall_at_10"] = float(np.mean([len(np.intersect1d(ia_[r][ia_[r] >= 0], exact_ids[r])) / k
                                                         for r in range(nq_)]))
                sweep.append(row)
        line["sweep"] = sweep

    # ---- CPU baseline on the same lists (bounded sample) ----------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ivf_c

        ivf_c.build()
        ivf_c.use_all_cores()
        cq = min(args.cpu_queries, nq)
        qs = qb[0][:cq].cpu().numpy()
        probes = g.probe(qs, nprobe)
        dt, (cd, ci), cprobes = cpu_search_sample(ivf_c, qs, g.get_centroids(), g.metric, nprobe, k, probes, g.export_list,
                                                  reps=args.cpu_reps)
        gdd, gii = g.search(qs, k, nprobe=nprobe)
        same = float(np.mean(np.all(gii == ci, axis=1)))
        line["cpu_baseline"] = {
            "value": cq / dt, "unit": UNIT, "cores": ivf_c.num_threads(), "kind": "port",
            "sample": f"{cq} of the {nq} step-0 queries, same centroids / lists / nprobe; median of {args.cpu_reps} repetitions "
                      f"after one warm-up ({args.cpu_reps * dt:.1f} s of CPU work); "
                      f"oracle/ivf_oracle.c (OpenMP over queries); ids identical to GPU for {same:.3f} of queries",
            "seconds": dt,
        }
    elif world == 1:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
