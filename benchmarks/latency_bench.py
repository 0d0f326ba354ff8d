"""Small-batch latency of one search call on the C2 index (10M x 768, nlist 16384): nq in {1, 4, 16, 128}, nprobe 32.
CUDA events around `reps` back-to-back calls (device-resident queries and outputs), plus the library's per-phase events.
With --once N it runs ONE search of nq = N after warm-up between cudaProfilerStart/Stop (for an ncu launch list)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import semcode_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--once", type=int, default=0)
ap.add_argument("--dataset", default="iid")
ap.add_argument("--nprobe", type=int, default=32)
a = ap.parse_args()
sys.argv = [sys.argv[0]]
args = bench.parse_args()
c = bench.Ctx(); c.torch = torch; c.dist = None; c.sb = sb; c.args = args; c.world = 1; c.rank = 0; c.local = 0
c.dev = torch.device("cuda", 0)
g, _ = bench.build_index(c, args.n, args.dim, args.nlist, a.dataset, "IP")
k = 10
for nq in ([a.once] if a.once else [1, 4, 16, 128]):
    q = bench.gen_rows(torch, 0, nq, args.dim, 4321, c.dev, a.dataset)
    od = torch.empty((nq, k), dtype=torch.float32, device=c.dev)
    oi = torch.empty((nq, k), dtype=torch.int64, device=c.dev)
    for _ in range(5):
        g.search(q, k, nprobe=a.nprobe, out=(od, oi))
    torch.cuda.synchronize()
    if a.once:
        torch.cuda.cudart().cudaProfilerStart()
        g.search(q, k, nprobe=a.nprobe, out=(od, oi))
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        continue
    reps = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.search(q, k, nprobe=a.nprobe, out=(od, oi))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    p = bench.profiled(c, g, q, k, a.nprobe, reps=5)
    print(json.dumps({"dataset": a.dataset, "nq": nq, "nprobe": a.nprobe, "us_per_call": round(us, 1),
                      "phases_us": {kk: round(p[kk] * 1e3, 1) for kk in ("coarse_ms", "select_ms", "plan_ms", "scan_ms", "topk_ms", "total_ms")},
                      "launches": p["total_launches"], "scan_GBps": round(p["scanned_rows"] * 4 * args.dim / max(p["scan_ms"], 1e-6) / 1e6)}), flush=True)
