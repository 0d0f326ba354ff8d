"""Where the end-to-end step (host buffers through the C ABI) spends its time beyond the device-resident step."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import semcode_b200 as sb

sys.argv = [sys.argv[0]]
args = bench.parse_args()
c = bench.Ctx()
c.torch, c.dist, c.sb, c.args, c.world, c.rank, c.local = torch, None, sb, args, 1, 0, 0
c.dev = torch.device("cuda", 0)
g, _ = bench.build_index(c, args.n, args.dim, args.nlist, "iid", "IP")
nq, k, npb = 1024, 10, 32
qs = [bench.gen_rows(torch, i * nq, (i + 1) * nq, args.dim, 4321, c.dev, "iid") for i in range(4)]
qh = [t.cpu().pin_memory() for t in qs]
qp = [t.cpu().numpy().copy() for t in qs]  # pageable
hd, hi = torch.empty((nq, k), dtype=torch.float32).pin_memory(), torch.empty((nq, k), dtype=torch.int64).pin_memory()
od, oi = torch.empty((nq, k), dtype=torch.float32, device=c.dev), torch.empty((nq, k), dtype=torch.int64, device=c.dev)


def wall(fn, reps=30):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


out = {}
out["device_resident_async_ms"] = wall(lambda i: g.search(qs[i % 4], k, nprobe=npb, out=(od, oi)))
out["device_resident_sync_each_ms"] = wall(lambda i: (g.search(qs[i % 4], k, nprobe=npb, out=(od, oi)), torch.cuda.synchronize()))
out["c_abi_pinned_host_ms"] = wall(lambda i: g.search(qh[i % 4], k, nprobe=npb, out=(hd, hi)))
out["c_abi_pageable_host_ms"] = wall(lambda i: g.search(qp[i % 4], k, nprobe=npb))


def manual(i):
    qd = qh[i % 4].to(c.dev, non_blocking=True)
    g.search(qd, k, nprobe=npb, out=(od, oi))
    hd.copy_(od, non_blocking=True)
    hi.copy_(oi, non_blocking=True)
    torch.cuda.current_stream().synchronize()


out["manual_copies_sync_each_ms"] = wall(manual)
out["h2d_3MB_pinned_ms"] = wall(lambda i: (qs[0].copy_(qh[i % 4], non_blocking=True), torch.cuda.synchronize()))
print(json.dumps({kk: round(v, 4) for kk, v in out.items()}))
