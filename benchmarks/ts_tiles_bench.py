"""List-major tile kernels on the C2 index (10M x 768, nlist 16384): shared-memory operands (lists_cfg 5) against list
rows from tensor memory (lists_cfg 0, the default) at large batches.  Prints one JSON line per (nq, nprobe, cfg)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import semcode_b200 as sb  # noqa: E402


def main():
    dataset = sys.argv[1] if len(sys.argv) > 1 else "iid"
    cfgs = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "5,0").split(",")]
    dev = torch.device("cuda", 0)
    n, d, nlist, k = 10_000_000, 768, 16384, 10
    g = sb.IVFFlatIndex(d, nlist=nlist, metric="IP")
    tr = bench.gen_rows(torch, 0, 1_000_000, d, 1234, dev, dataset)
    g.train(tr, niter=4, max_points_per_centroid=0)
    del tr
    for s in range(0, n, 1 << 20):
        e = min(n, s + (1 << 20))
        g.add(bench.gen_rows(torch, s, e, d, 1234, dev, dataset), torch.arange(s, e, device=dev, dtype=torch.int64))
    q = bench.gen_rows(torch, 0, 4096, d, 4321, dev, dataset)
    g.set_profiling(True)
    shapes = ((4096, 128), (4096, 32), (1024, 32))
    if len(sys.argv) > 3:
        shapes = tuple(tuple(int(v) for v in sh.split("x")) for sh in sys.argv[3].split(","))
    for nq, nprobe in shapes:
        ref = None
        for cfg in cfgs:
            g.set_param("scan_mode", 2)
            g.set_param("lists_cfg", cfg)
            best = None
            for _ in range(4):
                dd, ii = g.search(q[:nq], k, nprobe=nprobe)
                torch.cuda.synchronize()
                t = g.last_search_times()
                best = t if best is None or t.scan_ms < best.scan_ms else best
            if ref is None:
                ref = (dd.clone(), ii.clone())
            same = float((ii == ref[1]).all(dim=1).float().mean())
            err = float((dd - ref[0]).abs().max())
            prof = None
            if cfg == 0 and os.environ.get("SEMCODE_TS_PROF"):
                import ctypes as C

                from semcode_b200 import _capi

                buf = (C.c_uint64 * 16)()
                _capi.lib().scdbg_ts_prof(buf)  # reset
                g.search(q[:nq], k, nprobe=nprobe)
                _capi.lib().scdbg_ts_prof(buf)
                names = ["qprod.qfree", "qprod.staged", "qprod.total", "-", "conv.rfull", "conv.sfree", "conv.total", "iss.accempty",
                         "iss.aready", "iss.bready", "iss.total", "stager.bfree", "stager.total", "epi.accfull", "epi.total", "kernel"]
                prof = {n: round(buf[i] / 148 / 1e3, 1) for i, n in enumerate(names)}  # kcycles per CTA
            print(json.dumps({"dataset": dataset, "prof_kcyc_per_cta": prof, "nq": nq, "nprobe": nprobe, "cfg": cfg, "scan_ms": round(best.scan_ms, 3),
                              "topk_ms": round(best.topk_ms, 3), "coarse_ms": round(best.coarse_ms, 3),
                              "total_ms": round(best.total_ms, 3), "qps": round(nq / best.total_ms * 1e3),
                              "unique_GB": round(best.unique_rows * 4 * d / 1e9, 2),
                              "scan_TBps": round(best.unique_rows * 4 * d / best.scan_ms / 1e9, 3),
                              "ids_equal_first_cfg": same, "max_abs_diff_first_cfg": err}), flush=True)


if __name__ == "__main__":
    main()
