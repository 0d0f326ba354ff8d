"""Where the automatic query-major / list-major switch sits: step time of both routes around nq * nprobe = nlist / 8 .. nlist / 2
on the C2 index, iid and clustered sets.  One JSON line per (dataset, nq, nprobe)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import semcode_b200 as sb

sys.argv = [sys.argv[0]]
args = bench.parse_args()
c = bench.Ctx()
c.torch, c.dist, c.sb, c.args, c.world, c.rank, c.local = torch, None, sb, args, 1, 0, 0
c.dev = torch.device("cuda", 0)
for dataset in ("iid", "clustered"):
    g, _ = bench.build_index(c, args.n, args.dim, args.nlist, dataset, "IP")
    for nq, npb in ((64, 32), (128, 16), (256, 8), (128, 32), (256, 16), (512, 8), (32, 64)):
        q = bench.gen_rows(torch, 0, nq, args.dim, 4321, c.dev, dataset)
        out = {"dataset": dataset, "nq": nq, "nprobe": npb, "pairs_per_list": nq * npb / args.nlist}
        for mode, name in ((1, "query_major_ms"), (2, "list_major_ms"), (0, "auto_ms")):
            g.set_param("scan_mode", mode)
            ms, _ = bench.time_search(c, g, q, 10, npb, 10)
            out[name] = round(ms, 4)
        out["list_over_query"] = round(out["list_major_ms"] / out["query_major_ms"], 3)
        print(json.dumps(out), flush=True)
    g.close()
    del g
    torch.cuda.empty_cache()
