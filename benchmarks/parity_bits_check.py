"""How many distance VALUES differ between the list-major and query-major routes at full size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import semcode_b200 as sb

sys.argv = [sys.argv[0]]
args = bench.parse_args()
c = bench.Ctx(); c.torch = torch; c.dist = None; c.sb = sb; c.args = args; c.world = 1; c.rank = 0; c.local = 0
c.dev = torch.device("cuda", 0)
g, _ = bench.build_index(c, args.n, args.dim, args.nlist, "iid", "IP")
for nq, npb in ((1024, 32), (4096, 128)):
    q = bench.gen_rows(torch, 0, nq, args.dim, 4321, c.dev, "iid")
    for cfg in (0, 1, 5):
        g.set_param("lists_cfg", cfg); g.set_param("scan_mode", 0)
        g.set_profiling(True)
        d0, i0 = g.search(q, 10, nprobe=npb); torch.cuda.synchronize()
        t = g.last_search_times(); g.set_profiling(False)
        g.set_param("scan_mode", 1)
        d1, i1 = g.search(q, 10, nprobe=npb); torch.cuda.synchronize()
        a, b = d0.cpu().numpy(), d1.cpu().numpy()
        print(f"nq={nq} nprobe={npb} cfg={cfg}: scan_ms={t.scan_ms:.3f} launches={t.scan_launches} unique_rows={t.unique_rows} values differing={int((a != b).sum())}/{a.size} "
              f"max|diff|={np.abs(a - b).max():.3e} ids differing={int((i0 != i1).sum().item())}  sample {a[0, :3]} {b[0, :3]}", flush=True)
