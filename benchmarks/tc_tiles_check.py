"""Accuracy / parity probe for the tcgen05 list-major tiles (lists_cfg 0: list rows from tensor memory;
3 / 5: operands from shared memory, v1 / v2) against the
query-major fp32 scan on the same probes.  Prints the error statistics the default choice is based on."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import semcode_b200 as sb  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    for d, n, nlist, nq, nprobe in ((768, 40000, 16, 700, 4), (128, 20000, 8, 300, 3), (3072, 6000, 4, 200, 2), (1024, 30000, 32, 2000, 8)):
        x = rng.standard_normal((n, d)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        q = rng.standard_normal((nq, d)).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        g = sb.IVFFlatIndex(d, nlist=nlist, metric="IP")
        g.train(x, niter=3)
        g.add(x, np.arange(n, dtype=np.int64))
        g.remove_ids(np.arange(0, n, 17, dtype=np.int64))
        qd = torch.from_numpy(q).cuda()
        g.set_param("scan_mode", 1)
        d0, i0 = g.search(qd, 10, nprobe=nprobe)
        torch.cuda.synchronize()
        for cfg in (0, 3, 5):
            g.set_param("scan_mode", 2)
            g.set_param("lists_cfg", cfg)
            d1, i1 = g.search(qd, 10, nprobe=nprobe)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                g.search(qd, 10, nprobe=nprobe)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            same = float((i1 == i0).all(dim=1).float().mean())
            err = (d1 - d0).abs()
            rel = (err / d0.abs().clamp_min(1e-30)).max().item()
            print(f"d={d} nq={nq} nprobe={nprobe} cfg={cfg}: ids identical in {same:.4f} of queries, max |err| {err.max().item():.3e}, "
                  f"max rel {rel:.3e}, {ms:.3f} ms, score range [{d0.min().item():.3f}, {d0.max().item():.3f}]", flush=True)


if __name__ == "__main__":
    main()
