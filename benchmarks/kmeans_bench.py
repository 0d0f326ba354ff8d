#!/usr/bin/env python
"""k-means index build (BASELINE.json configs[3]: 50M x 768, nlist 65536, 20 Lloyd iterations, 1/2/4/8 GPUs).

    python benchmarks/kmeans_bench.py --rows-per-gpu 6250000 --iters 3            # one 8-GPU shard on 1 GPU
    torchrun --nproc-per-node 8 benchmarks/kmeans_bench.py --rows-per-gpu 6250000 --iters 20   # the full config

Per iteration: assignment = fused tcgen05 3xTF32 contraction + argmax (gemm_tc.cu), update = fp64 atomic
accumulation (kmeans.cu), all-reduce of sums/counts/objective across ranks, identical centroid update.
Reports seconds / iteration (CUDA events, max over ranks), fp32-equivalent TFLOP/s = 2*N*nlist*d / t and
the MMA-pipe rate (3 kind::tf32 MMAs per logical product).  One JSON line on stdout.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (gen_rows, measured_peaks)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--rows-per-gpu", type=int, default=6_250_000)
    p.add_argument("--total-rows", type=int, default=0, help="rows of the whole job (C4: 50000000): rows per GPU = total / world")
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--nlist", type=int, default=65536)
    p.add_argument("--iters", type=int, default=3)
    p.add_argument("--metric", default="IP")
    p.add_argument("--dataset", default="clustered")
    p.add_argument("--coarse-impl", type=int, default=0)
    a = p.parse_args()
    import torch
    import torch.distributed as dist

    import semcode_b200 as sb
    from semcode_b200.sharded import ShardedIVFFlat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries the one JSON line: whatever libraries print on fd 1 (NCCL's banner / log) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, nlist = (a.total_rows // world if a.total_rows else a.rows_per_gpu), a.dim, a.nlist
    x = bench.gen_rows(torch, rank * n, (rank + 1) * n, d, 1234, dev, a.dataset)
    if world > 1:
        sh = ShardedIVFFlat(d, nlist, a.metric, device=local)
        eng = sh.local
    else:
        eng = sb.IVFFlatIndex(d, nlist=nlist, metric=a.metric, device=local)
    if a.coarse_impl:
        eng.set_param("coarse_impl", a.coarse_impl)
    # initial centroids: seeded random rows of rank 0's shard
    from semcode_b200.index import kmeans_init_rows

    if rank == 0:
        init = x[torch.from_numpy(kmeans_init_rows(n, nlist, 1234)).to(dev)]
    else:
        init = None
    if world > 1:
        sh.set_centroids(init, src=0)
    else:
        eng.set_centroids(init)
    sums, counts, obj = eng.kmeans_buffers()
    times, objs, step_ms, red_ms = [], [], [], []
    for it in range(a.iters):
        sums.zero_(); counts.zero_(); obj.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        eng.kmeans_step(x, sums, counts, obj)
        e[1].record()
        if world > 1:
            dist.all_reduce(sums); dist.all_reduce(counts); dist.all_reduce(obj)
        e[2].record()
        eng.kmeans_update(sums, counts)
        e[3].record()
        torch.cuda.synchronize()
        t = torch.tensor([e[0].elapsed_time(e[3]), e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t[0])); step_ms.append(float(t[1])); red_ms.append(float(t[2]))
        objs.append(float(obj.item()))
    if rank == 0:
        peaks, src = bench.measured_peaks()
        it_s = min(times[1:] or times) / 1e3
        flop = 2.0 * n * world * nlist * d
        line = {
            "metric": "k-means seconds per Lloyd iteration", "value": it_s, "unit": "s/iteration", "n_gpus": world,
            "config": {"workload": f"k-means {n * world} x {d} fp32, nlist={nlist}, {a.dataset} set, metric={a.metric} "
                                   f"(BASELINE.json configs[3] is 50M rows x 20 iterations)",
                       "rows_per_gpu": n, "nlist": nlist, "dim": d, "iters_timed": a.iters},
            "ms_per_iteration_all": times, "assign_accumulate_ms": step_ms, "allreduce_ms": red_ms,
            "fp32_equiv_tflops_per_gpu": flop / world / it_s / 1e12,
            "mma_tf32_tflops_per_gpu": 3 * flop / world / it_s / 1e12,
            "tensor_peak_stated": {"bf16_dense_measured_tflops": peaks.get("bf16_tflops"), "tf32_stated_as_bf16_over_2": peaks.get("bf16_tflops", 0) / 2,
                                   "source": src},
            "mma_frac_of_tf32_stated": 3 * flop / world / it_s / 1e12 / (peaks.get("bf16_tflops", 1) / 2),
            "objective": objs,
            "total_s_all_iterations": sum(times) / 1e3,
            "extrapolated_20_iterations_s": 20 * it_s,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
