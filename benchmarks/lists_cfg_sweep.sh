mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "list_major" > gpurun_out/pytest6.log 2>&1; tail -5 gpurun_out/pytest6.log | cut -c1-300
python - <<'PY' > gpurun_out/cfg_sweep.log 2>&1
import sys, time, json
sys.path.insert(0, '.')
import torch, bench
import semcode_b200 as sb
dev = torch.device('cuda', 0)
n, d, nlist, k = 10_000_000, 768, 16384, 10
g = sb.IVFFlatIndex(d, nlist=nlist, metric='IP')
tr = bench.gen_rows(torch, 0, 1_000_000, d, 1234, dev)
g.train(tr, niter=4, max_points_per_centroid=0); del tr
for s in range(0, n, 1 << 20):
    e = min(n, s + (1 << 20))
    g.add(bench.gen_rows(torch, s, e, d, 1234, dev), torch.arange(s, e, device=dev, dtype=torch.int64))
q = bench.gen_rows(torch, 0, 4096, d, 4321, dev)
g.set_profiling(True)
ref = {}
for nq, nprobe in ((4096, 128), (4096, 32), (4096, 8), (1024, 32), (4096, 16)):
    for mode, cfg in ((1, 0), (2, 0)):
        if mode == 1 and nq * nprobe > 200000: 
            reps = 1
        else:
            reps = 3
        g.set_param('scan_mode', mode); g.set_param('lists_cfg', cfg)
        best = None
        for _ in range(reps + 1):
            dd, ii = g.search(q[:nq], k, nprobe=nprobe)
            torch.cuda.synchronize()
            t = g.last_search_times()
            best = t if best is None or t.scan_ms < best.scan_ms else best
        key = (nq, nprobe)
        if mode == 1: ref[key] = ii.clone()
        same = bool((ii == ref[key]).all()) if key in ref else None
        print(json.dumps({'nq': nq, 'nprobe': nprobe, 'mode': mode, 'cfg': cfg, 'scan_ms': round(best.scan_ms, 3), 'plan_ms': round(best.plan_ms, 3),
              'topk_ms': round(best.topk_ms, 3), 'coarse_ms': round(best.coarse_ms, 3), 'select_ms': round(best.probe_select_ms, 3), 'total_ms': round(best.total_ms, 3),
              'qps': round(nq / best.total_ms * 1e3), 'ids_equal_query_major': same}), flush=True)
PY
cat gpurun_out/cfg_sweep.log | tail -30
