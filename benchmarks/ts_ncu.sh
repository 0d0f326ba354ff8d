# DRAM traffic + duration of the tile kernels on the C2 index (one ncu pass, a handful of metrics)
mkdir -p gpurun_out/ts
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_active.avg,smsp__cycles_active.avg -k regex:"scan_lists_t" --clock-control none -c 6 --csv --log-file gpurun_out/ts/ncu_traffic.csv python benchmarks/ts_tiles_bench.py iid ${1:-5} 4096x128 > gpurun_out/ts/ncu_traffic.log 2>&1
python - <<'PY'
import csv
rows = list(csv.DictReader(l for l in open('gpurun_out/ts/ncu_traffic.csv') if not l.startswith('==')))
by = {}
for r in rows:
    by.setdefault((r['ID'], r['Kernel Name'][:40]), {})[r['Metric Name']] = r['Metric Value'] + ' ' + r['Metric Unit']
for k, v in by.items():
    print(k, v)
PY
