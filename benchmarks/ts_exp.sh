mkdir -p gpurun_out/ts
run() { # name, env...
  name=$1; shift
  env "$@" SEMCODE_TS_PROF=1 timeout -s KILL 200 python benchmarks/ts_tiles_bench.py iid 5 4096x128 > gpurun_out/ts/exp_$name.log 2>&1
  echo "== $name rc=$?"; tail -1 gpurun_out/ts/exp_$name.log | cut -c1-900
}
run cp_base SEMCODE_TS_NO_GATHER=1
run cp_noTMA SEMCODE_TS_NO_GATHER=1 SEMCODE_TS_ABLATE=1
run cp_noST SEMCODE_TS_NO_GATHER=1 SEMCODE_TS_ABLATE=2
run cp_noMMA SEMCODE_TS_NO_GATHER=1 SEMCODE_TS_ABLATE=8
run cp_noTMA_noST_noMMA SEMCODE_TS_NO_GATHER=1 SEMCODE_TS_ABLATE=11
run cp_noB SEMCODE_TS_NO_GATHER=1 SEMCODE_TS_ABLATE=4
run g4_noTMA SEMCODE_TS_ABLATE=1
