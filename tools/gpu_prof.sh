#!/bin/bash
# ncu captures of the trace kernel on the bench workload.  Outputs -> gpurun_out/
mkdir -p gpurun_out
W=${1:-C3}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$W.csv timeout -k 5 200 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_warp -s 3 -c 1 -f -o gpurun_out/prof_warp_$W timeout -k 5 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
