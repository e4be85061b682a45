#!/bin/bash
# Round-2 visit A (1 GPU): parity tests, smoke, the default bench line, partition simulation, ncu launch list + full capture.
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout -k 5 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout -k 5 400 python bench.py > gpurun_out/bench_C3.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_C3.json; tail -5 gpurun_out/bench.err
timeout -k 5 300 python tools/gpu_part.py C3 > gpurun_out/part_C3.log 2>&1; cat gpurun_out/part_C3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C3.csv timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_warp -s 30 -c 1 -f -o gpurun_out/prof_warp_C3 timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --schedule lpt > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -15
