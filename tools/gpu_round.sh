#!/bin/bash
# One GPU visit: tests, smoke, benches, ncu launch list + full capture of the trace kernels.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout -k 5 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout -k 5 200 python bench.py > gpurun_out/bench_C3.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-150 gpurun_out/bench_C3.json
timeout -k 5 200 python bench.py --workload C2 --no-cpu-baseline > gpurun_out/bench_C2.json 2>> gpurun_out/bench.err
timeout -k 5 200 python bench.py --workload C4 --no-cpu-baseline --steps 50 > gpurun_out/bench_C4.json 2>> gpurun_out/bench.err
timeout -k 5 200 python bench.py --baseline-kernel --no-cpu-baseline > gpurun_out/bench_C3_baseline_kernel.json 2>> gpurun_out/bench.err
timeout -k 5 200 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_C3_reference.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C3.csv timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_warp -s 3 -c 1 -f -o gpurun_out/prof_warp_C3 timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_ref -s 3 -c 1 -f -o gpurun_out/prof_ref_C3 timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --baseline-kernel > gpurun_out/ncu_full_ref.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:denoise -s 3 -c 1 -f -o gpurun_out/prof_denoise_1080p timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_denoise.log 2>&1
grep -i "error\|traceback" gpurun_out/bench.err | head -5
ls gpurun_out | head -40
