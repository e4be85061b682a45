#!/bin/bash
# Round-2 visit B (1 GPU): schedule-sort check, launch-bounds A/B on C3, TMA-staged masks A/B + ncu captures on C4
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "schedules or full_size or golden or mix" > gpurun_out/pytest_sched.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sched.log
timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_C3_quick.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_C3_quick.json')); print('C3', d['value'], d['step_ms'], d['config']['mode_candidates_ms'])"
VARIANTS="base t128b6 t128b7 t128b8" REPS=2 bash tools/gpu_ab.sh C3_bounds --schedule lpt
VARIANTS="base tma" REPS=2 bash tools/gpu_ab.sh C4_tma --workload C4 --schedule static
for name in base tma; do
  cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so; cp build/ab/libvrt_$name.so zig_vulkan_b200/libvrt.so
  ncu --set full --clock-control none --import-source on -k regex:trace_warp -s 10 -c 1 -f -o gpurun_out/prof_warp_C4_$name timeout -k 5 300 python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --schedule static > gpurun_out/ncu_C4_$name.log 2>&1
  cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
done
ls -la gpurun_out | tail -12
