#!/bin/bash
# visit H (1 GPU): lateral-distance shortcut: parity tests on that build, A/B, tile statistics, partition sim
mkdir -p gpurun_out
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so
cp build/ab/libvrt_lat.so zig_vulkan_b200/libvrt.so
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_lat.log 2>&1; echo "pytest(lat) rc=$?" | tee -a gpurun_out/pytest_lat.log; tail -3 gpurun_out/pytest_lat.log
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
VARIANTS="base lat" REPS=2 bash tools/gpu_ab.sh C3_lateral --schedule lpt
VARIANTS="base lat" REPS=1 bash tools/gpu_ab.sh C2_lateral --schedule lpt --workload C2
VARIANTS="base lat" REPS=1 bash tools/gpu_ab.sh C5_lateral --schedule lpt --workload C5 --steps 60
cp build/ab/libvrt_latstats.so zig_vulkan_b200/libvrt.so
timeout -k 5 200 python tools/gpu_tilestats.py C3 2>&1 | tee gpurun_out/tilestats_C3_lat.log
cp build/ab/libvrt_lat.so zig_vulkan_b200/libvrt.so
echo "== partition sim lat"; timeout -k 5 300 python tools/gpu_part.py C3 2>&1 | tail -4
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
