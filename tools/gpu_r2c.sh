#!/bin/bash
# Round-2 visit C (1 GPU): new tests (accel patch, host-assembled frames, pinned uploads), occupancy / prefetch A/B on C3, partition sim of the best
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_accel_update.py tests/test_gpu_parity.py -m gpu -x -q -k "accel or pinned or host_assembled or partial_uploads or insert" > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_c.log
VARIANTS="base t128b8 t128b9 t128b10 t64b16 pf1 pf2 pf2k1 pf2k8 b8pf2 b8pf1" REPS=1 bash tools/gpu_ab.sh C3_occ_pf --schedule lpt
VARIANTS="base t128b8 b8pf2" REPS=1 bash tools/gpu_ab.sh C2_occ_pf --schedule lpt --workload C2
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so
for name in t128b8 b8pf2; do
  cp build/ab/libvrt_$name.so zig_vulkan_b200/libvrt.so
  echo "== partition sim $name"; timeout -k 5 300 python tools/gpu_part.py C3 2>&1 | tail -4
done
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
