#!/bin/bash
# Round-2 visit E (1 GPU): what makes tiles expensive; rounds-with-slack A/B (+ partition sim of the best two)
mkdir -p gpurun_out
timeout -k 5 200 python tools/gpu_tilecost.py C3 2>&1 | tee gpurun_out/tilecost_C3.log
VARIANTS="base ks2 ks4 ks8 ks16 ks126" REPS=1 bash tools/gpu_ab.sh C3_kslack --schedule lpt
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so
for name in ks4 ks16; do
  cp build/ab/libvrt_$name.so zig_vulkan_b200/libvrt.so
  echo "== partition sim $name"; timeout -k 5 300 python tools/gpu_part.py C3 2>&1 | tail -4
done
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
