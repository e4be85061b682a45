#!/bin/bash
# A/B timing of kernel variants: every build/ab/libvrt_*.so is swapped in as zig_vulkan_b200/libvrt.so and benched (device-time
# line only), twice, alternating, on the same box.  Outputs -> gpurun_out/ab.log
mkdir -p gpurun_out; : > gpurun_out/ab.log
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so
for rep in 1 2 3; do
  for v in build/ab/libvrt_*.so; do
    cp "$v" zig_vulkan_b200/libvrt.so
    timeout -k 5 200 python bench.py --no-cpu-baseline --steps 300 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$v', 'rep$rep', 'ms %.4f' % d['ms_per_step'], 'Mrays/s %.0f' % d['value'])" | tee -a gpurun_out/ab.log
  done
done
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
