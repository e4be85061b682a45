#!/bin/bash
# A/B timing of kernel variants: every build/ab/libvrt_<name>.so named in $VARIANTS (default: all) is swapped in as
# zig_vulkan_b200/libvrt.so and benched (device-time line only), $REPS times, alternating, on the same box.
# usage: VARIANTS="base t128b7" REPS=2 tools/gpu_ab.sh <tag> [bench.py args]     -> gpurun_out/ab_<tag>.log
tag=$1; shift
mkdir -p gpurun_out; : > gpurun_out/ab_$tag.log
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so
for rep in $(seq 1 ${REPS:-2}); do
  for name in ${VARIANTS:-$(ls build/ab/libvrt_*.so | sed 's/.*libvrt_//; s/\.so//')}; do
    cp build/ab/libvrt_$name.so zig_vulkan_b200/libvrt.so
    timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras --steps 300 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['step_ms']; print('$name', 'rep$rep', 'mean %.4f median %.4f min %.4f' % (s['mean'], s['median'], s['min']), 'Mrays/s %.0f' % d['value'], d['config']['schedule'], 'crc', d['frame_crc']['value'])" | tee -a gpurun_out/ab_$tag.log
  done
done
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
