#!/bin/bash
# Round-2 visit D (1 GPU): full GPU test suite on the new default (128x8, restructured step loop), A/B of ticket prefetch / 2D-blocked
# distance planes / general-path occupancy, partition sim for the blocked layout
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
VARIANTS="base oldloop256 tk blk2d blk2dtk" REPS=2 bash tools/gpu_ab.sh C3_loop_tk_blk --schedule lpt
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so
for name in base gen7 gen8; do
  cp build/ab/libvrt_$name.so zig_vulkan_b200/libvrt.so
  python - <<PY
import sys; sys.path.insert(0, '.')
import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
R = scenes.REF_DEFAULT
g = scenes.build_ref_default_grid(); mats = zv.terrain_materials()
cam = scenes.camera(R["width"], R["height"], spp=R["spp"], max_bounce=R["max_bounce"], origin=(0.0, -5.0, 14.0), euler_deg=(25.0, 0.0, 0.0))
sun = scenes.sun(True, R["sun_radius"])
ctx = ffi.Context(R["width"], R["height"], len(g.brick_indices)); ctx.upload_grid(g, mats); ctx.set_schedule(ffi.VRT_SCHED_LPT, 8)
ms = []
for i in range(40):
    ctx.trace(cam, sun); ms.append(ctx.last_trace_kernel_ms())
print("$name REF general path ms: min %.4f median %.4f" % (min(ms[10:]), sorted(ms[10:])[15]))
PY
done
for name in base blk2d; do
  cp build/ab/libvrt_$name.so zig_vulkan_b200/libvrt.so
  echo "== partition sim $name"; timeout -k 5 300 python tools/gpu_part.py C3 2>&1 | tail -4
done
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
