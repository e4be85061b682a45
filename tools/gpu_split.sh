#!/bin/bash
# tile split (vrt_set_tile_split): schedule parity test, single-GPU timing with / without, partition simulation -> gpurun_out/split_*
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_schedules or explicit or golden or differential" > gpurun_out/split_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/split_pytest.log
for sc in lpt lpt+30 lpt+150; do
  timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras --steps 300 --schedule $sc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['step_ms']; print('$sc', 'mean %.4f median %.4f min %.4f' % (s['mean'], s['median'], s['min']), 'Mrays/s %.0f' % d['value'], 'crc', d['frame_crc']['value'])" | tee -a gpurun_out/split_time.log
done
timeout -k 5 400 python tools/gpu_part.py C3 2>&1 | grep world | tee gpurun_out/split_part_C3.txt
