#!/bin/bash
# visit I (1 GPU): blocked distance layout: parity, partition sim linear vs blocked, N=1 timing
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py tests/test_accel_update.py -m gpu -x -q -k "blocked or accel or golden or sweep" > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_i.log
echo "== linear"; timeout -k 5 300 python tools/gpu_part.py C3 2>&1 | tail -4
echo "== blocked"; timeout -k 5 300 python tools/gpu_part.py C3 --blocked 2>&1 | tail -4
echo "== C5 linear"; timeout -k 5 300 python tools/gpu_part.py C5 2>&1 | tail -4
echo "== C5 blocked"; timeout -k 5 300 python tools/gpu_part.py C5 --blocked 2>&1 | tail -4
