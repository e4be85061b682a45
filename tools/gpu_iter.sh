#!/bin/bash
# Fast iteration visit: parity tests + the C3 bench line (no CPU baseline).  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -k 5 200 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_iter.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_iter.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "clk", d["clocks"])
PY
