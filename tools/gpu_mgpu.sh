#!/bin/bash
# Multi-GPU visit: tests/test_multi_gpu.py for every world size the box has, then the bench line at that N.
# usage: [SKIP_TESTS=1] tools/gpu_mgpu.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ -z "$SKIP_TESTS" ]; then
  timeout -k 5 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -rs > gpurun_out/pytest_mgpu_n$N.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_mgpu_n$N.log; tail -6 gpurun_out/pytest_mgpu_n$N.log
fi
timeout -k 5 600 python bench.py --gpus $N > gpurun_out/bench_C3_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_C3_n$N.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), d["step_ms"], "mode", d["config"]["exchange"], d["config"]["schedule"])
print("candidates", {k: round(v,4) for k,v in d["config"]["mode_candidates_ms"].items()})
print("kernel", d.get("per_rank_kernel_ms"), "exchange", d.get("exchange_ms"), "crc", d["frame_crc"])
print("e2e", {k: (round(v,4) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k!="how"})
c=d.get("c5")
if c: print("c5 value", round(c["value"]), "ms", round(c["ms_per_step"],4), "mode", c["exchange"], c["schedule"], "kernel", c["per_rank_kernel_ms"], "exch", c["exchange_ms"], "crc", c["frame_crc"], "e2e", round(c["e2e"]["value"]))
PY
