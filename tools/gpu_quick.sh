#!/bin/bash
# quick parity + timing of both kernels (tools/gpu_check.py) -> gpurun_out/gpu_check.{log,json}
mkdir -p gpurun_out
timeout -k 5 300 python tools/gpu_check.py --full > gpurun_out/gpu_check.log 2>&1; echo rc=$?
grep -i "error\|Traceback" gpurun_out/gpu_check.log | head
python - <<PY
import json
d=json.load(open("gpurun_out/gpu_check.json"))
for k,v in d.items():
    if isinstance(v,dict) and "rgba_mismatch_px" in v:
        bad={a:b for a,b in v.items() if a!="counters" and a!="trace_ms" and b}
        ref=d.get(k.split("_")[0]+"_oracle_counters")
        if "counters" in v and ref and v["counters"]!=ref: bad["counters"]=v["counters"]
        print(k, "ms=%.3f"%v["trace_ms"], "BAD" if bad else "ok", bad)
    elif k.endswith("_ms"): print(k, "min %.4f mean %.4f"%(min(v), sum(v)/len(v)))
PY
