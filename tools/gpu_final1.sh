#!/bin/bash
# Final 1-GPU visit of the round: tests, smoke, bench lines (C3 default, C2, C4, baseline kernel, reference arm), ncu launch list + full
# captures (trace kernel on C3, present pass), tile statistics and partition simulation of the final kernel.
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout -k 5 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
ncu --set full --clock-control none --import-source on -k regex:trace_warp -s 20 -c 1 -f -o gpurun_out/prof_warp_C3 timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --schedule lpt > gpurun_out/ncu_full.log 2>&1
python - <<'PY'
# the bench line's issue roofline and `traffic` come from this capture of the kernel as built now
import json, subprocess
d = json.loads(subprocess.run(["python", "tools/ncu_summary.py", "gpurun_out/prof_warp_C3.ncu-rep"], capture_output=True, text=True).stdout)
t = json.load(open("profiles/traffic.json"))
t["C3"]["dram_bytes"] = int(d["dram_read"] + d["dram_write"]); t["C3"]["warp_instructions"] = int(d["warp_instructions"])
json.dump(t, open("profiles/traffic.json", "w"), indent=1); json.dump(d, open("gpurun_out/prof_warp_C3_summary.json", "w"), indent=1)
print("traffic.json <-", t["C3"]["warp_instructions"], "warp instructions,", t["C3"]["dram_bytes"], "DRAM bytes")
PY
cp profiles/traffic.json gpurun_out/traffic.json
timeout -k 5 400 python bench.py > gpurun_out/bench_C3.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_C3.json; tail -3 gpurun_out/bench.err
timeout -k 5 200 python bench.py --workload C2 --no-cpu-baseline --no-extras > gpurun_out/bench_C2.json 2>> gpurun_out/bench.err
timeout -k 5 300 python bench.py --workload C4 --no-cpu-baseline --no-extras --steps 60 > gpurun_out/bench_C4.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C3.csv timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so; cp build/ab/libvrt_stats.so zig_vulkan_b200/libvrt.so
timeout -k 5 200 python tools/gpu_tilestats.py C3 > gpurun_out/tilestats_C3.log 2>&1; head -4 gpurun_out/tilestats_C3.log
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
timeout -k 5 300 python tools/gpu_part.py C3 > gpurun_out/part_C3.log 2>&1; tail -4 gpurun_out/part_C3.log
grep -i "error\|traceback" gpurun_out/bench.err | head -5
ls gpurun_out | head -50
