#!/bin/bash
# present pass: parity tests, device time, one ncu capture of the main kernel -> gpurun_out/denoise_*
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_denoise.py tests/test_ref_shader.py -m gpu -x -q > gpurun_out/denoise_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/denoise_pytest.log
timeout -k 5 200 python tools/gpu_denoise_prof.py 2>&1 | tail -6 | tee gpurun_out/denoise_time.log
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:denoise_padded -s 2 -c 1 -f -o gpurun_out/denoise_padded python tools/gpu_denoise_prof.py > gpurun_out/denoise_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/denoise_padded.ncu-rep --page raw --csv > gpurun_out/denoise_padded_raw.csv 2>/dev/null
ncu -i gpurun_out/denoise_padded.ncu-rep --page details > gpurun_out/denoise_padded_details.txt 2>/dev/null
ls -la gpurun_out/denoise_padded*
