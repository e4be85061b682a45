#!/bin/bash
# visit G (1 GPU): per-cell record (mask + material start in one load) and park-time prefetch A/B; GPU tests on the new default
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
VARIANTS="prev base pp" REPS=2 bash tools/gpu_ab.sh C3_cellrec --schedule lpt
VARIANTS="prev base pp" REPS=1 bash tools/gpu_ab.sh C2_cellrec --schedule lpt --workload C2
