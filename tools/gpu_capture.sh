#!/bin/bash
# Profile visit: ncu launch list of the bench command + one `--set full` capture of the per-step kernels
# taken from bench.py itself (the product path: raw parameters, Adam inside the per-Gaussian backward).
#   gpurun --timeout 900 -- 'bash tools/gpu_capture.sh TAG'
TAG=${1:-run}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_kernel|gaussian_backward|emit_instances|onesweep_pass|pixel_loss" -s 130 -c 13 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
