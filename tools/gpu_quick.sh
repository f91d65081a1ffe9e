#!/bin/bash
# Short GPU-box visit: parity tests, the default bench, an A/B arm, the ncu launch list (+ optional full capture).
#   gpurun --timeout 1200 -- 'bash tools/gpu_quick.sh TAG [full]'
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
(timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log)
(timeout 400 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
(timeout 300 python bench.py --sync nccl --no-extra --no-cpu-baseline > gpurun_out/bench_dense_$TAG.json 2> gpurun_out/bench_dense_$TAG.err; echo "bench dense rc=$?"; cat gpurun_out/bench_dense_$TAG.json)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
if [ "$2" = "full" ]; then
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_kernel|gaussian_backward|emit_instances|onesweep_pass|radix_scatter|adam_kernel" -s 13 -c 14 -o gpurun_out/prof_$TAG -f python tests/prof_step.py c3 3 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
fi
