#!/bin/bash
# One GPU-box visit: parity tests, the bench (default + A/B arms), the ncu launch list of the bench
# command and one `--set full` capture of the rasteriser kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log)
(timeout 400 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
for arm in "--tile-cut 0" "--sync peer"; do
  name=$(echo $arm | tr -d ' -')
  (timeout 300 python bench.py $arm --no-extra --no-cpu-baseline > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err; echo "bench $arm rc=$?"; cat gpurun_out/bench_${name}_$TAG.json)
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_kernel|gaussian_backward|emit_instances|onesweep_pass|radix_scatter|adam_kernel" -s 13 -c 14 -o gpurun_out/prof_$TAG -f python tests/prof_step.py c3 3 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
