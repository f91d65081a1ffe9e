#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
(WAST3D_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log)
for v in "1 1" "0 0" "1 0" "0 1"; do
  set -- $v
  for cfg in c3 c2 c5; do
  echo "k6=$1 k7=$2: $(WAST3D_K6_MODE=$1 WAST3D_K7_MODE=$2 timeout 200 python tests/prof_step.py $cfg 10 stages 2>&1 | tail -1)"
  done
done 2>&1 | tee gpurun_out/k67_ab_$TAG.log
(timeout 300 python bench.py --no-cpu-baseline --no-extra --no-ref-cuda > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
