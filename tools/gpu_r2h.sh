#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
(WAST3D_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log)
for v in 1 0; do
  for cfg in c3 c2; do
  echo "k1vec=$v: $(WAST3D_K1_VEC=$v timeout 200 python tests/prof_step.py $cfg 10 stages 2>&1 | tail -1 | cut -c1-110)"
  done
done 2>&1 | tee gpurun_out/k1_ab_$TAG.log
(timeout 300 python bench.py --no-cpu-baseline --no-extra --no-ref-cuda > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
