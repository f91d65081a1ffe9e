#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_$TAG.log)
for v in "256 4" "256 5" "256 6" "512 4" "512 5" "128 4" "128 5"; do
  set -- $v
  echo "batch=$1 minb=$2: $(WAST3D_K7_BATCH=$1 WAST3D_K7_MINB=$2 timeout 200 python tests/prof_step.py c3 10 stages 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/k7_ab_$TAG.log
(timeout 300 python bench.py --no-cpu-baseline --no-extra --no-ref-cuda --async-forward 0 > gpurun_out/bench_sync_$TAG.json 2> gpurun_out/bench_sync_$TAG.err; echo "bench sync rc=$?"; cut -c1-2500 gpurun_out/bench_sync_$TAG.json; tail -3 gpurun_out/bench_sync_$TAG.err)
(timeout 300 python bench.py --no-cpu-baseline --no-extra --no-ref-cuda --async-forward 1 > gpurun_out/bench_async_$TAG.json 2> gpurun_out/bench_async_$TAG.err; echo "bench async rc=$?"; cut -c1-2500 gpurun_out/bench_async_$TAG.json; tail -3 gpurun_out/bench_async_$TAG.err)
