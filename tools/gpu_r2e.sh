#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
(WAST3D_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_$TAG.log)
(timeout 300 python bench.py --no-cpu-baseline --no-extra --no-ref-cuda > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cut -c1-3500 gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
