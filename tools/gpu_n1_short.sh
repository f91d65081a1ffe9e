#!/bin/bash
# Shortest single-GPU visit: rasteriser parity tests + the default bench.
#   gpurun --timeout 600 -- 'bash tools/gpu_n1_short.sh TAG'
TAG=${1:-run}
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_raster_gpu.py tests/test_python_api.py -m gpu -q > gpurun_out/pytest_raster_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_raster_$TAG.log)
(timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
