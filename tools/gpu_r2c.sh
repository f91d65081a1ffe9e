#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_$TAG.log)
(timeout 120 python tools/prof_match_knn.py 3 time > gpurun_out/match_knn_times_$TAG.log 2>&1; tail -4 gpurun_out/match_knn_times_$TAG.log)
(timeout 500 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
(timeout 400 python bench.py --config c5 --no-cpu-baseline > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err; echo "bench c5 rc=$?"; cat gpurun_out/bench_c5_$TAG.json; tail -3 gpurun_out/bench_c5_$TAG.err)
(timeout 300 python bench.py --config c2 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2_$TAG.json; tail -3 gpurun_out/bench_c2_$TAG.err)
