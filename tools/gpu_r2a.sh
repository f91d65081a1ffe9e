#!/bin/bash
# Round 2, visit A: full GPU test suite, the default bench (+ reference arm), match/kNN timings and ncu captures.
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log)
(timeout 120 python tools/prof_match_knn.py 3 time > gpurun_out/match_knn_times_$TAG.log 2>&1; cat gpurun_out/match_knn_times_$TAG.log | tail -5)
(timeout 500 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
(timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"match_kernel|knn_" -c 12 -o gpurun_out/prof_match_$TAG -f python tools/prof_match_knn.py 1 > gpurun_out/ncu_match_$TAG.log 2>&1; echo "ncu match rc=$?"
ls -la gpurun_out/*.ncu-rep
