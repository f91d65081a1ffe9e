#!/bin/bash
# Single-GPU visit: all GPU tests, the default bench, the single-GPU peer arm with and without overlapped features.
TAG=${1:-run}
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log)
(timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
for ov in 1 0; do
(timeout 300 python bench.py --sync peer --overlap $ov --no-cpu-baseline --no-extra > gpurun_out/bench_n1_peer${ov}_$TAG.json 2> gpurun_out/bench_n1_peer${ov}_$TAG.err; echo "bench peer overlap=$ov rc=$?"; cut -c1-330 gpurun_out/bench_n1_peer${ov}_$TAG.json; tail -3 gpurun_out/bench_n1_peer${ov}_$TAG.err)
done
