#!/bin/bash
# All GPU tests (no -x: every failure is listed) + the match/kNN timings.
TAG=${1:-t}
mkdir -p gpurun_out
(timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_$TAG.log)
(timeout 120 python tools/prof_match_knn.py 3 time > gpurun_out/match_knn_times_$TAG.log 2>&1; tail -4 gpurun_out/match_knn_times_$TAG.log)
