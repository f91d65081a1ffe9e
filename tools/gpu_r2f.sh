#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_normals_gpu.py -m gpu -q -x 2>&1 | tail -40)
(WAST3D_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_$TAG.log)
