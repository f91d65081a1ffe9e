#!/bin/bash
TAG=${1:-k7p}
mkdir -p gpurun_out
for rows in 0 2; do
WAST3D_K7_ROWS=$rows timeout 300 ncu --set full --clock-control none --import-source on -k regex:"render_backward_kernel" -s 2 -c 1 -o gpurun_out/prof_k7_rows${rows}_$TAG -f python tests/prof_step.py c3 3 > gpurun_out/ncu_k7_rows${rows}_$TAG.log 2>&1; echo "ncu rows=$rows rc=$?"
done
ls -la gpurun_out/*.ncu-rep
