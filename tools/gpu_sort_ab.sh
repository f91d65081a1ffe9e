#!/bin/bash
# Sort-chain A/B: correctness of the sort / scan primitives and the rasteriser, then stage times of C3 / C2 / C5 for the
# fused offsets scan, the tile-partition flavour and the look-back window; optional ncu capture.
TAG=${1:-s}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_sort_scan_gpu.py tests/test_raster_gpu.py -m gpu -q -x > gpurun_out/pytest_sort_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_sort_$TAG.log)
run() { echo "$1: $(env $1 timeout 200 python tests/prof_step.py $2 10 stages 2>&1 | tail -1)"; }
{
for cfg in c3 c2 c5; do
  run "WAST3D_SORT_TILE=1" $cfg
  run "WAST3D_SORT_TILE=1 WAST3D_EMIT_SCAN=0" $cfg
done
run "WAST3D_SORT_TILE=0" c3
run "WAST3D_SORT_TILE=1 WAST3D_LB_WINDOW=4" c3
run "WAST3D_SORT_TILE=1 WAST3D_LB_WINDOW=16" c3
} 2>&1 | tee gpurun_out/sort_ab_$TAG.log
if [ "$2" = "ncu" ]; then
  WAST3D_SORT_TILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass|emit_instances" -s 7 -c 7 \
    -o gpurun_out/prof_sort_$TAG -f python tests/prof_step.py c3 2 > gpurun_out/prof_sort_$TAG.log 2>&1
  tail -3 gpurun_out/prof_sort_$TAG.log
fi
