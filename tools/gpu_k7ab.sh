#!/bin/bash
# K7 reduction A/B: raster tests, then stage times of C3 / C2 / C5 for each WAST3D_K7_ROWS.
TAG=${1:-k7}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_raster_gpu.py tests/test_peer_gpu.py tests/test_staged_sh_records_gpu.py -m gpu -q -x > gpurun_out/pytest_raster_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_raster_$TAG.log)
for rows in 0 2 3 4; do
  for cfg in c3 c2 c5; do
    echo "rows=$rows: $(WAST3D_K7_ROWS=$rows timeout 200 python tests/prof_step.py $cfg 10 stages 2>&1 | tail -1)"
  done
done 2>&1 | tee gpurun_out/k7_ab_$TAG.log
