#!/bin/bash
# Side-stream overlap A/B (SH colour under the sort chain; culled-Gaussian Adam under the tile backward).
TAG=${1:-o}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_raster_gpu.py tests/test_graphed_gpu.py tests/test_peer_gpu.py -m gpu -q -x > gpurun_out/pytest_ov_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ov_$TAG.log)
run() { echo "$1: $(env $1 timeout 300 python bench.py --config $2 --no-extra --no-cpu-baseline --no-ref-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['implementation']['cuda_graph_replay'], {k:v for k,v in d['stages_ms'].items()})
")"; }
{
for cfg in c3 c2 c5; do
  run "WAST3D_COLOUR_OVERLAP=1" $cfg
  run "WAST3D_COLOUR_OVERLAP=0" $cfg
done
} 2>&1 | tee gpurun_out/overlap_ab_$TAG.log
