#!/bin/bash
TAG=${1:-r2j}
mkdir -p gpurun_out
(WAST3D_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log)
for pf in 0 1; do
(timeout 300 python bench.py --no-cpu-baseline --no-extra --no-ref-cuda --prefetch-projection $pf > gpurun_out/bench_pf${pf}_$TAG.json 2> gpurun_out/bench_pf${pf}_$TAG.err; echo "bench prefetch=$pf rc=$?"; python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/bench_pf${pf}_$TAG.json') if l.startswith('{')][-1])
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['stages_ms'], j['roofline']['frac'], j['roofline']['kernel_ms'])
PY
tail -3 gpurun_out/bench_pf${pf}_$TAG.err)
done
