#!/bin/bash
# Full GPU suite + smoke() + the default bench line (and c5 / c2 without the CPU baseline).
TAG=${1:-chk}
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log)
(timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n1_$TAG.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["stages_ms"], d["roofline"]["frac"], d["roofline"]["step_frac"])
PY
if [ "$2" = "more" ]; then
(timeout 400 python bench.py --config c5 --no-cpu-baseline > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err; echo "bench c5 rc=$?"; cut -c1-200 gpurun_out/bench_c5_$TAG.json)
(timeout 300 python bench.py --config c2 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"; cut -c1-200 gpurun_out/bench_c2_$TAG.json)
fi
