#!/bin/bash
TAG=${1:-g}
mkdir -p gpurun_out
for cfg in c3 c2 c5; do
(timeout 500 python bench.py --config $cfg --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_${cfg}_$TAG.json 2> gpurun_out/bench_${cfg}_$TAG.err; echo "bench $cfg rc=$?"; tail -2 gpurun_out/bench_${cfg}_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${cfg}_$TAG.json"))
print("$cfg", round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["implementation"]["cuda_graph_replay"], d["gpu_launches"], d["host_issue_ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["extra"].get("cuda_graph_step"))
PY
)
done
(timeout 500 python bench.py --cuda-graph 0 --no-extra --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_c3e_$TAG.json 2> gpurun_out/bench_c3e_$TAG.err; echo "bench eager rc=$?"; cut -c1-250 gpurun_out/bench_c3e_$TAG.json)
