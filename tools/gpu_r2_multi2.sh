#!/bin/bash
# N-GPU visit after the peer-kernel change: correctness checks under torchrun, then the bench with a late-grid sweep.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_multi2.sh TAG N "12 20"'
TAG=${1:-m}
N=${2:-2}
SWEEP=${3:-""}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 200 $TR tests/peer_check.py 300000 > gpurun_out/peer_check_n${N}_$TAG.log 2>&1; echo "peer_check rc=$?"; grep -v Warning gpurun_out/peer_check_n${N}_$TAG.log | tail -5)
show() { python - "$1" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["implementation"], "replicas", d["replicas_bit_identical"], d["stages_ms"])
PY
}
(timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; show gpurun_out/bench_n${N}_$TAG.json; tail -2 gpurun_out/bench_n${N}_$TAG.err)
for c in $SWEEP; do
(WAST3D_PEER_LATE_CTAS=$c timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_late${c}_$TAG.json 2> gpurun_out/bench_n${N}_late${c}_$TAG.err; echo "bench late$c rc=$?"; show gpurun_out/bench_n${N}_late${c}_$TAG.json)
done
