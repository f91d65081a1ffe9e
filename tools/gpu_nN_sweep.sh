#!/bin/bash
# Sweep of the late launch's grid cap with the diagnostic loop, then the bench with the default.
TAG=${1:-run}
N=${2:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift
  (env "$@" timeout 200 $TR tools/diag_overlap.py > gpurun_out/diag_n${N}_${name}_$TAG.log 2>&1; echo "diag $name rc=$?"; grep "^step\|^mean" gpurun_out/diag_n${N}_${name}_$TAG.log | tail -2)
}
run c37 WAST3D_PEER_LATE_CTAS=37
run c74 WAST3D_PEER_LATE_CTAS=74
run c20 WAST3D_PEER_LATE_CTAS=20
run ipc37 WAST3D_PEER_LATE_CTAS=37 WAST3D_PEER_BACKEND=ipc
run ipc20 WAST3D_PEER_LATE_CTAS=20 WAST3D_PEER_BACKEND=ipc
(timeout 300 $TR bench.py --gpus $N --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; grep '^{' gpurun_out/bench_n${N}_$TAG.json | cut -c1-300; tail -2 gpurun_out/bench_n${N}_$TAG.err)
