#!/bin/bash
TAG=${1:-run}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { # name, env...
  name=$1; shift
  (env "$@" timeout 200 $TR tools/diag_overlap.py > gpurun_out/diag_${name}_$TAG.log 2>&1; echo "diag $name rc=$?"; grep "^step\|^mean" gpurun_out/diag_${name}_$TAG.log | tail -4)
}
run base OVERLAP=1
run noov OVERLAP=0
run prio OVERLAP=1 WAST3D_PEER_SIDE_PRIORITY=-1
run prio74 OVERLAP=1 WAST3D_PEER_SIDE_PRIORITY=-1 WAST3D_PEER_MAX_CTAS=74
run cap74 OVERLAP=1 WAST3D_PEER_MAX_CTAS=74
run prio37 OVERLAP=1 WAST3D_PEER_SIDE_PRIORITY=-1 WAST3D_PEER_MAX_CTAS=37
