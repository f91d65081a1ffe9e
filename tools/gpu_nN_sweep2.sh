#!/bin/bash
# Sweep of the late launch's grid (fewer CTAs = fewer requests queued in the memory system beside the forward).
TAG=${1:-run}
N=${2:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for c in 12 8 5 3; do
  (WAST3D_PEER_LATE_CTAS=$c timeout 150 $TR tools/diag_overlap.py > gpurun_out/diag_n${N}_c${c}_$TAG.log 2>&1; echo "late_ctas=$c rc=$?"; grep "^step\|^mean" gpurun_out/diag_n${N}_c${c}_$TAG.log | tail -2)
done
