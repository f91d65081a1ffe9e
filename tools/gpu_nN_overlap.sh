#!/bin/bash
# Multi-GPU visit: peer optimizer with the late (side-stream) feature class, bench with and without overlap.
#   gpurun --gpus 4 --timeout 700 -- 'bash tools/gpu_nN_overlap.sh TAG 4'
TAG=${1:-run}
N=${2:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 200 $TR tests/peer_check.py 3000000 late > gpurun_out/peer_check_late_n${N}_$TAG.log 2>&1; echo "peer_check late rc=$?"; grep -v "Warning\|^\*\*\*\|OMP_NUM" gpurun_out/peer_check_late_n${N}_$TAG.log | tail -7)
for ov in 1 0; do
(timeout 300 $TR bench.py --gpus $N --overlap $ov --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_ov${ov}_$TAG.json 2> gpurun_out/bench_n${N}_ov${ov}_$TAG.err; echo "bench overlap=$ov rc=$?"; grep '^{' gpurun_out/bench_n${N}_ov${ov}_$TAG.json | cut -c1-300; tail -2 gpurun_out/bench_n${N}_ov${ov}_$TAG.err)
done
(WAST3D_PEER_BACKEND=ipc timeout 300 $TR bench.py --gpus $N --overlap 1 --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_ov1_ipc_$TAG.json 2> gpurun_out/bench_n${N}_ov1_ipc_$TAG.err; echo "bench overlap=1 ipc rc=$?"; grep '^{' gpurun_out/bench_n${N}_ov1_ipc_$TAG.json | cut -c1-300)
