#!/bin/bash
# Multi-GPU visit (N > 2): peer optimizer with the NVSwitch multicast path and with plain peer loads/stores.
#   gpurun --gpus 4 --timeout 600 -- 'bash tools/gpu_nN.sh TAG 4'
TAG=${1:-run}
N=${2:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 200 $TR tests/peer_check.py 3000000 > gpurun_out/peer_check_n${N}_$TAG.log 2>&1; echo "peer_check rc=$?"; grep -v Warning gpurun_out/peer_check_n${N}_$TAG.log | tail -7)
(WAST3D_PEER_BACKEND=ipc timeout 200 $TR tests/peer_check.py 3000000 > gpurun_out/peer_check_n${N}_ipc_$TAG.log 2>&1; echo "peer_check ipc rc=$?"; grep -v Warning gpurun_out/peer_check_n${N}_ipc_$TAG.log | tail -7)
(timeout 300 $TR bench.py --gpus $N --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n${N}_$TAG.json; tail -3 gpurun_out/bench_n${N}_$TAG.err)
(WAST3D_PEER_BACKEND=ipc timeout 300 $TR bench.py --gpus $N --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_ipc_$TAG.json 2> gpurun_out/bench_n${N}_ipc_$TAG.err; echo "bench ipc rc=$?"; cat gpurun_out/bench_n${N}_ipc_$TAG.json)
