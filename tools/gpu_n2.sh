#!/bin/bash
# Two-GPU visit: peer-memory optimizer against NCCL all-reduce + dense Adam, then the bench at N=2.
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_n2.sh TAG'
TAG=${1:-run}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 200 $TR tests/peer_check.py 3000000 > gpurun_out/peer_check_$TAG.log 2>&1; echo "peer_check rc=$?"; grep -v Warning gpurun_out/peer_check_$TAG.log | tail -8)
(timeout 300 $TR bench.py --gpus $N --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n${N}_$TAG.json; tail -3 gpurun_out/bench_n${N}_$TAG.err)
(timeout 300 $TR bench.py --gpus $N --sync nccl --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_nccl_$TAG.json 2> gpurun_out/bench_n${N}_nccl_$TAG.err; echo "bench nccl rc=$?"; cat gpurun_out/bench_n${N}_nccl_$TAG.json)
