#!/bin/bash
# N-GPU scaling arms: default (peer, overlapped features), colour records, late-grid sweep.
TAG=${1:-m8}
N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; grep '^{' gpurun_out/bench_n${N}_$TAG.json | cut -c1-700; tail -2 gpurun_out/bench_n${N}_$TAG.err)
(timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --sync records --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_records_$TAG.json 2> gpurun_out/bench_n${N}_records_$TAG.err; echo "bench records rc=$?"; grep '^{' gpurun_out/bench_n${N}_records_$TAG.json | cut -c1-700; tail -2 gpurun_out/bench_n${N}_records_$TAG.err)
(WAST3D_PEER_LATE_CTAS=12 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_late12_$TAG.json 2> gpurun_out/bench_n${N}_late12_$TAG.err; echo "bench late12 rc=$?"; grep '^{' gpurun_out/bench_n${N}_late12_$TAG.json | cut -c1-700)
