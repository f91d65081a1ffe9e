#!/bin/bash
# Round 2 multi-GPU visit: peer checks (incl. the colour-record optimizer), bench default vs --sync records.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r2_multi.sh TAG N'
TAG=${1:-m}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 200 $TR tests/peer_check.py 300000 > gpurun_out/peer_check_n${N}_$TAG.log 2>&1; echo "peer_check rc=$?"; grep -v Warning gpurun_out/peer_check_n${N}_$TAG.log | tail -5)
(timeout 200 $TR tests/peer_records_check.py 100000 > gpurun_out/peer_records_check_n${N}_$TAG.log 2>&1; echo "peer_records_check rc=$?"; grep -v Warning gpurun_out/peer_records_check_n${N}_$TAG.log | tail -12)
(timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; cut -c1-6000 gpurun_out/bench_n${N}_$TAG.json; tail -3 gpurun_out/bench_n${N}_$TAG.err)
(timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --sync records --no-extra --no-cpu-baseline > gpurun_out/bench_n${N}_records_$TAG.json 2> gpurun_out/bench_n${N}_records_$TAG.err; echo "bench records rc=$?"; cut -c1-3000 gpurun_out/bench_n${N}_records_$TAG.json; tail -5 gpurun_out/bench_n${N}_records_$TAG.err)
