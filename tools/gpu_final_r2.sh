#!/bin/bash
# Round-2 evidence visit: launch list + `ncu --set full` of the step's kernels (from bench.py itself, eager launches:
# --cuda-graph 0 issues the same kernels one by one) and of the match / kNN kernels; then the bench lines copied into profiles/.
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --cuda-graph 0 --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_kernel|gaussian_backward|emit_instances|onesweep_pass|pixel_loss_forward" -s 150 -c 16 -o gpurun_out/prof_step_$TAG -f python bench.py --cuda-graph 0 --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*_$TAG.ncu-rep
(timeout 500 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
(timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err)
(timeout 400 python bench.py --config c5 --no-cpu-baseline > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err; echo "bench c5 rc=$?"; cut -c1-300 gpurun_out/bench_c5_$TAG.json; tail -3 gpurun_out/bench_c5_$TAG.err)
(timeout 300 python bench.py --config c2 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench c2 rc=$?"; cut -c1-300 gpurun_out/bench_c2_$TAG.json; tail -3 gpurun_out/bench_c2_$TAG.err)
