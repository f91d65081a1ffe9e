#!/bin/bash
# Round-end style visit: all GPU tests, smoke, the default bench, the launch list and one full ncu capture of bench.py.
TAG=${1:-run}
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log)
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1)
(timeout 400 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_kernel|gaussian_backward" -s 40 -c 4 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
