#!/bin/bash
# A/B of the L2 bulk prefetch hints (K1, K8+K9 with and without the Adam epilogue), parity tests, bench.
TAG=${1:-run}
mkdir -p gpurun_out
for mode in backward dense; do for pf in 0 1 0 1; do WAST3D_L2_PREFETCH=$pf python tools/prof_adam_bwd.py c3 16 $mode 2>&1 | grep -o "preprocess=[0-9.]*\|gaussian_backward=[0-9.]*" | tr '\n' ' ' | sed "s/^/$mode prefetch=$pf /"; echo; done; done | tee gpurun_out/prefetch_ab_$TAG.log
(timeout 400 python -m pytest tests/test_raster_gpu.py -m gpu -q 2>&1 | tail -3)
(timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1_$TAG.json; tail -3 gpurun_out/bench_n1_$TAG.err)
