mkdir -p gpurun_out
for u in 1 2 4; do WAST3D_ADAM_UNROLL=$u python tools/prof_adam_bwd.py c3 10; done 2>&1 | tee gpurun_out/adam_unroll.log
python tools/prof_adam_bwd.py c3 10 dense 2>&1 | tee -a gpurun_out/adam_unroll.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gaussian_backward" -s 5 -c 1 -o gpurun_out/prof_gbadam -f python tools/prof_adam_bwd.py c3 3 > gpurun_out/ncu_gbadam.log 2>&1; echo "ncu rc=$?"
