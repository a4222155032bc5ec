#!/bin/bash
# ncu full capture of the experimental mode-7 kernel
out=gpurun_out; mkdir -p $out
export ABL_MODEL_DIR=build/exp/boids_m7 ABL_EXP_MODE7=1 ABL_EXP_BS=${1:-128} ABL_CUDA_FLAT=1
ncu --set full --clock-control none --import-source on -k regex:abl_kernel_update_boid -s 20 -c 1 -f \
  -o $out/prof_exp_m7 python tools/quick_step.py boids2d-1M-f64 --steps 20 > $out/exp_ncu_m7.log 2>&1
tail -3 $out/exp_ncu_m7.log
