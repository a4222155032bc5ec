#!/bin/bash
# Device timeline (ABL_CUDA_TRACE=1: globaltimer stamps between the stages) of boids2d 1 M per GPU on 1 and on 2 GPUs of one box.
set -u
out=gpurun_out
mkdir -p $out
ABL_CUDA_TRACE=1 timeout 300 python bench.py --steps 100 --warmup 10 --no-companion --no-cpu-baseline > $out/r2h_trace_n1.json 2> $out/r2h_trace_n1.err
grep "abl_cuda" $out/r2h_trace_n1.err | tail -n 6
ABL_CUDA_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 100 --warmup 10 --no-companion --no-cpu-baseline > $out/r2h_trace_n2.json 2> $out/r2h_trace_n2.err
grep "abl_cuda" $out/r2h_trace_n2.err | tail -n 12
