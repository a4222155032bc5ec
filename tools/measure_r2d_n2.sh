#!/bin/bash
# Round 2, second session: two B200, weak scaling of boids2d (1 M agents per GPU), halo exchange forked next to the
# step kernel (default) against in-stream (ABL_CUDA_HALO_ASYNC=0).   gpurun --gpus 2 -- 'bash tools/measure_r2d_n2.sh'
set -u
out=gpurun_out
mkdir -p $out
N=${1:-2}
for a in ${ASYNC_LIST:-1 0 1 0}; do
  ABL_CUDA_HALO_ASYNC=$a timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 10 --no-companion --no-cpu-baseline > $out/r2d_n${N}_async$a.json 2> $out/r2d_n${N}_async$a.err
  python - $out/r2d_n${N}_async$a.json $a <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("async", sys.argv[2], "N", d["n_gpus"], "G/s", round(d["value"] / 1e9, 2), "ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4),
              "kernel_ms", round(r["kernel_ms"], 4), "bin_ms", round(r["bin_ms"], 4), "commit", round(r["commit_ms"], 4), "e2e", round(d["e2e"]["value"] / 1e9, 2))
PY
done
tail -n 3 $out/r2d_n${N}_async1.err
