#!/bin/bash
# usage: measure_scale.sh N  — weak-scaling boids2d (1 M agents per GPU) and strong-scaling
# circle3d (16 M agents) on N GPUs of one box; JSON lines land in gpurun_out/.
set -u
N=$1
export ABL_CUDA_HALO_TIMEOUT_MS=30000
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline "$@"; }
run --steps 300 --warmup 20 > gpurun_out/scale_weak_$N.json 2> gpurun_out/scale_weak_$N.err
run --steps 30 --warmup 5 --workload circle3d-16M-f64 --strong > gpurun_out/scale_strong_circle3d_$N.json 2> gpurun_out/scale_strong_circle3d_$N.err
grep -il "error" gpurun_out/scale_*_$N.err
python - <<PY
import json
for f in ["gpurun_out/scale_weak_$N.json", "gpurun_out/scale_strong_circle3d_$N.json"]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.3f G/s" % (d["value"] / 1e9), "%.4f ms/step" % d["ms_per_step"], "e2e %.3f" % (d["e2e"]["value"] / 1e9), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
