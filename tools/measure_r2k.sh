#!/bin/bash
# Same-box A/B: bulk-tile kernels waiting for their tile with / without a suspend-time hint (ABL_CUDA_MBAR_HINT).
set -u
out=gpurun_out
mkdir -p $out
run() { # tag workload steps env...
  local tag=$1 w=$2 st=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $w --steps $st --warmup 5 --no-companion --no-cpu-baseline > $out/r2k_${w}_$tag.json 2> $out/r2k_${w}_$tag.err
  python - $out/r2k_${w}_$tag.json "$w $tag" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print(sys.argv[2], "ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4), "kernel_ms", round(r["kernel_ms"], 4),
              "stage", round(r["kernel_ms_stage_events"], 4), "bin_ms", round(r["bin_ms"], 4), "whole", round(r["whole_step_frac"], 3))
PY
}
for rep in 1 2; do
run h0 boids2d-1M-f64 100 ABL_CUDA_MBAR_HINT=0
run h1 boids2d-1M-f64 100 ABL_CUDA_MBAR_HINT=1
done
run h0 boids2d-16M-f64 20 ABL_CUDA_MBAR_HINT=0
run h1 boids2d-16M-f64 20 ABL_CUDA_MBAR_HINT=1
