#!/bin/bash
# Round 2, second session: same-box A/B of the explicit PDL trigger and of the record prefetch in k_bin_rank_move.
set -u
out=gpurun_out
mkdir -p $out
run() { # tag workload steps env...
  local tag=$1 w=$2 st=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $w --steps $st --warmup 5 --no-companion --no-cpu-baseline > $out/r2e_${w}_$tag.json 2> $out/r2e_${w}_$tag.err
  python - $out/r2e_${w}_$tag.json "$w $tag" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print(sys.argv[2], "ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4), "kernel_ms", round(r["kernel_ms"], 4),
              "bin_ms", round(r["bin_ms"], 4), "whole", round(r["whole_step_frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 2))
PY
}
for rep in 1 2; do
run t0p0 boids2d-1M-f64 100 ABL_CUDA_PDL_TRIGGER=0 ABL_CUDA_BIN_PREFETCH=0
run t1p0 boids2d-1M-f64 100 ABL_CUDA_PDL_TRIGGER=1 ABL_CUDA_BIN_PREFETCH=0
run t0p1 boids2d-1M-f64 100 ABL_CUDA_PDL_TRIGGER=0 ABL_CUDA_BIN_PREFETCH=1
run t1p1 boids2d-1M-f64 100 ABL_CUDA_PDL_TRIGGER=1 ABL_CUDA_BIN_PREFETCH=1
done
run t0p0 boids2d-16M-f64 20 ABL_CUDA_PDL_TRIGGER=0 ABL_CUDA_BIN_PREFETCH=0
run t1p1 boids2d-16M-f64 20 ABL_CUDA_PDL_TRIGGER=1 ABL_CUDA_BIN_PREFETCH=1
run t0p0 circle3d-1M-f64 10 ABL_CUDA_PDL_TRIGGER=0 ABL_CUDA_BIN_PREFETCH=0
run t1p1 circle3d-1M-f64 10 ABL_CUDA_PDL_TRIGGER=1 ABL_CUDA_BIN_PREFETCH=1
run t0p0 game_of_life-16M-f64 20 ABL_CUDA_PDL_TRIGGER=0 ABL_CUDA_BIN_PREFETCH=0
run t1p1 game_of_life-16M-f64 20 ABL_CUDA_PDL_TRIGGER=1 ABL_CUDA_BIN_PREFETCH=1
