#!/bin/bash
# Round 2, after the bulk-tile kernels (ABL_MODE 7) and the RED histogram became the default.
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/r2b_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2b_pytest_gpu.log
for w in boids2d-1M-f64 boids2d-1M-f32 game_of_life-16M-f64 boids2d-16M-f64; do
  for bulk in 1 0; do
    ABL_CUDA_BULK=$bulk timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 100 --warmup 10 \
      > $out/r2b_${w}_bulk$bulk.json 2> $out/r2b_${w}_bulk$bulk.err
  done
done
for bulk in 1 0; do ABL_CUDA_BULK=$bulk timeout 300 python tools/quick_step.py predator_prey-4M-f64 --steps 50 > $out/r2b_pp4M_bulk$bulk.txt 2>&1; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline", {})
            print(f.split("/")[-1], "%.3f G/s" % (d["value"] / 1e9), "ms", round(d.get("ms_per_step"), 4),
                  "steady", round(d.get("steady_state", {}).get("ms_per_step"), 4), "kernel_ms", round(r.get("kernel_ms"), 4),
                  "bin_ms", round(r.get("bin_ms"), 4), "whole", round(r.get("whole_step_frac"), 3), d["config"].get("candidate_loop_in_use"))
PY
tail -n 1 $out/r2b_pp4M_*.txt
tail -n 8 $out/r2b_pytest_gpu.log
