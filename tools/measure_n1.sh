#!/bin/bash
# Single-GPU measurement suite behind the tables of DESIGN.md §5 (run on the GPU box through gpurun;
# JSON lines land in gpurun_out/ and are copied to profiles/scaling/ afterwards).
set -u
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $out/clocks_n1.csv &
smi=$!
python bench.py > $out/bench_default_n1.json 2> $out/bench_default_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_reference_n1.json 2> $out/bench_reference_n1.err
for w in boids2d-1M-f32 boids2d-4M-f64; do
  python bench.py --workload $w --no-cpu-baseline > $out/sweep_$w.json 2> $out/sweep_$w.err
done
for w in boids2d-16M-f64 circle3d-1M-f64 circle3d-16M-f64 circle3d-16M-f32 game_of_life-16M-f64; do
  python bench.py --workload $w --no-cpu-baseline --steps 20 --warmup 5 > $out/sweep_$w.json 2> $out/sweep_$w.err
done
python bench.py --tile --no-cpu-baseline > $out/sweep_boids2d-1M-f64-tile.json 2> /dev/null
python tools/quick_step.py predator_prey-4M-f64 --steps 50 > $out/quick_predator_prey-4M.txt 2>&1
kill $smi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_n1.json") + glob.glob("gpurun_out/sweep_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline", {})
            print(f.split("/")[-1], "%.3f G/s" % (d["value"] / 1e9), d.get("ms_per_step"), "e2e %.3f" % (d["e2e"]["value"] / 1e9),
                  "kernel_ms", r.get("kernel_ms"), "bin_ms", r.get("bin_ms"), "frac", r.get("frac"), "whole", r.get("whole_step_frac"))
PY
cat gpurun_out/quick_predator_prey-4M.txt | tail -1
