#!/bin/bash
# Round 2, second session, final single-GPU measurements: GPU test tier, the driver's bench command, boids2d 16 M,
# ncu launch list of the bench command and full captures of the step kernel, the binning kernels and circle3d.
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $out/r2j_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2j_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2j_bench_default_n1.json 2> $out/r2j_bench_default_n1.err
for w in boids2d-16M-f64 boids2d-1M-f32; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-companion --steps 30 --warmup 10 > $out/r2j_$w.json 2> $out/r2j_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/r2j_launches_boids2d_1M.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-companion > $out/r2j_ncu_list.log 2>&1
cap() { # name workload kernel-regex skip count extra-args
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -f -o $out/prof_r2j_$1 \
    python bench.py --workload $2 --steps 6 --warmup 3 --no-cpu-baseline ${6:-} > $out/r2j_ncu_$1.log 2>&1
}
[ -n "${WITH_STEP_KERNEL:-}" ] && cap boids_f64 boids2d-1M-f64 abl_kernel_update_boid 8 1 --no-companion
cap boids_binning boids2d-1M-f64 'k_bin_|k_tile_' 24 4 --no-companion
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2j_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline", {})
            if not r: print(f, l[:300]); continue
            print(f.split("/")[-1], "%.3f G/s" % (d["value"] / 1e9), "ms", round(d.get("ms_per_step"), 4),
                  "steady", round(d.get("steady_state", {}).get("ms_per_step"), 4), "kernel_ms", round(r.get("kernel_ms"), 4), "(stage events", round(r.get("kernel_ms_stage_events"), 4), ")",
                  "bin_ms", round(r.get("bin_ms"), 4), "frac", round(r.get("frac"), 3), "whole", round(r.get("whole_step_frac"), 3), "e2e", round(d["e2e"]["value"]/1e9,3), d["config"].get("candidate_loop_in_use"))
            if "circle3d" in d: c=d["circle3d"]; print("   companion circle3d:", round(c["ms_per_step"],3), "ms", c.get("roofline_fp64",{}).get("frac"))
PY
tail -n 8 $out/r2j_pytest_gpu.log
