#!/bin/bash
# Round 2, final measurements on one B200: GPU test tier, bench lines, ncu launch list and full captures of the
# kernels the bench lines name.   gpurun --timeout 2700 -- 'bash tools/measure_r2c.sh'
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > $out/r2c_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2c_pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 10 > $out/r2c_bench_default_n1.json 2> $out/r2c_bench_default_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/r2c_bench_reference_n1.json 2> $out/r2c_bench_reference_n1.err
for w in boids2d-1M-f32 boids2d-4M-f64 boids2d-16M-f64 game_of_life-16M-f64 circle3d-1M-f64; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 30 --warmup 10 > $out/r2c_$w.json 2> $out/r2c_$w.err
done
timeout 300 python bench.py --workload game_of_life-16M-f64 --nlist --no-cpu-baseline --steps 30 --warmup 10 > $out/r2c_game_of_life-16M-f64-nlist.json 2> $out/r2c_nlist.err
timeout 300 python bench.py --workload circle3d-16M-f32 --no-cpu-baseline --steps 5 --warmup 3 > $out/r2c_circle3d-16M-f32.json 2> $out/r2c_c3d32.err
timeout 300 python tools/quick_step.py predator_prey-4M-f64 --steps 50 > $out/r2c_pp4M.txt 2>&1
# ncu: launch list of the default bench command, then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/r2c_launches_boids2d_1M.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-companion > $out/r2c_ncu_list.log 2>&1
cap() { # name workload kernel-regex skip count extra-args
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -f -o $out/prof_r2c_$1 \
    python bench.py --workload $2 --steps 6 --warmup 3 --no-cpu-baseline ${6:-} > $out/r2c_ncu_$1.log 2>&1
}
cap boids_f64 boids2d-1M-f64 abl_kernel_update_boid 8 1 --no-companion
cap boids_f32 boids2d-1M-f32 abl_kernel_update_boid 8 1
cap boids_binning boids2d-1M-f64 'k_bin_|k_tile_' 24 4 --no-companion
cap circle3d_1M circle3d-1M-f64 'abl_kernel_|abl_prefilter_' 12 2
cap gol_16M game_of_life-16M-f64 abl_kernel_ 6 1
cap gol_16M_nlist game_of_life-16M-f64 abl_kernel_ 8 1 --nlist
timeout 900 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_ -s 60 -c 13 -f -o $out/prof_r2c_pp_4M \
  python tools/quick_step.py predator_prey-4M-f64 --steps 5 > $out/r2c_ncu_pp.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline", {})
            if not r: print(f, l[:300]); continue
            print(f.split("/")[-1], "%.3f G/s" % (d["value"] / 1e9), "ms", round(d.get("ms_per_step"), 4),
                  "steady", round(d.get("steady_state", {}).get("ms_per_step"), 4), "kernel_ms", round(r.get("kernel_ms"), 4),
                  "bin_ms", round(r.get("bin_ms"), 4), "whole", round(r.get("whole_step_frac"), 3), "e2e", round(d["e2e"]["value"]/1e9,3), d["config"].get("candidate_loop_in_use"))
            if "circle3d" in d: c=d["circle3d"]; print("   companion circle3d:", round(c["ms_per_step"],3), "ms", c.get("roofline_fp64",{}).get("frac"))
PY
tail -n 1 $out/r2c_pp4M.txt
tail -n 25 $out/r2c_pytest_gpu.log
