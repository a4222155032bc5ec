#!/bin/bash
# Round 2, first GPU call: the full GPU test tier (with the new benchmark-size parity tests), the
# A/B of the candidate-loop variants that shipped untimed in round 1, and ncu captures of the
# kernels that actually run.   gpurun --timeout 2400 -- 'bash tools/measure_r2a.sh'
set -u
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2a_smi.txt 2>&1
nproc > $out/r2a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=25 > $out/r2a_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2a_pytest_gpu.log
for w in boids2d-1M-f64 boids2d-1M-f32 game_of_life-16M-f64; do
  for flat in 0 1 auto; do
    if [ $flat = auto ]; then env="ABL_CUDA_VERBOSE=1"; else env="ABL_CUDA_FLAT=$flat"; fi
    env $env timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 100 --warmup 10 \
      > $out/r2a_flat_${w}_$flat.json 2> $out/r2a_flat_${w}_$flat.err
  done
done
ABL_CUDA_VERBOSE=1 timeout 300 python bench.py --workload game_of_life-16M-f64 --nlist --no-cpu-baseline --steps 100 --warmup 10 \
  > $out/r2a_nlist_game_of_life-16M-f64.json 2> $out/r2a_nlist_game_of_life-16M-f64.err
for w in boids2d-4M-f64 boids2d-16M-f64 circle3d-1M-f64; do
  ABL_CUDA_VERBOSE=1 timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 30 --warmup 10 > $out/r2a_$w.json 2> $out/r2a_$w.err
done
ABL_CUDA_VERBOSE=1 timeout 300 python bench.py --workload circle3d-16M-f64 --no-cpu-baseline --steps 5 --warmup 3 > $out/r2a_circle3d-16M-f64.json 2> $out/r2a_circle3d-16M-f64.err
timeout 300 python tools/quick_step.py predator_prey-4M-f64 --steps 50 -C cuda.cull=false -C cuda.flat=false -C cuda.sqcmp=false > $out/r2a_pp4M_round1_kernels.txt 2>&1
ABL_CUDA_FLAT=0 timeout 300 python tools/quick_step.py predator_prey-4M-f64 --steps 50 > $out/r2a_pp4M_flat0.txt 2>&1
ABL_CUDA_VERBOSE=1 timeout 300 python tools/quick_step.py predator_prey-4M-f64 --steps 50 > $out/r2a_pp4M_tuned.txt 2>&1
timeout 300 python tools/quick_step.py boids2d-1M-f64 -C cuda.flat=false -C cuda.sqcmp=false > $out/r2a_boids1M_round1_kernels.txt 2>&1
# ncu: launch list, then full captures of the kernels the bench line names
ABL_CUDA_FLAT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/r2a_launches_boids2d_1M_flat.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/r2a_ncu_list.log 2>&1
ABL_CUDA_FLAT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_update_boid -s 20 -c 2 \
  -o $out/prof_r2a_boids_flat_f64 python bench.py --steps 30 --warmup 10 --no-cpu-baseline > $out/r2a_ncu_full_f64.log 2>&1
ABL_CUDA_FLAT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_update_boid -s 20 -c 2 \
  -o $out/prof_r2a_boids_flat_f32 python bench.py --workload boids2d-1M-f32 --steps 30 --warmup 10 --no-cpu-baseline > $out/r2a_ncu_full_f32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bin_\|k_tile_ -s 40 -c 8 \
  -o $out/prof_r2a_boids_binning python bench.py --steps 30 --warmup 10 --no-cpu-baseline > $out/r2a_ncu_full_bin.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_ -s 18 -c 1 \
  -o $out/prof_r2a_circle3d_1M python bench.py --workload circle3d-1M-f64 --steps 4 --warmup 3 --no-cpu-baseline > $out/r2a_ncu_full_c3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_ -s 18 -c 1 \
  -o $out/prof_r2a_gol_16M python bench.py --workload game_of_life-16M-f64 --steps 4 --warmup 3 --no-cpu-baseline > $out/r2a_ncu_full_gol.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_ -s 112 -c 7 \
  -o $out/prof_r2a_pp_4M python tools/quick_step.py predator_prey-4M-f64 --steps 5 > $out/r2a_ncu_full_pp.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline", {})
            print(f.split("/")[-1], "%.3f G/s" % (d["value"] / 1e9), "ms", round(d.get("ms_per_step"), 4),
                  "steady", round(d.get("steady_state", {}).get("ms_per_step"), 4), "kernel_ms", round(r.get("kernel_ms"), 4),
                  "bin_ms", round(r.get("bin_ms"), 4), "whole", round(r.get("whole_step_frac"), 3), d["config"].get("candidate_loop_in_use"))
PY
grep -h "variant" gpurun_out/r2a_*.err | head -20
tail -n 1 $out/r2a_pp4M_*.txt $out/r2a_boids1M_*.txt
tail -n 30 $out/r2a_pytest_gpu.log
