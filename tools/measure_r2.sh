#!/bin/bash
# First GPU call of round 2: times the flat candidate loop and the squared-distance comparisons
# (written after round 1's GPU budget had run out; verified bit-exact under the CPU emulator,
# tests/test_emu_kernels.py) against the cursor loop, then captures the mode-3 kernel with ncu.
#   gpurun --timeout 1500 -- 'bash tools/measure_r2.sh'
# JSON lines land in gpurun_out/r2_*.json; copy the ones quoted in DESIGN.md to profiles/scaling/.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r2_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2_pytest_gpu.log
for w in boids2d-1M-f64 boids2d-1M-f32 game_of_life-16M-f64; do
  for flat in 0 1 auto; do
    if [ $flat = auto ]; then env="ABL_CUDA_VERBOSE=1"; else env="ABL_CUDA_FLAT=$flat"; fi
    env $env python bench.py --workload $w --no-cpu-baseline --steps 100 --warmup 10 \
      > $out/r2_flat_${w}_$flat.json 2> $out/r2_flat_${w}_$flat.err
  done
done
# cached neighbour lists (game_of_life): guarded GPU tests first, then the A/B
python -m pytest tests/test_gpu_zz_nlist.py -m gpu -q -rxX > $out/r2_pytest_nlist.log 2>&1
ABL_CUDA_VERBOSE=1 python bench.py --workload game_of_life-16M-f64 --nlist --no-cpu-baseline --steps 100 --warmup 10 \
  > $out/r2_nlist_game_of_life-16M-f64.json 2> $out/r2_nlist_game_of_life-16M-f64.err
# predator_prey 4 M: round-1 kernels (no culling, no flat loop, square roots) / pinned cursor / tuned default
python tools/quick_step.py predator_prey-4M-f64 --steps 50 -C cuda.cull=false -C cuda.flat=false -C cuda.sqcmp=false > $out/r2_pp4M_round1_kernels.txt 2>&1
ABL_CUDA_FLAT=0 python tools/quick_step.py predator_prey-4M-f64 --steps 50 > $out/r2_pp4M_flat0.txt 2>&1
python tools/quick_step.py predator_prey-4M-f64 --steps 50 > $out/r2_pp4M_tuned.txt 2>&1
# round-1 kernels of the headline workload, same box, same run
python tools/quick_step.py boids2d-1M-f64 -C cuda.flat=false -C cuda.sqcmp=false > $out/r2_boids1M_round1_kernels.txt 2>&1
python tools/quick_step.py boids2d-1M-f64 > $out/r2_boids1M_tuned.txt 2>&1
python bench.py > $out/r2_bench_default_n1.json 2> $out/r2_bench_default_n1.err
# launch list + one full capture of the flat-loop kernel (numbers under ncu are never bench values)
ABL_CUDA_FLAT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/r2_launches_boids2d_1M_flat.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/r2_ncu_list.log 2>&1
ABL_CUDA_FLAT=1 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_update_boid -s 20 -c 2 \
  -o $out/prof_r2_boids_flat python bench.py --steps 30 --warmup 10 --no-cpu-baseline > $out/r2_ncu_full.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline", {})
            print(f.split("/")[-1], "%.3f G/s" % (d["value"] / 1e9), "ms", d.get("ms_per_step"),
                  "steady", d.get("steady_state", {}).get("ms_per_step"), "kernel_ms", r.get("kernel_ms"),
                  "bin_ms", r.get("bin_ms"), "whole", r.get("whole_step_frac"))
PY
grep -h "variant" gpurun_out/r2_flat_*_auto.err
tail -n 1 $out/r2_pp4M_*.txt $out/r2_boids1M_*.txt
