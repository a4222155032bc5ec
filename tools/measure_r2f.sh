#!/bin/bash
# Round 2, second session: GPU test tier + the default bench command (kernel time by abl_cuda_time_kernel).
set -u
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=6 > $out/r2f_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2f_pytest_gpu.log
tail -n 12 $out/r2f_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2f_bench_default_n1.json 2> $out/r2f_bench_default_n1.err
python - <<'PY'
import json
for l in open("gpurun_out/r2f_bench_default_n1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        for name, x in (("boids2d", d), ("circle3d", d.get("circle3d"))):
            if not x: continue
            r = x["roofline"]
            print(name, "ms", round(x["ms_per_step"], 4), "steady", round(x["steady_state"]["ms_per_step"], 4), "kernel_ms", round(r["kernel_ms"], 4),
                  "stage-events", round(r["kernel_ms_stage_events"], 4), "bin_ms", round(r["bin_ms"], 4), "frac", round(r["frac"], 3),
                  "whole", round(r["whole_step_frac"], 3), "e2e", round(x["e2e"]["value"] / 1e9, 3), x.get("roofline_fp64", {}).get("frac"))
PY
tail -n 3 $out/r2f_bench_default_n1.err
