#!/bin/bash
# Round 2, second session, call 1 (one B200): GPU test tier with the forked halo exchange and the scan-ranked
# append, a short bench line (stage pass behind a spin kernel), two slabs on one GPU with the exchange forked / in-stream.
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > $out/r2d_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2d_pytest_gpu.log
tail -n 12 $out/r2d_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-companion --no-cpu-baseline > $out/r2d_bench_n1_short.json 2> $out/r2d_bench_n1_short.err
python - <<'PY'
import json
for l in open("gpurun_out/r2d_bench_n1_short.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("boids2d 1M: ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4), "kernel_ms", round(r["kernel_ms"], 4),
              "bin_ms", round(r["bin_ms"], 4), "commit", round(r["commit_ms"], 4), "frac", round(r["frac"], 3), "whole", round(r["whole_step_frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 2))
PY
for a in 1 0; do
  echo "ABL_CUDA_HALO_ASYNC=$a"
  ABL_CUDA_HALO_ASYNC=$a timeout 200 python tools/quick_slabs.py boids2d-1M-f64 --slabs 2 --transport direct --steps 100 2>&1 | tail -n 2
done
