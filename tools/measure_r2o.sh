#!/bin/bash
# Same-box A/B of ABL_CUDA_BIN_SEGINFO (segment begin / size per agent from k_bin_scatter), then the GPU test tier.
set -u
out=gpurun_out
mkdir -p $out
for v in 0 1 0 1; do
  ABL_CUDA_BIN_SEGINFO=$v timeout 200 python bench.py --steps 60 --warmup 5 --no-companion --no-cpu-baseline > $out/r2o_seginfo$v.json 2> $out/r2o_seginfo$v.err
  python - $out/r2o_seginfo$v.json $v <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("seginfo", sys.argv[2], "ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4), "bin_ms", round(r["bin_ms"], 4), "whole", round(r["whole_step_frac"], 3))
PY
done
timeout 600 python -m pytest tests -m gpu -q -x > $out/r2o_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2o_pytest_gpu.log
tail -n 3 $out/r2o_pytest_gpu.log
