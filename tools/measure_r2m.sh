#!/bin/bash
# Final check of the round: GPU test tier, smoke(), one short bench line.
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $out/r2m_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2m_pytest_gpu.log
tail -n 9 $out/r2m_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 300 python bench.py --steps 20 --warmup 5 --no-companion --no-cpu-baseline > $out/r2m_bench_n1_short.json 2> $out/r2m_bench_n1_short.err
python - <<'PY'
import json
for l in open("gpurun_out/r2m_bench_n1_short.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("boids2d 1M: ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4), "kernel_ms", round(r["kernel_ms"], 4),
              "bin_ms", round(r["bin_ms"], 4), "frac", round(r["frac"], 3), "whole", round(r["whole_step_frac"], 3), "e2e", round(d["e2e"]["value"] / 1e9, 2))
PY
