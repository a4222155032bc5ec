#!/bin/bash
# 2 B200, final code: slab / multi-device tests, then the driver's command at N=2 (boids2d weak, no companion).
set -u
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_multi.py tests/test_gpu_lifecycle.py -m gpu -q -x > $out/r2l_pytest_slabs.log 2>&1; echo "slab tests rc=$?"; tail -n 2 $out/r2l_pytest_slabs.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 100 --warmup 10 --no-companion --no-cpu-baseline > $out/r2l_bench_n2.json 2> $out/r2l_bench_n2.err
python - $out/r2l_bench_n2.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("N", d["n_gpus"], "G/s", round(d["value"] / 1e9, 2), "ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4),
              "kernel_ms", round(r["kernel_ms"], 4), "bin_ms", round(r["bin_ms"], 4), "commit", round(r["commit_ms"], 4), "e2e", round(d["e2e"]["value"] / 1e9, 2))
PY
