#!/bin/bash
# 2 B200: owned-range report from k_bin_scatter (default) against from k_tile_scan, same box; slab tests first (one GPU).
set -u
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_multi.py tests/test_gpu_parity_fullsize.py -m gpu -q -x > $out/r2i_pytest_slabs.log 2>&1; echo "slab tests rc=$?"; tail -n 3 $out/r2i_pytest_slabs.log
for v in 0 1 0 1; do
  ABL_CUDA_REPORT_IN_SCAN=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 100 --warmup 10 --no-companion --no-cpu-baseline > $out/r2i_n2_scan$v.json 2> $out/r2i_n2_scan$v.err
  python - $out/r2i_n2_scan$v.json $v <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("report_in_scan", sys.argv[2], "G/s", round(d["value"] / 1e9, 2), "ms", round(d["ms_per_step"], 4), "steady", round(d["steady_state"]["ms_per_step"], 4),
              "kernel_ms", round(r["kernel_ms"], 4), "bin_ms", round(r["bin_ms"], 4), "commit", round(r["commit_ms"], 4), "e2e", round(d["e2e"]["value"] / 1e9, 2))
PY
done
