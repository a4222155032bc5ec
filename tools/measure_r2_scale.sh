#!/bin/bash
# Round 2, 8 GPUs of one box: the multi-GPU CLI tests on real devices, the generated ./main on 1 and 8 devices
# (same output file), and bench.py at N = 2, 4, 8 (boids2d weak + circle3d-16M strong companion line).
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/measure_r2_scale.sh'
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_multi.py -m gpu -q > $out/r2s_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
d=$(python - <<'PY'
import os, sys
sys.path.insert(0, ".")
from openabl_b200 import build
print(build.build_model(os.path.join("examples", "circle3d.abl"), {"num_agents": 4000000, "num_timesteps": 5}, {}))
PY
)
rm -rf /tmp/one /tmp/eight; mkdir -p /tmp/one /tmp/eight
( cd /tmp/one && /usr/bin/time -f "1 GPU: %e s" $d/main ) 2>&1 | tail -1
( cd /tmp/eight && ABL_CUDA_GPUS=8 /usr/bin/time -f "8 GPUs: %e s" $d/main ) 2>&1 | tail -1
if cmp -s /tmp/one/points.json /tmp/eight/points.json; then echo "circle3d 4M x 5 steps: points.json IDENTICAL on 1 and 8 GPUs ($(stat -c %s /tmp/one/points.json) bytes)"; else echo "points.json DIFFERS"; fi | tee $out/r2s_cli_8gpu.txt
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 \
    > $out/r2s_bench_n$n.json 2> $out/r2s_bench_n$n.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --impl reference --gpus $n --steps 3 --warmup 1 \
    > $out/r2s_reference_n$n.json 2> $out/r2s_reference_n$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s_bench_n*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); r = d["roofline"]; c = d.get("circle3d", {})
            print(f.split("/")[-1], "boids2d weak: %.2f G/s" % (d["value"] / 1e9), "ms", round(d["ms_per_step"], 4), "kernel", round(r["kernel_ms"], 4),
                  "bin", round(r["bin_ms"], 4), "commit", round(r["commit_ms"], 4), "e2e %.2f G/s" % (d["e2e"]["value"] / 1e9),
                  "| circle3d-16M strong: %.2f ms, %.3f G/s, e2e %.3f G/s" % (c.get("ms_per_step", 0), c.get("value", 0) / 1e9, c.get("e2e", {}).get("value", 0) / 1e9))
for f in sorted(glob.glob("gpurun_out/r2s_reference_n*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f.split("/")[-1], "%.0f agent-steps/s" % d["value"], d["cpu_baseline"]["sample"][-60:])
PY
tail -3 $out/r2s_pytest_multi.log
