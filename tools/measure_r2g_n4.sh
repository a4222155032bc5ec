#!/bin/bash
# Round 2, second session: the driver's scaling command on N GPUs (boids2d weak + circle3d-16M strong companion).
set -u
out=gpurun_out
mkdir -p $out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 20 --warmup 5 > $out/r2g_bench_n$N.json 2> $out/r2g_bench_n$N.err
python - $out/r2g_bench_n$N.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        for name, x in (("boids2d", d), ("circle3d", d.get("circle3d"))):
            if not x: continue
            r = x["roofline"]
            print(name, "N", d["n_gpus"], x["scaling"], "G/s", round(x["value"] / 1e9, 3), "ms", round(x["ms_per_step"], 4), "steady", round(x["steady_state"]["ms_per_step"], 4),
                  "kernel_ms", round(r["kernel_ms"], 4), "bin_ms", round(r["bin_ms"], 4), "commit", round(r["commit_ms"], 4), "e2e", round(x["e2e"]["value"] / 1e9, 3))
PY
tail -n 4 $out/r2g_bench_n$N.err
