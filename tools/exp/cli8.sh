#!/bin/bash
out=gpurun_out; mkdir -p $out
d=$(python - <<'PY'
import os, sys
sys.path.insert(0, ".")
from openabl_b200 import build
print(build.build_model(os.path.join("examples", "circle3d.abl"), {"num_agents": 4000000, "num_timesteps": 5}, {}))
PY
)
rm -rf /tmp/one /tmp/eight; mkdir -p /tmp/one /tmp/eight
( cd /tmp/one && s=$(date +%s.%N); $d/main; e=$(date +%s.%N); echo "1 GPU: $(echo "$e - $s" | bc) s" )
( cd /tmp/eight && s=$(date +%s.%N); ABL_CUDA_GPUS=8 $d/main; e=$(date +%s.%N); echo "8 GPUs: $(echo "$e - $s" | bc) s" )
ls -la /tmp/one /tmp/eight | grep points
if cmp -s /tmp/one/points.json /tmp/eight/points.json; then echo "circle3d 4M x 5 steps: points.json IDENTICAL on 1 and 8 GPUs ($(stat -c %s /tmp/one/points.json) bytes, md5 $(md5sum < /tmp/one/points.json | cut -c1-12))"; else echo "points.json DIFFERS"; fi | tee $out/r2s_cli_8gpu.txt
