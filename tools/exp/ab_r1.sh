#!/bin/bash
# same-box A/B: round-1 tree (build/r1) against the current one, N GPUs
N=${1:-2}
parse='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("ms",round(d["ms_per_step"],4),"steady",round(d["steady_state"]["ms_per_step"],4),"kernel",round(r["kernel_ms"],4),"bin",round(r["bin_ms"],4),"commit",round(r["commit_ms"],4),"e2e G/s",round(d["e2e"]["value"]/1e9,3),d["config"].get("candidate_loop_in_use"))'
run() { if [ $N = 1 ]; then python bench.py "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; fi; }
echo "== round 1"; (cd build/r1 && run --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "$parse")
echo "== now"; run --steps 20 --warmup 5 --no-cpu-baseline --no-companion 2>/dev/null | python -c "$parse"
echo "== now, warmup 16"; run --steps 20 --warmup 16 --no-cpu-baseline --no-companion 2>/dev/null | python -c "$parse"
