mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/B_pytest.log 2>&1
tail -12 gpurun_out/B_pytest.log
for w in boids2d-1M-f64 boids2d-1M-f32 circle3d-1M-f64 game_of_life-16M-f64; do
  for u in "" "--unroll"; do
    timeout 100 python bench.py --workload $w --no-cpu-baseline --steps 50 --warmup 5 $u > gpurun_out/B_${w}${u}.json 2> gpurun_out/B_${w}${u}.err
    python - "$w" "$u" <<'PY'
import json, sys
w, u = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/B_%s%s.json" % (w, u)).readline())
    print(w, u or "plain", "%.4f ms/step" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "bin_ms %.4f" % d["roofline"]["bin_ms"])
except Exception as e:
    print(w, u, "FAILED", e)
PY
  done
done
