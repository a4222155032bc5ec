mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/D_pytest.log 2>&1
tail -12 gpurun_out/D_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/D_clocks.csv &
smi=$!
(time timeout 200 python bench.py) > gpurun_out/D_bench.json 2> gpurun_out/D_bench.err
kill $smi
tail -3 gpurun_out/D_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/D_bench.json").readline())
    print("value %.3f G/s (%.4f ms/step, L2 flushed)  steady %.3f G/s (%.4f ms)  e2e %.3f  kernel_ms %.4f bin_ms %.4f frac %.3f whole %.3f launches %d" % (
        d["value"] / 1e9, d["ms_per_step"], d["steady_state"]["value"] / 1e9, d["steady_state"]["ms_per_step"], d["e2e"]["value"] / 1e9,
        d["roofline"]["kernel_ms"], d["roofline"]["bin_ms"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["gpu_launches"]))
    print(d["clocks"], d["cpu_baseline"]["value"])
except Exception as e:
    print("bench FAILED", e)
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/D_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/D_ncu_launch.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:abl_kernel_update -s 6 -c 2 -f -o gpurun_out/prof_r1d_boids python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/D_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
