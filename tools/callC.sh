mkdir -p gpurun_out
export ABL_CUDA_HALO_TIMEOUT_MS=20000
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline "$@"; }
show() { python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.3f G/s" % (d["value"] / 1e9), "%.4f ms/step" % d["ms_per_step"], "e2e %.3f" % (d["e2e"]["value"] / 1e9), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
            break
    else:
        print(f, "NO JSON LINE")
except Exception as e:
    print(f, "FAILED", e)
PY
}
ABL_CUDA_DEVICE_RANGE=1 run --steps 300 --warmup 20 > gpurun_out/C_weak2_dr1.json 2> gpurun_out/C_weak2_dr1.err; show gpurun_out/C_weak2_dr1.json
ABL_CUDA_DEVICE_RANGE=0 run --steps 300 --warmup 20 > gpurun_out/C_weak2_dr0.json 2> gpurun_out/C_weak2_dr0.err; show gpurun_out/C_weak2_dr0.json
run --steps 30 --warmup 5 --workload predator_prey-4M-f64 --strong > gpurun_out/C_pp2.json 2> gpurun_out/C_pp2.err; show gpurun_out/C_pp2.json
tail -3 gpurun_out/C_pp2.err
