#!/usr/bin/env python
"""Quick device-side timing of one workload: ms per timestep and per-stage times.
usage: quick_step.py WORKLOAD [--tile] [--steps K] [--bs B] [-C key=value ...]"""
import argparse, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402
from openabl_b200.model import Model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--tile", action="store_true")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--bs", type=int, default=0)
ap.add_argument("-C", dest="config", action="append", default=[], help="code-generator option, e.g. -C cuda.cull=false")
args = ap.parse_args()
config = {}
for kv in args.config:
    k, v = kv.split("=", 1)
    config[k] = {"true": True, "false": False}.get(v, v)
model_file, params, use_float, S, M, P = bench.WORKLOADS[args.workload]
m = Model(os.path.join(REPO, "examples", model_file), dict(params), use_float=use_float, config=config or None)
m.populate()
m.create_runtime(device=0, block_size=args.bs, tile=args.tile)
m.upload_host()
stream = torch.cuda.ExternalStream(m.rt.stream())
for _ in range(16):      # covers the launchers' tuning phase (12 trial launches per step function)
    m.timestep()
m.rt.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(args.steps):
    m.timestep()
e1.record(stream)
m.rt.synchronize()
ms = e0.elapsed_time(e1) / args.steps
m.rt.enable_timing(True)
st = {"bin_ms": 0.0, "kernel_ms": 0.0, "commit_ms": 0.0}
for _ in range(20):
    for s in range(m.n_steps):
        m.run_step(s)
        lt = m.rt.last_timing()
        for k in st:
            st[k] += lt[k] / 20
n = sum(m.host_count(t) for t in range(m.n_types))
modes = {m.step_names[s]: m.step_variant(s) for s in range(m.n_steps)}
print("%s tile=%s bs=%d %s: %.4f ms/step, %.2f G agent-steps/s; stages %s; ABL_MODE per step %s" % (
    args.workload, args.tile, args.bs, config, ms, n / ms / 1e6, {k: round(v, 4) for k, v in st.items()}, modes))
