#!/usr/bin/env python
"""Device-side timing of a decomposed run on ONE GPU (S slabs, one runtime each).
usage: quick_slabs.py WORKLOAD --slabs S [--transport staged|direct] [--steps K] [--scale F]
--scale multiplies num_agents (weak scaling: S slabs of the workload's size each)."""
import argparse, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402
from openabl_b200.model import Model  # noqa: E402
from openabl_b200.slab import LocalSlabs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--slabs", type=int, default=2)
ap.add_argument("--transport", default="staged")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--scale", type=int, default=1)
ap.add_argument("--stage-times", action="store_true")
args = ap.parse_args()
model_file, params, use_float, S, M, P = bench.WORKLOADS[args.workload]
params = dict(params)
params["num_agents"] *= args.scale
m = Model(os.path.join(REPO, "examples", model_file), params, use_float=use_float)
m.populate()
host = [m.host_agents(t) for t in range(m.n_types)]
ls = LocalSlabs(m, args.slabs, transport=args.transport)
ls.upload(host)
for _ in range(5):
    ls.timestep()
for rt in ls.rts:
    rt.synchronize()
t0 = time.perf_counter()
for _ in range(args.steps):
    ls.timestep()
for rt in ls.rts:
    rt.synchronize()
dt = (time.perf_counter() - t0) / args.steps
print("%s x%d, %d slabs (%s): %.4f ms per timestep of all slabs (wall)" % (args.workload, args.scale, args.slabs, args.transport, dt * 1e3))
if args.stage_times and args.transport == "staged":
    rt = ls.rts[0]
    rt.enable_timing(True)
    st = {"bin_ms": 0.0, "kernel_ms": 0.0, "commit_ms": 0.0}
    for _ in range(10):
        for s in range(m.n_steps):
            from openabl_b200.runtime import check
            for r in ls.rts:
                check(m.lib.abl_model_run_step(r.handle, s), "run_step")
            lt = rt.last_timing()
            for k in st:
                st[k] += lt[k] / 10
            ls._exchange(ls.step_pool(s))
    print("slab 0 stages:", {k: round(v, 4) for k, v in st.items()}, "owned", rt.pool_size(m.pool(0)))
ls.close()
