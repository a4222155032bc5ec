#!/usr/bin/env python
"""Experiment driver: hand-patched model dir (ABL_MODEL_DIR) — mode 7 (ABL_EXP_MODE7=1) against the flat loop:
bit-equality after 10 steps, then device time per step for several block sizes."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import numpy as np
import torch
from openabl_b200.model import Model

model, n = sys.argv[1], int(sys.argv[2])
use_float = len(sys.argv) > 3 and sys.argv[3] == "f32"

def run(steps, env):
    for k in ("ABL_EXP_MODE7", "ABL_EXP_BS", "ABL_CUDA_FLAT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    m = Model(os.path.join(REPO, "examples", model), {"num_agents": n}, use_float=use_float)
    m.populate()
    m.create_runtime(device=0)
    m.upload_host()
    for _ in range(steps):
        m.timestep()
    out = m.download(0)
    # timing
    stream = torch.cuda.ExternalStream(m.rt.stream())
    for _ in range(20):
        m.timestep()
    m.rt.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(100):
        m.timestep()
    e1.record(stream)
    m.rt.synchronize()
    ms = e0.elapsed_time(e1) / 100
    m.rt.enable_timing(True)
    for _ in range(50):
        for s in range(m.n_steps):
            m.run_step(s)
    lt = m.rt.last_timing()
    mode = m.step_variant(0)
    m.close()
    return out, ms, lt, mode

ref, ms, lt, mode = run(10, {"ABL_CUDA_FLAT": "1"})
print("flat     mode %d: %.4f ms/step kernel %.4f bin %.4f" % (mode, ms, lt["kernel_ms"], lt["bin_ms"]), flush=True)
for bs in [int(x) for x in os.environ.get("ABL_EXP_BSLIST", "128,256,64").split(",")]:
    got, ms, lt, mode = run(10, {"ABL_CUDA_FLAT": "1", "ABL_EXP_MODE7": "1", "ABL_EXP_BS": str(bs)})
    same = all(np.array_equal(ref[f], got[f]) for f in ref.dtype.names)
    print("mode7 bs=%d mode %d: %.4f ms/step kernel %.4f bin %.4f  bit-equal to flat: %s" % (bs, mode, ms, lt["kernel_ms"], lt["bin_ms"], same), flush=True)
