"""GPU parity against the plain-C oracle in grid mode (same neighbour order as the kernels).

The oracle is pinned to the real reference by tests/test_oracle.py.  In grid mode it
performs the same arithmetic in the same order as the CUDA step kernels, so in double
precision the comparison is BIT-EXACT, at population sizes the O(N^2) reference cannot
run.  With use_float the reference evaluates unsuffixed literals in double while the
kernels stay in float, so the bar is the north_star's 1e-4.
"""
import os

import numpy as np
import pytest

from oracle import GRID, Oracle
from openabl_b200.model import Model
from openabl_b200.state import F32_FLOOR_ULPS, exact_members_equal, max_rel_error

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (model, params, use_float, steps)
CASES = [
    ("circle.abl", {"num_agents": 1000, "num_timesteps": 100}, False, 100),   # BASELINE config 1
    ("circle.abl", {"num_agents": 50000}, False, 10),
    ("circle3d.abl", {"num_agents": 20000}, False, 5),
    ("boids2d.abl", {"num_agents": 100000}, False, 10),
    ("boids2d.abl", {"num_agents": 100000}, True, 10),
    ("game_of_life.abl", {"num_agents": 65536}, False, 10),
    ("game_of_life.abl", {"num_agents": 65536}, True, 10),
]


def gpu_run(model_file, params, use_float, steps, config=None, **rt_kw):
    m = Model(os.path.join(REPO, "examples", model_file), params, use_float=use_float, config=config)
    m.populate()
    init = m.host_agents(0)
    m.create_runtime(**rt_kw)
    m.upload_host()
    for _ in range(steps):
        m.timestep()
    out = m.download(0)
    m.close()
    return init, out


@pytest.mark.gpu
@pytest.mark.parametrize("tile", [False, True], ids=["global", "smem-tile"])
@pytest.mark.parametrize("model_file,params,use_float,steps", CASES,
                         ids=["%s-%d-%s" % (c[0][:-4], c[1]["num_agents"], "f32" if c[2] else "f64") for c in CASES])
def test_gpu_equals_grid_oracle(model_file, params, use_float, steps, tile):
    """Both neighbour-loop variants: the L1-cached global-memory loop (default) and the
    shared-memory tile path (cuda.tile=true)."""
    init, got = gpu_run(model_file, params, use_float, steps, tile=tile)
    o = Oracle(use_float)
    state = o.init_for(model_file, params)
    # the generated host program and the oracle build the same initial population
    for f in init.dtype.names:
        assert np.array_equal(init[f], state[f]), "initial %s differs" % f
    want = o.run_for(model_file, params, state, steps, GRID)
    assert len(got) == len(want)
    assert exact_members_equal(got, want), "integer/bool state differs"
    if use_float:
        assert max_rel_error(got, want, floor_ulps=F32_FLOOR_ULPS) <= 1e-4
    else:
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[f]), "member %s is not bit-equal (max rel err %.3e)" % (
                f, max_rel_error(got, want))


UNROLL_CASES = [CASES[1], CASES[2], CASES[3], CASES[4], CASES[5]]


@pytest.mark.gpu
@pytest.mark.parametrize("model_file,params,use_float,steps", UNROLL_CASES,
                         ids=["%s-%d-%s" % (c[0][:-4], c[1]["num_agents"], "f32" if c[2] else "f64") for c in UNROLL_CASES])
def test_unrolled_candidate_loop_equals_grid_oracle(model_file, params, use_float, steps):
    """-C cuda.unroll=true: the for-near loop unrolled by two with alternating prefetch
    registers visits the same candidates in the same order."""
    init, got = gpu_run(model_file, params, use_float, steps, config={"cuda.unroll": True})
    o = Oracle(use_float)
    state = o.init_for(model_file, params)
    want = o.run_for(model_file, params, state, steps, GRID)
    assert len(got) == len(want)
    assert exact_members_equal(got, want), "integer/bool state differs"
    if use_float:
        assert max_rel_error(got, want, floor_ulps=F32_FLOOR_ULPS) <= 1e-4
    else:
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[f]), "member %s is not bit-equal" % f
