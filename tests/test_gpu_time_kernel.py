"""abl_cuda_time_kernel (include/abl_cuda.h): the step function it times still runs exactly once — repeating
the launch on the same input buffers and clearing the fused histogram afterwards must leave no trace in the
simulation — and it refuses nothing silently: steps that add or remove agents report 0 ms and run normally."""
import os

import numpy as np
import pytest

from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(model_file, params, steps, timed):
    m = Model(os.path.join(REPO, "examples", model_file), params)
    m.populate()
    m.create_runtime()
    m.upload_host()
    ms = []
    for k in range(steps):
        if timed and k % 2 == 1:
            m.rt.begin_timestep()
            ms.append([m.rt.time_kernel(s, 5) for s in range(m.n_steps)])
            m.sequential_step()
            m.rt.end_timestep()     # (advances the counter the in-step random numbers are keyed by)
        else:
            m.timestep()
    out = [m.download(t) for t in range(m.n_types)]
    m.close()
    return out, ms


@pytest.mark.gpu
@pytest.mark.parametrize("model_file,n", [("boids2d.abl", 100000), ("circle3d.abl", 20000), ("game_of_life.abl", 65536)])
def test_timing_a_kernel_leaves_the_simulation_unchanged(model_file, n):
    plain, _ = _run(model_file, {"num_agents": n}, 6, False)
    timed, ms = _run(model_file, {"num_agents": n}, 6, True)
    assert len(ms) == 3 and all(v > 0 for row in ms for v in row)
    for a, b in zip(plain, timed):
        assert len(a) == len(b)
        for f in a.dtype.names:
            assert np.array_equal(a[f], b[f]), "%s differs after timed steps" % f


@pytest.mark.gpu
def test_steps_that_add_or_remove_agents_are_not_repeated():
    plain, _ = _run("predator_prey.abl", {"num_agents": 32000}, 4, False)
    timed, ms = _run("predator_prey.abl", {"num_agents": 32000}, 4, True)
    m = Model(os.path.join(REPO, "examples", "predator_prey.abl"), {"num_agents": 32000})
    flags = [m.step_flags(s) for s in range(m.n_steps)]
    assert any(flags)
    for row in ms:
        for s, v in enumerate(row):
            assert (v == 0) == bool(flags[s]), "step %d: flags %d, %.4f ms" % (s, flags[s], v)
    for a, b in zip(plain, timed):
        assert len(a) == len(b)
        for f in a.dtype.names:
            assert np.array_equal(a[f], b[f])
