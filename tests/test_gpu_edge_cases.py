"""Edge cases of the hot path on the GPU: cell size smaller than the radius (multi-cell
reach), empty and tiny populations, every agent in one cell, positions outside the
environment."""
import os

import numpy as np
import pytest

from oracle import BRUTE, GRID, Oracle
from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(model, steps, host=None):
    if host is None:
        model.populate()
        model.create_runtime()
        model.upload_host()
    else:
        model.create_runtime()
        for t, arr in enumerate(host):
            model.upload(t, arr)
    for _ in range(steps):
        model.timestep()
    out = [model.download(t) for t in range(model.n_types)]
    model.close()
    return out


@pytest.mark.gpu
def test_reach_two_cells_matches_oracle():
    """granularity 5 < radius 10: the iterator walks 5x5 cells through its on-demand row path."""
    n, steps = 20000, 10
    m = Model(os.path.join(REPO, "tests", "models", "circle_fine_grid.abl"), {"num_agents": n})
    got = run(m, steps)[0]
    o = Oracle(False)
    state = o.circle_init(2, n)
    want = o.circle_run(2, state, steps, GRID, granularity=5.0)
    assert np.array_equal(got["pos"], want["pos"])
    # and the finer grid only changes the summation order, not the physics
    ref = o.circle_run(2, state, steps, BRUTE)
    assert np.max(np.abs(got["pos"] - ref["pos"]) / np.maximum(np.abs(ref["pos"]), 1.0)) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 3])
def test_empty_and_tiny_populations(n):
    m = Model(os.path.join(REPO, "examples", "boids2d.abl"), {"num_agents": 100000})
    m.populate()
    full = m.host_agents(0)
    sub = np.ascontiguousarray(full[:n])
    got = run(m, 3, host=[sub])[0]
    assert len(got) == n
    if n:
        o = Oracle(False)
        want = o.boids_run(sub, 3, GRID, num_agents=100000)
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[f])


@pytest.mark.gpu
def test_all_agents_in_one_cell_and_outside_the_environment():
    """5000 agents inside one cell (ties ordered by id) plus agents beyond the environment
    bounds (clamped into the border cells): still exactly the oracle's result."""
    m = Model(os.path.join(REPO, "examples", "boids2d.abl"), {"num_agents": 100000})
    m.populate()
    rng = np.random.default_rng(7)
    n = 6000
    a = np.zeros(n, dtype=m.dtypes[0])
    a["pos"][:5000] = 5.0 + rng.random((5000, 2)) * 0.04          # one 0.05 x 0.05 cell
    a["pos"][5000:5500] = -3.0 + rng.random((500, 2))             # below the lower bound
    a["pos"][5500:] = 14.2 + rng.random((500, 2)) * 5.0           # beyond max_pos (= 14.14)
    a["velocity"] = rng.random((n, 2)) * 2 - 1
    got = run(m, 2, host=[a])[0]
    o = Oracle(False)
    want = o.boids_run(a, 2, GRID, num_agents=100000)
    for f in got.dtype.names:
        assert np.array_equal(got[f], want[f]), f
