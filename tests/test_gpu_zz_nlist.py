"""Cached neighbour lists (-C cuda.nlist=true) on the GPU.

The list kernels are verified bit for bit under the CPU emulator (tests/test_emu_kernels.py);
the runtime side (count pass -> sizing -> fill pass, validity tracking) was written after the
round's GPU budget had been spent and has not run on a device yet — hence the non-strict xfail
guard: a pass shows up as XPASS, a failure does not hide the rest of the suite.  Remove the guard
after the first green run."""
import os

import numpy as np
import pytest

from oracle import GRID, Oracle
from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
first_run = pytest.mark.xfail(reason="neighbour-list runtime path not yet validated on a GPU", strict=False)


def run(path, params, steps, config, use_float=False):
    m = Model(path, params, use_float=use_float, config=config)
    m.populate()
    m.create_runtime()
    m.upload_host()
    for _ in range(steps):
        m.timestep()
    out = [m.download(t) for t in range(m.n_types)]
    launches = m.rt.last_timing()["launches"]
    m.close()
    return out, launches


@pytest.mark.gpu
@first_run
def test_game_of_life_with_neighbour_lists_equals_grid_oracle():
    params = {"num_agents": 65536}
    got, _ = run(os.path.join(REPO, "examples", "game_of_life.abl"), params, 10, {"cuda.nlist": True})
    o = Oracle(False)
    want = o.run_for("game_of_life.abl", params, o.init_for("game_of_life.abl", params), 10, GRID)
    assert np.array_equal(got[0]["alive"], want["alive"])
    assert np.array_equal(got[0]["pos"], want["pos"])


@pytest.mark.gpu
@first_run
def test_lists_of_static_sites_survive_moving_walkers():
    """Site-Site lists are built once while the Walker pool is re-binned every timestep; the state
    equals the run without lists bit for bit, and re-uploading the population rebuilds the lists."""
    path = os.path.join(REPO, "tests", "models", "static_sites.abl")
    params = {"num_agents": 3000}
    plain, _ = run(path, params, 5, None)
    listed, _ = run(path, params, 5, {"cuda.nlist": True})
    for a, b in zip(plain, listed):
        assert len(a) == len(b)
        for f in a.dtype.names:
            assert np.array_equal(a[f], b[f]), "member %s differs" % f
    m = Model(path, params, config={"cuda.nlist": True})
    m.populate()
    m.create_runtime()
    for _ in range(2):          # second upload: same records, lists must be rebuilt for the new pool order
        m.upload_host()
        for _ in range(5):
            m.timestep()
        again = [m.download(t) for t in range(m.n_types)]
        for a, b in zip(plain, again):
            for f in a.dtype.names:
                assert np.array_equal(a[f], b[f]), "member %s differs after re-upload" % f
    m.close()
