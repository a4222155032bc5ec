"""Unit-level GPU tests of runtime stages through the C ABI: binning order, transfers,
reductions."""
import os

import numpy as np
import pytest

from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cell_keys(pos, origin, cell, n_cell):
    c = np.floor((pos - origin) * (1.0 / cell)).astype(np.int64)  # same formula as abl_cell_coord
    c = np.clip(c, 0, np.array(n_cell[: pos.shape[1]]) - 1)
    key = c[:, 1] * n_cell[0] + c[:, 0]
    if pos.shape[1] == 3:
        key = key + c[:, 2] * n_cell[0] * n_cell[1]
    return key


@pytest.mark.gpu
@pytest.mark.parametrize("model_file,n", [("boids2d.abl", 4000), ("circle3d.abl", 2000), ("circle.abl", 1000)])
def test_binning_is_sorted_by_cell_then_id(model_file, n):
    params = {"num_agents": n, "num_timesteps": 10 if model_file != "circle.abl" else 100}
    m = Model(os.path.join(REPO, "examples", model_file), params)
    m.populate()
    host = m.host_agents(0)
    m.create_runtime()
    m.upload_host()
    m.rt.bin(m.pool(0))
    cell_start, ids = m.rt.debug_binning(m.pool(0))
    n_cells, n_cell = m.rt.grid_cells()
    back = m.download(0)
    m.close()

    assert sorted(ids.tolist()) == list(range(n)), "ids are not a permutation"
    # expected order: ascending (cell key, id)
    cell = {"boids2d.abl": 0.05, "circle3d.abl": 10.0, "circle.abl": 10.0}[model_file]
    keys = cell_keys(host["pos"].astype(np.float64), 0.0, cell, n_cell)
    expect = np.lexsort((np.arange(n), keys))
    assert np.array_equal(ids, expect.astype(np.uint32))
    counts = np.bincount(keys, minlength=n_cells)
    assert np.array_equal(cell_start, np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32))
    # download un-permutes by id
    for f in back.dtype.names:
        assert np.array_equal(back[f], host[f])


@pytest.mark.gpu
def test_reductions_match_numpy():
    m = Model(os.path.join(REPO, "examples", "game_of_life.abl"), {"num_agents": 4096, "num_timesteps": 10})
    m.populate()
    host = m.host_agents(0)
    m.create_runtime()
    m.upload_host()
    p = m.pool(0)
    assert m.rt.count(p) == len(host)
    assert m.rt.sum_int(p, 1) == int(host["alive"].sum())
    assert m.rt.count_member_int(p, 1, 1) == int(host["alive"].sum())
    assert m.rt.count_member_int(p, 1, 0) == int((~host["alive"]).sum())
    sx = m.rt.sum_float(p, 0, 0)
    assert abs(sx - host["pos"][:, 0].sum()) <= 1e-9 * abs(sx)
    m.close()


@pytest.mark.gpu
def test_missing_device_or_library_fails_loudly(tmp_path):
    from openabl_b200 import runtime
    with pytest.raises(Exception):
        runtime._lib = None
        try:
            runtime.load_library(str(tmp_path / "nope.so"))
        finally:
            runtime._lib = None


@pytest.mark.gpu
def test_reupload_after_fused_binning_restarts_cleanly():
    """A step kernel that rewrites positions also produces the histogram of the next binning;
    a fresh upload must discard it (regression: stale counts corrupted cell_start)."""
    params = {"num_agents": 100000}
    m = Model(os.path.join(REPO, "examples", "boids2d.abl"), params)
    m.populate()
    m.create_runtime()
    runs = []
    for _ in range(2):
        m.upload_host()
        for _ in range(3):
            m.timestep()
        runs.append(m.download(0))
    m.close()
    for f in runs[0].dtype.names:
        assert np.array_equal(runs[0][f], runs[1][f])
