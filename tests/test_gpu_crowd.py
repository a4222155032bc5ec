"""Binning of populations the reference's brute-force loop handles like any other but a uniform grid does not:
thousands of agents in ONE cell and agents beyond the environment's bounds (clamped into edge cells).  Cells of
more than 128 agents are ranked by whole warps (k_bin_rank_move, asset/cuda/abl_runtime.cu); the result must be
the order every other part of the system assumes — ascending cell key, ascending id inside a cell — and the
records must come back untouched."""
import os

import numpy as np
import pytest

from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expected(pos, axes, cell):
    inv = 1.0 / cell
    key = np.zeros(len(pos), dtype=np.int64)
    mul = 1
    for a in range(2):
        c = np.clip(np.floor((pos[:, a] - 0.0) * inv), 0, axes[a] - 1).astype(np.int64)
        key += c * mul
        mul *= axes[a]
    ids = np.arange(len(pos), dtype=np.uint32)
    order = np.lexsort((ids, key))
    return key, ids[order]


@pytest.mark.gpu
@pytest.mark.parametrize("crowd", [200, 5000, 60000])
def test_crowded_cells_are_binned_in_cell_and_id_order(crowd):
    n = 100000
    m = Model(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": 50000})   # W = 1000, cells of 10
    m.populate()
    rng = np.random.default_rng(7)
    rec = np.zeros(n, dtype=m.dtypes[0])
    pos = rng.uniform(0.0, 1000.0, size=(n, 2))
    pos[:crowd] = rng.uniform(501.0, 509.0, size=(crowd, 2))                   # one cell (cells are 10 * (1 + 2^-20) wide)
    pos[crowd:crowd + 3000] = rng.uniform(-400.0, -1.0, size=(3000, 2))         # beyond the bounds: cell (0, 0)
    pos[crowd + 3000:crowd + 4000, 0] = rng.uniform(1001.0, 5000.0, size=1000)  # beyond the upper bound in x
    rng.shuffle(pos)
    rec["pos"] = pos
    m.create_runtime()
    m.upload(0, rec)
    m.rt.bin(m.pool(0))
    n_cells, axes = m.rt.grid_cells()
    cs, ids = m.rt.debug_binning(m.pool(0))
    key, want = _expected(pos, axes, 10.0 * (1.0 + 2.0 ** -20))
    counts = np.bincount(key, minlength=n_cells)
    assert counts.max() >= crowd
    assert np.array_equal(cs[1:n_cells + 1], np.cumsum(counts))
    assert np.array_equal(ids, want), "pool is not in (cell key, id) order"
    back = m.download(0)
    assert np.array_equal(back["pos"], pos)
    # two steps on the crowd: the neighbour loop sees the same candidates whichever block size steps them
    for _ in range(2):
        m.timestep()
    a = m.download(0)
    m.close()
    m2 = Model(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": 50000})
    m2.populate()
    m2.create_runtime(block_size=64)
    m2.upload(0, rec)
    for _ in range(2):
        m2.timestep()
    b = m2.download(0)
    m2.close()
    assert np.array_equal(a["pos"], b["pos"])
    assert np.isfinite(a["pos"]).all()
