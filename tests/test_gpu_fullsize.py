"""BASELINE-size consistency checks between the kernel variants: the shared-memory tile path, the plain
global-memory loop, the TMA-staged tile, the flat / cursor / timed choices and every block size must give
bit-identical results at 1 M agents; populations are conserved; boids stay inside their periodic world.
(Parity with the ORACLE at these sizes — boids2d 1 M, circle3d 1 M, game_of_life 16.7 M, predator_prey 4 M,
single GPU and slabs — is in tests/test_gpu_parity_fullsize.py.)"""
import os

import numpy as np
import pytest

from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(model_file, params, use_float, steps, **rt_kw):
    m = Model(os.path.join(REPO, "examples", model_file), params, use_float=use_float)
    m.populate()
    n = m.host_count(0)
    m.create_runtime(**rt_kw)
    m.upload_host()
    for _ in range(steps):
        m.timestep()
    out = m.download(0)
    m.close()
    return n, out


@pytest.mark.gpu
@pytest.mark.parametrize("use_float", [False, True], ids=["f64", "f32"])
def test_boids2d_1M_tile_and_plain_paths_agree_bitwise(use_float):
    params = {"num_agents": 1000000}
    n, tiled = run("boids2d.abl", params, use_float, 10, tile=True)
    _, plain = run("boids2d.abl", params, use_float, 10, tile=False)
    _, wide = run("boids2d.abl", params, use_float, 10, tile=True, block_size=256)
    assert len(tiled) == n == 1000000
    for f in tiled.dtype.names:
        assert np.array_equal(tiled[f], plain[f]), "tile path differs from the global-memory path in %s" % f
        assert np.array_equal(tiled[f], wide[f]), "result depends on the block size in %s" % f
    pos = tiled["pos"]
    assert np.isfinite(pos).all()
    assert pos.min() >= 0.0 and pos.max() <= 44.7214 + 1e-6   # boundPosition keeps boids inside [0, max_pos]
    speed = np.sqrt((tiled["velocity"].astype(np.float64) ** 2).sum(axis=1))
    assert speed.max() <= 1.0 + 1e-5                          # velocities are clipped to unit length


@pytest.mark.gpu
def test_circle3d_1M_is_independent_of_block_size_and_tiling():
    params = {"num_agents": 1000000}
    n, a = run("circle3d.abl", params, False, 2)
    _, b = run("circle3d.abl", params, False, 2, block_size=128, tile=False)
    assert len(a) == n == 1000000
    assert np.array_equal(a["pos"], b["pos"])
    assert np.isfinite(a["pos"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("model_file,n,steps", [("boids2d.abl", 1000000, 14), ("game_of_life.abl", 1048576, 14)],
                         ids=["boids2d-1M", "game_of_life-1M"])
def test_candidate_loop_variants_agree_bitwise_at_full_size(model_file, n, steps, monkeypatch):
    """Cursor loop (ABL_CUDA_FLAT=0), flat loop over global memory (ABL_CUDA_BULK=0), the default (flat
    loop over a TMA-staged shared-memory tile, ABL_MODE 7) and the timed choice (ABL_CUDA_TUNE=1: the
    first twelve launches rotate through cursor, flat and chunked loop) must give the same bits; 14
    timesteps take the timed run through all its trial launches and into the chosen variant."""
    params = {"num_agents": n}
    outs = {}
    settings = {"cursor": {"ABL_CUDA_FLAT": "0"}, "flat": {"ABL_CUDA_BULK": "0"}, "bulk": {}, "timed": {"ABL_CUDA_TUNE": "1"}}
    for name, env in settings.items():
        for k in ("ABL_CUDA_FLAT", "ABL_CUDA_BULK", "ABL_CUDA_TUNE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        outs[name] = run(model_file, params, False, steps)[1]
    for f in outs["cursor"].dtype.names:
        for name in ("flat", "bulk", "timed"):
            assert np.array_equal(outs["cursor"][f], outs[name][f]), "%s run differs from the cursor loop in %s" % (name, f)
