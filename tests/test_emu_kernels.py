"""CPU: the GENERATED step kernels, executed by the sequential CUDA emulator (tests/emu/).

The device program the code generator prints for a model is compiled for the host and run one
simulated thread after the other; binning is redone in numpy.  That covers, without a GPU,
what the kernels compute: candidate order (cell key, then id), the square-root-free radius
filter, the loop body, the stores and the fused histogram of the next binning — for every
neighbour-loop variant whose threads do not cooperate (global loop, chunked two-phase loop,
unrolled loop, flat loop).  Bars are those of the GPU tests: bit-equal to the grid-ordered
oracle, and within the north-star tolerances of the REAL reference's golden vectors.
"""
import os

import numpy as np
import pytest

import refgen
from emu.emu import EmuModel
from oracle import GRID, Oracle
from openabl_b200.state import exact_members_equal, max_rel_error

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (model, params, use_float, steps) — builds shared with tests/test_gpu_parity_oracle.py
CASES = [
    ("circle.abl", {"num_agents": 50000}, False, 3),
    ("circle3d.abl", {"num_agents": 20000}, False, 2),        # dense: chunked two-phase kernel
    ("boids2d.abl", {"num_agents": 100000}, False, 4),
    ("boids2d.abl", {"num_agents": 100000}, True, 4),
    ("game_of_life.abl", {"num_agents": 65536}, False, 5),
    ("game_of_life.abl", {"num_agents": 65536}, True, 5),
]
IDS = ["%s-%d-%s" % (c[0][:-4], c[1]["num_agents"], "f32" if c[2] else "f64") for c in CASES]
VARIANTS = [None, {"cuda.unroll": True}, {"cuda.flat": True}, {"cuda.sqcmp": True},
            {"cuda.flat": True, "cuda.sqcmp": True}]
VARIANT_IDS = ["default", "unroll", "flat", "sqcmp", "flat+sqcmp"]


def emulate(model_path, params, use_float, steps, config=None, block_size=0):
    m = EmuModel(model_path, params, use_float=use_float, config=config)
    m.block_size = block_size
    m.populate()
    init = [m.host_agents(t) for t in range(m.n_types)]
    for _ in range(steps):
        m.timestep()
    return m, init, [m.host_agents(t) for t in range(m.n_types)]


@pytest.mark.parametrize("config", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("model_file,params,use_float,steps", CASES, ids=IDS)
def test_generated_kernels_equal_grid_oracle(model_file, params, use_float, steps, config):
    m, init, got = emulate(os.path.join(REPO, "examples", model_file), params, use_float, steps, config)
    o = Oracle(use_float)
    state = o.init_for(model_file, params)
    for f in state.dtype.names:
        assert np.array_equal(init[0][f], state[f]), "initial %s differs" % f
    want = o.run_for(model_file, params, state, steps, GRID)
    assert len(got[0]) == len(want)
    assert exact_members_equal(got[0], want), "integer/bool state differs"
    # same arithmetic in the same order (x86-64 SSE2 without contraction == -fmad=false device
    # code for + - * / sqrt): bit-equal in double; use_float keeps the GPU tests' 1e-4 because the
    # oracle follows the reference in evaluating unsuffixed literals in double
    if use_float:
        assert max_rel_error(got[0], want) <= 1e-4
    else:
        for f in want.dtype.names:
            assert np.array_equal(got[0][f], want[f]), "member %s is not bit-equal (max rel err %.3e)" % (
                f, max_rel_error(got[0], want))
    if model_file != "game_of_life.abl":
        assert m.fused_checked == steps     # the fused histogram epilogue was verified every step


RUNS = [n for n, (_, p, _) in refgen.FIXTURES.items() if p["num_timesteps"] == 10]
EXTRA_RUNS = [n for n, (_, p, _) in refgen.EXTRA_FIXTURES.items() if p["num_timesteps"] == 10]


@pytest.mark.parametrize("config", [None, {"cuda.flat": True, "cuda.sqcmp": True}], ids=["default", "flat+sqcmp"])
@pytest.mark.parametrize("name", RUNS + EXTRA_RUNS)
def test_generated_kernels_match_reference_c_backend(name, config):
    """Golden vectors of the unmodified reference compiler + libabl (oracle/refgen.py), 10 steps:
    counts and integer/bool state exact, positions within 1e-9 (double) / 1e-4 (use_float)."""
    info, gold = refgen.load_fixture(name)
    params = dict(info["params"])
    m, _, got = emulate(refgen.model_path(info["model"]), params, info["use_float"], params["num_timesteps"], config)
    tol = 1e-4 if info["use_float"] else 1e-9
    for g, ref in zip(got, gold):
        assert len(g) == len(ref), "agent count differs"
        assert exact_members_equal(g, ref), "integer/bool state differs"
        err = max_rel_error(g, ref)
        assert err <= tol, "max relative error %.3e > %.1e" % (err, tol)


@pytest.mark.parametrize("block_size", [32, 96, 256])
def test_block_size_does_not_change_results(block_size):
    params = {"num_agents": 100000}
    path = os.path.join(REPO, "examples", "boids2d.abl")
    _, _, a = emulate(path, params, False, 2)
    _, _, b = emulate(path, params, False, 2, block_size=block_size)
    for f in a[0].dtype.names:
        assert np.array_equal(a[0][f], b[0][f])
