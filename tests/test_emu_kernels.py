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
from openabl_b200.state import F32_FLOOR_ULPS, exact_members_equal, max_rel_error

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
# (-C options, abl_step_launch.flat_loop).  Defaults of the code generator: cuda.sqcmp=true (distance
# comparisons on the squared distance) and cuda.flat=true (the flat loop is generated; whether a
# launch uses it is the runtime's flat_loop setting: 0 cursor loop, 1 flat loop, -1 timed).
VARIANTS = [(None, 0), (None, 1), ({"cuda.unroll": True}, 0), ({"cuda.sqcmp": False}, 0), ({"cuda.sqcmp": False}, 1),
            ({"cuda.flat": False, "cuda.sqcmp": False}, 1)]
VARIANT_IDS = ["cursor", "flat", "unroll", "cursor-sqrt", "flat-sqrt", "plain"]


def emulate(model_path, params, use_float, steps, config=None, block_size=0, flat_loop=0, costs=None):
    m = EmuModel(model_path, params, use_float=use_float, config=config)
    m.block_size = block_size
    m.flat_loop = flat_loop
    for mode in range(8):
        m.lib.emu_set_cost(mode, (costs or {}).get(mode, 1.0))
    m.populate()
    init = [m.host_agents(t) for t in range(m.n_types)]
    for _ in range(steps):
        m.timestep()
    return m, init, [m.host_agents(t) for t in range(m.n_types)]


@pytest.mark.parametrize("config,flat_loop", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("model_file,params,use_float,steps", CASES, ids=IDS)
def test_generated_kernels_equal_grid_oracle(model_file, params, use_float, steps, config, flat_loop):
    m, init, got = emulate(os.path.join(REPO, "examples", model_file), params, use_float, steps, config,
                           flat_loop=flat_loop)
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
        assert max_rel_error(got[0], want, floor_ulps=F32_FLOOR_ULPS) <= 1e-4
    else:
        for f in want.dtype.names:
            assert np.array_equal(got[0][f], want[f]), "member %s is not bit-equal (max rel err %.3e)" % (
                f, max_rel_error(got[0], want))
    if model_file != "game_of_life.abl":
        assert m.fused_checked == steps     # the fused histogram epilogue was verified every step


RUNS = [n for n, (_, p, _) in refgen.FIXTURES.items() if p["num_timesteps"] == 10]
EXTRA_RUNS = [n for n, (_, p, _) in refgen.EXTRA_FIXTURES.items() if p["num_timesteps"] == 10]


@pytest.mark.parametrize("flat_loop", [0, 1, -1], ids=["cursor", "flat", "timed"])
@pytest.mark.parametrize("name", RUNS + EXTRA_RUNS)
def test_generated_kernels_match_reference_c_backend(name, flat_loop):
    """Golden vectors of the unmodified reference compiler + libabl (oracle/refgen.py), 10 steps:
    counts and integer/bool state exact, positions within 1e-9 (double) / 1e-4 (use_float)."""
    info, gold = refgen.load_fixture(name)
    params = dict(info["params"])
    m, _, got = emulate(refgen.model_path(info["model"]), params, info["use_float"], params["num_timesteps"],
                        flat_loop=flat_loop)
    tol = 1e-4 if info["use_float"] else 1e-9
    for g, ref in zip(got, gold):
        assert len(g) == len(ref), "agent count differs"
        assert exact_members_equal(g, ref), "integer/bool state differs"
        err = max_rel_error(g, ref, floor_ulps=F32_FLOOR_ULPS if info["use_float"] else 0)
        assert err <= tol, "max relative error %.3e > %.1e" % (err, tol)


@pytest.mark.parametrize("block_size", [32, 96, 256])
def test_block_size_does_not_change_results(block_size):
    params = {"num_agents": 100000}
    path = os.path.join(REPO, "examples", "boids2d.abl")
    _, _, a = emulate(path, params, False, 2)
    _, _, b = emulate(path, params, False, 2, block_size=block_size)
    for f in a[0].dtype.names:
        assert np.array_equal(a[0][f], b[0][f])


def test_launchers_pick_the_expected_kernel_variant():
    """The emulated launch goes through the generated launcher: sparse 2-D neighbourhoods run the
    cursor loop (ABL_MODE 0) or the flat loop (3) as the runtime asks; dense ones the chunked
    two-phase loop (1); 3-D models never the flat loop."""
    from emu.emu import modes
    b = os.path.join(REPO, "examples", "boids2d.abl")
    c3 = os.path.join(REPO, "examples", "circle3d.abl")
    assert modes(emulate(b, {"num_agents": 100000}, False, 1)[0].kernels) == [0]
    assert modes(emulate(b, {"num_agents": 100000}, False, 1, flat_loop=1)[0].kernels) == [3]
    assert modes(emulate(b, {"num_agents": 100000}, False, 1, {"cuda.flat": False}, flat_loop=1)[0].kernels) == [0]
    assert modes(emulate(c3, {"num_agents": 20000}, False, 1)[0].kernels) == [1]
    assert modes(emulate(c3, {"num_agents": 20000}, False, 1, flat_loop=1)[0].kernels) == [1]


@pytest.mark.parametrize("faster", [0, 3, 1])
def test_run_time_tuner_times_the_variants_and_keeps_the_fastest(faster):
    """flat_loop = -1: the first twelve launches rotate through the cursor loop, the flat loop and
    the chunked loop (events on the launch stream; here a simulated clock), afterwards every launch
    uses the variant with the smallest time.  Results are those of any fixed choice, bit for bit."""
    import shutil
    import tempfile
    from emu.emu import modes
    params = {"num_agents": 4000, "num_timesteps": 10}
    path = os.path.join(REPO, "examples", "boids2d.abl")
    # (the tuner is static in the launcher: a private copy of the library gives this test its own)
    m = EmuModel(path, params)
    with tempfile.TemporaryDirectory() as tmp:
        private = os.path.join(tmp, "libmodel_emu_%d.so" % faster)
        shutil.copy(os.path.join(m.dir, "libmodel_emu.so"), private)
        m = EmuModel(path, params, lib_path=private)
        for mode in range(8):
            m.lib.emu_set_cost(mode, 1.0 if mode == faster else 2.0)
        m.flat_loop = -1
        m.populate()
        seen = []
        for _ in range(16):
            m.kernels = set()
            m.timestep()
            seen += modes(m.kernels)
        assert seen[:12] == [0, 3, 1] * 4
        assert seen[12:] == [faster] * 4
        assert m.lib.abl_model_step_variant(0) == faster      # what bench.py reports as candidate_loop_in_use
        _, _, want = emulate(path, params, False, 16)
        got = m.host_agents(0)
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[0][f])


def test_tuner_only_compares_plausible_variants():
    """Crowded rows (circle3d: 49 agents per cell) go straight to the chunked loop; nearly empty
    rows (predators: 0.04 per cell) are never tried on it."""
    from emu.emu import modes
    m, _, _ = emulate(os.path.join(REPO, "examples", "circle3d.abl"), {"num_agents": 20000}, False, 2, flat_loop=-1)
    assert modes(m.kernels) == [1]
    m, _, _ = emulate(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": 50000}, False, 3, flat_loop=-1)
    assert 0 in modes(m.kernels) and 3 in modes(m.kernels)


@pytest.mark.parametrize("use_float", [False, True], ids=["f64", "f32"])
def test_squared_distance_bounds_decide_like_the_square_root(use_float, tmp_path):
    """abl_near_sq_limit / abl_sq_cmp_limit (host code of the generated launchers): the bound found
    by bisection decides `s <= limit` / `s >= limit` exactly like `(abl_real)sqrtf((float)s) <op> C`
    for every representable s within 64 ulps of the bound, for special values (0, inf, NaN,
    denormals) and for 13 M random s, over 16 constants and the four operators."""
    import subprocess
    here = os.path.join(REPO, "tests", "emu")
    exe = str(tmp_path / "sq_limit_check")
    cmd = ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-w", "-I", here, "-I", os.path.join(REPO, "include"),
           "-I", os.path.join(REPO, "asset", "cuda"), os.path.join(here, "sq_limit_check.cpp"), "-o", exe, "-ldl"]
    if use_float:
        cmd.insert(1, "-DABL_USE_FLOAT")
    subprocess.run(cmd, check=True)
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout
    assert " 0 failures" in out.stdout


@pytest.mark.parametrize("flat_loop", [0, 1], ids=["cursor", "flat"])
def test_predator_prey_kernels_match_frozen_semantics(flat_loop):
    """BASELINE configs[3] on the CPU: the generated kernels of predator_prey (three agent types,
    eleven step functions, removeCurrent(), add(), in-step random numbers) under the emulator,
    with the commit rules of the runtime restated in numpy, against the frozen-semantics oracle:
    agent counts after every timestep, ids of run-time-added agents and every member, bit for bit."""
    from oracle import PredatorPreyOracle
    n, steps = 32000, 12
    m = EmuModel(os.path.join(REPO, "examples", "predator_prey.abl"), {"num_agents": n})
    m.flat_loop = flat_loop
    m.populate()
    o = PredatorPreyOracle(n)
    for t in range(3):
        _, rec = o.read(t)
        host = m.host_agents(t)
        assert len(host) == len(rec)
        for f in host.dtype.names:
            assert np.array_equal(host[f], rec[f]), "initial %s differs" % f
    initial = [len(m.pools[t].ids) for t in range(3)]
    removed_any = added_any = False
    for step in range(steps):
        before = [o.count(t) for t in range(3)]
        m.timestep()
        o.timestep(GRID)
        counts = [len(m.pools[t].ids) for t in range(3)]
        assert counts == [o.count(t) for t in range(3)], "agent counts differ after timestep %d" % step
        removed_any |= counts[1] < before[1]
    for t in range(3):
        ids, want = o.read(t)
        got = m.host_agents(t)
        assert np.array_equal(np.sort(m.pools[t].ids), ids), "agent ids differ"
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[f]), "type %d member %s differs" % (t, f)
        added_any |= bool(len(ids) and ids.max() >= initial[t])
    o.close()
    assert removed_any and added_any
    from emu.emu import modes
    assert (3 in modes(m.kernels)) == bool(flat_loop)


REF_EXAMPLES = "/root/reference/examples"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference tree not present")


@needs_ref
@pytest.mark.parametrize("example,params,steps", [
    ("ants.abl", {"num_agents": 3000}, 6),                # two agent types reading each other, in-step RNG
    ("sugarscape.abl", {"num_agents": 3000}, 6),          # four chained step functions, partial `out` writes
    ("boids2d_flockers.abl", {"num_agents": 3000}, 6),    # in-step random(), wraparound
    ("boids.abl", {"num_agents": 3000}, 4),               # 3-D: no flat loop, but squared-distance comparisons
    ("keratinocyte.abl", {}, 3),                          # constant tables, while, count(member, value)
])
def test_loop_variants_agree_on_the_reference_examples(example, params, steps):
    """Differential check on the models that have no oracle: the kernels printed with every
    optimisation off (cursor loop, distances through the square root) and the default kernels
    (flat loop where it applies, comparisons on the squared distance) give identical state —
    every agent type, every member, bit for bit."""
    path = os.path.join(REF_EXAMPLES, example)
    _, _, plain = emulate(path, params, False, steps, {"cuda.flat": False, "cuda.sqcmp": False}, flat_loop=0)
    m, _, tuned = emulate(path, params, False, steps, None, flat_loop=1)
    assert len(plain) == len(tuned)
    for a, b in zip(plain, tuned):
        assert len(a) == len(b)
        for f in (a.dtype.names or []):
            assert np.array_equal(a[f], b[f], equal_nan=True), "%s: member %s differs" % (example, f)


@pytest.mark.parametrize("use_float", [False, True], ids=["f64", "f32"])
def test_cached_neighbour_lists_equal_grid_oracle(use_float):
    """-C cuda.nlist=true on game_of_life (no step function moves a cell): the count pass, the fill
    pass and the list-walking kernel (ABL_MODE 5, 6, 4) against the grid-ordered oracle; the lists
    are built once and reused for every timestep."""
    from emu.emu import modes
    params = {"num_agents": 65536}
    m = EmuModel(os.path.join(REPO, "examples", "game_of_life.abl"), params, use_float=use_float,
                 config={"cuda.nlist": True})
    m.use_nlist = True
    m.populate()
    for _ in range(5):
        m.timestep()
    got = m.host_agents(0)
    o = Oracle(use_float)
    want = o.run_for("game_of_life.abl", params, o.init_for("game_of_life.abl", params), 5, GRID)
    assert np.array_equal(got["alive"], want["alive"]) and np.array_equal(got["pos"], want["pos"])
    assert m.nlist_builds == 1
    assert modes(m.kernels) == [4, 5, 6]


def test_neighbour_lists_are_only_offered_for_static_neighbourhoods():
    """boids2d moves its agents, predator_prey's steps remove/add or read moving agents: the code
    generator must not mark their steps; game_of_life without the option neither."""
    for model, cfg in [("boids2d.abl", {"cuda.nlist": True}), ("predator_prey.abl", {"cuda.nlist": True}),
                       ("circle3d.abl", {"cuda.nlist": True}), ("game_of_life.abl", None)]:
        m = EmuModel(os.path.join(REPO, "examples", model), {"num_agents": 4096}, config=cfg)
        assert not any(m.lib.emu_step_nlist(s) for s in range(m.n_steps)), model


def test_neighbour_lists_on_a_model_with_static_and_moving_types():
    """tests/models/static_sites.abl: Sites never move (their diffusion step reads neighbouring
    Sites through dist(), a float and a bool member, with `continue` and `break`), Walkers do.
    Only the Site-Site step is marked; its lists are built once although the other pools change
    every timestep; the state equals the plain kernels' bit for bit."""
    from emu.emu import modes
    path = os.path.join(REPO, "tests", "models", "static_sites.abl")
    _, _, plain = emulate(path, {"num_agents": 3000}, False, 5, {"cuda.flat": False, "cuda.sqcmp": False})
    m = EmuModel(path, {"num_agents": 3000}, config={"cuda.nlist": True})
    m.use_nlist = True
    m.populate()
    for _ in range(5):
        m.timestep()
    assert [m.lib.emu_step_nlist(s) for s in range(m.n_steps)] == [1, 0, 0]
    assert m.nlist_builds == 1 and {4, 5, 6} <= set(modes(m.kernels))
    for t in range(m.n_types):
        got = m.host_agents(t)
        for f in got.dtype.names:
            assert np.array_equal(got[f], plain[t][f]), "type %d member %s differs" % (t, f)
    assert plain[0]["visitors"].sum() > 0 and not np.array_equal(plain[0]["heat"], emulate(path, {"num_agents": 3000}, False, 0)[2][0]["heat"])


@pytest.mark.parametrize("me", [0, 1, 3])
def test_fused_halo_send_packs_what_the_decomposition_rules_say(me):
    """Slab decomposition, direct transport: the step kernel itself routes every owned agent by
    the cell layer of its NEW position and appends the record to the outgoing messages
    (abl_slab_epilogue / abl_slab_send / abl_slab_route).  Emulated for one of four slabs and
    compared with the rules restated in numpy (openabl_b200.slab.send_masks, the ones the gloo
    test runs): same agents per direction, word-major records equal to the stored columns."""
    import ctypes as C
    from openabl_b200.slab import send_masks, slab_layer, split_layers
    params = {"num_agents": 100000}
    m = EmuModel(os.path.join(REPO, "examples", "boids2d.abl"), params)
    m.populate()
    m.bin(0)
    g, p = m.grid, m.pools[0]
    n_layers = g.n_cell[1]
    bounds = split_layers(n_layers, 4)
    pairs = [(bounds[k], bounds[k + 1]) for k in range(4)] if not isinstance(bounds[0], (tuple, list)) else list(bounds)
    begin, end = pairs[me]
    layer_now = slab_layer(p.position()[1], g.origin[1], g.cell_size, n_layers)
    owned = np.nonzero((layer_now >= begin) & (layer_now < end))[0]
    ob, oe = int(owned[0]), int(owned[-1]) + 1
    assert oe - ob == len(owned)                      # sorted by cell key: the owned agents are contiguous
    N = 4
    lo = me - 1 if me > 0 else N - 1                  # peers form a ring (halo_fill_view)
    hi = me + 1 if me < N - 1 else 0
    capacity, rec_words = 20000, 4 + 4 + 1            # pos + velocity (2 doubles each) + id
    msgs = [np.zeros(rec_words * capacity, dtype=np.uint32) for _ in range(2)]
    counters = np.zeros(8, dtype=np.uint32)
    elem = (C.c_int * 2)(16, 16)
    m.lib.emu_set_slab.argtypes = [C.c_int] * 4 + [C.c_double] * 2 + [C.c_int] * 9 + [C.c_void_p] * 3 + [C.c_uint] * 2 + \
        [C.c_int, C.c_void_p]
    m.lib.emu_set_slab(1, 2, 0, n_layers, g.origin[1], g.inv_cell_size, begin, end, 1,
                       pairs[lo][0], pairs[lo][1], pairs[hi][0], pairs[hi][1], int(me > 0), int(me < N - 1),
                       msgs[0].ctypes.data, msgs[1].ctypes.data, counters.ctypes.data, capacity, rec_words, 2, elem)
    try:
        m.run_step(0, own_range=(ob, oe))
    finally:
        m.lib.emu_set_slab(*([0] * 4 + [0.0, 0.0] + [0] * 9 + [None] * 3 + [0, 0, 0, None]))
    new_pos, new_vel, ids = p.cols[0][ob:oe], p.cols[1][ob:oe], p.ids[ob:oe]
    layer_new = slab_layer(new_pos[:, 1], g.origin[1], g.cell_size, n_layers)
    want_lo, want_hi = send_masks(layer_new, pairs, me, 1)
    assert counters[2] == 0 and counters[3] == 0      # nobody moved farther than a neighbouring slab
    assert want_lo.sum() > 0 or want_hi.sum() > 0
    if me == 1:
        assert want_lo.sum() > 100 and want_hi.sum() > 100      # ghost copies for both true neighbours
    for d, want in enumerate((want_lo, want_hi)):
        cnt = int(counters[d])
        assert cnt == int(want.sum()), "direction %d: %d records, rules say %d" % (d, cnt, want.sum())
        words = msgs[d].reshape(rec_words, capacity)[:, :cnt]
        got_ids = words[8]
        order = np.argsort(got_ids)
        assert np.array_equal(got_ids[order], np.sort(ids[want]))
        rec = np.ascontiguousarray(words[:8, order].T).view(np.float64)      # pos.x pos.y vel.x vel.y
        sel = np.nonzero(want)[0][np.argsort(ids[want])]
        assert np.array_equal(rec[:, 0:2], new_pos[sel]) and np.array_equal(rec[:, 2:4], new_vel[sel])


@pytest.mark.parametrize("use_float", [False, True], ids=["f64", "f32"])
@pytest.mark.parametrize("outside", [False, True], ids=["inside", "beyond-the-bounds"])
def test_cell_range_culling_keeps_every_candidate_on_the_radius(outside, use_float):
    """Cell-range culling (radius well below the cell size: only the cells [cell(p - R), cell(p + R)]
    per axis are visited).  static_sites counts Walkers within 1.0 of each Site on a grid of 2.5:
    Walkers are placed exactly ON the radius around Sites that sit on / next to cell borders, in
    all directions and one ulp inside / outside; the kernel's counts must equal a brute-force
    count with the reference's predicate, and the culled kernel must visit fewer candidates."""
    import ctypes as C
    path = os.path.join(REPO, "tests", "models", "static_sites.abl")
    rng = np.random.default_rng(7)
    R, cell, W = 1.0, 2.5, float(np.sqrt(np.float64(3000) / 0.5))
    n_sites = 1500
    k = rng.integers(2, 28, size=(n_sites, 2)).astype(np.float64)
    off = rng.choice([0.0, 1e-13, -1e-13, 0.4, 1.0, 1.5, 2.4999999999], size=(n_sites, 2))
    site_pos = np.clip(k * cell + off, 0.0, W)
    if outside:
        # a third of the Sites beyond the environment (clamped into the border cells by the binning)
        site_pos[:500, 0] = rng.choice([-0.5, -3.0, W + 0.25, W + 7.0, 0.0, W], size=500)
        site_pos[250:750, 1] = rng.choice([-0.75, -1e-9, W + 1e-9, W + 2.0, 0.0, W], size=500)
    theta = np.concatenate([np.arange(8) * np.pi / 4, rng.uniform(0, 2 * np.pi, 4)])
    walkers = []
    for s in site_pos[:600]:
        for t in theta:
            q = s + R * np.array([np.cos(t), np.sin(t)])
            walkers += [q, np.nextafter(q, s), np.nextafter(q, q + (q - s))]
    walker_pos = np.array(walkers) if outside else np.clip(np.array(walkers), 0.0, W)
    real = np.float32 if use_float else np.float64
    site_pos, walker_pos = site_pos.astype(real), walker_pos.astype(real)
    if use_float:
        # one float ulp either side of the rounded positions as well
        extra = walker_pos[::3]
        walker_pos = np.concatenate([walker_pos, np.nextafter(extra, np.float32(0)), np.nextafter(extra, np.float32(1e9))])

    def run(cfg):
        m = EmuModel(path, {"num_agents": 3000}, use_float=use_float, config=cfg)
        m.lib.emu_load_count.restype = C.c_ulonglong
        m.populate()
        sites = np.zeros(n_sites, dtype=m.dtypes[0]); sites["pos"] = site_pos; sites["heat"] = 1.0
        wk = np.zeros(len(walker_pos), dtype=m.dtypes[1]); wk["pos"] = walker_pos
        m.pools[0].load(sites); m.pools[1].load(wk)
        before = m.lib.emu_load_count()
        m.run_step(1)                                  # site_count_walkers
        return m.host_agents(0)["visitors"], m.lib.emu_load_count() - before

    got, loads = run(None)
    plain, loads_plain = run({"cuda.cull": False})
    rowcull, loads_rowcull = run({"cuda.cull": False, "cuda.rowcull": True})     # rows narrowed along x to the cells within reach (dense rows)
    assert np.array_equal(rowcull, plain) and loads_rowcull < loads_plain
    # the reference's filter (CPrinter.cpp:166-169, libabl.h:156-168): skip if (double)sqrtf((float)d2) > R
    dx = walker_pos[None, :, 0] - site_pos[:, None, 0]
    dy = walker_pos[None, :, 1] - site_pos[:, None, 1]
    d2 = dx * dx + dy * dy
    want = (~(np.sqrt(d2.astype(np.float32)).astype(real) > real(R))).sum(axis=1)     # dx, dy, d2 in abl_float
    assert np.array_equal(plain, want)
    assert np.array_equal(got, want)
    assert want.max() >= 20 and loads < (1.0 if outside else 0.6) * loads_plain


# ---- the GPU edge cases (tests/test_gpu_edge_cases.py) under the emulator, for every loop choice ----
@pytest.mark.parametrize("flat_loop", [0, 1, -1], ids=["cursor", "flat", "timed"])
def test_edge_reach_two_cells_matches_oracle(flat_loop):
    """granularity 5 < radius 10: the on-demand row path of the iterator (no flat loop, no culling)."""
    from oracle import BRUTE
    n, steps = 20000, 4
    m, _, got = emulate(os.path.join(REPO, "tests", "models", "circle_fine_grid.abl"), {"num_agents": n}, False, steps,
                        flat_loop=flat_loop)
    o = Oracle(False)
    state = o.circle_init(2, n)
    want = o.circle_run(2, state, steps, GRID, granularity=5.0)
    assert np.array_equal(got[0]["pos"], want["pos"])
    ref = o.circle_run(2, state, steps, BRUTE)
    assert np.max(np.abs(got[0]["pos"] - ref["pos"]) / np.maximum(np.abs(ref["pos"]), 1.0)) <= 1e-9


@pytest.mark.parametrize("flat_loop", [0, 1, -1], ids=["cursor", "flat", "timed"])
@pytest.mark.parametrize("n", [0, 1, 2, 3])
def test_edge_empty_and_tiny_populations(n, flat_loop):
    m = EmuModel(os.path.join(REPO, "examples", "boids2d.abl"), {"num_agents": 100000})
    m.flat_loop = flat_loop
    m.populate()
    sub = np.ascontiguousarray(m.host_agents(0)[:n])
    m.pools[0].load(sub)
    for _ in range(3):
        m.timestep()
    got = m.host_agents(0)
    assert len(got) == n
    if n:
        want = Oracle(False).boids_run(sub, 3, GRID, num_agents=100000)
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[f])


@pytest.mark.parametrize("flat_loop", [0, 1, -1], ids=["cursor", "flat", "timed"])
def test_edge_one_cell_crowd_and_agents_outside_the_environment(flat_loop):
    m = EmuModel(os.path.join(REPO, "examples", "boids2d.abl"), {"num_agents": 100000})
    m.flat_loop = flat_loop
    m.populate()
    rng = np.random.default_rng(7)
    n = 6000
    a = np.zeros(n, dtype=m.dtypes[0])
    a["pos"][:5000] = 5.0 + rng.random((5000, 2)) * 0.04          # one 0.05 x 0.05 cell
    a["pos"][5000:5500] = -3.0 + rng.random((500, 2))             # below the lower bound
    a["pos"][5500:] = 14.2 + rng.random((500, 2)) * 5.0           # beyond max_pos (= 14.14)
    a["velocity"] = rng.random((n, 2)) * 2 - 1
    m.pools[0].load(a)
    for _ in range(2):
        m.timestep()
    got = m.host_agents(0)
    want = Oracle(False).boids_run(a, 2, GRID, num_agents=100000)
    for f in got.dtype.names:
        assert np.array_equal(got[f], want[f]), f


@pytest.mark.parametrize("flat_loop", [1, -1], ids=["flat", "timed"])
def test_break_and_continue_inside_the_loop_body_under_every_variant(flat_loop):
    """static_sites.site_diffuse leaves its for-near loop with `break` after twelve neighbours and
    skips itself with `continue`: the flat loop (two candidates per iteration, `continue` of the
    first must still reach the second, `break` of the first must not) and the chunked loop give
    the plain cursor loop's state bit for bit."""
    path = os.path.join(REPO, "tests", "models", "static_sites.abl")
    _, _, plain = emulate(path, {"num_agents": 3000}, False, 5, {"cuda.flat": False, "cuda.sqcmp": False, "cuda.cull": False})
    m, _, got = emulate(path, {"num_agents": 3000}, False, 5, None, flat_loop=flat_loop)
    from emu.emu import modes
    assert 3 in modes(m.kernels)
    for a, b in zip(plain, got):
        for f in a.dtype.names:
            assert np.array_equal(a[f], b[f]), "member %s differs" % f


@pytest.mark.parametrize("use_float", [False, True], ids=["f64", "f32"])
def test_neighbour_lists_and_culling_match_the_real_reference(use_float):
    """tests/models/sites_walkers.abl against the golden vectors of the unmodified reference compiler
    + libabl (10 timesteps): radii below the cell size (culled cell ranges), a distance compared
    with a constant, and — with cuda.nlist — cached lists for the static Sites.  Counts and
    integer/bool state exact, floating point within 1e-9 (double) / 1e-4 (use_float)."""
    name = "sites_walkers_n3000_t10" + ("_f32" if use_float else "")
    info, gold = refgen.load_fixture(name)
    params = dict(info["params"])
    tol = 1e-4 if use_float else 1e-9
    for config, lists in ((None, False), ({"cuda.nlist": True}, True)):
        m = EmuModel(refgen.model_path(info["model"]), params, use_float=use_float, config=config)
        m.use_nlist = lists
        m.flat_loop = 1
        m.populate()
        for _ in range(params["num_timesteps"]):
            m.timestep()
        assert m.nlist_builds == (1 if lists else 0)
        for t, ref in enumerate(gold):
            got = m.host_agents(t)
            assert len(got) == len(ref)
            assert exact_members_equal(got, ref), "integer/bool state differs"
            assert max_rel_error(got, ref, floor_ulps=F32_FLOOR_ULPS if tol > 1e-6 else 0) <= tol, "max relative error %.3e" % max_rel_error(got, ref)


def test_a_trial_variant_that_cannot_be_launched_falls_back_to_the_cursor_loop():
    """If a variant the tuner tries cannot even be launched (here: every launch of the chunked
    kernel fails), the launcher stops comparing, repeats the step with the cursor loop and stays
    there; the simulation goes on with correct results."""
    import shutil
    import tempfile
    from emu.emu import modes
    params = {"num_agents": 4000, "num_timesteps": 10}
    path = os.path.join(REPO, "examples", "boids2d.abl")
    m = EmuModel(path, params)
    with tempfile.TemporaryDirectory() as tmp:
        private = os.path.join(tmp, "libmodel_emu_fail.so")
        shutil.copy(os.path.join(m.dir, "libmodel_emu.so"), private)
        m = EmuModel(path, params, lib_path=private)
        m.lib.emu_set_fail_mode(1)
        m.flat_loop = -1
        m.populate()
        seen = []
        for _ in range(8):
            m.kernels = set()
            m.timestep()
            seen.append(modes(m.kernels))
        m.lib.emu_set_fail_mode(-1)
        assert seen[:2] == [[0], [3]] and all(s == [0] for s in seen[2:])    # third launch: chunked fails, cursor runs
        assert m.lib.abl_model_step_variant(0) == 0
        _, _, want = emulate(path, params, False, 8)
        got = m.host_agents(0)
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[0][f])


@pytest.mark.parametrize("flat_loop", [0, 1], ids=["cursor", "flat"])
def test_pairs_exactly_one_radius_apart_across_cell_borders_are_found(flat_loop):
    """Radius == granularity (the default: the granularity is the largest radius).  The reference
    (brute force, CPrinter.cpp:148-173) pairs up agents whose distance is exactly R — and, through
    its float-rounded filter, up to R(1 + 6e-8).  On a lattice of spacing R such partners sit one
    cell index apart only if the cells are a shade larger than R (ABL_CELL_PAD_*, abl_cuda.h):
    with cell == R the floor of the rounded product puts some of them TWO cells apart (9.999999999999998
    and 20.0 with R = 10: cells 0 and 2) and the 3 x 3 search misses them.  circle.abl (R = 2r = 10)
    on a lattice of spacing 10 plus one ulp either side, against the brute-force oracle."""
    from oracle import BRUTE
    R = 10.0
    k = np.arange(1, 25, dtype=np.float64)
    axis = np.concatenate([k * R, np.nextafter(k[::2] * R, 0.0), np.nextafter(k[1::3] * R, 1e9)])
    # the construction must contain what the test is about: accepted partners two UNPADDED cells apart
    cells = np.floor(axis * (1.0 / R))
    d = np.abs(axis[:, None] - axis[None, :])
    accepted = ~(np.sqrt((d * d).astype(np.float32)).astype(np.float64) > R)
    assert (accepted & (np.abs(cells[:, None] - cells[None, :]) >= 2)).any()
    xs, ys = np.meshgrid(axis, axis[:30])
    n_model = 50000                                       # W = sqrt(n / 0.05) = 1000: the lattice fits
    m = EmuModel(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": n_model})
    m.flat_loop = flat_loop
    m.populate()
    a = np.zeros(xs.size, dtype=m.dtypes[0])
    a["pos"][:, 0], a["pos"][:, 1] = xs.ravel(), ys.ravel()
    m.pools[0].load(a)
    m.timestep()
    got = m.host_agents(0)
    o = Oracle(False)
    want = o.circle_run(2, a, 1, BRUTE, num_agents=n_model)
    assert max_rel_error(got, want) <= 1e-9
    grid = o.circle_run(2, a, 1, GRID, num_agents=n_model)
    assert np.array_equal(got["pos"], grid["pos"])
