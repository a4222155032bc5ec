"""Slab decomposition on ONE GPU: S slabs (S runtimes) exchanging halo cells and migrating
agents must reproduce the undecomposed run bit for bit — through the staged in-process
transport and through the direct peer-memory transport (step kernels write their halo records
into the neighbouring slab's receive area; the same code path the multi-GPU runs use)."""
import os

import numpy as np
import pytest

from openabl_b200.model import Model
from openabl_b200.slab import LocalSlabs, split_layers

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ("boids2d.abl", {"num_agents": 100000}, False, 3, 10),
    ("boids2d.abl", {"num_agents": 100000}, True, 2, 10),
    ("circle.abl", {"num_agents": 50000}, False, 4, 10),
    ("circle3d.abl", {"num_agents": 20000}, False, 2, 5),
    ("game_of_life.abl", {"num_agents": 65536}, False, 8, 10),
]


def single_run(m, steps):
    m.create_runtime()
    m.upload_host()
    for _ in range(steps):
        m.timestep()
    single = m.download(0)
    m.close()
    return single


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["staged", "direct"])
@pytest.mark.parametrize("model_file,params,use_float,slabs,steps", CASES,
                         ids=["%s-%dslabs-%s" % (c[0][:-4], c[3], "f32" if c[2] else "f64") for c in CASES])
def test_decomposed_run_is_bit_identical(model_file, params, use_float, slabs, steps, transport):
    m = Model(os.path.join(REPO, "examples", model_file), params, use_float=use_float)
    m.populate()
    host = [m.host_agents(t) for t in range(m.n_types)]
    m.create_runtime()
    m.upload_host()
    for _ in range(steps):
        m.timestep()
    single = m.download(0)
    m.close()

    ls = LocalSlabs(m, slabs, transport=transport)
    ls.upload(host)
    assert sum(ls.owned_counts(0)) == len(host[0])
    for _ in range(steps):
        ls.timestep()
    assert sum(ls.owned_counts(0)) == len(host[0]), "agents lost or duplicated by migration"
    if model_file == "game_of_life.abl":
        # reductions are rank-local under decomposition; their sum is the global value
        pool = m.pool(0)
        assert sum(rt.sum_int(pool, 1) for rt in ls.rts) == int(single["alive"].sum())
        assert sum(rt.count(pool) for rt in ls.rts) == len(single)
    ids, rec = ls.download(0)
    ls.close()
    assert np.array_equal(ids, np.arange(len(host[0]), dtype=np.uint32))
    for f in rec.dtype.names:
        assert np.array_equal(rec[f], single[f]), "member %s differs from the single-slab run" % f


@pytest.mark.gpu
def test_stage_timing_keeps_the_forked_exchange_and_the_results():
    """abl_cuda_enable_timing brackets the stages of every step with events while the halo exchange still runs next
    to the step kernel on its side stream: results stay bit-identical, and every stage reports a time."""
    model_file, params, use_float, slabs, steps = CASES[0]
    m = Model(os.path.join(REPO, "examples", model_file), params, use_float=use_float)
    m.populate()
    host = [m.host_agents(t) for t in range(m.n_types)]
    single = single_run(m, steps)
    ls = LocalSlabs(m, slabs, transport="direct")
    ls.upload(host)
    for rt in ls.rts:
        rt.enable_timing(True)
    for _ in range(steps):
        ls.timestep()
    for rt in ls.rts:
        t = rt.last_timing()
        assert t["kernel_ms"] > 0 and t["bin_ms"] > 0 and t["commit_ms"] >= 0
        rt.enable_timing(False)
    ids, rec = ls.download(0)
    ls.close()
    assert np.array_equal(ids, np.arange(len(single), dtype=ids.dtype))
    for f in single.dtype.names:
        assert np.array_equal(rec[f], single[f])


@pytest.mark.gpu
@pytest.mark.parametrize("model_file,params,use_float,slabs,steps", [CASES[0], CASES[3], CASES[4]],
                         ids=["%s-%dslabs" % (c[0][:-4], c[3]) for c in (CASES[0], CASES[3], CASES[4])])
def test_scalable_upload_is_bit_identical(model_file, params, use_float, slabs, steps):
    """Slab r uploads the r-th part of the population by index; the records are grouped by owner on the
    device (abl_cuda_partition_upload), routed and adopted (abl_cuda_adopt_records): same state as the
    upload-everything-and-crop path, hence the same results as the undecomposed run."""
    m = Model(os.path.join(REPO, "examples", model_file), params, use_float=use_float)
    m.populate()
    host = [m.host_agents(t) for t in range(m.n_types)]
    single = single_run(m, steps)
    ls = LocalSlabs(m, slabs, transport="direct")
    ls.upload(host, scalable=True)
    assert sum(ls.owned_counts(0)) == len(host[0])
    for _ in range(steps):
        ls.timestep()
    ids, rec = ls.download(0)
    ls.close()
    assert np.array_equal(ids, np.arange(len(host[0]), dtype=np.uint32))
    for f in rec.dtype.names:
        assert np.array_equal(rec[f], single[f]), "member %s differs from the single-slab run" % f


@pytest.mark.gpu
def test_direct_transport_recovers_from_underestimated_padding(monkeypatch):
    """The host bins `owned + pad` records without knowing how many arrive; when more arrive
    than it assumed, the next binning notices and bins again over the true range."""
    monkeypatch.setenv("ABL_CUDA_HALO_PAD", "256")   # far fewer than a ghost layer holds
    m = Model(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": 50000})
    m.populate()
    host = [m.host_agents(0)]
    single = single_run(m, 6)
    ls = LocalSlabs(m, 3, transport="direct")
    ls.upload(host)
    for _ in range(6):
        ls.timestep()
    ids, rec = ls.download(0)
    ls.close()
    assert np.array_equal(ids, np.arange(len(host[0]), dtype=np.uint32))
    assert np.array_equal(rec["pos"], single["pos"])


@pytest.mark.gpu
def test_direct_transport_reports_message_overflow():
    """A receive area too small for a ghost layer is an error, not silent loss of agents."""
    from openabl_b200.runtime import AblError
    m = Model(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": 50000})
    m.populate()
    host = [m.host_agents(0)]
    ls = LocalSlabs(m, 2, transport="direct", halo_records=64)
    with pytest.raises(AblError, match="exceeds the capacity"):
        ls.upload(host)
        for _ in range(2):
            ls.timestep()
        ls.download(0)
    ls.close()


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["staged", "direct"])
@pytest.mark.parametrize("slabs", [2, 3])
def test_predator_prey_decomposed_add_remove_is_bit_identical(slabs, transport):
    """Run-time add()/removeCurrent() under slab decomposition (three agent types reading each
    other): removal compacts the owned range of the owning slab, new agents get the ids of the
    undecomposed run (rank of the parent among the parents of all slabs), and every member of
    every survivor equals the single-slab run bit for bit."""
    steps = 12
    m = Model(os.path.join(REPO, "examples", "predator_prey.abl"), {"num_agents": 32000})
    m.populate()
    host = [m.host_agents(t) for t in range(m.n_types)]
    m.create_runtime()
    m.upload_host()
    counts_single = []
    for _ in range(steps):
        m.timestep()
        counts_single.append([m.rt.count(m.pool(t)) for t in range(m.n_types)])
    single_ids = [m.rt.download_ids(m.pool(t)) for t in range(m.n_types)]
    single = [m.download(t) for t in range(m.n_types)]
    m.close()
    assert any(len(single[t]) != len(host[t]) for t in range(m.n_types)), "population never changed"
    assert any(len(single_ids[t]) and single_ids[t].max() >= len(host[t]) for t in range(m.n_types)), "nothing was added"

    ls = LocalSlabs(m, slabs, transport=transport)
    ls.upload(host)
    for k in range(steps):
        ls.timestep()
        got = [sum(ls.owned_counts(t)) for t in range(m.n_types)]
        assert got == counts_single[k], "agent counts differ from the single-slab run after timestep %d" % k
    # rank-local reductions add up to the global value
    grass = m.pool(2)
    assert sum(rt.sum_int(grass, 2) for rt in ls.rts) == int(single[2]["avail"].sum())
    for t in range(m.n_types):
        ids, rec = ls.download(t)
        assert np.array_equal(ids, single_ids[t]), "%s: ids differ from the single-slab run" % m.names[t]
        for f in rec.dtype.names:
            assert np.array_equal(rec[f], single[t][f]), "%s.%s differs from the single-slab run" % (m.names[t], f)
    ls.close()


def test_global_add_ranks():
    from openabl_b200.slab import global_add_ranks
    ranks, total = global_add_ranks([np.array([3, 10, 42], dtype=np.uint32), np.zeros(0, dtype=np.uint32),
                                     np.array([1, 11], dtype=np.uint32)])
    assert total == 5
    assert [r.tolist() for r in ranks] == [[1, 2, 4], [], [0, 3]]
    ranks, total = global_add_ranks([np.zeros(0, dtype=np.uint32)] * 2)
    assert total == 0 and all(len(r) == 0 for r in ranks)


def test_split_layers():
    assert split_layers(69, 8) == [(0, 8), (8, 17), (17, 25), (25, 34), (34, 43), (43, 51), (51, 60), (60, 69)]
    assert split_layers(4, 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    with pytest.raises(ValueError):
        split_layers(3, 4)
