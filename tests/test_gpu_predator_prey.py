"""predator_prey: run-time add/remove (stream compaction), in-step RNG, reductions.

The reference cannot run this model on any backend available here (its `c` backend exits
with BackendError), so parity is UNPINNED by the reference; the semantics are frozen in
oracle/abl_oracle.c (see the header of that section) and the CUDA path must match that
restatement exactly: agent counts, integer/bool state and — in double precision — every
floating-point member, bit for bit."""
import os

import numpy as np
import pytest

from oracle import GRID, PredatorPreyOracle
from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("n,steps,append_scan", [(32000, 20, None), (200000, 6, None), (32000, 12, "0"), (32000, 12, "1000000000")])
def test_predator_prey_matches_frozen_semantics(n, steps, append_scan, monkeypatch):
    # append_scan: ABL_CUDA_APPEND_SCAN pins how new agents are ranked by parent id (commit_adds in
    # abl_runtime.cu): "0" = always through the scan over per-id presence flags, huge = always by counting;
    # None = the runtime's rule (counting up to 2048 adds per step function).  Same ids either way.
    if append_scan is not None:
        monkeypatch.setenv("ABL_CUDA_APPEND_SCAN", append_scan)
    m = Model(os.path.join(REPO, "examples", "predator_prey.abl"), {"num_agents": n})
    m.populate()
    o = PredatorPreyOracle(n)
    # same initial population (host init code vs. restatement)
    for t in range(3):
        host = m.host_agents(t)
        _, rec = o.read(t)
        assert len(host) == len(rec)
        for f in host.dtype.names:
            assert np.array_equal(host[f], rec[f]), "initial %s.%s differs" % (m.names[t], f)
    m.create_runtime()
    m.upload_host()
    removed_any = added_any = False
    for step in range(steps):
        before = [o.count(t) for t in range(3)]
        m.timestep()
        o.timestep(GRID)
        counts = [m.rt.count(m.pool(t)) for t in range(3)]
        assert counts == [o.count(t) for t in range(3)], "agent counts differ after step %d" % step
        removed_any |= counts[1] < before[1]
        assert m.rt.sum_int(m.pool(2), 2) == o.sum_avail()
    for t in range(3):
        got = m.download(t)
        ids, want = o.read(t)
        assert len(got) == len(want)
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[f]), "%s.%s differs" % (m.names[t], f)
        added_any |= bool(len(ids) and ids.max() >= len(m.host_agents(t)))
    m.close()
    o.close()
    assert removed_any and added_any, "test population too small to exercise add/remove"


@pytest.mark.gpu
def test_predator_prey_cli_end_to_end(tmp_path):
    """The generated ./main runs the whole program: init, simulate with a sequential step that
    logs reductions to log.csv, save() to JSON."""
    import json
    import subprocess
    from openabl_b200 import build
    out = build.build_model(os.path.join(REPO, "examples", "predator_prey.abl"),
                            {"num_agents": 32000, "num_timesteps": 5})
    work = tmp_path / "run"
    work.mkdir()
    subprocess.run([os.path.join(out, "main")], cwd=str(work), check=True)
    rows = (work / "log.csv").read_text().strip().splitlines()
    assert len(rows) == 5
    first = rows[0].split(",")
    assert len(first) == 4 and int(first[0]) > 0 and int(first[1]) > 0
    data = json.loads((work / "agents.json").read_text())
    assert set(data) == {"Predator", "Prey", "Grass"}
    assert len(data["Prey"]) == int(rows[-1].split(",")[0])
