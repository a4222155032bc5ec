"""Run-time add / remove semantics against an INDEPENDENT restatement.

The reference defines these constructs only in its Mason / FlameGPU printers (its `c` backend exits with
BackendError): an agent that calls removeCurrent() is gone after its step function
(MasonPrinter.cpp:177-178, 358-367), an agent created by add() takes part from the next step function on
(MasonPrinter.cpp:159-174), members a step function does not assign are carried over
(MasonPrinter.cpp:480-486), step functions run one after the other within a timestep (:535-541).  The
repository's other checker for them, oracle/abl_oracle.c, was written by the same hand as the kernels;
this test restates the rules once more in plain Python for a model small and deterministic enough to trace
by hand (tests/models/lifecycle.abl: no random numbers), including the two conventions the backend adds:
new agents get the next free ids in the order of their parents' ids, and removal keeps the survivors'
order.  Checked: the population after every timestep and every member of the final state, exactly —
under the CPU kernel emulator here and on the GPU (tests/test_gpu_lifecycle.py imports this module)."""
import os

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = os.path.join(REPO, "tests", "models", "lifecycle.abl")
PARAMS = {"num_agents": 64, "num_timesteps": 8}


def trace(n, timesteps):
    """-> (population after every timestep, final agents as dicts in id order)."""
    cells = [{"id": i, "pos": ((i % 8) * 2 + 0.5, (i // 8) * 2 + 0.5), "age": i % 7, "kind": 0} for i in range(n)]
    next_id = n
    counts = []
    for _ in range(timesteps):
        # grow: only `age` is written
        cells = [dict(c, age=c["age"] + 1) for c in cells]
        # divide: every cell sees the population as it was when the step function started
        born = []
        for c in cells:
            crowd = 0
            for o in cells:
                dx, dy = o["pos"][0] - c["pos"][0], o["pos"][1] - c["pos"][1]
                # the reference's filter: skip if (double)sqrtf((float)d2) > radius (libabl.h:156-168)
                if not (float(np.sqrt(np.float32(dx * dx + dy * dy))) > 1.5):
                    crowd += 1
            if c["age"] % 3 == 0 and crowd < 4:
                born.append((c["id"], {"pos": (c["pos"][0] + 0.25, c["pos"][1] + 0.125), "age": 0, "kind": c["kind"] + 1}))
        for _, child in sorted(born, key=lambda b: b[0]):     # ids in the order of the parents' ids
            child["id"] = next_id
            next_id += 1
            cells.append(child)
        # die: the daughters of this timestep take part already
        cells = [c for c in cells if not c["age"] > 5]
        counts.append(len(cells))
    return counts, sorted(cells, key=lambda c: c["id"])


def check(counts, ids, rec, timesteps=PARAMS["num_timesteps"], n=PARAMS["num_agents"]):
    want_counts, want = trace(n, timesteps)
    assert counts == want_counts, "population per timestep: %s, hand trace: %s" % (counts, want_counts)
    order = np.argsort(ids)
    assert [int(i) for i in ids[order]] == [c["id"] for c in want]
    rec = rec[order]
    assert [int(a) for a in rec["age"]] == [c["age"] for c in want]
    assert [int(k) for k in rec["kind"]] == [c["kind"] for c in want]
    assert np.array_equal(rec["pos"], np.array([c["pos"] for c in want], dtype=rec["pos"].dtype))
    assert max(c["kind"] for c in want) >= 2 and want_counts[-1] != n, "the scenario exercises neither add nor remove"


def test_hand_trace_is_what_the_docstring_says():
    counts, cells = trace(PARAMS["num_agents"], PARAMS["num_timesteps"])
    # timestep 1 by hand: ages 0..6 become 1..7; cells of age 3 and 6 with fewer than 4 neighbours within 1.5
    # (lattice spacing 2: only themselves) divide: i % 7 in (2, 5) -> 9 + 9 = 18 daughters; cells of age 6
    # and 7 die (i % 7 in (5, 6): 9 + 9 = 18, among them 9 mothers)
    assert counts[0] == 64 + 18 - 18
    assert len({c["id"] for c in cells}) == len(cells)


@pytest.mark.parametrize("flat_loop", [0, 1], ids=["cursor", "flat"])
def test_generated_kernels_under_the_emulator_follow_the_hand_trace(flat_loop):
    import sys
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from emu.emu import EmuModel
    m = EmuModel(MODEL, PARAMS)
    m.flat_loop = flat_loop
    m.populate()
    counts = []
    for _ in range(PARAMS["num_timesteps"]):
        m.timestep()
        counts.append(len(m.pools[0].ids))
    check(counts, np.sort(np.asarray(m.pools[0].ids)), m.host_agents(0))      # host_agents(): ascending id order
