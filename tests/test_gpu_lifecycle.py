"""tests/test_lifecycle_trace.py on the GPU: removal, run-time add, partial writes and a neighbour count of
tests/models/lifecycle.abl against the hand trace restated in Python — undecomposed and in two slabs."""
import os

import numpy as np
import pytest

from openabl_b200.model import Model
from openabl_b200.slab import LocalSlabs
from test_lifecycle_trace import MODEL, PARAMS, check


@pytest.mark.gpu
def test_gpu_follows_the_hand_trace():
    m = Model(MODEL, PARAMS)
    m.populate()
    m.create_runtime()
    m.upload_host()
    counts = []
    for _ in range(PARAMS["num_timesteps"]):
        m.timestep()
        counts.append(m.rt.count(m.pool(0)))
    ids, rec = m.rt.download_ids(m.pool(0)), m.download(0)
    m.close()
    check(counts, ids, rec)


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["staged", "direct"])
def test_gpu_slabs_follow_the_hand_trace(transport):
    m = Model(MODEL, PARAMS)
    m.populate()
    host = [m.host_agents(0)]
    ls = LocalSlabs(m, 2, transport=transport)
    ls.upload(host)
    counts = []
    for _ in range(PARAMS["num_timesteps"]):
        ls.timestep()
        counts.append(sum(ls.owned_counts(0)))
    ids, rec = ls.download(0)
    ls.close()
    check(counts, ids, rec)
