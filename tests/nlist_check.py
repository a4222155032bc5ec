"""Bodies of tests/test_gpu_zz_nlist.py, run in a child process (`python tests/nlist_check.py <case>`)
so that a device fault on this not-yet-validated path cannot poison the CUDA context of the
test session."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

from oracle import GRID, Oracle          # noqa: E402
from openabl_b200.model import Model     # noqa: E402


def run(path, params, steps, config, use_float=False):
    m = Model(path, params, use_float=use_float, config=config)
    m.populate()
    m.create_runtime()
    m.upload_host()
    for _ in range(steps):
        m.timestep()
    out = [m.download(t) for t in range(m.n_types)]
    m.close()
    return out


def game_of_life():
    params = {"num_agents": 65536}
    got = run(os.path.join(REPO, "examples", "game_of_life.abl"), params, 10, {"cuda.nlist": True})
    o = Oracle(False)
    want = o.run_for("game_of_life.abl", params, o.init_for("game_of_life.abl", params), 10, GRID)
    assert np.array_equal(got[0]["alive"], want["alive"])
    assert np.array_equal(got[0]["pos"], want["pos"])


def static_sites():
    path = os.path.join(REPO, "tests", "models", "static_sites.abl")
    params = {"num_agents": 3000}
    plain = run(path, params, 5, None)
    listed = run(path, params, 5, {"cuda.nlist": True})
    for a, b in zip(plain, listed):
        assert len(a) == len(b)
        for f in a.dtype.names:
            assert np.array_equal(a[f], b[f]), "member %s differs" % f
    m = Model(path, params, config={"cuda.nlist": True})
    m.populate()
    m.create_runtime()
    for _ in range(2):          # second upload: same records, the lists must be rebuilt for the new pool order
        m.upload_host()
        for _ in range(5):
            m.timestep()
        again = [m.download(t) for t in range(m.n_types)]
        for a, b in zip(plain, again):
            for f in a.dtype.names:
                assert np.array_equal(a[f], b[f]), "member %s differs after re-upload" % f
    m.close()


def sites_walkers():
    """Lists for the static Sites, against the golden vector of the real reference `c` backend."""
    import refgen
    from openabl_b200.state import exact_members_equal, max_rel_error
    info, gold = refgen.load_fixture("sites_walkers_n3000_t10")
    params = dict(info["params"])
    got = run(refgen.model_path(info["model"]), params, params["num_timesteps"], {"cuda.nlist": True})
    for g, ref in zip(got, gold):
        assert len(g) == len(ref) and exact_members_equal(g, ref)
        assert max_rel_error(g, ref) <= 1e-9, max_rel_error(g, ref)


if __name__ == "__main__":
    {"game_of_life": game_of_life, "static_sites": static_sites, "sites_walkers": sites_walkers}[sys.argv[1]]()
    print("ok")
