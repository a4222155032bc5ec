"""tools/abl_sweep.py writes the result files the reference's bench/plot.py reads
(bench_<model>_<backend>.txt with an `n,t` header, bench/bench.py:93-136, plot.py:43-85)."""
import csv
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))
import abl_sweep  # noqa: E402


class FakeRunner:
    def __init__(self):
        self.calls = []

    def get_exec_time(self, model, backend, params, config, out_dir=None):
        self.calls.append((model, backend, dict(params), dict(config)))
        if params["num_agents"] > 4000:
            raise abl_sweep.InvocationFailed("too large")
        return params["num_agents"] * 1e-6


def test_rows_double_and_stop_at_the_first_failure():
    r = FakeRunner()
    text = abl_sweep.run_bench(r, "cuda", "circle", (250, 32000), None, {"use_float": "true"}, log=lambda s: None)
    rows = list(csv.reader(text.splitlines()))
    assert rows[0] == ["n", "t"]
    assert [int(x[0]) for x in rows[1:]] == [250, 500, 1000, 2000, 4000]
    assert all(len(x) == 2 and float(x[1]) > 0 for x in rows[1:])      # what plot.py's reader requires
    assert r.calls[0][2] == {"num_timesteps": 100, "num_agents": 250}
    assert r.calls[0][3] == {"use_float": "true"}


def test_file_name_matches_the_pattern_plot_py_parses():
    pattern = re.compile(r"^bench_(.+?)_([^_]+)\.txt$")     # reference bench/plot.py:71
    m = pattern.match("bench_%s_%s.txt" % ("game_of_life", "cuda"))
    assert m and m.group(1) == "game_of_life" and m.group(2) == "cuda"


def test_command_line_is_the_reference_cli():
    r = abl_sweep.OpenAbl("/x/OpenABL", "/x/examples", "/x/asset")
    cmd = r.command("boids2d", "cuda", {"num_agents": 1000}, {"use_float": "true"})
    assert cmd == ["/x/OpenABL", "-i", "/x/examples/boids2d.abl", "-b", "cuda", "-A", "/x/asset", "-R",
                   "-P", "num_agents=1000", "-C", "use_float=true"]
