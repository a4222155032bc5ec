"""tools/abl_sweep.py against a real device: the population sweep of the reference's benchmark harness
(bench/bench.py:93-136) through the product CLI (`OpenABL -i … -b cuda -R`, generate + build + run) writes
bench_<model>_cuda.txt in the format bench/plot.py:43-85 reads."""
import csv
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sweep_runs_the_cli_and_writes_the_reference_result_format(tmp_path):
    from openabl_b200 import build
    build.build_compiler()
    build.build_runtime()
    proc = subprocess.run([sys.executable, os.path.join(REPO, "tools", "abl_sweep.py"), "-r", str(tmp_path),
                           "-m", "circle", "-n", "250-500"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stdout
    path = tmp_path / "bench_circle_cuda.txt"
    assert path.exists(), proc.stdout
    rows = list(csv.reader(path.read_text().splitlines()))
    assert rows[0] == ["n", "t"]
    assert [int(r[0]) for r in rows[1:]] == [250, 500], proc.stdout
    assert all(float(r[1]) > 0 for r in rows[1:])     # the `Execution time: <s>s` line of the generated program
