"""The OUTPUT FILE of a generated program against the file the reference writes (north_star: "the output
files (e.g. points.json) identical").  Golden text: tests/golden/text/*.txt, the reference's own save()
(libabl.c:46-124: `%f` with six decimals, `[x,y(,z)]` vectors, true/false, agents in array order) from
runs of the real reference `c` backend (oracle/refgen.py TEXT_FIXTURES).

  * game_of_life (integer / bool state, positions that never change): BYTE-IDENTICAL;
  * floating-point models: the same text skeleton (keys, order, punctuation, number of agents) and every
    printed number within one unit of the last printed digit — positions agree to ~1e-12 relative, so a
    digit can only differ where the value sits on a rounding boundary of `%f`."""
import os
import re
import subprocess

import pytest

import refgen
from openabl_b200 import build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NUMBER = re.compile(r"-?\d+\.\d+")


def run_program(model, params, use_float, work):
    out = build.build_model(refgen.model_path(model), params, {"use_float": True} if use_float else {})
    work.mkdir()
    proc = subprocess.run([os.path.join(out, "main")], cwd=str(work), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:]
    files = [f for f in os.listdir(str(work)) if not f.endswith(".bin")]
    assert len(files) == 1, files
    return (work / files[0]).read_text()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(refgen.TEXT_FIXTURES))
def test_output_file_matches_the_reference(name, tmp_path):
    model, params, use_float = refgen.TEXT_FIXTURES[name]
    with open(refgen.text_fixture_path(name)) as f:
        want = f.read()
    got = run_program(model, params, use_float, tmp_path / "run")
    if model == "game_of_life.abl":
        assert got == want, "game_of_life output is not byte-identical to the reference's file"
        return
    assert NUMBER.sub("#", got) == NUMBER.sub("#", want), "text skeleton (keys, order, punctuation) differs"
    a, b = NUMBER.findall(got), NUMBER.findall(want)
    assert len(a) == len(b)
    unit = 2e-6 * (50 if use_float else 1)      # one unit of the sixth decimal; float runs agree to 1e-4 relative only
    worst = max(abs(float(x) - float(y)) for x, y in zip(a, b))
    same = sum(x == y for x, y in zip(a, b))
    assert worst <= unit * (1 if not use_float else max(1.0, max(abs(float(y)) for y in b))), \
        "printed value differs by %.3g" % worst
    if not use_float:
        assert same >= 0.999 * len(a), "only %d of %d printed numbers are identical" % (same, len(a))


def test_text_fixtures_are_present():
    for name in refgen.TEXT_FIXTURES:
        assert os.path.getsize(refgen.text_fixture_path(name)) > 1000
