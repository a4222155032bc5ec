"""Front-end conformance against the reference tree (only where /root/reference exists):
the 23 golden diagnostics, and identical folded constants / overload numbering as the real
reference compiler on every example."""
import os
import re
import subprocess

import pytest

from openabl_b200 import build
from openabl_b200.paths import ASSET_DIR, COMPILER, REPO_ROOT

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(REPO_ROOT, "oracle", "_ref", "OpenABL_ref")
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _tests():
    if not os.path.isdir(REF):
        return []
    return sorted(f[:-4] for f in os.listdir(os.path.join(REF, "test")) if f.endswith(".abl"))


@needs_ref
@pytest.mark.parametrize("name", _tests())
def test_reference_lint_golden(name):
    build.build_compiler()
    proc = subprocess.run([COMPILER, "--lint-only", "-A", ASSET_DIR, "-i", os.path.join(REF, "test", name + ".abl")],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    with open(os.path.join(REF, "test", name + ".exp")) as f:
        expected = f.read()
    assert proc.stderr.split() == expected.split()  # diff -b, as reference test.sh:41


def _examples():
    if not os.path.isdir(REF):
        return []
    return sorted(f for f in os.listdir(os.path.join(REF, "examples")) if f.endswith(".abl"))


@needs_ref
@pytest.mark.parametrize("example", _examples())
def test_reference_examples_compile_unchanged(example, tmp_path):
    """Every model of the reference distribution is accepted unchanged by the cuda backend."""
    build.build_compiler()
    out = tmp_path / "out"
    proc = subprocess.run([COMPILER, "-A", ASSET_DIR, "-i", os.path.join(REF, "examples", example),
                           "-b", "cuda", "-o", str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout
    assert (out / "model_kernels.cu").exists() and (out / "build.sh").exists() and (out / "run.sh").exists()


CONST_RE = re.compile(r"^(?:double|float|int|bool|abl_real)\s+(\w+)\s*=\s*([^;{]+);", re.M)


@needs_ref
@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built")
@pytest.mark.parametrize("example,params", [
    ("circle.abl", {"num_agents": 1000}), ("circle3d.abl", {"num_agents": 16000000}),
    ("boids2d.abl", {"num_agents": 1000000}), ("game_of_life.abl", {"num_agents": 16777216}),
    ("boids2d.abl", {"num_agents": 1200}), ("sugarscape.abl", {}), ("ants.abl", {}),
])
def test_folded_constants_and_function_names_match_reference(example, params, tmp_path):
    build.build_compiler()
    model = os.path.join(REF, "examples", example)
    pargs = []
    for k, v in params.items():
        pargs += ["-P", "%s=%s" % (k, v)]
    mine = tmp_path / "mine"
    subprocess.run([COMPILER, "-A", ASSET_DIR, "-i", model, "-b", "cuda", "-o", str(mine)] + pargs, check=True)
    backend = "c"
    theirs = tmp_path / "theirs"
    proc = subprocess.run([REF_BIN, "-A", os.path.join(REF, "asset"), "-i", model, "-b", backend, "-o", str(theirs)] + pargs,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout
    host = (mine / "model_host.c").read_text()
    if backend == "c":
        ref = (theirs / "main.c").read_text()
        mine_consts = dict(CONST_RE.findall(host))
        ref_consts = dict(CONST_RE.findall(ref))
        assert ref_consts, "no constants found in the reference output"
        for name, value in ref_consts.items():
            assert mine_consts.get(name, "").strip() == value.strip(), "constant %s: %r vs reference %r" % (
                name, mine_consts.get(name), value)
        ref_funcs = set(re.findall(r"^\w[\w\*]*\s+(\w+)\(", ref, re.M)) - {"main"}
        mine_funcs = set(re.findall(r"^\w[\w\* ]*\s+\*?(\w+)\(", host, re.M))
        step_funcs = set(re.findall(r"^void (\w+)\(\w+\* in, \w+\* out\)", ref, re.M))
        assert (ref_funcs - step_funcs) <= mine_funcs, sorted(ref_funcs - step_funcs - mine_funcs)


@needs_ref
@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built")
def test_reference_c_backend_rejects_add_remove(tmp_path):
    """predator_prey uses add()/removeCurrent(): the reference `c` backend exits with code 2
    (BackendError, CBackend.cpp:30-32), so that model has no reference-pinned numbers."""
    proc = subprocess.run([REF_BIN, "-A", os.path.join(REF, "asset"), "-i",
                           os.path.join(REF, "examples", "predator_prey.abl"), "-b", "c", "-o", str(tmp_path / "o")],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 2


ARRAY_RE = re.compile(r"^(\w+)\s+(\w+)\[\]\s*=\s*\{([^}]*)\};", re.M)


@needs_ref
@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built")
@pytest.mark.parametrize("model", ["tests/models/table_cells.abl", "tests/models/flock3d.abl",
                                   "tests/models/two_species.abl", os.path.join(REF, "examples", "keratinocyte.abl")])
def test_constants_and_tables_of_feature_models_match_reference(model, tmp_path):
    """Scalars fold to the reference's 6-digit text, and constant TABLES keep the reference's
    element type: the reference prints the DSL base type for arrays (CPrinter.cpp:20-24), so a
    `float` table is single precision even in a double-precision build."""
    build.build_compiler()
    model = model if os.path.isabs(model) else os.path.join(REPO, model)
    mine, theirs = tmp_path / "mine", tmp_path / "theirs"
    subprocess.run([COMPILER, "-A", ASSET_DIR, "-i", model, "-b", "cuda", "-o", str(mine)], check=True)
    proc = subprocess.run([REF_BIN, "-A", os.path.join(REF, "asset"), "-i", model, "-b", "c", "-o", str(theirs)],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode == 2:
        pytest.skip("the reference c backend rejects this model (run-time add/remove)")
    assert proc.returncode == 0, proc.stdout
    host, ref = (mine / "model_host.c").read_text(), (theirs / "main.c").read_text()
    dev = (mine / "model_kernels.cu").read_text()
    mine_consts, ref_consts = dict(CONST_RE.findall(host)), dict(CONST_RE.findall(ref))
    for name, value in ref_consts.items():
        assert mine_consts.get(name, "").strip() == value.strip(), "constant %s" % name
    norm = lambda t: re.sub(r"\s+", "", t)
    mine_tabs = {n: (ty, norm(v)) for ty, n, v in ARRAY_RE.findall(host)}
    for ty, name, values in ARRAY_RE.findall(ref):
        assert mine_tabs.get(name) == (ty, norm(values)), "table %s: %r vs reference %r" % (name, mine_tabs.get(name), (ty, norm(values)))
        assert re.search(r"__device__ const %s %s\[\]" % (ty, name), dev), "device copy of %s has another element type" % name
