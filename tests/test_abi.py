"""CPU tests: the C-ABI library loads and exports every symbol include/abl_cuda.h declares
(no compute calls without a GPU), and the front end reproduces the reference's lint goldens."""
import ctypes
import os
import re
import subprocess

import pytest

from openabl_b200 import build
from openabl_b200.paths import COMPILER, INCLUDE_DIR, REPO_ROOT, RUNTIME_LIB
from openabl_b200.runtime import ABI


def declared_symbols():
    with open(os.path.join(INCLUDE_DIR, "abl_cuda.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abl_cuda_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(ABI), "openabl_b200.runtime.ABI is out of sync with include/abl_cuda.h"


def test_library_exports_every_declared_symbol():
    build.build_runtime()
    lib = ctypes.CDLL(RUNTIME_LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), "libabl_cuda.so does not export %s" % name
    assert lib.abl_cuda_abi_version() == 1


def test_create_fails_without_gpu_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from openabl_b200.runtime import AblError, Runtime
    with pytest.raises(AblError):
        Runtime()


LINT_CASES = {
    # a few of the reference's diagnostics, restated as inline programs (the full golden set
    # lives in the reference tree and is run by tests/test_frontend.py when it is present)
    "agent Agent {}\nvoid foo() {\n  add(Agent {});\n}\nvoid main() {}\n":
        "add() can only be used in main() or a step function on line 3\n",
    "void main() {\n  if (1) {}\n  while (1.0) {}\n}\n":
        "if() condition must be bool, but received int on line 2\nwhile() condition must be bool, but received float on line 3\n",
    "void notMain() {}\n": "Script must have a main function on line 1\n",
}


@pytest.mark.parametrize("src,expected", list(LINT_CASES.items()))
def test_lint_diagnostics(tmp_path, src, expected):
    build.build_compiler()
    f = tmp_path / "m.abl"
    f.write_text(src)
    proc = subprocess.run([COMPILER, "--lint-only", "-A", os.path.join(REPO_ROOT, "asset"), "-i", str(f)],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert proc.returncode == 1
    assert proc.stderr == expected
