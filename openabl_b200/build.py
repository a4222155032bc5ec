"""In-tree builds: compiler driver, device runtime, and compiled models.

Everything lands inside the repository (OpenABL, asset/cuda/libabl_cuda.so,
build/models/<key>/libmodel.so) so that it travels to the GPU box with the snapshot.
"""
import hashlib
import os
import shutil
import subprocess

from .paths import ASSET_DIR, COMPILER, MODEL_CACHE, REPO_ROOT, RUNTIME_LIB


class BuildError(RuntimeError):
    pass


def _run(cmd, cwd=None):
    proc = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise BuildError("command failed (%d): %s\n%s" % (proc.returncode, " ".join(cmd), proc.stdout))
    return proc.stdout


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _sources(*dirs, exts=(".cpp", ".hpp", ".cu", ".cuh", ".h", ".c")):
    out = []
    for d in dirs:
        for root, _, files in os.walk(d):
            out += [os.path.join(root, f) for f in files if f.endswith(exts)]
    return out


def build_compiler(force=False):
    srcs = _sources(os.path.join(REPO_ROOT, "src"))
    if force or not _newer(COMPILER, srcs):
        _run(["make", "-s", "-B" if force else "-s", "OpenABL"], cwd=REPO_ROOT)
    return COMPILER


def build_runtime(force=False):
    srcs = _sources(os.path.join(ASSET_DIR, "cuda"), os.path.join(REPO_ROOT, "include"))
    if force or not _newer(RUNTIME_LIB, srcs):
        _run(["make", "-s", "runtime"], cwd=REPO_ROOT)
    return RUNTIME_LIB


def model_key(abl_path, params, config):
    h = hashlib.sha1()
    with open(abl_path, "rb") as f:
        h.update(f.read())
    for name in ("lib.abl", "cuda/abl_device.cuh", "cuda/abl_slab.cuh", "cuda/abl_host.h", "cuda/abl_host.c"):
        with open(os.path.join(ASSET_DIR, name), "rb") as f:
            h.update(f.read())
    with open(os.path.join(REPO_ROOT, "include", "abl_cuda.h"), "rb") as f:
        h.update(f.read())
    for src in sorted(_sources(os.path.join(REPO_ROOT, "src"))):
        with open(src, "rb") as f:
            h.update(f.read())
    h.update(repr(sorted((params or {}).items())).encode())
    h.update(repr(sorted((config or {}).items())).encode())
    base = os.path.splitext(os.path.basename(abl_path))[0]
    return "%s-%s" % (base, h.hexdigest()[:12])


def build_model(abl_path, params=None, config=None, force=False):
    """Compiles `abl_path` with `OpenABL -b cuda -B`; returns the output directory, which
    contains libmodel.so, ./main and the generated sources."""
    params = dict(params or {})
    config = dict(config or {})
    out = os.path.join(MODEL_CACHE, model_key(abl_path, params, config))
    lib = os.path.join(out, "libmodel.so")
    if os.path.exists(lib) and not force:
        return out
    build_compiler()
    build_runtime()
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(out)
    cmd = [COMPILER, "-i", abl_path, "-b", "cuda", "-o", out, "-A", ASSET_DIR, "-B"]
    for k, v in params.items():
        cmd += ["-P", "%s=%s" % (k, _fmt(v))]
    for k, v in config.items():
        cmd += ["-C", "%s=%s" % (k, _fmt(v))]
    try:
        _run(cmd, cwd=REPO_ROOT)
    except BuildError:
        shutil.rmtree(out, ignore_errors=True)
        raise
    return out


def _fmt(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    return str(v)
