"""A compiled `.abl` model: generated libmodel.so driven through its C entry points.

Mirrors what the generated ./main does (populate -> setup -> upload -> timesteps ->
download -> save), but leaves every stage callable on its own so tests and the benchmark
can look at raw state and time the device-resident part separately.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .runtime import AblError, AgentDesc, Runtime, check, load_library
from .state import agent_dtype, parse_agents


class _HostArray(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


class _HostType(C.Structure):
    _fields_ = [("desc", AgentDesc), ("agents", C.POINTER(_HostArray)), ("pool", C.c_int)]


class Model:
    def __init__(self, abl_path, params=None, use_float=False, config=None):
        self.abl_path = abl_path
        self.params = dict(params or {})
        self.use_float = bool(use_float)
        cfg = dict(config or {})
        if use_float:
            cfg["use_float"] = True
        self.dir = _build.build_model(abl_path, self.params, cfg)
        # A/B experiments: load a hand-modified build of the same model instead
        self.dir = os.environ.get("ABL_MODEL_DIR", self.dir)
        load_library()
        lib_path = os.path.join(self.dir, "libmodel.so")
        if not os.path.exists(lib_path):
            raise AblError("model library missing: %s" % lib_path)
        self.lib = C.CDLL(lib_path, mode=C.RTLD_LOCAL)
        with open(abl_path) as f:
            self.agents = parse_agents(f.read())
        self.dtypes = [agent_dtype(m, self.use_float) for _, m in self.agents]
        self.names = [n for n, _ in self.agents]
        self.n_types = C.c_int.in_dll(self.lib, "abl_model_n_types").value
        assert self.n_types == len(self.agents), "agent table mismatch"
        self._types = (_HostType * (self.n_types + 1)).in_dll(self.lib, "abl_model_types")
        for t in range(self.n_types):
            assert self._types[t].desc.stride == self.dtypes[t].itemsize, "layout mismatch for %s" % self.names[t]
        self.n_steps = C.c_int.in_dll(self.lib, "abl_model_n_steps").value
        self.lib.abl_model_step_name.restype = C.c_char_p
        self.step_names = [self.lib.abl_model_step_name(s).decode() for s in range(self.n_steps)]
        for fn in ("abl_model_setup", "abl_model_timestep", "abl_model_parallel_steps",
                   "abl_model_upload", "abl_model_download", "abl_model_unpin"):
            getattr(self.lib, fn).argtypes = [C.c_void_p]
            getattr(self.lib, fn).restype = C.c_int
        self.lib.abl_model_run_step.argtypes = [C.c_void_p, C.c_int]
        self.lib.abl_model_step_variant.argtypes = [C.c_int]
        self.lib.abl_model_step_variant.restype = C.c_int
        self.lib.abl_model_step_flags.argtypes = [C.c_int]
        self.lib.abl_model_step_flags.restype = C.c_int
        self.lib.abl_model_sequential_step.argtypes = [C.c_void_p]
        self.lib.abl_model_sequential_step.restype = C.c_int
        self.lib.abl_model_set_runtime.argtypes = [C.c_void_p]
        self.rt = None

    # ---- host side ----------------------------------------------------------------------
    def populate(self):
        """Runs the model's own initialisation code (statements of main() before simulate)."""
        self.lib.abl_model_populate()

    def write_frame(self, path):
        """`-C visualize=true` builds only: one picture (binary PPM) of the host arrays as they are — after
        populate() or download_host() — painted with the model's getColor / getSize hooks."""
        if not hasattr(self.lib, "abl_model_write_frame"):
            raise AblError("model was not generated with -C visualize=true")
        self.lib.abl_model_write_frame.argtypes = [C.c_char_p]
        self.lib.abl_model_write_frame.restype = C.c_int
        if self.lib.abl_model_write_frame(os.fsencode(path)) != 0:
            raise AblError("could not write frame %s" % path)

    def host_agents(self, t):
        """Copy of the host records of agent type index `t` as a structured array."""
        arr = self._types[t].agents.contents
        n = arr.len
        if n == 0:
            return np.zeros(0, dtype=self.dtypes[t])
        buf = (C.c_char * (n * self.dtypes[t].itemsize)).from_address(arr.data)
        return np.frombuffer(buf, dtype=self.dtypes[t]).copy()

    def host_view(self, t):
        """The host records of agent type index `t` WITHOUT a copy (valid until the array is resized)."""
        arr = self._types[t].agents.contents
        n = arr.len
        if n == 0:
            return np.zeros(0, dtype=self.dtypes[t])
        buf = (C.c_char * (n * self.dtypes[t].itemsize)).from_address(arr.data)
        return np.frombuffer(buf, dtype=self.dtypes[t])

    def host_count(self, t):
        """Number of host records of agent type index `t` (no copy)."""
        return self._types[t].agents.contents.len

    def pool(self, t):
        return self._types[t].pool

    # ---- device side --------------------------------------------------------------------
    def create_runtime(self, **kw):
        self.rt = Runtime(use_float=self.use_float, **kw)
        check(self.lib.abl_model_setup(self.rt.handle), "abl_model_setup")
        self.lib.abl_model_set_runtime(self.rt.handle)
        return self.rt

    def upload_host(self):
        check(self.lib.abl_model_upload(self.rt.handle), "abl_model_upload")

    def upload(self, t, array):
        assert array.dtype == self.dtypes[t]
        self.rt.upload(self.pool(t), np.ascontiguousarray(array))

    def download(self, t):
        return self.rt.download(self.pool(t), self.dtypes[t])

    def download_host(self):
        """Device -> the model's own (page-locked) host arrays, as the generated program does."""
        check(self.lib.abl_model_download(self.rt.handle), "abl_model_download")

    def timestep(self):
        check(self.lib.abl_model_timestep(self.rt.handle), "abl_model_timestep")

    def run_step(self, s):
        check(self.lib.abl_model_run_step(self.rt.handle, s), "abl_model_run_step")

    def step_flags(self, s):
        """bit 0: step function `s` removes agents, bit 1: it adds agents at run time."""
        return self.lib.abl_model_step_flags(s)

    def step_variant(self, s):
        """ABL_MODE of the kernel the latest launch of step function `s` used (-1 before the first):
        0 cursor loop, 1 chunked, 2 shared-memory tile, 3 flat loop, 4 neighbour-list walk, 7 TMA-staged tile,
        8 single-precision shadow pre-filter, 9 split pre-filter.  Chosen by the launcher's rule (with
        ABL_CUDA_TUNE=1: by its run-time tuner, after the tuning phase)."""
        return self.lib.abl_model_step_variant(s)

    def sequential_step(self):
        check(self.lib.abl_model_sequential_step(self.rt.handle), "abl_model_sequential_step")

    def close(self):
        if self.rt is not None:
            self.lib.abl_model_unpin(self.rt.handle)
            self.rt.close()
            self.rt = None
