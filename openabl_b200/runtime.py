"""ctypes binding of the runtime C ABI (include/abl_cuda.h).

Fails loudly when the CUDA library is missing or no device is present: there is no CPU
fallback on the product path.
"""
import ctypes as C
import os

from .paths import RUNTIME_LIB

ABL_MAX_COLUMNS = 32


class AblError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("use_float", C.c_int), ("seed", C.c_uint64),
                ("deterministic", C.c_int), ("tile_neighbours", C.c_int), ("block_size", C.c_int)]


class MemberDesc(C.Structure):
    _fields_ = [("type", C.c_int), ("offset", C.c_uint), ("name", C.c_char_p), ("is_pos", C.c_int)]


class AgentDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("members", C.POINTER(MemberDesc)), ("n_members", C.c_int),
                ("stride", C.c_uint)]


class StepTiming(C.Structure):
    _fields_ = [("bin_ms", C.c_float), ("kernel_ms", C.c_float), ("commit_ms", C.c_float),
                ("launches", C.c_uint)]


# every symbol include/abl_cuda.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
ABI = {
    "abl_cuda_abi_version": (C.c_int, []),
    "abl_cuda_default_config": (None, [C.POINTER(Config)]),
    "abl_cuda_create": (C.c_int, [C.POINTER(_VP), C.POINTER(Config)]),
    "abl_cuda_destroy": (C.c_int, [_VP]),
    "abl_cuda_last_error": (C.c_char_p, []),
    "abl_cuda_set_environment": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double]),
    "abl_cuda_add_pool": (C.c_int, [_VP, C.POINTER(AgentDesc), C.POINTER(C.c_int)]),
    "abl_cuda_upload": (C.c_int, [_VP, C.c_int, _VP, C.c_size_t]),
    "abl_cuda_upload_with_ids": (C.c_int, [_VP, C.c_int, _VP, _VP, C.c_size_t, C.c_uint]),
    "abl_cuda_download": (C.c_int, [_VP, C.c_int, _VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "abl_cuda_pool_size": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_size_t)]),
    "abl_cuda_transit_record_bytes": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_size_t)]),
    "abl_cuda_partition_upload": (C.c_int, [_VP, C.c_int, _VP, C.c_size_t, C.c_uint, _VP, C.POINTER(C.c_uint)]),
    "abl_cuda_adopt_records": (C.c_int, [_VP, C.c_int, _VP, C.c_size_t, C.c_uint]),
    "abl_cuda_group_simulate": (C.c_int, [_VP, C.c_int, _VP, _VP, C.c_int, _VP]),
    "abl_cuda_download_ids": (C.c_int, [_VP, C.c_int, _VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "abl_cuda_pin_host": (C.c_int, [_VP, _VP, C.c_size_t]),
    "abl_cuda_unpin_host": (C.c_int, [_VP, _VP]),
    "abl_cuda_register_step": (C.c_int, [_VP, _VP, C.POINTER(C.c_int)]),
    "abl_cuda_step": (C.c_int, [_VP, C.c_int]),
    "abl_cuda_begin_timestep": (C.c_int, [_VP]),
    "abl_cuda_end_timestep": (C.c_int, [_VP]),
    "abl_cuda_synchronize": (C.c_int, [_VP]),
    "abl_cuda_count": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_int)]),
    "abl_cuda_count_member_int": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "abl_cuda_count_member_float": (C.c_int, [_VP, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int)]),
    "abl_cuda_sum_int": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "abl_cuda_sum_float": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "abl_cuda_last_exec_time": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "abl_cuda_bin": (C.c_int, [_VP, C.c_int]),
    "abl_cuda_debug_binning": (C.c_int, [_VP, C.c_int, _VP, C.c_size_t, _VP, C.c_size_t]),
    "abl_cuda_grid_cells": (C.c_int, [_VP, C.POINTER(C.c_uint), C.POINTER(C.c_int * 3)]),
    "abl_cuda_enable_timing": (C.c_int, [_VP, C.c_int]),
    "abl_cuda_time_kernel": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "abl_cuda_last_timing": (C.c_int, [_VP, C.POINTER(StepTiming)]),
    "abl_cuda_stream": (_VP, [_VP]),
    "abl_cuda_nccl_unique_id": (C.c_int, [_VP]),
    "abl_cuda_comm_init_nccl": (C.c_int, [_VP, _VP, C.c_int, C.c_int]),
    "abl_cuda_set_slab": (C.c_int, [_VP, C.POINTER(C.c_int), C.c_int, C.c_int]),
    "abl_cuda_slab_axis_layers": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "abl_cuda_exchange": (C.c_int, [_VP, C.c_int]),
    "abl_cuda_owned_size": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_size_t)]),
    "abl_cuda_halo_setup": (C.c_int, [_VP, C.c_int, C.c_size_t, _VP]),
    "abl_cuda_halo_connect": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "abl_cuda_halo_connect_local": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "abl_cuda_set_local_peers": (C.c_int, [_VP, _VP, _VP]),
    "abl_cuda_exchange_begin": (C.c_int, [_VP, C.c_int]),
    "abl_cuda_exchange_end": (C.c_int, [_VP, C.c_int]),
    "abl_cuda_pending_adds": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_uint)]),
    "abl_cuda_pending_add_parents": (C.c_int, [_VP, _VP, C.c_size_t]),
    "abl_cuda_resolve_adds": (C.c_int, [_VP, _VP, C.c_uint]),
    "abl_cuda_set_reduce_hook": (C.c_int, [_VP, _VP, _VP]),
}

# int (*abl_reduce_hook)(void *user, long long *ints, int n_ints, double *reals, int n_reals)
REDUCE_HOOK = C.CFUNCTYPE(C.c_int, _VP, C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_double), C.c_int)

_lib = None


def load_library(path=None):
    """Loads libabl_cuda.so (RTLD_GLOBAL so model libraries resolve against it)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or RUNTIME_LIB
    if not os.path.exists(path):
        raise AblError("CUDA runtime library %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load_library().abl_cuda_last_error().decode(errors="replace")
        raise AblError("%s failed (code %d): %s" % (what or "abl_cuda call", rc, msg))


class Runtime:
    """One device runtime (one GPU)."""

    def __init__(self, use_float=False, device=-1, block_size=0, tile=False, seed=None):
        self.lib = load_library()
        cfg = Config()
        self.lib.abl_cuda_default_config(C.byref(cfg))
        cfg.device = device
        cfg.use_float = 1 if use_float else 0
        cfg.block_size = block_size
        cfg.tile_neighbours = 1 if tile else 0
        if seed is not None:
            cfg.seed = seed
        self.handle = _VP()
        check(self.lib.abl_cuda_create(C.byref(self.handle), C.byref(cfg)), "abl_cuda_create")
        self.use_float = use_float

    def close(self):
        if self.handle:
            check(self.lib.abl_cuda_destroy(self.handle), "abl_cuda_destroy")
            self.handle = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # thin wrappers -------------------------------------------------------------------
    def set_environment(self, env_min, env_max, granularity):
        dim = len(env_max)
        lo = (C.c_double * 3)(*(list(env_min) + [0.0] * (3 - dim)))
        hi = (C.c_double * 3)(*(list(env_max) + [0.0] * (3 - dim)))
        check(self.lib.abl_cuda_set_environment(self.handle, dim, lo, hi, granularity), "set_environment")

    def upload(self, pool, array):
        check(self.lib.abl_cuda_upload(self.handle, pool, array.ctypes.data_as(_VP), len(array)), "upload")

    def pool_size(self, pool):
        n = C.c_size_t()
        check(self.lib.abl_cuda_pool_size(self.handle, pool, C.byref(n)), "pool_size")
        return n.value

    def download(self, pool, dtype):
        import numpy as np
        n = self.pool_size(pool)
        out = np.zeros(n, dtype=dtype)
        got = C.c_size_t()
        check(self.lib.abl_cuda_download(self.handle, pool, out.ctypes.data_as(_VP), n, C.byref(got)), "download")
        return out

    def download_ids(self, pool):
        import numpy as np
        n = self.pool_size(pool)
        out = np.zeros(n, dtype=np.uint32)
        got = C.c_size_t()
        check(self.lib.abl_cuda_download_ids(self.handle, pool, out.ctypes.data_as(_VP), n, C.byref(got)), "download_ids")
        return out

    def pin_host(self, array):
        """Page-locks the memory of a numpy array for direct DMA (idempotent; best effort)."""
        check(self.lib.abl_cuda_pin_host(self.handle, array.ctypes.data_as(_VP), array.nbytes), "pin_host")

    def download_into(self, pool, out):
        """abl_cuda_download into a caller-provided (ideally page-locked) structured array; -> records written."""
        got = C.c_size_t()
        check(self.lib.abl_cuda_download(self.handle, pool, out.ctypes.data_as(_VP), len(out), C.byref(got)), "download")
        return got.value

    # ---- scalable upload under slab decomposition (abl_cuda.h) ----------------------------
    def transit_record_bytes(self, pool):
        n = C.c_size_t()
        check(self.lib.abl_cuda_transit_record_bytes(self.handle, pool, C.byref(n)), "transit_record_bytes")
        return n.value

    def partition_upload(self, pool, array, first_id, dev_out, n_slabs):
        """Host records `array` (ids first_id ...) -> device buffer `dev_out` (a device pointer with room for
        len(array) transit records), grouped by owning slab.  -> records per slab."""
        counts = (C.c_uint * n_slabs)()
        check(self.lib.abl_cuda_partition_upload(self.handle, pool, array.ctypes.data_as(_VP), len(array), int(first_id),
                                                 C.c_void_p(dev_out), counts), "partition_upload")
        return [int(c) for c in counts]

    def adopt_records(self, pool, dev_records, n, next_id):
        check(self.lib.abl_cuda_adopt_records(self.handle, pool, C.c_void_p(dev_records), int(n), int(next_id)), "adopt_records")

    # ---- slab decomposition ------------------------------------------------------------
    def slab_layers(self):
        n = C.c_int()
        check(self.lib.abl_cuda_slab_axis_layers(self.handle, C.byref(n)), "slab_axis_layers")
        return n.value

    def set_slab(self, bounds, my_slab):
        """bounds: [(begin, end)] per slab (contiguous)."""
        table = [b for b, _ in bounds] + [bounds[-1][1]]
        arr = (C.c_int * len(table))(*table)
        check(self.lib.abl_cuda_set_slab(self.handle, arr, len(bounds), my_slab), "set_slab")

    def set_local_peers(self, lower, upper):
        check(self.lib.abl_cuda_set_local_peers(self.handle, lower.handle if lower else None,
                                                upper.handle if upper else None), "set_local_peers")

    def exchange(self, pool):
        check(self.lib.abl_cuda_exchange(self.handle, pool), "exchange")

    # direct halo transport (peer memory): records are written by the step kernels themselves
    def halo_setup(self, pool, capacity_records=0):
        """Allocates the receive area of `pool`; returns its 64-byte CUDA IPC handle."""
        buf = (C.c_ubyte * 64)()
        check(self.lib.abl_cuda_halo_setup(self.handle, pool, capacity_records, buf), "halo_setup")
        return bytes(buf)

    def halo_connect(self, pool, lower_handle, upper_handle):
        lo = (C.c_ubyte * 64).from_buffer_copy(lower_handle) if lower_handle else None
        hi = (C.c_ubyte * 64).from_buffer_copy(upper_handle) if upper_handle else None
        check(self.lib.abl_cuda_halo_connect(self.handle, pool, lo, hi), "halo_connect")

    def halo_connect_local(self, pool, lower, upper):
        check(self.lib.abl_cuda_halo_connect_local(self.handle, pool, lower.handle if lower else None,
                                                   upper.handle if upper else None), "halo_connect_local")

    def exchange_begin(self, pool):
        check(self.lib.abl_cuda_exchange_begin(self.handle, pool), "exchange_begin")

    def exchange_end(self, pool):
        check(self.lib.abl_cuda_exchange_end(self.handle, pool), "exchange_end")

    def init_nccl(self, unique_id, rank, world):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        check(self.lib.abl_cuda_comm_init_nccl(self.handle, buf, rank, world), "comm_init_nccl")

    def nccl_unique_id(self):
        buf = (C.c_ubyte * 128)()
        check(self.lib.abl_cuda_nccl_unique_id(buf), "nccl_unique_id")
        return bytes(buf)

    def step(self, step_id):
        check(self.lib.abl_cuda_step(self.handle, step_id), "step")

    def begin_timestep(self):
        check(self.lib.abl_cuda_begin_timestep(self.handle), "begin_timestep")

    def end_timestep(self):
        check(self.lib.abl_cuda_end_timestep(self.handle), "end_timestep")

    # run-time add() under slab decomposition: ids of new agents are resolved across slabs
    def pending_add_parents(self):
        """-> ascending ids of the local parents of the open adding step, or None if no step is open."""
        import numpy as np
        is_open, m = C.c_int(), C.c_uint()
        check(self.lib.abl_cuda_pending_adds(self.handle, C.byref(is_open), C.byref(m)), "pending_adds")
        if not is_open.value:
            return None
        ids = np.zeros(m.value, dtype=np.uint32)
        check(self.lib.abl_cuda_pending_add_parents(self.handle, ids.ctypes.data_as(_VP), len(ids)), "pending_add_parents")
        return ids

    def resolve_adds(self, global_rank, global_total):
        import numpy as np
        r = np.ascontiguousarray(global_rank, dtype=np.uint32)
        check(self.lib.abl_cuda_resolve_adds(self.handle, r.ctypes.data_as(_VP), int(global_total)), "resolve_adds")

    def set_reduce_hook(self, fn):
        """fn(ints: list[int], reals: list[float]) -> (ints, reals) summed over all slabs; None removes the hook."""
        if fn is None:
            self._hook = None
            check(self.lib.abl_cuda_set_reduce_hook(self.handle, None, None), "set_reduce_hook")
            return

        def trampoline(_user, ints, n_ints, reals, n_reals):
            try:
                i_in = [ints[k] for k in range(n_ints)]
                r_in = [reals[k] for k in range(n_reals)]
                i_out, r_out = fn(i_in, r_in)
                for k in range(n_ints):
                    ints[k] = int(i_out[k])
                for k in range(n_reals):
                    reals[k] = float(r_out[k])
                return 0
            except Exception:   # must not propagate through the C frame
                import traceback
                traceback.print_exc()
                return 1
        self._hook = REDUCE_HOOK(trampoline)   # keep the thunk alive
        check(self.lib.abl_cuda_set_reduce_hook(self.handle, C.cast(self._hook, _VP), None), "set_reduce_hook")

    def bin(self, pool):
        check(self.lib.abl_cuda_bin(self.handle, pool), "bin")

    def synchronize(self):
        check(self.lib.abl_cuda_synchronize(self.handle), "synchronize")

    def grid_cells(self):
        n = C.c_uint()
        axes = (C.c_int * 3)()
        check(self.lib.abl_cuda_grid_cells(self.handle, C.byref(n), C.byref(axes)), "grid_cells")
        return n.value, list(axes)

    def debug_binning(self, pool):
        import numpy as np
        ncells, _ = self.grid_cells()
        n = self.pool_size(pool)
        cs = np.zeros(ncells + 1, dtype=np.uint32)
        ids = np.zeros(n, dtype=np.uint32)
        check(self.lib.abl_cuda_debug_binning(self.handle, pool, cs.ctypes.data_as(_VP), len(cs),
                                              ids.ctypes.data_as(_VP), len(ids)), "debug_binning")
        return cs, ids

    def enable_timing(self, on=True):
        check(self.lib.abl_cuda_enable_timing(self.handle, 1 if on else 0), "enable_timing")

    def time_kernel(self, step, reps=20):
        """One call of step function `step`; -> average ms of one launch of its kernel over `reps` repetitions."""
        ms = C.c_float()
        check(self.lib.abl_cuda_time_kernel(self.handle, step, reps, C.byref(ms)), "time_kernel")
        return ms.value

    def last_timing(self):
        t = StepTiming()
        check(self.lib.abl_cuda_last_timing(self.handle, C.byref(t)), "last_timing")
        return {"bin_ms": t.bin_ms, "kernel_ms": t.kernel_ms, "commit_ms": t.commit_ms, "launches": t.launches}

    def stream(self):
        return self.lib.abl_cuda_stream(self.handle)

    def count(self, pool):
        r = C.c_int()
        check(self.lib.abl_cuda_count(self.handle, pool, C.byref(r)), "count")
        return r.value

    def sum_int(self, pool, member):
        r = C.c_int()
        check(self.lib.abl_cuda_sum_int(self.handle, pool, member, C.byref(r)), "sum_int")
        return r.value

    def count_member_int(self, pool, member, value):
        r = C.c_int()
        check(self.lib.abl_cuda_count_member_int(self.handle, pool, member, int(value), C.byref(r)), "count_member_int")
        return r.value

    def sum_float(self, pool, member, component=0):
        r = C.c_double()
        check(self.lib.abl_cuda_sum_float(self.handle, pool, member, component, C.byref(r)), "sum_float")
        return r.value
