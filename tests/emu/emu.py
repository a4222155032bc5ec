"""Sequential CPU emulation of the generated step kernels — TEST INFRASTRUCTURE.

`EmuModel` compiles the device program the code generator printed for a model
(build/models/<key>/model_kernels.cu + asset/cuda/abl_device.cuh) with g++ against the CUDA
stand-in header in this directory and runs its step launchers one simulated thread after the
other.  Binning (cell keys, (key, id) order, cell_start) is redone here in numpy with the
arithmetic of asset/cuda/abl_runtime.cu (cell_coord, abl_cuda_set_environment), so only the
generated kernels are under test.  Nothing on the product path imports this module.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

from openabl_b200 import build as _build
from openabl_b200.paths import ASSET_DIR
from openabl_b200.runtime import ABL_MAX_COLUMNS, AgentDesc, load_library
from openabl_b200.state import agent_dtype, parse_agents

HERE = os.path.dirname(os.path.abspath(__file__))


class _HostArray(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


class _HostType(C.Structure):
    _fields_ = [("desc", AgentDesc), ("agents", C.POINTER(_HostArray)), ("pool", C.c_int)]


class PoolView(C.Structure):      # abl_pool_view
    _fields_ = [("n", C.c_uint), ("cin", C.c_void_p * ABL_MAX_COLUMNS), ("cout", C.c_void_p * ABL_MAX_COLUMNS),
                ("id", C.c_void_p), ("cell_start", C.c_void_p)]


class GridView(C.Structure):      # abl_grid_view
    _fields_ = [("dim", C.c_int), ("n_cell", C.c_int * 3), ("origin", C.c_double * 3), ("cell_size", C.c_double),
                ("inv_cell_size", C.c_double), ("n_cells", C.c_uint), ("axis_lo", C.c_int), ("axis_hi", C.c_int),
                ("key_base", C.c_uint)]


def _generate(abl_path, params, config):
    """Generated sources + host objects of a model WITHOUT the nvcc build (the emulator does not
    need the device code): reuses a full build when one exists, else runs the code generator and
    gcc into build/emu/<key>/."""
    key = _build.model_key(abl_path, params, config)
    full = os.path.join(_build.MODEL_CACHE, key)
    if os.path.exists(os.path.join(full, "libmodel.so")):
        return full
    out = os.path.join(os.path.dirname(_build.MODEL_CACHE), "emu", key)
    if os.path.exists(os.path.join(out, "abl_host.o")):
        return out
    _build.build_compiler()
    _build.build_runtime()
    os.makedirs(out, exist_ok=True)
    cmd = [_build.COMPILER, "-i", abl_path, "-b", "cuda", "-o", out, "-A", ASSET_DIR]
    for k, v in params.items():
        cmd += ["-P", "%s=%s" % (k, _build._fmt(v))]
    for k, v in config.items():
        cmd += ["-C", "%s=%s" % (k, _build._fmt(v))]
    _build._run(cmd, cwd=_build.REPO_ROOT)
    define = ["-DLIBABL_USE_FLOAT=1", "-DABL_USE_FLOAT=1"] if config.get("use_float") else []
    with open(os.path.join(out, "build.sh")) as f:
        script = f.read()
    # the host objects exactly as build.sh compiles them (same flags and defines)
    for line in script.splitlines():
        if line.startswith("$CC ") and ("-DABL_MODEL_NO_MAIN" in line or "abl_host.c" in line):
            _build._run(["sh", "-c", "CC=gcc; " + line], cwd=out)
    return out


def _build_emu(model_dir):
    lib = os.path.join(model_dir, "libmodel_emu.so")
    srcs = [os.path.join(HERE, "cuda_runtime.h"), os.path.join(HERE, "emu_kernels.cpp"),
            os.path.join(model_dir, "model_kernels.cu"), os.path.join(model_dir, "abl_device.cuh")]
    if os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in srcs):
        return lib
    rt_dir = os.path.join(ASSET_DIR, "cuda")
    tmp = "%s.%d.tmp" % (lib, os.getpid())   # (parallel test workers: never expose a half-written library)
    cmd = ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-fPIC", "-w", "-I", HERE, "-I", model_dir, "-shared",
           "-o", tmp, "-x", "c++", os.path.join(HERE, "emu_kernels.cpp"), "-x", "none",
           os.path.join(model_dir, "model_host_lib.o"), os.path.join(model_dir, "abl_host.o"),
           "-L", rt_dir, "-labl_cuda", "-ldl", "-Wl,-rpath," + rt_dir]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + proc.stdout)
    os.replace(tmp, lib)
    return lib


class _Pool:
    """SoA columns of one agent type, in the column order of the device runtime: bool -> u8,
    int -> i32, float -> real, float2 -> one packed (n, 2) column, float3 -> three scalar columns."""

    def __init__(self, members, dtype, real):
        self.members, self.dtype, self.real = members, dtype, real
        self.cols, self.ids = [], np.zeros(0, np.uint32)
        self.cell_start = None
        self.pos_member = next((k for k, m in enumerate(members) if m[2]), -1)
        self.first_col = []
        c = 0
        for _, ty, _ in members:
            self.first_col.append(c)
            c += 3 if ty == "float3" else 1
        self.n_cols = c

    def load(self, records, ids=None):
        self.cols = []
        for name, ty, _ in self.members:
            v = records[name]
            if ty == "bool":
                self.cols.append(np.ascontiguousarray(v.astype(np.uint8)))
            elif ty == "float3":
                for k in range(3):
                    self.cols.append(np.ascontiguousarray(v[:, k]))
            else:
                self.cols.append(np.ascontiguousarray(v))
        self.ids = np.arange(len(records), dtype=np.uint32) if ids is None else ids.astype(np.uint32)
        self.next_id = int(self.ids.max()) + 1 if len(self.ids) else 0
        self.cell_start = None

    def records(self):
        """Host records in ascending id order (what abl_cuda_download returns)."""
        order = np.argsort(self.ids, kind="stable")
        out = np.zeros(len(self.ids), dtype=self.dtype)
        for k, (name, ty, _) in enumerate(self.members):
            c = self.first_col[k]
            if ty == "bool":
                out[name] = self.cols[c][order] != 0
            elif ty == "float3":
                for q in range(3):
                    out[name][:, q] = self.cols[c + q][order]
            else:
                out[name] = self.cols[c][order]
        return out

    def position(self):
        k = self.pos_member
        name, ty, _ = self.members[k]
        c = self.first_col[k]
        if ty == "float2":
            return [self.cols[c][:, 0], self.cols[c][:, 1]]
        return [self.cols[c], self.cols[c + 1], self.cols[c + 2]]

    def permute(self, order):
        self.cols = [np.ascontiguousarray(c[order]) for c in self.cols]
        self.ids = np.ascontiguousarray(self.ids[order])


def modes(kernels):
    """ABL_MODE template arguments of a set of mangled kernel names."""
    import re
    return sorted({int(m.group(1)) for k in kernels for m in [re.search(r"ILi(\d+)E", k)] if m})


class EmuModel:
    def __init__(self, abl_path, params=None, use_float=False, config=None, lib_path=None):
        cfg = dict(config or {})
        if use_float:
            cfg["use_float"] = True
        self.use_float = bool(use_float)
        self.real = np.float32 if use_float else np.float64
        self.dir = _generate(abl_path, dict(params or {}), cfg)
        load_library()   # the host part of the model resolves against libabl_cuda.so (never called here)
        self.lib = C.CDLL(lib_path or _build_emu(self.dir), mode=C.RTLD_LOCAL)
        with open(abl_path) as f:
            self.agents = parse_agents(f.read())
        self.dtypes = [agent_dtype(m, self.use_float) for _, m in self.agents]
        self.n_types = C.c_int.in_dll(self.lib, "abl_model_n_types").value
        self._types = (_HostType * (self.n_types + 1)).in_dll(self.lib, "abl_model_types")
        assert self.lib.emu_real_size() == (4 if use_float else 8)
        assert self.lib.emu_setup() == 0
        self.n_steps = C.c_int.in_dll(self.lib, "abl_model_n_steps").value
        self.lib.emu_run_step.argtypes = [C.c_int, C.POINTER(PoolView), C.POINTER(PoolView), C.POINTER(GridView),
                                          C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p * ABL_MAX_COLUMNS),
                                          C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_ulonglong, C.c_uint, C.c_int, C.c_int]
        self.lib.emu_set_cost.argtypes = [C.c_int, C.c_float]
        self.lib.emu_set_nlist.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
        self.use_nlist = False     # drive steps registered with desc.nlist through cached neighbour lists
        self.nlists = {}           # step -> (ids of self pool, ids of nbr pool, cnt, idx, stride)
        self.nlist_builds = 0
        self.steps = []
        for s in range(self.n_steps):
            sp, nb, ap, ur = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            rad, wr = C.c_double(), C.c_uint()
            assert self.lib.emu_step_info(s, C.byref(sp), C.byref(nb), C.byref(rad), C.byref(wr), C.byref(ur), C.byref(ap)) == 0
            self.steps.append({"self": sp.value, "nbr": nb.value, "radius": rad.value, "written": wr.value,
                               "removal": ur.value, "added": ap.value})
        dim, cell = C.c_int(), C.c_double()
        lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
        self.grid = None
        if self.lib.emu_environment(C.byref(dim), lo, hi, C.byref(cell)) == 0:
            self.grid = self._grid(dim.value, list(lo), list(hi), cell.value)
        self.pools = [_Pool(m, dt, self.real) for (_, m), dt in zip(self.agents, self.dtypes)]
        self.timestep_no = 0
        self.seed = 0x0123456789abcdef     # abl_cuda_default_config
        self.block_size = 0
        self.flat_loop = 0         # abl_step_launch.flat_loop: 0 cursor loop, 1 flat loop, -1 timed by the launcher
        self.check_fused_histogram = True
        self.fused_checked = 0
        self.lib.emu_last_kernel_name.restype = C.c_char_p
        self.kernels = set()     # mangled names of the kernels launched (…ILi<ABL_MODE>EE…)

    # abl_cuda_set_environment (asset/cuda/abl_runtime.cu)
    def _grid(self, dim, lo, hi, cell):
        g = GridView()
        g.dim = dim
        cell = cell * (1.0 + (2.0 ** -10 if self.use_float else 2.0 ** -20))   # ABL_CELL_PAD_* (abl_cuda.h)
        g.cell_size = cell
        g.inv_cell_size = 1.0 / cell if not self.use_float else float(np.float32(1.0) / np.float32(cell))
        cells = 1
        for a in range(3):
            if a < dim:
                g.n_cell[a] = max(1, int(math.ceil((hi[a] - lo[a]) / cell)))
                g.origin[a] = lo[a]
            else:
                g.n_cell[a] = 1
                g.origin[a] = 0.0
            cells *= g.n_cell[a]
        g.n_cells = cells
        g.axis_lo, g.axis_hi, g.key_base = 0, g.n_cell[dim - 1], 0
        return g

    def populate(self):
        self.lib.abl_model_populate()
        for t in range(self.n_types):
            arr = self._types[t].agents.contents
            n = arr.len
            if n:
                buf = (C.c_char * (n * self.dtypes[t].itemsize)).from_address(arr.data)
                rec = np.frombuffer(buf, dtype=self.dtypes[t]).copy()
            else:
                rec = np.zeros(0, dtype=self.dtypes[t])
            self.pools[t].load(rec)

    def host_agents(self, t):
        return self.pools[t].records()

    # cell_coord + key of asset/cuda/abl_runtime.cu (bin_count_one), in the precision of abl_real
    def keys(self, pos):
        g, R = self.grid, self.real
        key = np.zeros(len(pos[0]), dtype=np.int64)
        mul = 1
        for a in range(g.dim):
            c = np.floor((pos[a] - R(g.origin[a])) * R(g.inv_cell_size))
            c = np.clip(np.nan_to_num(c, nan=0.0, posinf=2.0 ** 31, neginf=-2.0 ** 31), 0, g.n_cell[a] - 1).astype(np.int64)
            key += c * mul
            mul *= g.n_cell[a]
        return key

    def bin(self, t):
        p = self.pools[t]
        if p.pos_member < 0 or self.grid is None:
            return
        key = self.keys(p.position())
        order = np.lexsort((p.ids, key))
        p.permute(order)
        counts = np.bincount(key, minlength=self.grid.n_cells)
        p.cell_start = np.zeros(self.grid.n_cells + 2, dtype=np.uint32)
        p.cell_start[1:self.grid.n_cells + 1] = np.cumsum(counts)
        p.cell_start[self.grid.n_cells + 1] = p.cell_start[self.grid.n_cells]

    def _column_protos(self, p):
        """Empty arrays with the dtype / trailing shape of each column of pool `p`."""
        protos = []
        for _, ty, _ in p.members:
            if ty == "bool":
                protos.append(np.zeros(0, np.uint8))
            elif ty == "int":
                protos.append(np.zeros(0, np.int32))
            elif ty == "float":
                protos.append(np.zeros(0, self.real))
            elif ty == "float2":
                protos.append(np.zeros((0, 2), self.real))
            else:
                protos += [np.zeros(0, self.real)] * 3
        return protos

    def _view(self, p, out_cols):
        v = PoolView()
        v.n = len(p.ids)
        for c in range(p.n_cols):
            v.cin[c] = p.cols[c].ctypes.data
            v.cout[c] = (out_cols[c] if out_cols is not None else p.cols[c]).ctypes.data
        v.id = p.ids.ctypes.data
        v.cell_start = p.cell_start.ctypes.data if p.cell_start is not None else None
        return v

    def run_step(self, s, own_range=None):
        """own_range = (begin, end): step only that part of the (binned) pool, like the runtime does
        under slab decomposition (views offset to the first owned record); no commit stages."""
        st = self.steps[s]
        me = self.pools[st["self"]]
        nb = self.pools[st["nbr"]] if st["nbr"] >= 0 else None
        if nb is not None:
            self.bin(st["nbr"])
        n = len(me.ids)
        out = [c.copy() for c in me.cols]
        dead = np.zeros(n, dtype=np.uint8) if st["removal"] else None
        # run-time add(): one staging slot per parent in the column layout of the added type
        target = self.pools[st["added"]] if st["added"] >= 0 else None
        add_flag, staging, staging_ptrs = None, None, None
        if target is not None:
            add_flag = np.zeros(n, dtype=np.uint8)
            staging = [np.zeros((n,) + c.shape[1:], dtype=c.dtype) for c in self._column_protos(target)]
            staging_ptrs = (C.c_void_p * ABL_MAX_COLUMNS)()
            for c, col in enumerate(staging):
                staging_ptrs[c] = col.ctypes.data
        writes_pos = me.pos_member >= 0 and (st["written"] >> me.pos_member) & 1
        fuse = bool(writes_pos and self.grid is not None and not st["removal"] and self.check_fused_histogram)
        bk = bl = bc = None
        if fuse:
            bk = np.full(n, 0xffffffff, dtype=np.uint32)
            bl = np.full(n, 0xffffffff, dtype=np.uint32)
            bc = np.zeros(self.grid.n_cells + 2, dtype=np.uint32)
        reach = 1
        if nb is not None:
            reach = max(1, int(math.ceil(st["radius"] / self.grid.cell_size - 1e-12)))
        sv = self._view(me, out)
        if own_range is not None:
            b, e = own_range
            sv.n = e - b
            for c in range(me.n_cols):
                sv.cin[c] = me.cols[c][b:e].ctypes.data
                sv.cout[c] = out[c][b:e].ctypes.data
            sv.id = me.ids[b:e].ctypes.data
            fuse, bk, bl, bc = False, None, None, None
        nv = self._view(nb, None) if nb is not None else None
        grid = self.grid if self.grid is not None else GridView()
        listed = self.use_nlist and nb is not None and self.lib.emu_step_nlist(s)
        if listed:
            # the runtime's protocol (abl_cuda_step): count pass, size the lists by the largest count,
            # fill pass; rebuilt whenever the order of either pool changed
            cached = self.nlists.get(s)
            if cached is None or not (np.array_equal(cached[0], me.ids) and np.array_equal(cached[1], nb.ids)):
                cnt = np.zeros(n, dtype=np.uint32)
                mx = np.zeros(1, dtype=np.uint32)
                self.lib.emu_set_nlist(1, cnt.ctypes.data, None, 0, mx.ctypes.data)
                assert self.lib.emu_run_step(s, C.byref(sv), C.byref(nv), C.byref(grid), reach, None, None, None,
                                             None, None, None, self.seed, self.timestep_no, self.block_size, 0) == 0
                self.kernels.add(self.lib.emu_last_kernel_name().decode())
                assert int(mx[0]) == int(cnt.max()) if n else True
                stride = n
                idx = np.full(max(1, int(mx[0])) * stride, 0xffffffff, dtype=np.uint32)
                self.lib.emu_set_nlist(2, cnt.ctypes.data, idx.ctypes.data, stride, mx.ctypes.data)
                assert self.lib.emu_run_step(s, C.byref(sv), C.byref(nv), C.byref(grid), reach, None, None, None,
                                             None, None, None, self.seed, self.timestep_no, self.block_size, 0) == 0
                self.kernels.add(self.lib.emu_last_kernel_name().decode())
                for c, o in zip(me.cols, out):
                    assert np.array_equal(c, o, equal_nan=True) if c.dtype.kind == "f" else np.array_equal(c, o), \
                        "a list-building launch stored something"
                cached = (me.ids.copy(), nb.ids.copy(), cnt, idx, stride)
                self.nlists[s] = cached
                self.nlist_builds += 1
            self.lib.emu_set_nlist(0, cached[2].ctypes.data, cached[3].ctypes.data, cached[4], None)
        else:
            self.lib.emu_set_nlist(0, None, None, 0, None)
        rc = self.lib.emu_run_step(s, C.byref(sv), C.byref(nv) if nv is not None else None, C.byref(grid), reach,
                                   dead.ctypes.data if dead is not None else None,
                                   add_flag.ctypes.data if add_flag is not None else None,
                                   C.byref(staging_ptrs) if staging_ptrs is not None else None,
                                   bk.ctypes.data if fuse else None, bl.ctypes.data if fuse else None,
                                   bc.ctypes.data if fuse else None, self.seed, self.timestep_no, self.block_size, self.flat_loop)
        assert rc == 0, "emulated launch of step %d failed (%d)" % (s, rc)
        if n:
            self.kernels.add(self.lib.emu_last_kernel_name().decode())
        me.cols = out
        me.cell_start = None
        if fuse:
            # the fused epilogue must produce the next binning's keys and histogram (the slots inside a
            # cell segment are drawn by k_bin_scatter, not here)
            key = self.keys(me.position())
            assert np.array_equal(bk.astype(np.int64), key), "fused cell keys differ from k_bin_count's"
            counts = np.bincount(key, minlength=self.grid.n_cells)
            assert np.array_equal(bc[:self.grid.n_cells].astype(np.int64), counts), "fused histogram differs"
            self.fused_checked += 1
        # commit (asset/cuda/abl_runtime.cu: commit_adds, then commit_removals): new agents are
        # appended in ascending parent-id order with ids next_id + rank; removal is a stable compaction
        if target is not None and add_flag.any():
            parents = np.nonzero(add_flag)[0]
            parents = parents[np.argsort(me.ids[parents], kind="stable")]
            m = len(parents)
            target.cols = [np.ascontiguousarray(np.concatenate([tc, sc[parents]])) for tc, sc in zip(target.cols, staging)]
            target.ids = np.concatenate([target.ids, (target.next_id + np.arange(m)).astype(np.uint32)])
            target.next_id += m
            target.cell_start = None
            if target is me and dead is not None:
                dead = np.concatenate([dead, np.zeros(m, dtype=np.uint8)])
        if dead is not None and dead.any():
            keep = dead == 0
            me.cols = [np.ascontiguousarray(c[keep]) for c in me.cols]
            me.ids = np.ascontiguousarray(me.ids[keep])

    def timestep(self):
        for s in range(self.n_steps):
            self.run_step(s)
        self.timestep_no += 1
