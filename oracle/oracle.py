"""ctypes front end of the plain-C oracle (oracle/abl_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of abl_oracle.c for who may use it.
Each model exposes init() (the reference's sequential main() set-up) and run()
(T timesteps of the reference simulate loop) on numpy structured arrays whose layout is
the host record layout (openabl_b200.state.agent_dtype).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BRUTE, GRID = 0, 1


def _lib(use_float):
    path = os.path.join(HERE, "_build", "liboracle_f32.so" if use_float else "liboracle_f64.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    return C.CDLL(path)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, use_float=False):
        self.use_float = use_float
        self.lib = _lib(use_float)
        self.real = np.float32 if use_float else np.float64
        self.lib.oracle_fold6.restype = C.c_double
        self.lib.oracle_fold6.argtypes = [C.c_double]

    def set_threads(self, n):
        """OpenMP threads of the parallel loops (torchrun exports OMP_NUM_THREADS=1); -> threads in use."""
        self.lib.oracle_set_threads(C.c_int(int(n)))
        return int(self.lib.oracle_omp_threads())

    def threads(self):
        return int(self.lib.oracle_omp_threads())

    def reset_rng(self):
        self.lib.oracle_rng_reset()

    # ---- circle / circle3d ------------------------------------------------------------
    def circle_dtype(self, dim):
        return np.dtype([("pos", self.real, (dim,))], align=True)

    def circle_init(self, dim, n, rho=0.05):
        self.reset_rng()
        a = np.zeros(n, dtype=self.circle_dtype(dim))
        self.lib.oracle_circle_init(C.c_int(dim), C.c_int(n), C.c_double(rho), _ptr(a))
        return a

    def circle_run(self, dim, state, steps, mode=BRUTE, rho=0.05, k_rep=0.05, k_att=0.01, r=5.0,
                   sample=None, num_agents=None, granularity=0.0):
        n = len(state)
        num_agents = n if num_agents is None else num_agents
        cur, nxt = state.copy(), state.copy()
        i0, i1 = sample if sample else (0, n)
        for _ in range(steps):
            self.lib.oracle_circle_step_g(C.c_int(dim), C.c_int(num_agents), C.c_double(rho), C.c_double(k_rep),
                                          C.c_double(k_att), C.c_double(r), _ptr(cur), _ptr(nxt),
                                          C.c_int(n), C.c_int(mode), C.c_int(i0), C.c_int(i1),
                                          C.c_double(granularity))
            cur, nxt = nxt, cur
        return cur

    # ---- boids2d ----------------------------------------------------------------------
    def boids_dtype(self):
        return np.dtype([("pos", self.real, (2,)), ("velocity", self.real, (2,))], align=True)

    def boids_init(self, n, agent_density=500):
        self.reset_rng()
        a = np.zeros(n, dtype=self.boids_dtype())
        self.lib.oracle_boids2d_init(C.c_int(n), C.c_int(agent_density), _ptr(a))
        return a

    def boids_run(self, state, steps, mode=BRUTE, agent_density=500, interaction_radius=0.05,
                  separation_radius=0.005, sample=None, num_agents=None):
        n = len(state)
        num_agents = n if num_agents is None else num_agents
        cur, nxt = state.copy(), state.copy()
        i0, i1 = sample if sample else (0, n)
        for _ in range(steps):
            self.lib.oracle_boids2d_step(C.c_int(num_agents), C.c_int(agent_density),
                                         C.c_double(interaction_radius), C.c_double(separation_radius),
                                         _ptr(cur), _ptr(nxt), C.c_int(n), C.c_int(mode),
                                         C.c_int(i0), C.c_int(i1))
            cur, nxt = nxt, cur
        return cur

    # ---- game_of_life -----------------------------------------------------------------
    def gol_dtype(self):
        dt = np.dtype([("pos", self.real, (2,)), ("alive", np.bool_)], align=True)
        assert dt.itemsize == self.lib.oracle_gol_record_size()
        return dt

    def gol_init(self, num_agents, alive_fraction=0.1):
        self.reset_rng()
        size = self.lib.oracle_gol_size(C.c_int(num_agents))
        a = np.zeros(size * size, dtype=self.gol_dtype())
        self.lib.oracle_gol_init(C.c_int(num_agents), C.c_double(alive_fraction), _ptr(a))
        return a

    def gol_run(self, state, steps, mode=BRUTE, num_agents=None, sample=None):
        n = len(state)
        num_agents = n if num_agents is None else num_agents
        cur, nxt = state.copy(), state.copy()
        i0, i1 = sample if sample else (0, n)
        for _ in range(steps):
            self.lib.oracle_gol_step(C.c_int(num_agents), _ptr(cur), _ptr(nxt), C.c_int(n), C.c_int(mode),
                                     C.c_int(i0), C.c_int(i1))
            cur, nxt = nxt, cur
        return cur

    # ---- by fixture name ---------------------------------------------------------------
    def init_for(self, model, params):
        n = params["num_agents"]
        if model == "circle.abl":
            return self.circle_init(2, n)
        if model == "circle3d.abl":
            return self.circle_init(3, n)
        if model == "boids2d.abl":
            return self.boids_init(n)
        if model == "game_of_life.abl":
            return self.gol_init(n)
        raise KeyError(model)

    def run_for(self, model, params, state, steps, mode):
        if model == "circle.abl":
            return self.circle_run(2, state, steps, mode)
        if model == "circle3d.abl":
            return self.circle_run(3, state, steps, mode)
        if model == "boids2d.abl":
            return self.boids_run(state, steps, mode)
        if model == "game_of_life.abl":
            return self.gol_run(state, steps, mode, num_agents=params["num_agents"])
        raise KeyError(model)


class PredatorPreyOracle:
    """predator_prey.abl — semantics frozen by this repository (parity unpinned by the
    reference, see abl_oracle.c).  Types: 0 Predator, 1 Prey, 2 Grass."""

    def __init__(self, num_agents, use_float=False):
        self.o = Oracle(use_float)
        self.lib = self.o.lib
        self.lib.oracle_pp_create.restype = C.c_void_p
        self.lib.oracle_pp_create.argtypes = [C.c_int]
        self.lib.oracle_pp_destroy.argtypes = [C.c_void_p]
        self.lib.oracle_pp_timestep.argtypes = [C.c_void_p, C.c_int]
        self.lib.oracle_pp_count.argtypes = [C.c_void_p, C.c_int]
        self.lib.oracle_pp_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.oracle_pp_sum_avail.argtypes = [C.c_void_p]
        self.w = self.lib.oracle_pp_create(num_agents)
        real = self.o.real
        animal = np.dtype([("pos", real, (2,)), ("dir", real, (2,)), ("steer", real, (2,)), ("life", np.int32)], align=True)
        grass = np.dtype([("pos", real, (2,)), ("dead_cycles", np.int32), ("avail", np.bool_)], align=True)
        self.dtypes = [animal, animal, grass]
        for t in range(3):
            assert self.dtypes[t].itemsize == self.lib.oracle_pp_record_size(t)

    def timestep(self, mode=GRID):
        self.lib.oracle_pp_timestep(self.w, mode)

    def count(self, t):
        return self.lib.oracle_pp_count(self.w, t)

    def sum_avail(self):
        return self.lib.oracle_pp_sum_avail(self.w)

    def read(self, t):
        n = self.count(t)
        rec = np.zeros(n, dtype=self.dtypes[t])
        ids = np.zeros(n, dtype=np.uint32)
        self.lib.oracle_pp_read(self.w, t, _ptr(rec), _ptr(ids))
        return ids, rec

    def close(self):
        if self.w:
            self.lib.oracle_pp_destroy(self.w)
            self.w = None
