#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native OpenABL `cuda` backend.

Metric (BASELINE.json): agent-steps/s.  A "step" is one simulation timestep of the model:
binning of the population (cell-key counting sort + cell ranges) + every step kernel +
commit.  Workload at N=1: boids2d.abl with 1 M agents in double precision (BASELINE
configs[1]); `--workload` selects the other configurations.

  value      device-resident throughput: population already in HBM, K timesteps timed with
             CUDA events on the runtime's stream.
  e2e        the same metric through the C-ABI calls the generated program makes for
             `simulate(K)`: upload of the host AoS records, K timesteps, download back to
             host records in id order — all inside the timed region.
  roofline   step kernel: algorithmic bytes (S + M + S per agent) / average
             kernel time (CUDA events), against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the plain-C oracle (restatement of the reference `c` backend, brute-force
             O(N^2) loop, OpenMP on all host cores) on a bounded sample of the same workload.

`--impl reference` times only that CPU path (there is no GPU work in it).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (model, params, use_float, S bytes/agent, M bytes/agent read per neighbour, P pos bytes)
    # S/M/P as defined in SURVEY.md §8d; whole-timestep algorithmic bytes = P + M + 4S + 32,
    # except game_of_life whose step does not move agents: binning is hoisted out of the
    # loop (the runtime detects it) and only the step kernel's S + M + S remain.
    "boids2d-1M-f64": ("boids2d.abl", {"num_agents": 1000000}, False, 32, 32, 16),
    "boids2d-1M-f32": ("boids2d.abl", {"num_agents": 1000000}, True, 16, 16, 8),
    "boids2d-4M-f64": ("boids2d.abl", {"num_agents": 4000000}, False, 32, 32, 16),
    "boids2d-16M-f64": ("boids2d.abl", {"num_agents": 16000000}, False, 32, 32, 16),
    "circle3d-1M-f64": ("circle3d.abl", {"num_agents": 1000000}, False, 24, 24, 24),
    "circle3d-16M-f64": ("circle3d.abl", {"num_agents": 16000000}, False, 24, 24, 24),
    "circle3d-16M-f32": ("circle3d.abl", {"num_agents": 16000000}, True, 12, 12, 12),
    "game_of_life-16M-f64": ("game_of_life.abl", {"num_agents": 16777216}, False, 17, 17, 16),
    "circle-1000-f64": ("circle.abl", {"num_agents": 1000}, False, 16, 16, 16),
    # BASELINE configs[3]: three agent types, run-time add/remove; S/M/P of the Prey type (the
    # roofline object is only indicative for this workload)
    "predator_prey-4M-f64": ("predator_prey.abl", {"num_agents": 4000000}, False, 52, 16, 16),
}
BINNING_HOISTED = {"game_of_life-16M-f64"}


def whole_step_bytes(workload):
    _, _, _, S, M, P = WORKLOADS[workload]
    return (S + M + S) if workload in BINNING_HOISTED else (P + M + 4 * S + 32)
DEFAULT_WORKLOAD = "boids2d-1M-f64"


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML every few milliseconds while the
    timed region runs (same fields as the nvidia-smi line of B200_PROFILING.md)."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.samples = []
        self.max_mhz = None
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device]) if visible else device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        nv = self.nvml
        if nv is None:
            return
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.002)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}


def ncu_traffic(workload, variant):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel VARIANT this run
    used, from the `ncu --set full` captures summarised under profiles/ (profiles/traffic.json,
    written by profiles/summarize.py --traffic); (None, None) when that variant was never captured."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        table = json.load(f)
    hit = table.get("%s/%s" % (workload, variant))
    return (hit["bytes"], hit["source"]) if hit else (None, None)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(workload, budget_pairs=1.5e11, num_agents=None):
    """Times the oracle's brute-force (reference-order) step on a bounded sample of agents.
    `num_agents`: population of the run this baseline stands next to (weak scaling multiplies the
    workload's by the number of GPUs).  The OpenMP thread count is SET here (torchrun exports
    OMP_NUM_THREADS=1 to its children) and the number reported is the one the library uses."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from oracle import BRUTE, Oracle
    model, params, use_float, _, _, _ = WORKLOADS[workload]
    if model not in ("boids2d.abl", "circle.abl", "circle3d.abl", "game_of_life.abl"):
        raise SystemExit("no CPU baseline for %s: the reference `c` backend rejects run-time add/remove "
                         "(use --no-cpu-baseline)" % model)
    if num_agents:
        params = dict(params, num_agents=int(num_agents))
    n = params["num_agents"]
    o = Oracle(use_float)
    cores = o.set_threads(host_threads())
    if model == "game_of_life.abl":
        size = int(n ** 0.5)
        n = size * size
    state = o.init_for(model, params)
    n = len(state)
    sample = int(max(256, min(n, budget_pairs / n)))
    t0 = time.perf_counter()
    if model == "boids2d.abl":
        o.boids_run(state, 1, BRUTE, sample=(0, sample))
    elif model == "circle.abl":
        o.circle_run(2, state, 1, BRUTE, sample=(0, sample))
    elif model == "circle3d.abl":
        o.circle_run(3, state, 1, BRUTE, sample=(0, sample))
    else:
        o.gol_run(state, 1, BRUTE, num_agents=params["num_agents"], sample=(0, sample))
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": "reference brute-force step (all %d candidates per agent) for the first %d of %d agents, "
                      "1 timestep, OpenMP on %d threads, %.1f s" % (n, sample, n, cores, dt)}, dt


def run_reference_arm(args):
    """The reference's CPU path for the same metric and population as the repo arm (weak scaling:
    workload population x GPUs), rank 0 only, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    model, params, use_float, S, M, P = WORKLOADS[workload]
    n_agents = params["num_agents"] * (1 if args.strong else max(1, args.gpus))
    vals, times, last = [], [], None
    total = args.warmup + args.steps
    # each "step" is a bounded sample; the whole run stays within about half a minute of CPU work
    # at ~1e10 candidate tests per second (16 threads)
    budget = max(2e9, 3e11 / max(1, total))
    for i in range(total):
        base, dt = cpu_baseline(workload, budget_pairs=budget, num_agents=n_agents)
        if i >= args.warmup:
            vals.append(base["value"])
            times.append(dt)
        last = base
    value = sum(vals) / len(vals)
    last["value"] = value
    line = {"impl": "reference", "metric": "agent-steps/s", "value": value, "unit": "agent-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times),   # one step = one bounded sample (see cpu_baseline.sample)
            "higher_is_better": True, "scaling": "strong" if (args.strong and args.gpus > 1) else "weak", "vs_baseline": None,
            "dtype": "f32" if use_float else "f64", "data": "synthetic",
            "config": {"workload": workload, "model": model, "num_agents": n_agents},
            "cpu_baseline": last,
            "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


MODE_NAMES = {0: "cursor", 1: "chunked", 2: "tile", 3: "flat", 4: "neighbour-list walk", 7: "bulk-tile", 8: "shadow-prefilter", 9: "split-prefilter", -1: "not launched"}

# FP64 operations per agent-step of circle3d's step kernel (SURVEY.md 8d asks for this view next to
# the HBM one): per candidate 3 sub + 3 mul + 2 add + 1 compare = 9, per accepted candidate the
# force (normalize: sqrt + 3 IEEE divisions, scale, accumulate) ~ 80; candidates = 27 cells x
# agents per cell, accepted = (4/3 pi) / 27 of them.  Peak: 64 FP64 lanes per SM and clock.
def fp64_roofline(workload, n_agents, kernel_ms, clocks):
    if not workload.startswith("circle3d") or kernel_ms <= 0:
        return None
    candidates, accepted = 27 * 49.0, 27 * 49.0 * 0.1551
    ops = 9 * candidates + 80 * accepted
    mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = 148 * 64 * mhz * 1e6 / 1e12          # T lane-ops/s
    achieved = ops * n_agents / (kernel_ms / 1e3) / 1e12
    return {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "T FP64 lane-ops/s", "frac": achieved / peak,
            "ops_per_agent_step": ops,
            "definition": "ALGORITHMIC double-precision operations of the reference formulation (9 per candidate, 1323 candidates per "
                          "agent, + ~80 per accepted candidate, 205 per agent) / kernel time, against 148 SMs x 64 FP64 lanes x SM clock "
                          "(an FMA counted as one operation).  The shadow pre-filter variant evaluates the 9 per candidate in single "
                          "precision, so fewer FP64 instructions are EXECUTED than this figure counts (ncu: FP64 pipe 16 % busy)"}


def measure(args, workload, env, strong, steps, warmup, with_cpu_baseline):
    """One workload through the device-resident, per-stage and end-to-end passes -> JSON object."""
    import torch
    from openabl_b200.model import Model
    from openabl_b200.slab import RankSlab
    rank, local_rank, world, dist = env
    model_file, params, use_float, S, M, P = WORKLOADS[workload]
    params = dict(params)
    if not strong:
        # weak scaling: the per-GPU population stays fixed, the world (and the environment,
        # which the model derives from num_agents) grows with the number of GPUs
        params["num_agents"] = params["num_agents"] * world
    m = Model(os.path.join(REPO, "examples", model_file), params, use_float=use_float,
              config=dict(([("cuda.unroll", True)] if args.unroll else []) + ([("cuda.nlist", True)] if args.nlist else [])) or None)
    m.populate()
    n_agents = sum(m.host_count(t) for t in range(m.n_types))          # whole job, all ranks
    host_bytes = sum(m.host_count(t) * m.dtypes[t].itemsize for t in range(m.n_types))
    slab = None
    if world > 1:
        slab = RankSlab(m, rank, world, dist, device=local_rank, transport=args.transport,
                        block_size=args.block_size, tile=args.tile)
    else:
        m.create_runtime(device=local_rank, block_size=args.block_size, tile=args.tile)
    rt = m.rt
    stream = torch.cuda.ExternalStream(rt.stream(), device=torch.device("cuda", local_rank))

    def upload():
        if slab:
            slab.upload()
        else:
            m.upload_host()

    # (models with run-time add() are driven step by step under decomposition: the ids of new
    # agents are resolved across the ranks, see RankSlab.timestep)
    timestep = slab.timestep if slab else m.timestep
    mutating_slabs = bool(slab and slab._mutating)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        rt.synchronize()

    # ---- device-resident throughput ------------------------------------------------------
    # (ABL_CUDA_TUNE=1: at least 16 warm-up timesteps, because the launchers' run-time tuner then times up
    # to three candidate-loop variants over the first 12 launches of a step function; by default the
    # variant follows from a rule and the requested warm-up, at least 3 timesteps, is used as is)
    n_warmup = max(16 if os.environ.get("ABL_CUDA_TUNE", "0") not in ("", "0") else 3, warmup)
    upload()
    for _ in range(n_warmup):
        timestep()
    barrier()
    launches0 = rt.last_timing()["launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    # (1) steady state: K timesteps back to back, one event pair.  A timestep consumes the
    # previous one's output, so part of the state is still in the 126 MB L2.
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(steps):
        timestep()
    ev1.record(stream)
    barrier()
    ms_steady = ev0.elapsed_time(ev1)
    launches = rt.last_timing()["launches"] - launches0
    # (2) the headline: the same K timesteps with the L2 flushed before every one of them (a
    # 256 MB write on the runtime's stream, outside the timed intervals), each timestep between
    # its own event pair on that stream — the timing rule for inputs smaller than the L2.
    flush_bytes = 256 << 20
    if args.no_l2_flush:
        ms = ms_steady
    else:
        scratch = torch.empty(flush_bytes, dtype=torch.uint8, device=torch.device("cuda", local_rank))
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        with torch.cuda.stream(stream):
            for k, (a, b) in enumerate(pairs):
                scratch.fill_(k & 0xff)
                a.record(stream)
                timestep()
                b.record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in pairs)
        del scratch
    clocks = sampler.summary()
    t = torch.tensor([ms, ms_steady], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, ms_steady_max = float(t[0].item()), float(t[1].item())
    value = n_agents * steps / (ms_max / 1e3)
    steady = {"value": n_agents * steps / (ms_steady_max / 1e3), "unit": "agent-steps/s",
              "ms_per_step": ms_steady_max / steps,
              "definition": "the same %d timesteps back to back without the L2 flush (what a simulation run sees)" % steps}

    # ---- per-stage device times (separate pass over the same population) --------------------
    # The runtime queues four events per abl_cuda_step and evaluates them when asked, so the
    # steps of this pass run back to back like those of the timed region (no host sync, no idle
    # GPU between them); last_timing() returns the mean per step-function call.
    rt.enable_timing(True)
    stage = {"bin_ms": 0.0, "kernel_ms": 0.0, "commit_ms": 0.0}
    reps = max(10, min(200, int(200.0 / max(ms_steady_max / steps, 1e-3))))   # ~0.2 s of device time, 10..200 timesteps
    if not mutating_slabs:
        # Four event records + up to five launches per step function: the host cannot always queue them as
        # fast as the device runs them, and then every interval between two events also counts the time
        # the device waited for the host.  A spin kernel of ~20 ms ahead of the pass lets the host run ahead.
        with torch.cuda.stream(stream):
            torch.cuda._sleep(40_000_000)
        for _ in range(reps):
            for s in range(m.n_steps):
                m.run_step(s)
        lt = rt.last_timing()
        for k in stage:
            stage[k] = lt[k] * m.n_steps      # per timestep
    rt.enable_timing(False)
    # ---- the step kernel's own launch duration --------------------------------------------------
    # An event between two kernels of the per-step chain keeps the next kernel from being set up while its
    # predecessor drains (programmatic dependent launch) and adds its own idle time, so the stage intervals
    # above overstate every stage (their sum exceeds the steady-state timestep).  abl_cuda_time_kernel repeats
    # the launch of one step function 20 times between ONE pair of events on the runtime's stream, on the
    # buffers the chain has just produced (L2-warm as in a timestep, where k_bin_rank_move writes them right
    # before the kernel) and with the fused histogram epilogue; median of 7 such measurements.  Not available
    # for steps that add or remove agents and under slab decomposition (the stage interval is used there).
    stage["kernel_ms_stage_events"] = stage["kernel_ms"]
    kernel_src = "interval between two events around the launch, inside the per-step chain (stage pass)"
    if world == 1 and not any(m.step_flags(s) for s in range(m.n_steps)):
        samples = []
        for _ in range(7):
            samples.append(sum(rt.time_kernel(s, 20) for s in range(m.n_steps)))
        samples.sort()
        if samples[len(samples) // 2] > 0:
            stage["kernel_ms"] = samples[len(samples) // 2]
            kernel_src = ("abl_cuda_time_kernel: 20 launches of the kernel between one pair of CUDA events on the runtime's stream, "
                          "inputs as the chain leaves them; median of 7")
    peak, peak_src = load_peaks()
    n_local = rt.pool_size(m.pool(0)) if world > 1 else n_agents
    kernel_bytes = (S + M + S) * n_local
    achieved = kernel_bytes / (stage["kernel_ms"] / 1e3) / 1e9 if stage["kernel_ms"] > 0 else 0.0
    step_bytes = whole_step_bytes(workload) * n_agents / world
    variants = {m.step_names[s]: MODE_NAMES.get(m.step_variant(s), str(m.step_variant(s))) for s in range(m.n_steps)}
    traffic, traffic_src = ncu_traffic(workload, variants[m.step_names[0]]) if world == 1 else (None, None)
    roofline = {"bound": "hbm", "kernel": "abl_kernel_%s<%s>" % (m.step_names[0], variants[m.step_names[0]]), "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_agent": S + M + S,
                "kernel_ms": stage["kernel_ms"], "kernel_ms_source": kernel_src,
                "kernel_ms_stage_events": stage["kernel_ms_stage_events"],
                "bin_ms": stage["bin_ms"], "commit_ms": stage["commit_ms"],
                "commit_ms_definition": ("what the stream waits for after the step kernel: the halo exchange kernel runs next to the "
                                         "step kernel on a side stream (ABL_CUDA_HALO_ASYNC), this is the join" if world > 1 and args.transport == "direct"
                                         else "interval between the event behind the step kernel and the end of the step function"),
                "stage_pass_timesteps": reps,
                "whole_step_algorithmic_bytes_per_agent": whole_step_bytes(workload),
                "whole_step_frac": step_bytes * steps / (ms_max / 1e3) / 1e9 / peak}

    # ---- end to end through the C ABI with host buffers -------------------------------------
    e2e_reps = 3
    d2h = h2d = 0
    # (one untimed call first: page-locked staging buffers are allocated, host arrays registered and NCCL
    # connections made once per process, not once per simulate())
    for rep in range(e2e_reps + 1):
        if rep == 1:
            barrier()
            t0 = time.perf_counter()
        upload()
        h2d = slab.last_upload_bytes if slab else host_bytes
        for _ in range(steps):
            timestep()
        if slab:
            out = [slab.download_owned(tt) for tt in range(m.n_types)]   # this rank's owned agents, page-locked buffers
            d2h = sum(o.nbytes for o in out)
        else:
            m.download_host()
            d2h = sum(m.host_count(tt) * m.dtypes[tt].itemsize for tt in range(m.n_types))
    rt.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_reps
    te = torch.tensor([e2e_s, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = te.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        e2e_s, h2d, d2h = float(tm[0].item()), float(te[1].item()), float(te[2].item())   # bytes: all ranks
    e2e_value = n_agents * steps / e2e_s
    m.close()

    line = {"metric": "agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": steps, "warmup": n_warmup, "ms_per_step": ms_max / steps,
            "higher_is_better": True, "scaling": "strong" if (strong and world > 1) else "weak",
            "vs_baseline": None,
            "dtype": "f32" if use_float else "f64", "data": "synthetic",
            "config": {"workload": workload, "model": model_file, "num_agents": n_agents,
                       "agents_per_gpu": n_agents // world,
                       "parallelism": ("slab%d (cell layers along the slowest axis; halo + migration %s)" % (world, "written into the neighbour's HBM over NVLink by the step kernel, no host sync" if args.transport == "direct" else "over NCCL send/recv")) if world > 1 else "single",
                       "block_size": args.block_size, "neighbour_lists": bool(args.nlist),
                       "candidate_loop_in_use": variants,     # ABL_MODE of each step function's latest launch (rank 0)
                       "candidate_loop": {"0": "cursor loop (ABL_CUDA_FLAT=0)", "1": "flat loop (ABL_CUDA_FLAT=1)"}.get(
                           os.environ.get("ABL_CUDA_FLAT", ""), "chosen by the launcher (rule; timed at run time where two variants are plausible)"),
                       "l2": ("no flush: state (%.0f MB per GPU) stays partly L2-resident between timesteps" % (n_agents * S / 1e6 / world))
                             if args.no_l2_flush else
                             ("flushed: a %d MB write before every timed timestep (state: %.0f MB per GPU, L2: 126 MB); "
                              "`steady_state` is the same run without the flush" % (flush_bytes >> 20, n_agents * S / 1e6 / world))},
            "steady_state": steady,
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d / steps,
                    "d2h_bytes_per_step": d2h / steps,
                    "definition": "upload + %d timesteps + download per simulate() call; bytes summed over all ranks" % steps},
            "roofline": roofline}
    fp64 = fp64_roofline(workload, n_local, stage["kernel_ms"], clocks)
    if fp64:
        line["roofline_fp64"] = fp64
    if with_cpu_baseline and rank == 0:
        base, _ = cpu_baseline(workload, num_agents=n_agents)
        line["cpu_baseline"] = base
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true", help="report the back-to-back (steady state) time as `value`")
    ap.add_argument("--block-size", type=int, default=0, help="threads per CTA of step kernels (0 = automatic)")
    ap.add_argument("--tile", action="store_true", help="stage neighbour rows in shared memory (ABL_MODE 2 kernels)")
    ap.add_argument("--unroll", action="store_true", help="-C cuda.unroll=true: for-near candidate loop unrolled by two")
    ap.add_argument("--nlist", action="store_true", help="-C cuda.nlist=true: cached neighbour lists for step functions whose "
                    "neighbourhoods never change (game_of_life)")
    ap.add_argument("--strong", action="store_true", help="N>1: keep the total population fixed (strong scaling)")
    ap.add_argument("--no-companion", action="store_true",
                    help="skip the circle3d-16M line (the other half of BASELINE.json's metric) that the default run "
                         "adds to the boids2d line under the key `circle3d`")
    ap.add_argument("--transport", default="direct", choices=["direct", "nccl"],
                    help="N>1 halo/migration exchange: step kernels write into the neighbour's memory over "
                         "NVLink (direct) or grouped ncclSend/ncclRecv (nccl)")
    args = ap.parse_args()
    companion = args.workload is None and not args.no_companion
    if args.workload is None:
        args.workload = DEFAULT_WORKLOAD
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cuda backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = (rank, local_rank, world, dist)
    line = measure(args, args.workload, env, args.strong or world == 1, args.steps, args.warmup, not args.no_cpu_baseline)
    if companion:
        # BASELINE.json's metric names boids2d AND circle3d; configs[2] is circle3d 16 M agents under
        # slab decomposition at 1/2/4/8 GPUs with the population fixed (strong scaling).  One timestep
        # takes tens of milliseconds, so a few of them are enough.
        c = measure(args, "circle3d-16M-f64", env, True, max(3, min(args.steps, 10)), 3, False)
        line["circle3d"] = {k: c[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "dtype", "config",
                                               "steady_state", "e2e", "roofline", "roofline_fp64", "gpu_launches") if k in c}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
