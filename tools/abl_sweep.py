#!/usr/bin/env python
"""Population sweep in the file format of the reference's benchmark harness.

The reference's bench/bench.py (lines 93-136) runs every (backend, model) pair over doubling
agent counts with `OpenABL -i <model> -b <backend> -A <assets> -R -P num_agents=N -P
num_timesteps=100`, reads the `Execution time: <s>s` line (bench/openabl.py:58-63) and writes
`bench_<model>_<backend>.txt` holding `n,t` rows, which bench/plot.py:71-85 picks up by file
name.  This script does the same for the `cuda` backend, so its results drop into a result
directory next to the reference's files for the other backends and plot.py draws them together.

    python tools/abl_sweep.py -r results/ [-m circle,boids2d] [-n 250-16000000] [-M 600]
                              [-e /path/to/examples] [-C use_float=true]
"""
import argparse
import os
import re
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DEFAULT_MODELS = ["circle", "boids2d", "game_of_life", "sugarscape", "ants", "predator_prey"]  # bench.py:26-29
DEFAULT_RANGE = (250, 16384000)   # the reference stops at 32 000 (c) ... 1 024 000 (flamegpu), bench.py:31-38
NUM_TIMESTEPS = 100               # bench.py:90


class InvocationFailed(Exception):
    pass


class OpenAbl:
    """Same contract as the reference's bench/openabl.py: run() returns the tool's output,
    get_exec_time() the seconds of its `Execution time` line."""

    def __init__(self, openabl_bin, example_dir, asset_dir):
        self.openabl_bin, self.example_dir, self.asset_dir = openabl_bin, example_dir, asset_dir

    def command(self, model, backend, params, config, out_dir=None):
        args = [self.openabl_bin, "-i", os.path.join(self.example_dir, model + ".abl"), "-b", backend,
                "-A", self.asset_dir, "-R"]
        if out_dir:
            args += ["-o", out_dir]
        for k, v in params.items():
            args += ["-P", "%s=%s" % (k, v)]
        for k, v in config.items():
            args += ["-C", "%s=%s" % (k, v)]
        return args

    def run(self, model, backend, params, config, out_dir=None):
        args = self.command(model, backend, params, config, out_dir)
        try:
            return subprocess.check_output(args, stderr=subprocess.STDOUT).decode("utf-8", "replace")
        except subprocess.CalledProcessError as err:
            raise InvocationFailed("Invocation of command\n%s\nexited with exit code %d and the following output:\n%s"
                                   % (" ".join(args), err.returncode, err.output.decode("utf-8", "replace")))

    def get_exec_time(self, model, backend, params, config, out_dir=None):
        m = re.search(r"Execution time: (.*)s", self.run(model, backend, params, config, out_dir))
        if m is None:
            raise RuntimeError("Failed to extract execution time")
        return float(m.group(1))


def agent_counts(lo, hi, factor=2):
    n = lo
    while n <= hi:
        yield n
        n *= factor


def run_bench(runner, backend, model, num_agents_range, max_time, config, out_dir=None, log=print):
    """-> the text of bench_<model>_<backend>.txt (`n,t` header + one row per population)."""
    lo, hi = num_agents_range
    result = "n,t\n"
    log("Running %s on %s backend with %d-%d agents" % (model, backend, lo, hi))
    start = time.time()
    for n in agent_counts(lo, hi):
        t0 = time.time()
        try:
            t = runner.get_exec_time(model, backend, {"num_timesteps": NUM_TIMESTEPS, "num_agents": n}, config, out_dir)
        except InvocationFailed as err:
            log(str(err))
            break
        row = "%d,%s" % (n, t)
        result += row + "\n"
        log(row)
        # the time of the current run is the estimate for the next one (bench.py:124-128)
        if max_time is not None and (time.time() - start) + (time.time() - t0) > max_time:
            break
    return result


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-b", "--backends", default="cuda", help="Backends to benchmark (comma separated)")
    ap.add_argument("-m", "--models", help="Models to benchmark (comma separated)")
    ap.add_argument("-n", "--num-agents", help="Number of agent range (min-max)")
    ap.add_argument("-r", "--result-dir", help="Directory for benchmark results")
    ap.add_argument("-M", "--max-time", metavar="SEC", type=int, help="(Approximate) maximal time per backend per model")
    ap.add_argument("-e", "--example-dir", default=os.path.join(REPO, "examples"))
    ap.add_argument("-C", "--config", action="append", default=[], help="name=value passed through to OpenABL -C")
    args = ap.parse_args(argv)
    if args.result_dir:
        os.makedirs(args.result_dir, exist_ok=True)
    else:
        print("WARNING: No result directory specified")
    openabl_bin = os.path.join(REPO, "OpenABL")
    if not os.path.isfile(openabl_bin):
        sys.exit("OpenABL binary not found. Tried: " + openabl_bin)
    runner = OpenAbl(openabl_bin, args.example_dir, os.path.join(REPO, "asset"))
    rng = DEFAULT_RANGE
    if args.num_agents:
        spec = args.num_agents.split("-")
        if len(spec) != 2:
            sys.exit("Invalid agent number specification (min-max)")
        rng = (int(spec[0]), int(spec[1]))
    config = dict(c.split("=", 1) for c in args.config)
    models = args.models.split(",") if args.models else [
        m for m in DEFAULT_MODELS if os.path.exists(os.path.join(args.example_dir, m + ".abl"))]
    for backend in args.backends.split(","):
        for model in models:
            text = run_bench(runner, backend, model, rng, args.max_time, config)
            if args.result_dir:
                with open(os.path.join(args.result_dir, "bench_%s_%s.txt" % (model, backend)), "w") as f:
                    f.write(text)


if __name__ == "__main__":
    main()
