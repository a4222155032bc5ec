#!/usr/bin/env python
"""Agent-count sweep in the format of the reference harness.

The reference's bench/bench.py doubles `num_agents` per run at 100 timesteps, takes the
wall-clock `Execution time: <s>s` the driver prints around ./run.sh (reference
src/main.cpp:300-309, bench/openabl.py:57-62) and writes `bench_<model>_<backend>.txt` as CSV
`n,t` (bench/bench.py:93-136), which its plot.py consumes.  This script produces the same
file for `-b cuda`, so results drop into the reference's plots.

    python tools/sweep.py examples/boids2d.abl --min 250 --max 1024000 [--float]
"""
import argparse
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--min", type=int, default=250)
    ap.add_argument("--max", type=int, default=1024000)
    ap.add_argument("--timesteps", type=int, default=100)
    ap.add_argument("--float", action="store_true")
    ap.add_argument("--out-dir", default=".")
    args = ap.parse_args()
    name = os.path.splitext(os.path.basename(args.model))[0]
    out_path = os.path.join(args.out_dir, "bench_%s_cuda.txt" % name)
    with open(out_path, "w") as out:
        out.write("n,t\n")
        n = args.min
        while n <= args.max:
            cmd = [os.path.join(REPO, "OpenABL"), "-A", os.path.join(REPO, "asset"), "-i", args.model, "-b", "cuda",
                   "-P", "num_agents=%d" % n, "-P", "num_timesteps=%d" % args.timesteps, "-R"]
            if args.float:
                cmd += ["-C", "use_float=true"]
            proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            m = re.search(r"Execution time: ([0-9.]+)s", proc.stdout)
            if proc.returncode != 0 or not m:
                sys.stderr.write(proc.stdout)
                sys.exit("run failed for n=%d" % n)
            out.write("%d,%s\n" % (n, m.group(1)))
            out.flush()
            print(n, m.group(1))
            n *= 2
    print("wrote", out_path)


if __name__ == "__main__":
    main()
