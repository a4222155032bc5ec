#!/bin/bash
# Last check of the round: GPU test tier and smoke() with the code as committed.
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=4 > $out/r2n_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/r2n_pytest_gpu.log
tail -n 9 $out/r2n_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
