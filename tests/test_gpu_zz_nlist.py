"""Cached neighbour lists (-C cuda.nlist=true) on the GPU.

The list kernels are verified bit for bit under the CPU emulator (tests/test_emu_kernels.py);
the runtime side (count pass -> sizing -> fill pass, validity tracking) passed on a B200 at the
end of round 1 (GPUTEST_r01.json: three XPASS), so the xfail guard of that round is gone and a
regression fails.  The checks still run in a child process (a device fault cannot poison the CUDA
context of the session)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def check(case):
    proc = subprocess.run([sys.executable, os.path.join(HERE, "nlist_check.py"), case], stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0 and proc.stdout.strip().endswith("ok"), proc.stdout[-3000:]


@pytest.mark.gpu
def test_game_of_life_with_neighbour_lists_equals_grid_oracle():
    check("game_of_life")


@pytest.mark.gpu
def test_lists_of_static_sites_survive_moving_walkers():
    """Site-Site lists are built once while the Walker pool is re-binned every timestep; the state
    equals the run without lists bit for bit, and re-uploading the population rebuilds the lists."""
    check("static_sites")


@pytest.mark.gpu
def test_lists_match_the_real_reference():
    """sites_walkers with lists for the static Sites against the reference `c` backend's golden vector."""
    check("sites_walkers")
