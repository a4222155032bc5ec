"""`-C cuda.gpus=N` / ABL_CUDA_GPUS=N: the generated ./main runs its simulate statement on N devices (one
runtime and one host thread per device, abl_cuda_group_simulate) and must write the same points.json as a
single-device run — the reference's output contract (main.cpp:281-303, libabl.c:46-124) does not know
about the decomposition.  On a single-GPU box the slabs share the device (ABL_CUDA_OVERSUBSCRIBE=1)."""
import os
import subprocess

import pytest

from openabl_b200 import build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ("circle3d.abl", {"num_agents": 20000, "num_timesteps": 5}, "points.json", 3),
    ("boids2d.abl", {"num_agents": 100000, "num_timesteps": 10}, "boids.out", 4),
    ("game_of_life.abl", {"num_agents": 65536, "num_timesteps": 10}, None, 2),
]


def run_main(out_dir, work, env_extra):
    work.mkdir()
    env = dict(os.environ)
    env.update(env_extra)
    proc = subprocess.run([os.path.join(out_dir, "main")], cwd=str(work), env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:]
    files = sorted(p for p in os.listdir(str(work)))
    assert files, "the program wrote no output file"
    return {f: (work / f).read_bytes() for f in files}


@pytest.mark.gpu
@pytest.mark.parametrize("model,params,_out,gpus", CASES, ids=[c[0][:-4] for c in CASES])
def test_generated_program_on_several_devices_writes_the_same_file(model, params, _out, gpus, tmp_path):
    import torch
    out = build.build_model(os.path.join(REPO, "examples", model), params)
    single = run_main(out, tmp_path / "one", {"ABL_CUDA_GPUS": "1"})
    env = {"ABL_CUDA_GPUS": str(gpus)}
    if torch.cuda.device_count() < gpus:
        env["ABL_CUDA_OVERSUBSCRIBE"] = "1"
    multi = run_main(out, tmp_path / "many", env)
    assert sorted(single) == sorted(multi)
    for f in single:
        assert single[f] == multi[f], "%s differs between 1 and %d devices" % (f, gpus)


@pytest.mark.gpu
def test_compile_time_option_selects_the_device_count(tmp_path):
    import torch
    out = build.build_model(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": 5000, "num_timesteps": 5},
                            {"cuda.gpus": 2})
    one = build.build_model(os.path.join(REPO, "examples", "circle.abl"), {"num_agents": 5000, "num_timesteps": 5})
    env = {"ABL_CUDA_OVERSUBSCRIBE": "1"} if torch.cuda.device_count() < 2 else {}
    a = run_main(out, tmp_path / "two", env)
    b = run_main(one, tmp_path / "one", {})
    assert a == b


def test_models_that_need_the_host_between_steps_are_refused(tmp_path):
    """predator_prey has a sequential step and run-time add(): exit code 2 (BackendError, the reference's
    convention for "this backend cannot do that", main.cpp:270-274)."""
    proc = subprocess.run([build.build_compiler(), "-i", os.path.join(REPO, "examples", "predator_prey.abl"), "-b", "cuda",
                           "-A", os.path.join(REPO, "asset"), "-o", str(tmp_path / "o"), "-C", "cuda.gpus=2"],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert proc.returncode == 2 and "cuda.gpus" in proc.stderr
