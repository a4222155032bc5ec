"""Runs the REAL reference (`oracle/_ref/OpenABL_ref -b c` + the reference's own libabl,
compiled by gcc with the reference's build line) and collects raw agent state.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container);
the GPU box uses the fixtures this script commits under tests/golden/ and the prebuilt
binaries under oracle/_ref/.

    python oracle/refgen.py            # (re)generate every fixture listed in FIXTURES
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("OPENABL_REFERENCE", "/root/reference")
REF_BIN = os.path.join(HERE, "_ref", "OpenABL_ref")
GOLDEN = os.path.join(REPO, "tests", "golden")

sys.path.insert(0, REPO)
from openabl_b200.state import agent_dtype, parse_agents, read_raw  # noqa: E402


def reference_available():
    return os.path.isdir(REF) and os.path.exists(REF_BIN)


def build_reference_program(model, params, use_float, out_dir, threads_flag=True):
    """Generates C with the reference compiler and builds it with the reference's build line
    plus the save() wrapper that dumps raw state.  Returns the path of the executable."""
    os.makedirs(out_dir, exist_ok=True)
    cmd = [REF_BIN, "-A", os.path.join(REF, "asset"), "-i", model, "-b", "c", "-o", out_dir]
    for k, v in params.items():
        cmd += ["-P", "%s=%s" % (k, v)]
    if use_float:
        cmd += ["-C", "use_float=true"]
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    # reference build line (src/backend/CBackend.cpp:21-27) + wrapper object
    gcc = ["gcc", "-O2", "-std=c99"] + (["-DLIBABL_USE_FLOAT=1"] if use_float else []) + [
        "main.c", "libabl.c", os.path.join(HERE, "shim", "save_wrap.c"), "-I.", "-lm", "-fopenmp",
        "-Wl,--wrap=save", "-o", "main"]
    subprocess.run(gcc, cwd=out_dir, check=True)
    return os.path.join(out_dir, "main")


def strip_lifecycle(source):
    """predator_prey.abl without what the reference's `c` back end cannot print (CBackend.cpp:30-32 rejects run-time
    add / remove; CPrinter has no count / sum / log_csv): removeCurrent() and the add() calls of the step functions
    become empty blocks, the sequential step goes.  main() — the population set-up — is untouched, which is all the
    `_t0` fixture of this model pins."""
    import re
    cut = source.index("void main()")
    head, tail = source[:cut], source[cut:]
    head = head.replace("removeCurrent();", "{ }")
    head = re.sub(r"add\((Prey|Predator) \{.*?\}\s*\);", "{ }", head, flags=re.S)
    head = re.sub(r"sequential step gather_stats\(\) \{.*?\n\}\n", "", head, flags=re.S)
    tail = tail.replace("grass_growth,\n    gather_stats", "grass_growth")
    return head + tail


def run_reference(model, params, use_float=False, keep=None, transform=None):
    """-> (list of structured arrays per agent type, json text)"""
    tmp = keep or tempfile.mkdtemp(prefix="ablref_")
    try:
        with open(model) as f:
            source = f.read()
        if transform is not None:
            scratch = os.path.join(tmp, "scratch_" + os.path.basename(model))
            with open(scratch, "w") as f:
                f.write(transform(source))
            exe = build_reference_program(scratch, params, use_float, tmp)
        elif "save(" not in source:
            # a model that never writes its state (sugarscape.abl): the reference runs a scratch copy whose main()
            # ends with save(); the device side is compared through abl_cuda_download, the model itself is untouched
            cut = source.rstrip().rfind("}")
            scratch = os.path.join(tmp, "with_save_" + os.path.basename(model))
            with open(scratch, "w") as f:
                f.write(source[:cut] + '  save("state.json");\n' + source[cut:])
            exe = build_reference_program(scratch, params, use_float, tmp)
        else:
            exe = build_reference_program(model, params, use_float, tmp)
        subprocess.run([exe], cwd=tmp, check=True)
        agents = parse_agents(source)
        dtypes = [agent_dtype(m, use_float) for _, m in agents]
        raw = [p for p in os.listdir(tmp) if p.endswith(".bin")]
        assert len(raw) == 1, raw
        state = read_raw(os.path.join(tmp, raw[0]), dtypes)
        with open(os.path.join(tmp, raw[0][:-4])) as f:
            text = f.read()
        return state, text
    finally:
        if not keep:
            shutil.rmtree(tmp, ignore_errors=True)


# name -> (model file under examples/, params, use_float)
FIXTURES = {
    "circle_n1000_t100": ("circle.abl", {"num_agents": 1000, "num_timesteps": 100}, False),
    "circle_n1000_t10": ("circle.abl", {"num_agents": 1000, "num_timesteps": 10}, False),
    "circle_n1000_t0": ("circle.abl", {"num_agents": 1000, "num_timesteps": 0}, False),
    "circle_n2000_t10_f32": ("circle.abl", {"num_agents": 2000, "num_timesteps": 10}, True),
    "circle3d_n2000_t10": ("circle3d.abl", {"num_agents": 2000, "num_timesteps": 10}, False),
    "circle3d_n2000_t0": ("circle3d.abl", {"num_agents": 2000, "num_timesteps": 0}, False),
    "boids2d_n4000_t10": ("boids2d.abl", {"num_agents": 4000, "num_timesteps": 10}, False),
    "boids2d_n4000_t0": ("boids2d.abl", {"num_agents": 4000, "num_timesteps": 0}, False),
    "boids2d_n4000_t10_f32": ("boids2d.abl", {"num_agents": 4000, "num_timesteps": 10}, True),
    "game_of_life_n4096_t10": ("game_of_life.abl", {"num_agents": 4096, "num_timesteps": 10}, False),
    "game_of_life_n4096_t0": ("game_of_life.abl", {"num_agents": 4096, "num_timesteps": 0}, False),
}


# Feature-test models (tests/models/): constructs of the reference's other examples — 3-D
# flocking, two agent types reading each other over several step functions, constant tables /
# while / float modulo.  Golden vectors come from the real reference as above; there is no
# plain-C restatement of these models (tests/test_oracle.py covers FIXTURES only), the GPU is
# compared with the reference's output directly (tests/test_gpu_parity_reference.py).
EXTRA_FIXTURES = {
    "flock3d_n2000_t10": ("tests/models/flock3d.abl", {"num_agents": 2000, "num_timesteps": 10}, False),
    "flock3d_n2000_t10_f32": ("tests/models/flock3d.abl", {"num_agents": 2000, "num_timesteps": 10}, True),
    "two_species_n3000_t10": ("tests/models/two_species.abl", {"num_agents": 3000, "num_timesteps": 10}, False),
    "two_species_n3000_t0": ("tests/models/two_species.abl", {"num_agents": 3000, "num_timesteps": 0}, False),
    "table_cells_n2500_t10": ("tests/models/table_cells.abl", {"num_agents": 2500, "num_timesteps": 10}, False),
    "table_cells_n2500_t10_f32": ("tests/models/table_cells.abl", {"num_agents": 2500, "num_timesteps": 10}, True),
    # radii below the cell size (cell-range culling), a static type next to a moving one (neighbour lists),
    # a distance compared with a constant
    "sites_walkers_n3000_t10": ("tests/models/sites_walkers.abl", {"num_agents": 3000, "num_timesteps": 10}, False),
    "sites_walkers_n3000_t10_f32": ("tests/models/sites_walkers.abl", {"num_agents": 3000, "num_timesteps": 10}, True),
}


# The reference examples the `c` backend can run besides the four above (keratinocyte, ants, sugarscape,
# boids2d_flockers and predator_prey use run-time add/remove or in-step random(), which it rejects or
# leaves racy): boids.abl, 3-D flocking, pinned on the real model instead of a look-alike.
EXTRA_FIXTURES.update({
    "boids_n2000_t10": ("boids.abl", {"num_agents": 2000, "num_timesteps": 10}, False),
    "boids_n2000_t10_f32": ("boids.abl", {"num_agents": 2000, "num_timesteps": 10}, True),
})

# The three remaining examples of the distribution compile under the reference's `c` backend but cannot be pinned
# beyond their INITIAL populations: ants.abl and boids2d_flockers.abl draw random numbers inside their step
# functions (one global xorshift stream advanced from an OpenMP loop: the reference's own output is not
# reproducible), and sugarscape.abl (i) assigns `out` members only conditionally — the `c` backend's out buffer then
# keeps the value of two step functions ago (CPrinter.cpp:210-228) where Mason semantics copy `in` — and (ii) lets
# the FIRST matching neighbour win, i.e. depends on the visiting order.  Their host initialisation (randomInt,
# exp, int conversions, nested loops) is pinned bit for bit.
EXTRA_FIXTURES.update({
    "ants_n500_t0": ("ants.abl", {"num_agents": 500, "num_timesteps": 0}, False),
    "boids2d_flockers_n2000_t0": ("boids2d_flockers.abl", {"num_agents": 2000, "num_timesteps": 0}, False),
    "sugarscape_n4096_t0": ("sugarscape.abl", {"num_agents": 4096, "num_timesteps": 0}, False),
})

# predator_prey.abl: run-time add / remove keep the reference's `c` back end from generating anything; its INITIAL
# population (three agent types, randomDirection(), randomInt) is pinned through a scratch copy without them.
TRANSFORMS = {"predator_prey_n32000_t0": strip_lifecycle}
EXTRA_FIXTURES.update({
    "predator_prey_n32000_t0": ("predator_prey.abl", {"num_agents": 32000, "num_timesteps": 0}, False),
})

# save() text of the reference (libabl.c:46-124), byte for byte: name -> (model, params, use_float, output file).
# Stored under tests/golden/text/<name>.txt; tests/test_gpu_save_text.py compares the file the generated
# ./main writes with it (integer / bool models byte-identical, floating point within the last printed digit).
TEXT_FIXTURES = {
    "game_of_life_n1024_t10": ("game_of_life.abl", {"num_agents": 1024, "num_timesteps": 10}, False),
    "circle_n500_t10": ("circle.abl", {"num_agents": 500, "num_timesteps": 10}, False),
    "boids2d_n1000_t10": ("boids2d.abl", {"num_agents": 1000, "num_timesteps": 10}, False),
    # (single precision on the 2-D model: circle3d with use_float is order-sensitive beyond any tolerance after 10 steps —
    # 200 neighbours per agent and a force that jumps at distance r; tests/test_oracle.py::test_circle3d_float_is_order_sensitive)
    "circle_n500_t10_f32": ("circle.abl", {"num_agents": 500, "num_timesteps": 10}, True),
}


def text_fixture_path(name):
    return os.path.join(GOLDEN, "text", name + ".txt")


def model_path(model):
    """Fixture model names are relative to examples/ unless they carry a directory."""
    return os.path.join(REPO, model) if os.sep in model or "/" in model else os.path.join(REPO, "examples", model)


def fixture_paths(name):
    return os.path.join(GOLDEN, name + ".npz"), os.path.join(GOLDEN, name + ".json")


def load_fixture(name):
    npz, meta = fixture_paths(name)
    with open(meta) as f:
        info = json.load(f)
    data = np.load(npz)
    return info, [data["type%d" % i] for i in range(info["n_types"])]


def main():
    if not reference_available():
        sys.exit("reference not available: build oracle/_ref first (make -C oracle ref)")
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])
    for name, (model, params, use_float) in list(FIXTURES.items()) + list(EXTRA_FIXTURES.items()):
        if only and name not in only:
            continue
        path = model_path(model)
        state, text = run_reference(path, params, use_float, transform=TRANSFORMS.get(name))
        npz, meta = fixture_paths(name)
        np.savez_compressed(npz, **{"type%d" % i: s for i, s in enumerate(state)})
        with open(meta, "w") as f:
            json.dump({"model": model, "params": params, "use_float": use_float,
                       "n_types": len(state), "counts": [len(s) for s in state],
                       "generator": "oracle/refgen.py (reference c backend, gcc -O2 -std=c99 -fopenmp)"},
                      f, indent=1)
        print(name, [len(s) for s in state])
    os.makedirs(os.path.join(GOLDEN, "text"), exist_ok=True)
    for name, (model, params, use_float) in TEXT_FIXTURES.items():
        if only and name not in only:
            continue
        _, text = run_reference(model_path(model), params, use_float)
        with open(text_fixture_path(name), "w") as f:
            f.write(text)
        print(name, len(text), "bytes of save() text")


if __name__ == "__main__":
    main()
