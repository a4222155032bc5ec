"""Generates the golden files of tests/test_save_formats.py with the REAL reference runtime:
the reference `c` backend's program (oracle/_ref/OpenABL_ref -b c + the reference's own
libabl.c, see refgen.py) is run with ABL_REF_XML=1, which makes the link-time save() wrapper
(oracle/shim/save_wrap.c) call the reference's save() twice more with SAVE_FLAME_XML and
SAVE_FLAMEGPU_XML (reference asset/c/libabl.c:126-213).  The raw records the files were written
from are kept next to them.

TEST INFRASTRUCTURE ONLY; works only where /root/reference exists.

    python oracle/xmlgen.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
import refgen  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden", "xml")
CASES = {
    "circle_n40": ("circle.abl", {"num_agents": 40, "num_timesteps": 2}),
    "circle3d_n30": ("circle3d.abl", {"num_agents": 30, "num_timesteps": 1}),
    "game_of_life_n36": ("game_of_life.abl", {"num_agents": 36, "num_timesteps": 1}),
}


def main():
    if not refgen.reference_available():
        sys.exit("reference not available: build oracle/_ref first (make -C oracle ref)")
    os.environ["ABL_REF_XML"] = "1"
    os.makedirs(OUT, exist_ok=True)
    for name, (model, params) in CASES.items():
        tmp = tempfile.mkdtemp(prefix="ablxml_")
        try:
            exe = refgen.build_reference_program(os.path.join(REPO, "examples", model), params, False, tmp)
            subprocess.run([exe], cwd=tmp, check=True)
            for f in os.listdir(tmp):
                if f.endswith(".xml"):
                    shutil.copy(os.path.join(tmp, f), os.path.join(OUT, "%s.%s.xml" % (name, f.split(".")[-2])))
                elif f.endswith(".bin"):
                    shutil.copy(os.path.join(tmp, f), os.path.join(OUT, name + ".bin"))
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        print(name)


if __name__ == "__main__":
    main()
