"""`-C visualize=true`: the model's getColor / getSize hooks — which the reference analyses for every
back end and emits only for mason / dmason (AnalysisVisitor.cpp:215-222) — paint one picture per frame the way
the reference's display does (MasonPrinter.cpp:718-780: 500 x 500, white backdrop, an oval of 4 * getSize pixels
in Color(getColor(agent)); getColor = 0 and getSize = 1 when the model has no hooks, :608-613).
CPU part: the generated host functions against the hooks restated in Python, on the initial population."""
import os
import subprocess

import numpy as np
import pytest

from openabl_b200 import build
from openabl_b200.model import Model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_ppm(path):
    with open(path, "rb") as f:
        data = f.read()
    assert data[:3] == b"P6\n"
    header, rest = data.split(b"\n255\n", 1)
    w, h = (int(v) for v in header.split(b"\n")[1].split())
    return np.frombuffer(rest, dtype=np.uint8).reshape(h, w, 3)


def pixel_of(pos, scale):
    return np.floor(pos * scale).astype(int)


def rgb(c):
    return np.array([(c >> 16) & 255, (c >> 8) & 255, c & 255], dtype=np.uint8)


def test_game_of_life_frame_shows_getColor_at_every_cell(tmp_path):
    m = Model(os.path.join(REPO, "examples", "game_of_life.abl"), {"num_agents": 4096}, config={"visualize": True})
    m.populate()
    path = str(tmp_path / "f.ppm")
    m.write_frame(path)
    img = read_ppm(path)
    assert img.shape == (500, 500, 3)
    cells = m.host_agents(0)
    scale = 500.0 / 64.0
    px = pixel_of(cells["pos"], scale)
    want = np.where(cells["alive"][:, None], rgb(0)[None, :], rgb(0xbbbbbb)[None, :])    # getColor of game_of_life.abl
    assert np.array_equal(img[px[:, 1], px[:, 0]], want)
    assert 0 < cells["alive"].sum() < len(cells)
    # between four cells (more than the 2-pixel radius away from every centre) the white backdrop shows
    assert np.array_equal(img[0, 0], rgb(0xffffff))
    corner = pixel_of(np.array([[8.0, 8.0]]), scale)[0]
    assert np.array_equal(img[corner[1], corner[0]], rgb(0xffffff))


def test_sugarscape_frame_follows_its_three_way_getColor(tmp_path):
    m = Model(os.path.join(REPO, "examples", "sugarscape.abl"), {"num_agents": 4096}, config={"visualize": True})
    m.populate()
    path = str(tmp_path / "f.ppm")
    m.write_frame(path)
    img = read_ppm(path)
    a = m.host_agents(0)
    # getColor of sugarscape.abl, restated
    frac = np.minimum(1.0, a["env_sugar_level"].astype(np.float64) / 50.0)
    c = (255 * (1 - frac)).astype(np.int64)
    colour = np.where(a["state"] == 0, 0xff00ff | (c << 8), np.where(a["state"] == 1, 0x000000, 0xff0000))
    want = np.stack([(colour >> 16) & 255, (colour >> 8) & 255, colour & 255], axis=1).astype(np.uint8)
    px = pixel_of(a["pos"], 500.0 / 64.0)
    assert np.array_equal(img[px[:, 1], px[:, 0]], want)
    assert len(np.unique(colour)) > 10


def test_a_model_without_hooks_is_drawn_in_black_dots_of_four_pixels(tmp_path):
    m = Model(os.path.join(REPO, "examples", "boids2d.abl"), {"num_agents": 200}, config={"visualize": True})
    m.populate()
    path = str(tmp_path / "f.ppm")
    m.write_frame(path)
    img = read_ppm(path)
    b = m.host_agents(0)
    side = float(np.sqrt(200 / 500.0))
    px = pixel_of(b["pos"], 500.0 / side)
    inside = (px >= 0).all(axis=1) & (px < 500).all(axis=1)
    assert np.array_equal(img[px[inside, 1], px[inside, 0]], np.zeros((inside.sum(), 3), np.uint8))
    black = (img == 0).all(axis=2).sum()
    assert 200 * 6 <= black <= 200 * 16      # discs of diameter 4: about 12 pixels each, overlaps and borders aside


def test_visualize_is_refused_on_several_gpus():
    proc = subprocess.run([build.build_compiler(), "-i", os.path.join(REPO, "examples", "circle.abl"), "-b", "cuda",
                           "-A", os.path.join(REPO, "asset"), "-C", "visualize=true", "-C", "cuda.gpus=2", "-o", "/tmp/abl_vis_refused"],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 2 and "visualize" in proc.stdout      # BackendError, the reference's exit code


@pytest.mark.gpu
def test_generated_program_writes_one_frame_per_interval(tmp_path):
    out = build.build_model(os.path.join(REPO, "examples", "game_of_life.abl"), {"num_agents": 4096, "num_timesteps": 6},
                            {"visualize": True, "cuda.frame_interval": 2})
    proc = subprocess.run([os.path.join(out, "main")], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout
    frames = sorted(os.listdir(tmp_path / "frames"))
    assert frames == ["frame_00000.ppm", "frame_00002.ppm", "frame_00004.ppm", "frame_00006.ppm"]
    imgs = [read_ppm(str(tmp_path / "frames" / f)) for f in frames]
    assert all(i.shape == (500, 500, 3) for i in imgs)
    assert not np.array_equal(imgs[0], imgs[1]) and not np.array_equal(imgs[1], imgs[3])     # the cells live
    # the last frame is the state save() writes: dark discs exactly where the file says alive
    import json
    with open(tmp_path / "result.json") as f:
        cells = json.load(f)["Cell"]
    pos = np.array([c["pos"] for c in cells])
    alive = np.array([c["alive"] for c in cells])
    px = pixel_of(pos, 500.0 / 64.0)
    assert np.array_equal((imgs[3][px[:, 1], px[:, 0]] == 0).all(axis=1), alive)
