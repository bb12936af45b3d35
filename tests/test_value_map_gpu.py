"""Parity of the CUDA value map (through the C-ABI, via the reference-shaped class)
against the oracle and the golden fixtures.  Confidence grid: bit-exact.  Value grid:
<= 1e-6 absolute (float32 storage of what the reference holds in float64; the
north-star tolerance is 1e-4)."""
import glob
import os

import numpy as np
import pytest

from test_oracle_value_map import FOV, dense, load_case
from oracle.value_map_oracle import ValueMapOracle
from vlfm_b200.utils.synthetic import trajectory

pytestmark = pytest.mark.gpu
VAL_TOL = 1e-6


def _gpu_map(*a, **k):
    from vlfm_b200.mapping.value_map import ValueMap

    return ValueMap(*a, **k)


def test_golden_fixtures(golden_dir):
    for path in sorted(glob.glob(os.path.join(golden_dir, "vm_*.npz"))):
        z, frames = load_case(path)
        size, ch = int(z["size"]), int(z["channels"])
        g = _gpu_map(ch, size=size, use_max_confidence=bool(z["use_max_confidence"]), fusion_type=str(z["fusion"]))
        for f, v in zip(frames, z["values"]):
            g.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
        conf = dense(z["conf_idx"], z["conf_val"], (size, size), np.float32)
        val = dense(z["value_idx"], z["value_val"], (size, size, ch), np.float64)
        assert np.array_equal(g._map, conf), path
        assert np.abs(g._value_map.astype(np.float64) - val).max() <= VAL_TOL, path
        red = (lambda s: [max(t) for t in s]) if ch > 1 else None
        sw, sv = g.sort_waypoints(z["waypoints"], 0.5, reduce_fn=red)
        assert np.array_equal(sw, z["sorted_wp"]), path
        assert np.abs(np.asarray(sv, float) - z["sorted_val"]).max() <= VAL_TOL, path


@pytest.mark.parametrize("cfg", [
    dict(ch=1, maxc=False, fus="default", size=1000, hw=(480, 640), seed=31, steps=12, bound=18.0),
    dict(ch=3, maxc=False, fus="default", size=640, hw=(240, 320), seed=32, steps=10, bound=9.0),
    dict(ch=1, maxc=True, fus="default", size=1000, hw=(480, 640), seed=33, steps=10, bound=18.0),
    dict(ch=1, maxc=False, fus="default", size=333, hw=(97, 131), seed=34, steps=10, bound=8.2),   # ragged: W%4!=0, G%4!=0, clipping
    dict(ch=2, maxc=True, fus="equal_weighting", size=400, hw=(120, 160), seed=35, steps=8, bound=4.0),
])
def test_trajectory_vs_oracle(cfg):
    o = ValueMapOracle(cfg["ch"], size=cfg["size"], use_max_confidence=cfg["maxc"], fusion_type=cfg["fus"], prims="cv2")
    g = _gpu_map(cfg["ch"], size=cfg["size"], use_max_confidence=cfg["maxc"], fusion_type=cfg["fus"])
    rng = np.random.default_rng(cfg["seed"])
    h, w = cfg["hw"]
    for i, f in enumerate(trajectory(cfg["seed"], cfg["steps"], h=h, w=w, bound_m=cfg["bound"])):
        v = rng.random(cfg["ch"])
        o.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
        g.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
        if i % 4 == 3 or i == cfg["steps"] - 1:
            assert np.array_equal(g._map, o._map), f"conf differs at step {i}"
            assert np.abs(g._value_map.astype(np.float64) - o._value_map.astype(np.float64)).max() <= VAL_TOL


def test_adversarial_depth_profiles():
    """noise / constant / step / all-zero / all-one depth: exercises self-intersecting
    occlusion polygons and long polygon edges."""
    rng = np.random.default_rng(7)
    h, w = 64, 640
    imgs = [rng.random((h, w), dtype=np.float32), np.zeros((h, w), np.float32), np.ones((h, w), np.float32),
            np.full((h, w), 0.5, np.float32), np.tile((np.arange(w) % 50 < 25).astype(np.float32), (h, 1)),
            np.tile(np.linspace(0, 1, w, dtype=np.float32), (h, 1))]
    from vlfm_b200.utils.synthetic import tf_from_pose

    o = ValueMapOracle(1, size=600, use_max_confidence=False, prims="cv2")
    g = _gpu_map(1, size=600, use_max_confidence=False)
    for i, d in enumerate(imgs * 2):
        tf = tf_from_pose(0.37 * i - 2, 0.21 * i, 0.88, 0.7 * i)
        o.update_map(np.array([0.3 + 0.05 * i]), d, tf, 0.5, 5.0, FOV)
        g.update_map(np.array([0.3 + 0.05 * i]), d, tf, 0.5, 5.0, FOV)
        assert np.array_equal(g._map, o._map), f"profile {i}"
    assert np.abs(g._value_map - o._value_map).max() <= VAL_TOL


def test_batched_engine_matches_per_env():
    import torch
    from vlfm_b200.mapping.value_map import ValueMapBatch

    B, size = 5, 500
    eng = ValueMapBatch(B, 1, size=size, use_max_confidence=False)
    oracles = [ValueMapOracle(1, size=size, use_max_confidence=False) for _ in range(B)]
    trajs = [trajectory(40 + b, 6, h=120, w=160, bound_m=5.0) for b in range(B)]
    for s in range(6):
        depth = torch.from_numpy(np.stack([trajs[b][s].depth for b in range(B)])).cuda()
        tf = torch.from_numpy(np.stack([trajs[b][s].tf for b in range(B)])).cuda()
        vals = torch.full((B, 1), 0.25 + 0.1 * s, dtype=torch.float64).cuda()
        eng.update(vals, depth, tf, 0.5, 5.0, FOV)
        for b in range(B):
            oracles[b].update_map(np.array([0.25 + 0.1 * s]), trajs[b][s].depth, trajs[b][s].tf, 0.5, 5.0, FOV)
    conf = eng.conf.cpu().numpy()
    for b in range(B):
        assert np.array_equal(conf[b], oracles[b]._map)


def test_error_behaviour():
    g = _gpu_map(2, size=200)
    with pytest.raises(AssertionError):
        g.update_map(np.array([0.1]), np.zeros((8, 8), np.float32), np.eye(4), 0.5, 5.0, FOV)  # wrong len(values)
    tf = np.eye(4); tf[0, 3] = 50.0
    with pytest.raises(AssertionError):
        g.update_map(np.array([0.1, 0.2]), np.zeros((8, 8), np.float32), tf, 0.5, 5.0, FOV)  # camera off grid
    g.reset()
    assert g._map.sum() == 0


def test_full_size_properties():
    """BASELINE full size (640x480, G=1000): size-independent properties instead of a
    slow oracle run -- idempotence of max-confidence fusion and conf bounds."""
    g = _gpu_map(1, size=1000, use_max_confidence=True)
    f = trajectory(50, 1)[0]
    g.update_map(np.array([0.5]), f.depth, f.tf, 0.5, 5.0, FOV)
    a = g._map.copy(); va = g._value_map.copy()
    g.update_map(np.array([0.5]), f.depth, f.tf, 0.5, 5.0, FOV)
    assert np.array_equal(a, g._map) and np.array_equal(va, g._value_map)  # same view twice changes nothing
    assert a.max() <= 1.0 and a.min() >= 0.0 and 0 < (a > 0).sum() <= 7081 + 600


def test_explored_area_masking_matches_oracle():
    """value_map.py:365-375: with an obstacle map attached, unexplored cells are zeroed in the new observation,
    the confidence grid and the value grid (whole grid, every step)."""
    import torch

    class FakeObstacleMap:
        pixels_per_meter, size = 20, 500

        def __init__(self):
            self.dev = torch.zeros((1, 500, 500), dtype=torch.uint8, device="cuda")
            self.host = np.zeros((500, 500), dtype=bool)

        def explored_device(self):
            return self.dev

        def set(self, mask):
            self.host = mask
            self.dev.copy_(torch.from_numpy(mask.astype(np.uint8))[None])

    om = FakeObstacleMap()
    o = ValueMapOracle(1, size=500, use_max_confidence=False, explored_fn=lambda: om.host)
    g = _gpu_map(1, size=500, use_max_confidence=False, obstacle_map=om)
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:500, 0:500]
    for i, f in enumerate(trajectory(61, 8, h=120, w=160, bound_m=4.0)):
        cx, cy = 250 + 60 * np.cos(i), 250 + 60 * np.sin(i)
        om.set(((yy - cy) ** 2 + (xx - cx) ** 2) < (90 + 10 * i) ** 2)     # a moving, growing explored disc
        v = rng.random(1)
        o.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
        g.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
        assert np.array_equal(g._map, o._map), f"step {i}"
        assert np.abs(g._value_map - o._value_map).max() <= VAL_TOL


def test_other_resolution_ppm40():
    """configs[4]-style geometry: 0.025 m cells (ppm 40 -> R = 401), square depth image.  The reference class
    needs the documented patch (pixels_per_meter attribute + cone-cache clear); the oracle takes ppm."""
    o = ValueMapOracle(1, size=1400, use_max_confidence=False, pixels_per_meter=40)
    g = _gpu_map(1, size=1400, use_max_confidence=False, pixels_per_meter=40)
    rng = np.random.default_rng(9)
    for i, f in enumerate(trajectory(62, 5, h=256, w=256, bound_m=10.0)):
        v = rng.random(1)
        o.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
        g.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
    assert np.array_equal(g._map, o._map)
    assert np.abs(g._value_map - o._value_map).max() <= VAL_TOL


@pytest.mark.parametrize("cfg", [(79.0, 5.0, 20), (79.0, 5.0, 40), (79.0, 5.0, 50), (90.0, 3.5, 20), (60.0, 10.0, 20), (120.0, 2.0, 20)])
def test_cone_template_built_on_device_is_bit_exact(cfg):
    """vlfm_value_cone_template (device) against the reference's construction (cv2.ellipse sector x python-float cos^2
    falloff, value_map.py:321-355) as restated in oracle.value_map_oracle.cone_template: every float32 identical."""
    import torch

    from oracle.value_map_oracle import cone_template
    from vlfm_b200.mapping.value_map import build_cone_template

    fov_deg, max_depth, ppm = cfg
    fov = float(np.deg2rad(fov_deg))
    got = build_cone_template(fov, max_depth, ppm, torch.device("cuda")).cpu().numpy()
    want = cone_template(fov, max_depth, ppm)
    assert got.shape == want.shape and got.dtype == np.float32
    bad = np.flatnonzero(got.view(np.uint32) != np.asarray(want, dtype=np.float32).view(np.uint32))
    assert bad.size == 0, f"{bad.size} of {got.size} template cells differ (first at {np.unravel_index(bad[0], got.shape)})"
