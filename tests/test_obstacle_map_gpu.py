"""Obstacle half of ObstacleMap on the GPU: bit-exact against the oracle."""
import numpy as np
import pytest

from oracle.obstacle_map_oracle import ObstacleMapOracle
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [
    dict(seed=0, hw=(480, 640), size=1000, ppm=20, steps=8, bound=15.0),
    dict(seed=1, hw=(240, 320), size=600, ppm=20, steps=8, bound=5.0),
    dict(seed=2, hw=(480, 640), size=2500, ppm=50, steps=5, bound=10.0),   # ActionReplayPolicy resolution
    dict(seed=3, hw=(97, 131), size=300, ppm=20, steps=10, bound=1.5),     # ragged image, window hits the border
])
def test_obstacle_half_vs_oracle(cfg):
    from vlfm_b200.mapping.obstacle_map import ObstacleMap

    h, w = cfg["hw"]
    fx = focal_from_hfov(w)
    o = ObstacleMapOracle(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=-1, size=cfg["size"], pixels_per_meter=cfg["ppm"])
    g = ObstacleMap(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=-1, size=cfg["size"], pixels_per_meter=cfg["ppm"])
    for i, f in enumerate(trajectory(cfg["seed"], cfg["steps"], h=h, w=w, bound_m=cfg["bound"])):
        o.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79), explore=False)
        g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79), explore=False)
        assert np.array_equal(g._map, o._map), f"obstacles differ at step {i}"
        assert np.array_equal(g._navigable_map, np.asarray(o._navigable_map)), f"navigable differs at step {i}"
    assert o._map.sum() > 0
    g.reset()
    assert g._map.sum() == 0 and g._navigable_map.sum() == 0


def test_out_of_range_scatter_raises_index_error():
    from vlfm_b200.mapping.obstacle_map import ObstacleMap
    from vlfm_b200.utils.synthetic import tf_from_pose

    g = ObstacleMap(0.0, 2.0, 0.18, hole_area_thresh=-1, size=200)
    depth = np.full((60, 80), 0.5, np.float32)
    # camera at x=+4.9 m looking along +x: points beyond x=5 m give row >= G -> IndexError in numpy
    with pytest.raises(IndexError):
        g.update_map(depth, tf_from_pose(4.9, 0.0, 0.88, 0.0), 0.5, 5.0, 80.0, 80.0, np.deg2rad(79), explore=False)


def _speckled_depth(seed, h, w):
    """Depth with zero regions of every kind fill_small_holes distinguishes: speckle, blobs, blobs enclosing valid
    islands (hole borders), islands holding zeros again, and regions touching the frame."""
    rng = np.random.default_rng(seed)
    d = rng.uniform(0.05, 1.0, (h, w)).astype(np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(int(rng.integers(3, 9))):
        cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(3, max(4, min(h, w) // 4))
        blob = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        d[blob] = 0
        if r > 6:                                                # valid island inside, with a zero core
            d[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 2) ** 2] = 0.5
            d[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 4) ** 2] = 0
    d[rng.random((h, w)) < 0.01] = 0
    d[h // 3, : w // 2] = 0                                       # a one-pixel-wide line
    return d


@pytest.mark.parametrize("hw", [(480, 640), (97, 131), (64, 64)])
@pytest.mark.parametrize("thresh", [5, 40, 200, 100000])
def test_fill_small_holes_vs_oracle(hw, thresh):
    import ctypes
    import torch

    from oracle.obstacle_map_oracle import fill_holes, fill_holes_numpy
    from vlfm_b200 import _lib

    lib = _lib.load()
    h, w = hw
    nb = ctypes.c_size_t(0)
    _lib.check(lib.vlfm_holes_workspace_bytes(h, w, ctypes.byref(nb)), "ws")
    ws = torch.zeros((nb.value + 3) // 4, dtype=torch.int32, device="cuda")
    fill = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    for seed in range(4):
        d = _speckled_depth(seed * 7 + thresh, h, w)
        dev = torch.from_numpy(d).cuda()
        _lib.check(lib.vlfm_fill_small_holes(_lib.ptr(dev), h, w, float(thresh), _lib.ptr(fill), _lib.ptr(ws), _lib.ptr(status),
                                             _lib.stream_ptr()), "fill")
        assert int(status.item()) == 0
        got = np.where(fill.cpu().numpy() == 1, np.float32(1), d)
        want = fill_holes(d, thresh)
        assert np.array_equal(got, want), f"seed {seed}: {np.count_nonzero(got != want)} pixels differ from cv2"
        if h * w <= 97 * 131:
            assert np.array_equal(got, fill_holes_numpy(d, thresh))


def test_obstacle_half_with_hole_filling_vs_oracle():
    """ObstacleMap's default hole_area_thresh (vlfm/policy/base_objectnav_policy.py:70 passes 100000)."""
    from vlfm_b200.mapping.obstacle_map import ObstacleMap

    h, w = 240, 320
    fx = focal_from_hfov(w)
    for thresh in (100000, 60):
        o = ObstacleMapOracle(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=thresh, size=600)
        g = ObstacleMap(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=thresh, size=600)
        for i, f in enumerate(trajectory(11, 6, h=h, w=w, bound_m=5.0)):
            depth = f.depth.copy()
            depth[_speckled_depth(i, h, w) == 0] = 0
            o.update_map(depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79), explore=False)
            g.update_map(depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79), explore=False)
            assert np.array_equal(g._map, o._map), f"obstacles differ at step {i} (thresh {thresh})"
            assert np.array_equal(g._navigable_map, np.asarray(o._navigable_map))
        assert o._map.sum() > 0


def test_fill_small_holes_overflow_is_loud():
    from vlfm_b200 import _lib
    from vlfm_b200.mapping.obstacle_map import ObstacleMap

    g = ObstacleMap(0.61, 0.88, 0.18, hole_area_thresh=100000, size=600)
    rng = np.random.default_rng(0)
    depth = rng.uniform(0.1, 1.0, (480, 640)).astype(np.float32)
    depth[::2, ::2] = 0                                           # 76800 isolated zero pixels > 65536 contours
    from vlfm_b200.utils.synthetic import tf_from_pose
    g.update_map(depth, tf_from_pose(0.0, 0.0, 0.88, 0.0), 0.5, 5.0, 300.0, 300.0, np.deg2rad(79), explore=False)
    with pytest.raises(_lib.VlfmError):
        _ = g._map
