"""Explore half of ObstacleMap on the GPU vs the restated oracle (bit-exact explored area, identical ORDERED frontier
lists).  Parity is unpinned w.r.t. the absent frontier_exploration package (see oracle/explore_oracle.py); every OpenCV
primitive under it is pinned against cv2 (tests/test_oracle_*.py)."""
import numpy as np
import pytest

from oracle.obstacle_map_oracle import ObstacleMapOracle
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory

pytestmark = pytest.mark.gpu
FOV = np.deg2rad(79)


def _pair(size, ppm=20, hole=-1, area=1.5):
    from vlfm_b200.mapping.obstacle_map import ObstacleMap

    o = ObstacleMapOracle(0.61, 0.88, 0.18, area_thresh=area, hole_area_thresh=hole, size=size, pixels_per_meter=ppm)
    g = ObstacleMap(0.61, 0.88, 0.18, area_thresh=area, hole_area_thresh=hole, size=size, pixels_per_meter=ppm)
    return o, g


def _same(o, g, i):
    assert np.array_equal(g.explored_area, o.explored_area), f"explored area differs at step {i}: {g.explored_area.sum()} vs {o.explored_area.sum()}"
    fo, fg = np.asarray(o._frontiers_px), np.asarray(g._frontiers_px)
    assert fo.shape == fg.shape, f"frontier count differs at step {i}: {fg.shape} vs {fo.shape}"
    if fo.size:
        assert np.array_equal(fo, fg), f"frontier px differ at step {i}"
        assert np.array_equal(np.asarray(o.frontiers), np.asarray(g.frontiers))


@pytest.mark.parametrize("cfg", [
    dict(seed=0, hw=(120, 160), size=400, steps=10, bound=4.0),
    dict(seed=3, hw=(240, 320), size=600, steps=10, bound=5.0),
    dict(seed=5, hw=(480, 640), size=1000, steps=8, bound=12.0),
    dict(seed=6, hw=(480, 640), size=1000, steps=6, bound=12.0, hole=100000),          # the policies' hole_area_thresh
    dict(seed=7, hw=(480, 640), size=2500, ppm=50, steps=5, bound=10.0, hole=100000),  # action_replay_policy.py:53-60
    dict(seed=8, hw=(480, 640), size=2000, steps=5, bound=20.0),                        # BASELINE configs[3] grid
    dict(seed=9, hw=(256, 256), size=4000, ppm=40, steps=4, bound=20.0),                # BASELINE configs[4] grid (R = 401)
])
def test_explore_vs_oracle(cfg):
    h, w = cfg["hw"]
    fx = focal_from_hfov(w)
    o, g = _pair(cfg["size"], cfg.get("ppm", 20), cfg.get("hole", -1))
    for i, f in enumerate(trajectory(cfg["seed"], cfg["steps"], h=h, w=w, bound_m=cfg["bound"])):
        o.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, FOV)
        g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, FOV)
        _same(o, g, i)
    assert o.explored_area.sum() > 100


@pytest.mark.parametrize("seed,start", [(0, (8.2, 8.2)), (1, (-8.2, 8.2)), (2, (8.2, -8.2)), (3, (-8.2, -8.2)), (4, (0.0, 8.3)), (5, (-8.3, 0.5))])
def test_explore_at_the_map_border(seed, start):
    """The agent is within max_depth of the grid edge: cv2 clips the cone and the occlusion rays (the reference accepts any
    agent cell, obstacle_map.py:114-127).  Walls are close enough that no obstacle cell leaves the grid (that would be the
    reference's IndexError)."""
    fx = focal_from_hfov(160)
    o, g = _pair(400)
    for i, f in enumerate(trajectory(seed, 8, h=120, w=160, bound_m=0.4, start_xy=start)):
        d = f.depth * np.float32(0.15)
        o.update_map(d, f.tf, 0.5, 5.0, fx, fx, FOV)
        g.update_map(d, f.tf, 0.5, 5.0, fx, fx, FOV)
        _same(o, g, i)
    assert o.explored_area.sum() > 50


def test_walk_into_the_border_raises_index_error_like_the_reference():
    """Walking outwards until an obstacle cell leaves the grid: both raise IndexError at the same step (the caller turns it
    into STOP, base_objectnav_policy.py:157-162); every step before it is identical."""
    fx = focal_from_hfov(160)
    o, g = _pair(400)
    frames = trajectory(11, 60, h=120, w=160, bound_m=30.0, start_xy=(6.0, 0.0))
    raised = False
    for i, f in enumerate(frames):
        eo = eg = None
        try:
            o.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, FOV)
        except IndexError as e:
            eo = e
        try:
            g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, FOV)
        except IndexError as e:
            eg = e
        assert (eo is None) == (eg is None), f"step {i}: oracle {eo!r} vs gpu {eg!r}"
        if eo is not None:
            raised = True
            break
        _same(o, g, i)
    assert raised, "the walk never reached the border"


def test_several_obstacle_updates_then_one_explore():
    """reality_policies.py:114-138: obstacle updates from several camera poses (explore=False), then a single explore call
    at another pose without depth (update_obstacles=False)."""
    fx = focal_from_hfov(160)
    o, g = _pair(600)
    fr = trajectory(4, 24, h=120, w=160, bound_m=6.0)
    for k in range(0, 24, 4):
        for f in fr[k:k + 3]:
            for m in (o, g):
                m.update_map(f.depth, f.tf, 0.5, 5.0 if k % 8 else 3.5, fx, fx, FOV, explore=False)
        f = fr[k + 3]
        for m in (o, g):
            m.update_map(None, f.tf, 0.5, 5.0, fx, fx, FOV, explore=True, update_obstacles=False)
        _same(o, g, k)
        assert np.array_equal(g._navigable_map, np.asarray(o._navigable_map).astype(np.int64))
