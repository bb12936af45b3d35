"""Explore half of ObstacleMap on the GPU vs the restated oracle (bit-exact explored area, identical frontier
lists).  Parity is unpinned w.r.t. the absent frontier_exploration package (see oracle/explore_oracle.py)."""
import numpy as np
import pytest

from oracle.obstacle_map_oracle import ObstacleMapOracle
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [
    dict(seed=0, hw=(120, 160), size=400, steps=10, bound=4.0),
    dict(seed=3, hw=(240, 320), size=600, steps=10, bound=5.0),
    dict(seed=5, hw=(480, 640), size=1000, steps=8, bound=12.0),
])
def test_explore_vs_oracle(cfg):
    from vlfm_b200.mapping.obstacle_map import ObstacleMap

    h, w = cfg["hw"]
    fx = focal_from_hfov(w)
    o = ObstacleMapOracle(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=-1, size=cfg["size"])
    g = ObstacleMap(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=-1, size=cfg["size"])
    for i, f in enumerate(trajectory(cfg["seed"], cfg["steps"], h=h, w=w, bound_m=cfg["bound"])):
        o.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
        g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
        assert np.array_equal(g.explored_area, o.explored_area), f"explored area differs at step {i}: {g.explored_area.sum()} vs {o.explored_area.sum()}"
        fo, fg = np.asarray(o._frontiers_px), np.asarray(g._frontiers_px)
        assert fo.shape == fg.shape, f"frontier count differs at step {i}: {fg.shape} vs {fo.shape}"
        if fo.size:
            assert np.array_equal(fo, fg), f"frontier px differ at step {i}"
            assert np.array_equal(np.asarray(o.frontiers), np.asarray(g.frontiers))
    assert o.explored_area.sum() > 100
