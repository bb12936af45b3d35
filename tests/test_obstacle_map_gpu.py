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
