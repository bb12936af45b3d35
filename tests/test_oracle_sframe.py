"""The S-frame restriction of the explore step (vlfm_b200/mapping/sframe.py, csrc/explore.cu) is EXACT: running component
selection and the frontier search of the oracle inside the S frame -- with the one extra rule that unexplored components
touching a non-grid edge of the frame are never absorbed -- gives the same explored area and the same ordered frontier list as
the whole-grid oracle, step after step, including near the map edge.  Pure CPU (numpy / cv2)."""
import math

import numpy as np
import pytest

import oracle.explore_oracle as ex
from oracle.obstacle_map_oracle import ObstacleMapOracle
from vlfm_b200.mapping import sframe as sf
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory

FOV = np.deg2rad(79)


def _frontiers_in_frame(nav_s, grown_s, area_thresh, ext, origin):
    """detect_frontier_waypoints on the S frame; `ext` = (left, top, right, bottom) flags of non-grid edges."""
    h, w = nav_s.shape
    unexplored = nav_s.copy()
    unexplored[grown_s > 0] = 0
    out = grown_s.copy()
    for c in ex._contours(unexplored):
        p = c.reshape(-1, 2)
        touches = (ext[0] and (p[:, 0] == 0).any()) or (ext[1] and (p[:, 1] == 0).any()) or \
                  (ext[2] and (p[:, 0] == w - 1).any()) or (ext[3] and (p[:, 1] == h - 1).any())
        if touches:
            continue                                   # the exterior: never absorbed
        if ex._ct.contour_area(c) < area_thresh:
            m = ex._fill_contour(np.zeros_like(grown_s), c, 1)
            vals = set(unexplored[m.astype(bool)].tolist())
            if 1 in vals and len(vals) == 1:
                ex._fill_contour(out, c, 255)
    contours = ex._contours(out, simple=False)
    un = np.where(out > 0, 0, nav_s)
    import cv2

    blur = cv2.blur(np.where(un > 0, 255, un).astype(np.uint8), (3, 3))
    fronts = []
    for c in contours:
        fronts.extend(ex._split(ex._interpolate(c), blur))
    fronts = [f for f in fronts if len(f) >= 2]
    if not fronts:
        return np.array([])
    res = []
    for f in fronts:
        p = f.reshape(-1, 2).astype(np.float64) + np.array(origin, dtype=np.float64)        # grid coordinates BEFORE the arithmetic
        seg = np.sqrt(((p[1:] - p[:-1]) ** 2).sum(1))
        cum = np.cumsum(seg)
        half = cum[-1] / 2
        i = int(np.argmax(cum > half))
        before = cum[i - 1] if i > 0 else 0.0
        res.append(p[i] + (half - before) / seg[i] * (p[i + 1] - p[i]))
    return np.array(res)


class FramedOracle(ObstacleMapOracle):
    """ObstacleMapOracle whose explore half works in the S frame (the algorithm the GPU kernels implement)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.cover = None
        self.nav_valid = False
        self.frames = []

    def update_map(self, depth, tf, min_depth, max_depth, fx, fy, topdown_fov, explore=True, update_obstacles=True):
        g, ppm = self.size, self.ppm
        agent = self.xy_to_px(tf[:2, 3].reshape(1, 2))[0]
        if update_obstacles:
            self.update_obstacles(depth, tf, min_depth, max_depth, fx, fy)
            w = depth.shape[1]
            half = int(math.ceil(max_depth * ppm * math.sqrt(1.0 + (w / 2.0 / fx) ** 2))) + self.k // 2 + 2
            self.cover = sf.cover_add(self.cover, sf.obstacle_window(int(agent[0]), int(agent[1]), half, g), g)
        if not explore:
            return
        L = int(max_depth * ppm)
        self.cover = sf.cover_add(self.cover, sf.fog_window(int(agent[0]), int(agent[1]), L), g)
        x0, y0, x1, y1 = sf.sframe(self.cover, g, self.area_thresh_px)
        self.frames.append((x0, y0, x1, y1))
        yaw = float(np.arctan2(tf[1, 0], tf[0, 0]))
        nav = np.asarray(self._navigable_map).astype(np.uint8)
        new = ex.reveal_fog_of_war(nav, np.zeros_like(self._map, dtype=np.uint8), agent[::-1], -yaw, np.rad2deg(topdown_fov), max_depth * ppm)
        new = ex._dilate(new, 3)
        ex_s = ((self.explored_area[y0:y1, x0:x1] | (new[y0:y1, x0:x1] > 0)) & (nav[y0:y1, x0:x1] > 0)).astype(np.uint8)
        assert not (self.explored_area.sum() - self.explored_area[y0:y1, x0:x1].sum()), "explored cells outside the S frame"
        nav_s = nav[y0:y1, x0:x1]
        cs = ex._contours(ex_s)
        if len(cs) > 1:
            best, bd = 0, np.inf
            pt = (int(agent[0] - x0), int(agent[1] - y0))
            for i, c in enumerate(cs):
                d = ex._ppt(c, pt)
                if d >= 0:
                    best = i
                    break
                if abs(d) < bd:
                    best, bd = i, abs(d)
            ex_s = ex._fill_contour(np.zeros_like(ex_s), cs[best], 1)
        self.explored_area[:] = False
        self.explored_area[y0:y1, x0:x1] = ex_s.astype(bool)
        grown = ex._dilate(ex_s, 5)
        ext = (x0 > 0, y0 > 0, x1 < g, y1 < g)
        self._frontiers_px = _frontiers_in_frame(nav_s, grown, self.area_thresh_px, ext, (x0, y0))
        self.frontiers = self.px_to_xy(self._frontiers_px) if len(self._frontiers_px) else np.array([])


@pytest.mark.parametrize("cfg", [
    dict(seed=0, hw=(120, 160), size=1000, steps=14, bound=12.0, start=(0.0, 0.0)),
    dict(seed=1, hw=(120, 160), size=1000, steps=14, bound=12.0, start=(6.0, -9.0)),
    dict(seed=2, hw=(120, 160), size=600, steps=12, bound=6.0, start=(0.0, 0.0)),
    dict(seed=3, hw=(120, 160), size=1000, steps=10, bound=2.0, start=(17.5, 18.0), scale=0.3),      # corner of the map, close walls
    dict(seed=4, hw=(120, 160), size=1000, steps=10, bound=2.0, start=(-18.5, 3.0), scale=0.2),      # low edge (wrap-around rule)
    dict(seed=5, hw=(120, 160), size=1000, steps=12, bound=12.0, start=(0.0, 0.0), area=40.0),       # huge absorb threshold
])
def test_framed_explore_equals_whole_grid(cfg):
    h, w = cfg["hw"]
    fx = focal_from_hfov(w)
    kw = dict(area_thresh=cfg.get("area", 1.5), hole_area_thresh=-1, size=cfg["size"])
    a, b = ObstacleMapOracle(0.61, 0.88, 0.18, **kw), FramedOracle(0.61, 0.88, 0.18, **kw)
    small = 0
    for i, f in enumerate(trajectory(cfg["seed"], cfg["steps"], h=h, w=w, bound_m=cfg["bound"], start_xy=cfg["start"])):
        d = f.depth * np.float32(cfg.get("scale", 1.0))
        a.update_map(d, f.tf, 0.5, 5.0, fx, fx, FOV)
        b.update_map(d, f.tf, 0.5, 5.0, fx, fx, FOV)
        assert np.array_equal(a.explored_area, b.explored_area), f"explored area differs at step {i}"
        fa, fb = np.asarray(a._frontiers_px), np.asarray(b._frontiers_px)
        assert fa.shape == fb.shape and np.array_equal(fa, fb), f"frontiers differ at step {i}: {fa} vs {fb}"
        fr = b.frames[-1]
        small += (fr[2] - fr[0]) * (fr[3] - fr[1]) < cfg["size"] ** 2
    assert a.explored_area.sum() > 50
    if cfg["start"] == (0.0, 0.0):
        assert small == cfg["steps"]          # the frame really is a sub-rectangle on these runs
