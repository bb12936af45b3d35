"""OpenCV's behaviour for shapes that LEAVE the image (the fog-of-war cone and the occlusion rays of an agent close to
the map border), pinned against cv2 itself: clipLine before every line walk, PolyEdges built from clipped end points,
thickness-2 segments clipped to the image grown by two pixels."""
import cv2
import numpy as np

from oracle import cv_prims as P
from oracle.cv_draw import ellipse_sector, thick_line2


def test_clipped_thin_line():
    rng = np.random.default_rng(0)
    h, w = 90, 130
    for _ in range(1500):
        p0 = (int(rng.integers(-100, 250)), int(rng.integers(-100, 200)))
        p1 = (int(rng.integers(-300, 450)), int(rng.integers(-300, 400)))
        ref = cv2.line(np.zeros((h, w), np.uint8), p0, p1, 1, 1, 8) > 0
        xs, ys = P.line8_clipped(w, h, p0, p1)
        assert ((xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)).all()          # a clipped walk never leaves the image
        m = np.zeros((h, w), bool)
        m[ys, xs] = True
        assert np.array_equal(m, ref)


def test_sector_leaving_the_image():
    rng = np.random.default_rng(1)
    for t in range(300):
        h, w = int(rng.integers(120, 260)), int(rng.integers(120, 260))
        c = (int(rng.integers(-30, w + 30)), int(rng.integers(-30, h + 30)))
        r = int(rng.integers(15, 130))
        head, fov = float(rng.uniform(-200, 400)), [79.0, 90.0, 42.0, 60.5][t % 4]
        ref = cv2.ellipse(np.zeros((h, w), np.uint8), c, (r, r), 0, head - fov / 2, head + fov / 2, 1, -1)
        assert np.array_equal(ref > 0, ellipse_sector(h, w, c, r, head - fov / 2, head + fov / 2)), (h, w, c, r, head, fov)


def test_thick_segment_leaving_the_image():
    rng = np.random.default_rng(2)
    h, w = 150, 110
    for t in range(1500):
        p0 = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        if t % 4 == 0:                                                        # start point outside as well
            p0 = (int(rng.integers(-40, w + 40)), int(rng.integers(-40, h + 40)))
        p1 = (int(rng.integers(-300, 400)), int(rng.integers(-300, 450)))
        ref = np.zeros((h, w), np.uint8)
        cv2.polylines(ref, np.array([[p0, p1]], dtype=np.int32), isClosed=False, color=1, thickness=2)
        got = np.zeros((h, w), bool)
        thick_line2(got, p0, p1)
        assert np.array_equal(ref > 0, got), (p0, p1)


def test_drawcontours_polygon_leaving_the_image():
    rng = np.random.default_rng(3)
    R = 120
    for _ in range(200):
        n = int(rng.integers(3, 40))
        pts = np.stack([rng.integers(-60, 180, n), rng.integers(-60, 180, n)], 1).astype(np.int32)
        img = np.ones((R, R), np.uint8)
        cv2.drawContours(img, [pts], -1, 0, -1)
        assert np.array_equal(P.fill_polygon(R, R, pts), img == 0)
