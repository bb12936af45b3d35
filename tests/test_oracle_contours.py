"""oracle/contours.py and oracle/cv_draw.py (the OpenCV rules of the explore half) pinned against cv2."""
import cv2
import numpy as np

from oracle.contours import contour_area, find_external_contours, is_convex, point_polygon_distance
from oracle.cv_draw import blur3, ellipse_sector, thick_line2


def _random_masks(n, seed):
    rng = np.random.default_rng(seed)
    for t in range(n):
        h, w = int(rng.integers(5, 40)), int(rng.integers(5, 40))
        img = (rng.random((h, w)) < [0.2, 0.45, 0.6, 0.8][t % 4]).astype(np.uint8)
        if t % 5 == 0:
            img = cv2.dilate(img, np.ones((3, 3), np.uint8))
        yield rng, img


def test_find_external_contours_none_and_simple():
    for _, img in _random_masks(120, 0):
        for simple, flag in ((False, cv2.CHAIN_APPROX_NONE), (True, cv2.CHAIN_APPROX_SIMPLE)):
            ref, _h = cv2.findContours(img, cv2.RETR_EXTERNAL, flag)
            got = find_external_contours(img, simple)
            assert len(ref) == len(got)
            for a, b in zip(ref, got):      # same order, same start point, same direction
                assert np.array_equal(a, b)


def test_area_convexity_and_point_distance():
    for rng, img in _random_masks(80, 1):
        h, w = img.shape
        for c in cv2.findContours(img, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]:
            assert abs(cv2.contourArea(c) - contour_area(c)) < 1e-9
            assert bool(cv2.isContourConvex(c)) == is_convex(c)
            for _ in range(4):
                pt = (int(rng.integers(-3, w + 3)), int(rng.integers(-3, h + 3)))
                assert cv2.pointPolygonTest(c, pt, True) == point_polygon_distance(c, pt)


def test_ellipse_sector_inside_image():
    """circle inside the image; sectors clipped by the image border: tests/test_oracle_cv_draw.py"""
    rng = np.random.default_rng(2)
    for t in range(80):
        G, r = int(rng.integers(260, 400)), int(rng.integers(15, 120))
        c = (int(rng.integers(r, G - r)), int(rng.integers(r, G - r)))
        head, fov = float(rng.uniform(-200, 400)), [79.0, 90.0, 42.0, 60.5][t % 4]
        ref = cv2.ellipse(np.zeros((G, G), np.uint8), c, (r, r), 0, head - fov / 2, head + fov / 2, 1, -1)
        assert np.array_equal(ref > 0, ellipse_sector(G, G, c, r, head - fov / 2, head + fov / 2))


def test_thick_line_and_blur():
    rng = np.random.default_rng(3)
    G = 120
    for t in range(400):
        p0 = (int(rng.integers(5, G - 5)), int(rng.integers(5, G - 5)))
        p1 = (p0[0] + int(rng.integers(-4, 5)), p0[1] + int(rng.integers(-4, 5))) if t % 3 == 0 else (int(rng.integers(5, G - 5)), int(rng.integers(5, G - 5)))
        p1 = (min(max(p1[0], 3), G - 4), min(max(p1[1], 3), G - 4))
        ref = np.zeros((G, G), np.uint8)
        cv2.polylines(ref, np.array([[p0, p1]], dtype=np.int32), isClosed=False, color=1, thickness=2)
        got = np.zeros((G, G), bool)
        thick_line2(got, p0, p1)
        assert np.array_equal(ref > 0, got)
    for _ in range(10):
        im = (rng.integers(0, 2, (40, 50)) * 255).astype(np.uint8)
        assert np.array_equal(cv2.blur(im, (3, 3)), blur3(im))
