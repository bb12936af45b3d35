"""The numpy restatement of OpenCV's rasterisation rules (the rules the CUDA kernels
implement) is pinned against cv2 itself."""
import cv2
import numpy as np

from oracle import cv_prims as P


def test_line8_matches_cv2():
    rng = np.random.default_rng(1)
    for _ in range(1500):
        p0, p1 = rng.integers(0, 60, 2), rng.integers(0, 60, 2)
        img = np.zeros((60, 60), np.uint8)
        cv2.line(img, tuple(int(v) for v in p0), tuple(int(v) for v in p1), 1, 1, 8)
        xs, ys = P.line8(p0, p1)
        m = np.zeros((60, 60), bool)
        m[ys, xs] = True
        assert np.array_equal(m, img > 0)


def test_fill_polygon_matches_drawcontours():
    rng = np.random.default_rng(2)
    R = 201
    for t in range(80):
        n = int(rng.integers(5, 300))
        if t % 2 == 0:
            xs, ys = rng.integers(0, R, n), rng.integers(0, R, n)
        else:
            ys = np.clip((150 + 30 * np.sin(np.linspace(0, 6, n)) + rng.normal(0, 8 if t % 4 == 1 else 0.5, n)).astype(int), 0, R - 1)
            xs = np.sort(rng.integers(18, 183, n))
        pts = np.concatenate([[[0, R - 1]], np.stack([xs, ys], 1), [[R - 1, R - 1]]], 0).astype(np.int32)
        img = np.ones((R, R), np.float64)
        cv2.drawContours(img, [pts], -1, 0, -1)
        assert np.array_equal(P.fill_polygon(R, R, pts), img == 0)


def test_rotate_bilinear_bit_exact():
    rng = np.random.default_rng(3)
    src = rng.random((201, 201)).astype(np.float32).astype(np.float64)
    for ang in list(rng.uniform(-7, 7, 25)) + [0.0, np.pi / 2, -np.pi / 2, np.pi, 1e-9]:
        m = cv2.getRotationMatrix2D((100, 100), np.degrees(ang), 1.0)
        ref = cv2.warpAffine(src, m, (201, 201), borderValue=0)
        assert np.array_equal(ref, P.rotate_bilinear(src, ang))


def test_dilate_box():
    rng = np.random.default_rng(4)
    for t in range(12):
        k = [3, 5, 7][t % 3]
        im = (rng.random((120, 90)) < 0.02).astype(np.uint8)
        assert np.array_equal(cv2.dilate(im, np.ones((k, k), np.uint8)), P.dilate_box(im, k))
