"""numpy restatement of the OpenCV rasterisation rules on the VLFM hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against cv2 4.13 itself by
tests/test_oracle_cv_prims.py (random segments / polygons / angles, zero mismatches).

These are the rules the CUDA kernels re-implement, written so the kernels can be read
against them:

* ``line8``            cv2.line(thickness=1, lineType=8) pixel set
* ``clip_line``        cv::clipLine, applied by cv2 before any line is walked (``line8_clipped``)
                       (used by cv2.drawContours outline; reference call site
                       vlfm/mapping/value_map.py:260).
* ``fill_polygon``     cv2.drawContours(img, [poly], -1, c, -1): outline + even-odd
                       16.16 fixed-point scanline interior (value_map.py:260).
* ``rotate_bilinear``  vlfm/utils/img_utils.py:9-28 (cv2.getRotationMatrix2D +
                       cv2.warpAffine INTER_LINEAR, border 0) with OpenCV's 10-bit
                       fixed-point coordinates and 5-bit interpolation weights.
* ``dilate_box``       cv2.dilate with an all-ones k x k kernel (zero padded)
                       (vlfm/mapping/obstacle_map.py:105-109,125,159-163).
"""
from __future__ import annotations

import math

import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS


def line_minor_steps(k, major: int, minor: int):
    """Number of minor-axis steps taken before the k-th plotted pixel of an OpenCV
    LineIterator (8-connected).  Closed form of the err<0 recurrence:
    s_k = ceil(minor*k/major - 1/2) = floor((2*minor*k + major - 1) / (2*major))."""
    if major == 0:
        return k * 0
    return (2 * minor * k + major - 1) // (2 * major)


def line8(p0, p1):
    """Pixels (x, y) of cv2.line(img, p0, p1, c, 1, 8), unclipped, in plot order."""
    x0, y0 = int(p0[0]), int(p0[1])
    x1, y1 = int(p1[0]), int(p1[1])
    if x1 < x0:  # iterator always walks towards +x
        x0, y0, x1, y1 = x1, y1, x0, y0
    dx = x1 - x0
    dy = y1 - y0
    sy = 1 if dy >= 0 else -1
    ady = abs(dy)
    k = np.arange(max(dx, ady) + 1, dtype=np.int64)
    if ady > dx:  # y is the major axis
        s = line_minor_steps(k, ady, dx)
        xs = x0 + s
        ys = y0 + sy * k
    else:
        s = line_minor_steps(k, dx, ady)
        xs = x0 + k
        ys = y0 + sy * s
    return xs, ys


def clip_line(w: int, h: int, p1, p2):
    """cv::clipLine(Size2l(w, h), pt1, pt2) (drawing.cpp): Cohen-Sutherland with the intersection computed in double and
    truncated toward zero.  Returns (visible, (x1, y1), (x2, y2)).  cv2 clips every line to the image BEFORE walking it
    (LineIterator for thickness-1 lines, Line2 for the fixed-point outlines of FillConvexPoly), so the pixels of a line
    that leaves the image are those of the CLIPPED segment, not the visible part of the unclipped one."""
    x1, y1, x2, y2 = int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1])
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return False, (x1, y1), (x2, y2)

    def tr(a: float) -> int:            # (int64)double
        return int(a)

    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += tr(float(a - y1) * float(x2 - x1) / float(y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += tr(float(a - y2) * float(x2 - x1) / float(y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += tr(float(a - x1) * float(y2 - y1) / float(x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += tr(float(a - x2) * float(y2 - y1) / float(x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def line8_clipped(w: int, h: int, p0, p1):
    """Pixels of cv2.line(img[h, w], p0, p1, c, 1, 8): LineIterator clips to the image first, then walks the clipped segment."""
    ok, a, b = clip_line(w, h, p0, p1)
    if not ok:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return line8(a, b)


def fill_polygon(h: int, w: int, pts: np.ndarray) -> np.ndarray:
    """Boolean (h, w) mask of the cells cv2.drawContours(img,[pts],-1,c,-1) writes.

    pts: (n, 2) integer (x=col, y=row) vertices of one closed contour.
    """
    pts = np.asarray(pts, dtype=np.int64).reshape(-1, 2)
    n = len(pts)
    out = np.zeros((h, w), dtype=bool)
    nxt = np.roll(np.arange(n), -1)
    # --- outline: every edge, including the closing one, as an 8-connected line
    for i in range(n):
        xs, ys = line8_clipped(w, h, pts[i], pts[nxt[i]])
        ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
        out[ys[ok], xs[ok]] = True
    # --- interior: even-odd scanline with 16.16 intercepts.  An edge with an end point outside the image is built from the
    # CLIPPED end points (cv2 4.13 CollectPolyEdges, see oracle/cv_draw.py::poly_edge)
    toggle = np.zeros((h, w + 1), dtype=np.int64)
    exact = np.zeros((h, w), dtype=bool)
    for i in range(n):
        ax, ay, bx, by = int(pts[i, 0]), int(pts[i, 1]), int(pts[nxt[i], 0]), int(pts[nxt[i], 1])
        if ay == by:
            continue
        c0, c1 = (ax << 16, ay), (bx << 16, by)
        if not (0 <= ax < w and 0 <= bx < w and 0 <= ay < h and 0 <= by < h):
            _, a, b = clip_line(w, h, (ax, ay), (bx, by))
            if a[1] != b[1]:
                c0, c1 = (a[0] << 16, a[1]), (b[0] << 16, b[1])
            else:
                c0, c1 = (a[0] << 16, ay), (b[0] << 16, by)
        num, den = c1[0] - c0[0], c1[1] - c0[1]
        dxe = abs(num) // abs(den) * (1 if (num >= 0) == (den > 0) else -1)  # C truncating division
        if ay < by:
            ya, yb, xs16 = ay, by, c0[0] + (ay - c0[1]) * dxe
        else:
            ya, yb, xs16 = by, ay, c1[0] + (by - c1[1]) * dxe
        r = np.arange(max(ya, 0), min(yb, h), dtype=np.int64)
        if r.size == 0:
            continue
        X = xs16 + dxe * (r - ya)
        t = np.clip((X >> 16) + 1, 0, w)  # first column c with (c<<16) > X
        np.add.at(toggle, (r, t), 1)
        hit = ((X & 0xFFFF) == 0) & ((X >> 16) >= 0) & ((X >> 16) < w)
        exact[r[hit], (X >> 16)[hit]] = True
    less_cnt = np.cumsum(toggle[:, :w], axis=1)  # #{intercepts < c<<16}
    out |= exact | ((less_cnt & 1) == 1)
    return out


def rotation_inverse_matrix(cx: float, cy: float, radians: float) -> np.ndarray:
    """The 2x3 matrix cv2.warpAffine ends up using for rotate_image(img, radians):
    invert(getRotationMatrix2D((cx,cy), degrees(radians), 1.0)), same operation order
    as OpenCV (double precision, no fused multiply-add)."""
    angle = float(np.degrees(radians)) * (math.pi / 180.0)
    a = math.cos(angle)
    b = math.sin(angle)
    m00, m01, m02 = a, b, (1.0 - a) * cx - b * cy
    m10, m11, m12 = -b, a, b * cx + (1.0 - a) * cy
    det = m00 * m11 - m01 * m10
    det = 1.0 / det if det != 0.0 else 0.0
    a11 = m11 * det
    a22 = m00 * det
    i00 = a11
    i01 = m01 * (-det)
    i10 = m10 * (-det)
    i11 = a22
    b1 = -i00 * m02 - i01 * m12
    b2 = -i10 * m02 - i11 * m12
    return np.array([[i00, i01, b1], [i10, i11, b2]], dtype=np.float64)


def warp_tables(mi: np.ndarray, w: int, h: int):
    """Fixed-point coordinate tables of cv2.warpAffine (AB_BITS=10, round delta 16)."""
    xs = np.arange(w, dtype=np.float64)
    ys = np.arange(h, dtype=np.float64)
    adelta = np.rint(mi[0, 0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(mi[1, 0] * xs * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // INTER_TAB // 2
    x0 = np.rint((mi[0, 1] * ys + mi[0, 2]) * AB_SCALE).astype(np.int64) + rd
    y0 = np.rint((mi[1, 1] * ys + mi[1, 2]) * AB_SCALE).astype(np.int64) + rd
    return x0, y0, adelta, bdelta


def rotate_bilinear(img: np.ndarray, radians: float) -> np.ndarray:
    """rotate_image(img, radians) for a 2-D float image, bit-exact vs cv2."""
    h, w = img.shape
    mi = rotation_inverse_matrix(w // 2, h // 2, radians)
    x0, y0, ad, bd = warp_tables(mi, w, h)
    X = (x0[:, None] + ad[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (y0[:, None] + bd[None, :]) >> (AB_BITS - INTER_BITS)
    sx = X >> INTER_BITS
    sy = Y >> INTER_BITS
    ax = (X & (INTER_TAB - 1)).astype(np.float64) / INTER_TAB
    ay = (Y & (INTER_TAB - 1)).astype(np.float64) / INTER_TAB
    src = np.asarray(img, dtype=np.float64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], 0.0)

    out = (
        tap(sy, sx) * ((1 - ax) * (1 - ay))
        + tap(sy, sx + 1) * (ax * (1 - ay))
        + tap(sy + 1, sx) * ((1 - ax) * ay)
        + tap(sy + 1, sx + 1) * (ax * ay)
    )
    return out.astype(img.dtype)


def dilate_box(img: np.ndarray, k: int) -> np.ndarray:
    """cv2.dilate(img, ones((k,k))) for a 2-D 0/1 (or uint8) image, zero padded."""
    r = k // 2
    h, w = img.shape
    p = np.zeros((h + 2 * r, w + 2 * r), dtype=img.dtype)
    p[r : r + h, r : r + w] = img
    rows = p[:, 0:w].copy()
    for d in range(1, k):
        np.maximum(rows, p[:, d : d + w], out=rows)
    out = rows[0:h].copy()
    for d in range(1, k):
        np.maximum(out, rows[d : d + h], out=out)
    return out
