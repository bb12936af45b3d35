"""numpy/python restatement of the OpenCV contour primitives used by the explore half
(obstacle_map.py:114-169 and the frontier_exploration functions it calls).

TEST INFRASTRUCTURE.  Pinned against cv2 4.13 by tests/test_oracle_contours.py.
These are the rules a GPU implementation of the explore half must follow (round-2 work):

* ``find_external_contours``   cv2.findContours(img, RETR_EXTERNAL, CHAIN_APPROX_NONE | SIMPLE):
  Suzuki-Abe border following.  Raster scan; a pixel starts an outer border when it is non-zero, its
  west neighbour is zero and it has not been traced yet; only borders whose parent is the frame
  (not nested inside a hole of another component) are returned; contours come out in REVERSE order
  of discovery; each starts at its raster-first pixel and runs counter-clockwise in image
  coordinates (first step towards +y).
* ``approx_simple``            CHAIN_APPROX_SIMPLE: keep the points where the chain direction changes.
* ``contour_area``             cv2.contourArea (shoelace, absolute value).
* ``is_convex``                cv2.isContourConvex (sign consistency of consecutive cross products).
* ``point_polygon_distance``   cv2.pointPolygonTest(cnt, pt, True) for integer contours.
"""
from __future__ import annotations

from typing import List

import numpy as np

# 8-neighbourhood in clockwise order starting at west (dx, dy), image coordinates (y down)
_CW = [(-1, 0), (-1, -1), (0, -1), (1, -1), (1, 0), (1, 1), (0, 1), (-1, 1)]
_IDX = {d: i for i, d in enumerate(_CW)}


def _trace(f: np.ndarray, x0: int, y0: int, fx: int, fy: int, nbd: int) -> List[tuple]:
    """Suzuki-Abe steps 3.1-3.5 on the working image f (int32, zero padded by 1): follow the border that
    starts at (x0, y0) entered from the zero pixel (fx, fy); marks visited pixels with +-nbd."""
    start_dir = _IDX[(fx - x0, fy - y0)]
    # 3.1 clockwise from the entry pixel: first non-zero neighbour
    first = None
    for k in range(1, 9):
        d = _CW[(start_dir + k) % 8]
        if f[y0 + d[1], x0 + d[0]] != 0:
            first = (x0 + d[0], y0 + d[1])
            break
    if first is None:
        f[y0, x0] = -nbd
        return [(x0, y0)]
    pts = []
    x2, y2 = first
    x3, y3 = x0, y0
    while True:
        # 3.3 counter-clockwise around (x3,y3) starting after (x2,y2)
        d0 = _IDX[(x2 - x3, y2 - y3)]
        east_zero_examined = False
        for k in range(1, 9):
            di = (d0 - k) % 8
            d = _CW[di]
            if f[y3 + d[1], x3 + d[0]] != 0:
                x4, y4 = x3 + d[0], y3 + d[1]
                break
            if d == (1, 0):
                east_zero_examined = True
        # 3.4
        if east_zero_examined:
            f[y3, x3] = -nbd
        elif f[y3, x3] == 1:
            f[y3, x3] = nbd
        pts.append((x3, y3))
        # 3.5
        if (x4, y4) == (x0, y0) and (x3, y3) == first:
            break
        x2, y2 = x3, y3
        x3, y3 = x4, y4
    return pts


def find_all_contours(img: np.ndarray) -> List[np.ndarray]:
    """Every border cv2.findContours(img, RETR_TREE, CHAIN_APPROX_NONE) returns (outer borders of all
    components and all hole borders), as a set -- order is not restated (fill_small_holes takes a union)."""
    return find_external_contours(img, False, _all=True)


def find_external_contours(img: np.ndarray, simple: bool = False, _all: bool = False) -> List[np.ndarray]:
    """-> list of (n,1,2) int32 arrays exactly like cv2.findContours(img, RETR_EXTERNAL, ...)."""
    h, w = img.shape
    f = np.zeros((h + 2, w + 2), dtype=np.int32)
    f[1:-1, 1:-1] = (img != 0).astype(np.int32)
    nbd = 1
    borders = {1: ("hole", 0)}          # id -> (kind, parent); the frame is a hole border with id 1
    out = []
    for y in range(1, h + 1):
        lnbd = 1
        for x in range(1, w + 1):
            v = f[y, x]
            if v == 0:
                continue
            kind = None
            if v == 1 and f[y, x - 1] == 0:
                kind, fx, fy = "outer", x - 1, y
            elif v >= 1 and f[y, x + 1] == 0:
                kind, fx, fy = "hole", x + 1, y
                if v > 1:
                    lnbd = v
            if kind is not None:
                nbd += 1
                pk, pp = borders[abs(lnbd)]
                parent = (pp if pk == "outer" else abs(lnbd)) if kind == "outer" else (abs(lnbd) if pk == "outer" else pp)
                borders[nbd] = (kind, parent)
                pts = _trace(f, x, y, fx, fy, nbd)
                if _all or (kind == "outer" and parent == 1):
                    out.append(np.array([(px - 1, py - 1) for px, py in pts], dtype=np.int32).reshape(-1, 1, 2))
            if f[y, x] != 1:
                lnbd = abs(f[y, x])
    out.reverse()
    if simple:
        out = [approx_simple(c) for c in out]
    return out


def approx_simple(chain: np.ndarray) -> np.ndarray:
    p = chain.reshape(-1, 2)
    n = len(p)
    if n <= 2:
        return chain
    nxt = np.roll(p, -1, axis=0) - p
    prv = p - np.roll(p, 1, axis=0)
    keep = np.any(nxt != prv, axis=1)
    return p[keep].reshape(-1, 1, 2)


def contour_area(c: np.ndarray) -> float:
    p = c.reshape(-1, 2).astype(np.float64)
    if len(p) == 0:
        return 0.0
    q = np.roll(p, 1, axis=0)
    return float(abs(np.sum(q[:, 0] * p[:, 1] - p[:, 0] * q[:, 1]) * 0.5))


def is_convex(c: np.ndarray) -> bool:
    """cv2.isContourConvex for an integer contour: every consecutive turn has the same strict sign
    (a collinear triple makes the contour non-convex)."""
    p = c.reshape(-1, 2).astype(np.int64)
    n = len(p)
    if n == 0:
        return False
    prev, cur = p[(n - 2) % n], p[n - 1]
    dx0, dy0 = cur[0] - prev[0], cur[1] - prev[1]
    orient = 0
    for i in range(n):
        prev, cur = cur, p[i]
        dx, dy = cur[0] - prev[0], cur[1] - prev[1]
        dxdy0, dydx0 = dx * dy0, dy * dx0
        orient |= 1 if dydx0 > dxdy0 else (2 if dydx0 < dxdy0 else 3)
        if orient == 3:
            return False
        dx0, dy0 = dx, dy
    return True


def point_polygon_distance(c: np.ndarray, pt) -> float:
    """cv2.pointPolygonTest(c, pt, True): signed distance (>0 inside) to an integer contour, float32 points,
    float64 accumulation exactly as OpenCV evaluates it."""
    p = c.reshape(-1, 2)
    n = len(p)
    if n == 0:
        return -np.finfo(np.float64).max
    px, py = np.float32(pt[0]), np.float32(pt[1])
    v = p[n - 1].astype(np.float32)
    min_num, min_den = float(np.finfo(np.float32).max), 1.0
    counter = 0
    for i in range(n):
        v0, v = v, p[i].astype(np.float32)
        dx, dy = float(v[0] - v0[0]), float(v[1] - v0[1])
        dx1, dy1 = float(px - v0[0]), float(py - v0[1])
        dx2, dy2 = float(px - v[0]), float(py - v[1])
        den = 1.0
        if dx1 * dx + dy1 * dy <= 0:
            num = dx1 * dx1 + dy1 * dy1
        elif dx2 * dx + dy2 * dy >= 0:
            num = dx2 * dx2 + dy2 * dy2
        else:
            num = dy1 * dx - dx1 * dy
            num *= num
            den = dx * dx + dy * dy
        if num * min_den < min_num * den:
            min_num, min_den = num, den
            if min_num == 0:
                break
        if (v0[1] <= py and v[1] <= py) or (v0[1] > py and v[1] > py):
            continue
        cross = dy1 * dx - dx1 * dy
        if dy < 0:
            cross = -cross
        counter += cross > 0
    r = float(np.sqrt(min_num / min_den))
    return r if counter % 2 else -r
