"""numpy/python restatement of the OpenCV drawing primitives used by the explore half
(fog-of-war cone and occlusion rays).  TEST INFRASTRUCTURE; pinned to cv2 4.13 by
tests/test_oracle_cv_draw.py.

* ``fill_poly_fixed``   CollectPolyEdges + FillEdgeCollection for vertices in 16.16 fixed point
                        (rows rounded, columns kept fractional), incl. the 8-connected outline.
* ``ellipse_sector``    cv2.ellipse(img, c, (r,r), 0, a0, a1, color, -1): integer-degree ellipse2Poly
                        polygon (5-degree steps for r >= 15) + centre, filled with fill_poly_fixed.
* ``thick_line2``       cv2.line / cv2.polylines with thickness=2: FillConvexPoly of the 2-px wide
                        rectangle (fixed-point normal) + radius-1 discs (plus shapes) at both ends.
* ``blur3``             cv2.blur(img, (3,3)) on uint8 (BORDER_REFLECT_101, round half up).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from .cv_prims import clip_line, line8_clipped

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT


def cv_round(x: float) -> int:
    """cvRound: round half to even (lrint)."""
    return int(np.rint(x))


def _plot(mask: np.ndarray, xs, ys) -> None:
    h, w = mask.shape
    xs, ys = np.asarray(xs), np.asarray(ys)
    ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
    mask[ys[ok], xs[ok]] = True


def poly_edge(w: int, h: int, p0: Tuple[int, int], p1: Tuple[int, int]):
    """One PolyEdge of CollectPolyEdges (cv2 4.13): p0, p1 = (x in 16.16, integer row).  Returns (y0, y1, x_at_y0, dx)
    or None for a horizontal edge.  When a (pixel-rounded) end point lies outside the image, cv2 clips the segment
    (clipLine, integer pixels) and builds the edge from the CLIPPED columns -- always -- and from the clipped rows when
    they differ, extrapolating back to the original first row (pinned by tests/test_oracle_cv_draw.py)."""
    t0 = ((p0[0] + (XY_ONE >> 1)) >> XY_SHIFT, p0[1])
    t1 = ((p1[0] + (XY_ONE >> 1)) >> XY_SHIFT, p1[1])
    c0, c1 = p0, p1
    if not (0 <= t0[0] < w and 0 <= t1[0] < w and 0 <= t0[1] < h and 0 <= t1[1] < h):
        _, a, b = clip_line(w, h, t0, t1)
        if a[1] != b[1]:
            c0, c1 = (a[0] << XY_SHIFT, a[1]), (b[0] << XY_SHIFT, b[1])
        else:
            c0, c1 = (a[0] << XY_SHIFT, p0[1]), (b[0] << XY_SHIFT, p1[1])
    if p0[1] == p1[1]:
        return None
    num, den = c1[0] - c0[0], c1[1] - c0[1]
    dx = abs(num) // abs(den) * (1 if (num >= 0) == (den > 0) else -1)           # C truncating division
    if p0[1] < p1[1]:
        return (p0[1], p1[1], c0[0] + (p0[1] - c0[1]) * dx, dx)
    return (p1[1], p0[1], c1[0] + (p1[1] - c1[1]) * dx, dx)


def fill_poly_fixed(mask: np.ndarray, v: List[Tuple[int, int]]) -> None:
    """v: closed polygon, (x, y) in 16.16 fixed point.  Sets the cells cv2 would write."""
    h, w = mask.shape
    pts = [(int(x), (int(y) + (XY_ONE >> 1)) >> XY_SHIFT) for x, y in v]       # x: 16.16, y: rounded row
    edges = []
    p0 = pts[-1]
    for p1 in pts:
        t0 = ((p0[0] + (XY_ONE >> 1)) >> XY_SHIFT, p0[1])
        t1 = ((p1[0] + (XY_ONE >> 1)) >> XY_SHIFT, p1[1])
        _plot(mask, *line8_clipped(w, h, t0, t1))
        e = poly_edge(w, h, p0, p1)
        if e is not None:
            edges.append(e)
        p0 = p1
    if not edges:
        return
    y_lo = max(min(e[0] for e in edges), 0)
    y_hi = min(max(e[1] for e in edges), h)
    for y in range(y_lo, y_hi):
        xs = sorted(x + dx * (y - y0) for (y0, y1, x, dx) in edges if y0 <= y < y1)
        for a, b in zip(xs[0::2], xs[1::2]):
            x1 = (a + XY_ONE - 1) >> XY_SHIFT
            x2 = b >> XY_SHIFT
            if x1 <= x2 and x2 >= 0 and x1 < w:
                mask[y, max(x1, 0) : min(x2, w - 1) + 1] = True


# OpenCV's SinTable: float literals with seven decimals (0.0174524f, 0.0348995f, ...)
_SIN = [float(np.float32(round(math.sin(math.radians(a)), 7))) for a in range(0, 451)]


def ellipse_sector(h: int, w: int, center: Tuple[int, int], radius: int, start_deg: float, end_deg: float) -> np.ndarray:
    """Boolean mask of cv2.ellipse(zeros(h,w), center, (radius,radius), 0, start_deg, end_deg, 1, -1)."""
    mask = np.zeros((h, w), dtype=bool)
    a0, a1 = cv_round(start_deg), cv_round(end_deg)
    cx, cy, ax = center[0] << XY_SHIFT, center[1] << XY_SHIFT, abs(radius) << XY_SHIFT
    delta = (ax + (XY_ONE >> 1)) >> XY_SHIFT
    delta = 90 if delta < 3 else 30 if delta < 10 else 18 if delta < 15 else 5
    if a0 > a1:
        a0, a1 = a1, a0
    while a0 < 0:
        a0 += 360; a1 += 360
    while a1 > 360:
        a1 -= 360; a0 -= 360
    if a1 - a0 > 360:
        a0, a1 = 0, 360
    pts = []
    i = a0
    while i < a1 + delta:
        ang = min(i, a1)
        if ang < 0:
            ang += 360
        x = ax * _SIN[450 - ang]
        y = ax * _SIN[ang]
        pts.append((cx + x * 1.0 - y * 0.0, cy + x * 0.0 + y * 1.0))      # alpha = cos(0) = 1, beta = sin(0) = 0
        i += delta
    if len(pts) == 1:
        pts = [(float(cx), float(cy))] * 2
    v, prev = [], None
    for fx, fy in pts:
        px = cv_round(fx / XY_ONE) << XY_SHIFT
        py = cv_round(fy / XY_ONE) << XY_SHIFT
        px += cv_round(fx - px)
        py += cv_round(fy - py)
        if (px, py) != prev:
            v.append((px, py)); prev = (px, py)
    if len(v) <= 1:
        v = [(cx, cy)] * 2
    if a1 - a0 >= 360:
        raise NotImplementedError("full ellipse (FillConvexPoly path) is not needed by the explore half")
    v.append((cx, cy))
    fill_poly_fixed(mask, v)
    return mask


def _line2(mask: np.ndarray, p1: Tuple[int, int], p2: Tuple[int, int]) -> None:
    """drawing.cpp Line2: clipLine against the image scaled to 16.16, then a DDA between the clipped endpoints."""
    h, w = mask.shape
    ok, (x1, y1), (x2, y2) = clip_line(w << XY_SHIFT, h << XY_SHIFT, p1, p2)
    if not ok:
        return
    dx, dy = x2 - x1, y2 - y1
    ax, ay = abs(dx), abs(dy)

    def put(x, y):
        if 0 <= x < w and 0 <= y < h:
            mask[y, x] = True

    def cdiv(a, b):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b > 0) else -q

    if ax > ay:
        if dx < 0:
            x1, y1, x2, y2 = x2, y2, x1, y1
            dy = -dy
        x_step, y_step = XY_ONE, cdiv(dy << XY_SHIFT, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        if dy < 0:
            x1, y1, x2, y2 = x2, y2, x1, y1
            dx = -dx
        x_step, y_step = cdiv(dx << XY_SHIFT, ay | 1), XY_ONE
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += XY_ONE >> 1
    y1 += XY_ONE >> 1
    put((x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT)
    if ax > ay:
        x = x1 >> XY_SHIFT
        y = y1
        while ecount >= 0:
            put(x, y >> XY_SHIFT)
            x += 1; y += y_step; ecount -= 1
    else:
        y = y1 >> XY_SHIFT
        x = x1
        while ecount >= 0:
            put(x >> XY_SHIFT, y)
            x += x_step; y += 1; ecount -= 1


def fill_convex_fixed(mask: np.ndarray, v: List[Tuple[int, int]]) -> None:
    """drawing.cpp FillConvexPoly(..., shift=XY_SHIFT), line_type 8: Line2 outline + two-edge scan."""
    h, w = mask.shape
    n = len(v)
    delta = XY_ONE >> 1
    p0 = v[-1]
    for p in v:
        _line2(mask, p0, p)
        p0 = p
    ys = [p[1] for p in v]
    xs = [p[0] for p in v]
    imin = int(np.argmin(ys))          # first minimum, as the `<` scan in OpenCV
    ymin = (min(ys) + delta) >> XY_SHIFT
    ymax = (max(ys) + delta) >> XY_SHIFT
    xmin = (min(xs) + delta) >> XY_SHIFT
    xmax = (max(xs) + delta) >> XY_SHIFT
    if n < 3 or xmax < 0 or ymax < 0 or xmin >= w or ymin >= h:
        return
    ymax = min(ymax, h - 1)
    edge = [dict(idx=imin, di=1, x=-XY_ONE, dx=0, ye=ymin), dict(idx=imin, di=n - 1, x=-XY_ONE, dx=0, ye=ymin)]
    edges_left = n
    y = ymin
    while True:
        for e in edge:
            if y >= e["ye"]:
                idx0, di = e["idx"], e["di"]
                idx = (idx0 + di) % n
                while True:
                    edges_left -= 1
                    if edges_left < 0:
                        break
                    ty = (v[idx][1] + delta) >> XY_SHIFT
                    if ty > y:
                        xs_, xe_ = v[idx0][0], v[idx][0]
                        num, den = (xe_ - xs_) * 2 + (ty - y), 2 * (ty - y)
                        e["dx"] = abs(num) // den * (1 if num >= 0 else -1)
                        e["ye"], e["x"], e["idx"] = ty, xs_, idx
                        break
                    idx0 = idx
                    idx = (idx + di) % n
        if edges_left < 0:
            break
        if y >= 0:
            l, r = (edge[0], edge[1]) if edge[0]["x"] <= edge[1]["x"] else (edge[1], edge[0])
            xx1 = (l["x"] + delta) >> XY_SHIFT
            xx2 = (r["x"] + delta) >> XY_SHIFT
            if xx2 >= 0 and xx1 < w:
                mask[y, max(xx1, 0) : min(xx2, w - 1) + 1] = True
        edge[0]["x"] += edge[0]["dx"]
        edge[1]["x"] += edge[1]["dx"]
        y += 1
        if y > ymax:
            break


def thick_line2(mask: np.ndarray, p0: Tuple[int, int], p1: Tuple[int, int]) -> None:
    """cv2.line / cv2.polylines(img, p0, p1, c, thickness=2, lineType=8) (drawing.cpp ThickLine).

    cv2 4.13 first clips the integer centre line to the image rectangle grown by the thickness (2 px) on every side
    (pinned empirically by tests/test_oracle_cv_draw.py: no other margin reproduces cv2 for segments that leave the image);
    the 2-px rectangle and the end discs are then built from the CLIPPED end points."""
    h, w = mask.shape
    t = 2
    ok, a, b = clip_line(w + 2 * t, h + 2 * t, (p0[0] + t, p0[1] + t), (p1[0] + t, p1[1] + t))
    if not ok:
        return
    p0, p1 = (a[0] - t, a[1] - t), (b[0] - t, b[1] - t)
    x0, y0, x1, y1 = p0[0] << XY_SHIFT, p0[1] << XY_SHIFT, p1[0] << XY_SHIFT, p1[1] << XY_SHIFT
    dx, dy = (x0 - x1) / XY_ONE, (y1 - y0) / XY_ONE
    r = dx * dx + dy * dy
    thickness = 2 << (XY_SHIFT - 1)
    if abs(r) > np.finfo(np.float64).eps:
        r = thickness / math.sqrt(r)
        dpx, dpy = cv_round(dy * r), cv_round(dx * r)
        fill_convex_fixed(mask, [(x0 + dpx, y0 + dpy), (x0 - dpx, y0 - dpy), (x1 - dpx, y1 - dpy), (x1 + dpx, y1 + dpy)])
    for cx, cy in (p0, p1):        # Circle(center, radius = 1, filled): a plus shape
        _plot(mask, [cx - 1, cx, cx + 1, cx, cx], [cy, cy, cy, cy - 1, cy + 1])


def blur3(img: np.ndarray) -> np.ndarray:
    """cv2.blur(img, (3, 3)) for uint8: BORDER_REFLECT_101, (sum + 4) // 9 ... rounding as cv2 (saturate_cast of sum/9)."""
    p = np.pad(img.astype(np.int64), 1, mode="reflect")
    s = sum(p[dy : dy + img.shape[0], dx : dx + img.shape[1]] for dy in range(3) for dx in range(3))
    return np.rint(s / 9.0).astype(np.uint8)
