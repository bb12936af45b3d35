"""The S frame: the grid rectangle in which the explore step's whole-grid operations are carried out (csrc/explore.cu).

The reference runs ``explored[nav == 0] = 0``, ``cv2.findContours`` (component selection) and ``detect_frontier_waypoints`` on
the whole G x G grid every step (vlfm/mapping/obstacle_map.py:127-169).  Everything those operations can see lies inside the
COVER rectangle of the episode -- the union of the obstacle updates' dilation windows and of the fog-of-war windows:

  * explored cells are only ever set inside a fog-of-war window;
  * obstacle cells (hence non-navigable cells) only appear inside an obstacle-update window; outside the cover the navigable map
    is uniformly 1 once the first obstacle update has run (uniformly 0 before it);

so the unexplored mask ``nav & ~dilate5(explored)`` is uniformly 1 outside cover + 2 cells.  Restricting the operations to
S = cover + margin is exact provided (tests/test_oracle_sframe.py checks all of this against the whole-grid oracle):

  1. the margin ring is wide enough that no contour of the explored / grown masks touches a non-grid edge of S and the 3x3 blur
     never reflects there (margin 8 >= 2 + 1 + slack);
  2. an unexplored component that touches a non-grid edge of S is treated as the exterior: never absorbed by the small-pocket rule
     (F1).  Its true contourArea is at least (D - 1) * (G - 1) >= area_thresh, because every non-grid side of S is kept at least
     D cells away from the grid edge (a side closer than that is snapped onto the edge);
  3. frontier coordinates are produced in grid coordinates (S origin added before the interpolation arithmetic).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

Rect = Tuple[int, int, int, int]      # (col0, row0, col1, row1), half-open
S_MARGIN = 8


def cover_add(cover: Optional[Rect], r: Rect, g: int) -> Optional[Rect]:
    r = (max(r[0], 0), max(r[1], 0), min(r[2], g), min(r[3], g))
    if r[2] <= r[0] or r[3] <= r[1]:
        return cover
    if cover is None:
        return r
    return (min(cover[0], r[0]), min(cover[1], r[1]), max(cover[2], r[2]), max(cover[3], r[3]))


def obstacle_window(col: int, row: int, half: int, g: int) -> Rect:
    """cells an obstacle update at camera cell (col, row) can change (scatter reach + dilation radius).  Near the LOW edges numpy's
    negative indices wrap around (base_map.py:44-46 + fancy indexing): anything may change."""
    r = (col - half, row - half, col + half + 1, row + half + 1)
    return (0, 0, g, g) if (r[0] < 0 or r[1] < 0) else r


def fog_window(col: int, row: int, line_len: int) -> Rect:
    return (col - line_len - 4, row - line_len - 4, col + line_len + 5, row + line_len + 5)


def snap_distance(area_thresh_px: float, g: int) -> int:
    return int(math.ceil(area_thresh_px / max(g - 1, 1))) + 2


def sframe(cover: Optional[Rect], g: int, area_thresh_px: float) -> Rect:
    if cover is None:
        return (0, 0, g, g)
    d = snap_distance(area_thresh_px, g)
    x0, y0, x1, y1 = cover[0] - S_MARGIN, cover[1] - S_MARGIN, cover[2] + S_MARGIN, cover[3] + S_MARGIN
    return (0 if x0 < d else x0, 0 if y0 < d else y0, g if x1 > g - d else x1, g if y1 > g - d else y1)
