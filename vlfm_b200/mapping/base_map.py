"""Grid conventions shared by the maps (reference: vlfm/mapping/base_map.py:10-60).

The episodic frame is metres with +x forward / +y left; a map is a G x G grid whose cell (row, col) = (G//2, G//2) is the
episode origin, rows growing with +x and columns growing with -y.  The reference exposes the conversion in cv2 (col, row)
order; the formulae below are its arithmetic written out per axis:

    col = G - (rint(y * ppm) + G//2)        row = rint(x * ppm) + G//2
    x   = (row - G//2) / ppm                y   = ((G - col) - G//2) / ppm
"""
from __future__ import annotations

from typing import Any, List

import numpy as np


class BaseMap:
    """Host-side bookkeeping of a map: geometry (size, pixels_per_meter, origin), the metre <-> cell conversions and the
    camera trajectory used by ``visualize``.  The grids themselves live on the GPU in the subclasses."""

    _camera_positions: List[np.ndarray] = []
    _last_camera_yaw: float = 0.0

    def __init__(self, size: int = 1000, pixels_per_meter: int = 20, *args: Any, **kwargs: Any):
        self.size = int(size)
        self.pixels_per_meter = pixels_per_meter
        half = self.size // 2
        self._episode_pixel_origin = np.array([half, half])
        self._camera_positions = []

    # ---- episode bookkeeping (base_map.py:26-33)
    def reset(self) -> None:
        self._camera_positions = []

    def update_agent_traj(self, robot_xy: np.ndarray, robot_heading: float) -> None:
        self._last_camera_yaw = robot_heading
        self._camera_positions.append(robot_xy)

    # ---- conversions (base_map.py:35-60); same floating-point operations per element as the reference
    def _xy_to_px(self, points: np.ndarray) -> np.ndarray:
        """[(x, y)] metres -> [(col, row)] integer cells (np.rint: half to even)."""
        pts = np.asarray(points)
        o_col, o_row = self._episode_pixel_origin
        col = self.size - (np.rint(pts[:, 1] * self.pixels_per_meter) + o_col)
        row = np.rint(pts[:, 0] * self.pixels_per_meter) + o_row
        return np.stack((col, row), axis=1).astype(int)

    def _px_to_xy(self, px: np.ndarray) -> np.ndarray:
        """[(col, row)] cells (integer or fractional, e.g. frontier midpoints) -> [(x, y)] metres."""
        cells = np.asarray(px)
        o_col, o_row = self._episode_pixel_origin
        x = (cells[:, 1] - o_row) / self.pixels_per_meter
        y = ((self.size - cells[:, 0]) - o_col) / self.pixels_per_meter
        return np.stack((x, y), axis=1)
