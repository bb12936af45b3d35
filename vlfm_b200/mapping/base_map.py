"""Grid conventions shared by the maps (reference: vlfm/mapping/base_map.py:10-60)."""
from __future__ import annotations

from typing import Any, List

import numpy as np


class BaseMap:
    """Host-side mirror of the reference ``BaseMap``: size, pixels_per_meter, origin and
    the metre <-> cell conversions.  The grids themselves live on the GPU in subclasses."""

    _camera_positions: List[np.ndarray] = []
    _last_camera_yaw: float = 0.0

    def __init__(self, size: int = 1000, pixels_per_meter: int = 20, *args: Any, **kwargs: Any):
        self.pixels_per_meter = pixels_per_meter
        self.size = size
        self._episode_pixel_origin = np.array([size // 2, size // 2])
        self._camera_positions = []

    def reset(self) -> None:  # base_map.py:26-29
        self._camera_positions = []

    def update_agent_traj(self, robot_xy: np.ndarray, robot_heading: float) -> None:  # :31-33
        self._camera_positions.append(robot_xy)
        self._last_camera_yaw = robot_heading

    def _xy_to_px(self, points: np.ndarray) -> np.ndarray:
        """(x, y) metres -> (col, row) cells; np.rint, y flipped (base_map.py:35-46)."""
        px = np.rint(points[:, ::-1] * self.pixels_per_meter) + self._episode_pixel_origin
        px[:, 0] = self.size - px[:, 0]
        return px.astype(int)

    def _px_to_xy(self, px: np.ndarray) -> np.ndarray:
        """(col, row) cells -> (x, y) metres (base_map.py:48-60)."""
        q = px.copy()
        q[:, 0] = self.size - q[:, 0]
        return ((q - self._episode_pixel_origin) / self.pixels_per_meter)[:, ::-1]
