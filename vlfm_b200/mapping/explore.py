"""Host side of the explore half (vlfm/mapping/obstacle_map.py:114-169): scalar pose parameters exactly as the
reference derives them, one C-ABI call (csrc/explore.cu), lazily fetched frontier list."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib

MAX_FRONTIERS = 4096


def wrap_heading(h: float) -> float:
    return (h + np.pi) % (2 * np.pi) - np.pi


class ExploreEngine:
    def __init__(self, omap) -> None:
        self.m = omap
        self.lib = omap.lib
        n = ctypes.c_size_t(0)
        _lib.check(self.lib.vlfm_explore_workspace_bytes(omap.size, ctypes.byref(n)), "vlfm_explore_workspace_bytes")
        dev = omap.device
        self.ws = torch.zeros((n.value + 3) // 4, dtype=torch.int32, device=dev)
        self.frontiers = torch.zeros((MAX_FRONTIERS, 2), dtype=torch.float64, device=dev)
        self.count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)

    def update(self, tf: np.ndarray, max_depth: float, topdown_fov: float, nav_half: int) -> None:
        m = self.m
        agent = m._xy_to_px(np.asarray(tf[:2, 3], dtype=np.float64).reshape(1, 2))[0]      # (col, row), obstacle_map.py:115-116
        yaw = float(np.arctan2(tf[1, 0], tf[0, 0]))
        heading = float(np.rad2deg(wrap_heading(yaw + np.pi / 2)))                          # current_angle = -yaw (:121)
        rc = self.lib.vlfm_explore_update(m.size, m._explored.data_ptr(), m._nav.data_ptr(), int(agent[0]), int(agent[1]), heading,
                                          float(np.rad2deg(topdown_fov)), float(max_depth * m.pixels_per_meter),
                                          float(m._area_thresh_in_pixels), int(nav_half), self.frontiers.data_ptr(),
                                          self.count.data_ptr(), self.ws.data_ptr(), self.status.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "vlfm_explore_update")

    def fetch_frontiers_px(self) -> np.ndarray:
        n = int(self.count.item())
        if int(self.status.item()) != 0:
            raise _lib.VlfmError("explore: a device scratch buffer overflowed (too many contours / points)")
        if n == 0:
            return np.array([])
        return self.frontiers[:n].cpu().numpy()
