"""GPU obstacle / explored-area map behind the reference's ``ObstacleMap`` surface.

Reference: vlfm/mapping/obstacle_map.py (class :15, update_map :55, reset :48).
Grids are uint8 [1,G,G] tensors in HBM: ``obst`` (ObstacleMap._map), ``nav``
(_navigable_map as 0/1) and ``explored`` (explored_area).
"""
from __future__ import annotations

import ctypes
import math
from typing import Any, Optional, Union

import numpy as np
import torch

from .. import _lib
from .base_map import BaseMap


class ObstacleMap(BaseMap):
    radius_padding_color: tuple = (100, 100, 100)

    def __init__(self, min_height: float, max_height: float, agent_radius: float, area_thresh: float = 3.0,
                 hole_area_thresh: int = 100000, size: int = 1000, pixels_per_meter: int = 20,
                 device: Union[str, torch.device, None] = None):
        super().__init__(size, pixels_per_meter)
        if not torch.cuda.is_available():
            raise _lib.VlfmError("vlfm_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        self._obst = torch.zeros((1, size, size), dtype=torch.uint8, device=self.device)
        self._nav = torch.zeros((1, size, size), dtype=torch.uint8, device=self.device)
        self._explored = torch.zeros((1, size, size), dtype=torch.uint8, device=self.device)
        self._status = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._min_height = min_height
        self._max_height = max_height
        self._area_thresh_in_pixels = area_thresh * (self.pixels_per_meter**2)  # obstacle_map.py:41
        self._hole_area_thresh = hole_area_thresh
        kernel_size = self.pixels_per_meter * agent_radius * 2  # :43-46
        self._kernel = int(kernel_size) + (int(kernel_size) % 2 == 0)
        self._nav_valid = False  # before the first obstacle update the reference's navigable map is all 0
        # cells whose navigable value may have changed since the last explore step, as a grid rectangle (col0, row0, col1, row1):
        # the union of the obstacle updates' dilation windows.  The reference clears explored cells on non-navigable cells over
        # the WHOLE grid (:127); outside this rectangle nothing changed, so masking it is identical (several update_map(explore=
        # False) calls with different cameras followed by one explore call: reality_policies.py:114-138)
        self._dirty: Optional[tuple] = None
        self._front_cache: Optional[np.ndarray] = np.array([])
        self._pin: Optional[torch.Tensor] = None
        self._pin_tf: Optional[torch.Tensor] = None
        self._dev_depth: Optional[torch.Tensor] = None
        self._dev_tf = torch.empty((16,), dtype=torch.float64, device=self.device)
        self._ev: Optional[torch.cuda.Event] = None
        self._explore_impl = None
        self._fill = None

    def _check_fill(self) -> None:
        # fill_small_holes reports scratch exhaustion (> 65536 borders or > 1M border points in one depth image) through a
        # sticky device flag; it is read wherever the host synchronises anyway
        if self._fill is not None and int(self._fill_status.item()) != 0:
            self._fill_status.zero_()
            raise _lib.VlfmError("fill_small_holes: a device scratch buffer overflowed (too many contours in the depth image)")

    # ---- numpy views
    @property
    def _map(self) -> np.ndarray:
        self._check_fill()
        return self._obst[0].cpu().numpy().astype(bool)

    @property
    def _navigable_map(self) -> np.ndarray:
        self._check_fill()
        return self._nav[0].cpu().numpy().astype(np.int64)  # the reference's is int64 0/1 (:105-109)

    @property
    def explored_area(self) -> np.ndarray:
        return self._explored[0].cpu().numpy().astype(bool)

    # frontier waypoints are produced on the device; the host copy is fetched (and the stream synchronised) on access
    @property
    def _frontiers_px(self) -> np.ndarray:
        if self._front_cache is None:
            self._check_fill()
            self._front_cache = self._explore_impl.fetch_frontiers_px()
        return self._front_cache

    @property
    def frontiers(self) -> np.ndarray:
        px = self._frontiers_px
        return np.array([]) if len(px) == 0 else self._px_to_xy(px)      # obstacle_map.py:149-153

    def explored_device(self) -> torch.Tensor:
        return self._explored

    def reset(self) -> None:  # obstacle_map.py:48-53
        super().reset()
        self._obst.zero_(); self._nav.zero_(); self._explored.zero_(); self._status.zero_()
        self._nav_valid = False
        self._dirty = None
        self._front_cache = np.array([])

    def _mark_dirty(self, tf: np.ndarray, half: int) -> None:
        g = self.size
        if not self._nav_valid:                       # first update: the navigable map changes everywhere (0 -> 1)
            self._dirty = (0, 0, g, g)
            return
        col, row = (int(v) for v in self._xy_to_px(np.asarray(tf[:2, 3], dtype=np.float64).reshape(1, 2))[0])
        r = (col - half, row - half, col + half + 1, row + half + 1)
        if r[0] < 0 or r[1] < 0 or r[2] > g or r[3] > g:   # near the edge the scatter may wrap (numpy negative indices): whole grid
            r = (0, 0, g, g)
        d = self._dirty
        self._dirty = r if d is None else (min(d[0], r[0]), min(d[1], r[1]), max(d[2], r[2]), max(d[3], r[3]))

    def _upload(self, depth: Optional[np.ndarray], tf: np.ndarray) -> None:
        if self._pin_tf is None:
            self._pin_tf = torch.empty((16,), dtype=torch.float64).pin_memory()
            self._ev = torch.cuda.Event()
        else:
            self._ev.synchronize()
        self._pin_tf.numpy()[:] = np.asarray(tf, dtype=np.float64).reshape(16)
        if depth is not None:
            h, w = depth.shape
            if self._pin is None or self._pin.shape != (1, h, w):
                self._pin = torch.empty((1, h, w), dtype=torch.float32).pin_memory()
                self._dev_depth = torch.empty((1, h, w), dtype=torch.float32, device=self.device)
            direct = None
            if isinstance(depth, np.ndarray) and depth.dtype == np.float32 and depth.flags.c_contiguous:
                t = torch.from_numpy(depth)
                # page-locked caller frame and an idle stream: DMA straight from it and wait (the caller may reuse the frame);
                # with work queued ahead the wait would stall the host, so the frame is staged instead
                if t.is_pinned() and torch.cuda.current_stream(self.device).query():
                    direct = t
            if direct is None:
                self._pin[0].numpy()[...] = depth
            self._dev_depth.copy_(self._pin if direct is None else direct[None], non_blocking=True)
            self._dev_tf.copy_(self._pin_tf, non_blocking=True)
            self._ev.record()
            if direct is not None:
                self._ev.synchronize()
            return
        self._dev_tf.copy_(self._pin_tf, non_blocking=True)
        self._ev.record()

    def update_map(self, depth: Union[np.ndarray, Any], tf_camera_to_episodic: np.ndarray, min_depth: float,
                   max_depth: float, fx: float, fy: float, topdown_fov: float, explore: bool = True,
                   update_obstacles: bool = True) -> None:
        """obstacle_map.py:55-153."""
        with torch.cuda.device(self.device):
            if update_obstacles:
                if depth.ndim == 3:
                    depth = depth.squeeze(2)
                self._upload(depth, tf_camera_to_episodic)
                h, w = depth.shape
                half = int(math.ceil(max_depth * self.pixels_per_meter * math.sqrt(1.0 + (w / 2.0 / fx) ** 2))) + self._kernel // 2 + 2
                self._mark_dirty(tf_camera_to_episodic, half)
                p = _lib.ObstacleParams(h, w, self.size, self.pixels_per_meter,
                                        float(np.float32(max_depth - min_depth)), float(np.float32(min_depth)),
                                        float(np.float32(max_depth)), float(fx), float(fy),
                                        float(self._min_height), float(self._max_height), self._kernel,
                                        0 if self._nav_valid else 1, half)
                fill = None
                if self._hole_area_thresh != -1:          # fill_small_holes (img_utils.py:361-390) on the device
                    if self._fill is None or self._fill.shape != (1, h, w):
                        nb = ctypes.c_size_t(0)
                        _lib.check(self.lib.vlfm_holes_workspace_bytes(h, w, ctypes.byref(nb)), "vlfm_holes_workspace_bytes")
                        self._fill = torch.zeros((1, h, w), dtype=torch.uint8, device=self.device)
                        self._fill_ws = torch.zeros((nb.value + 3) // 4, dtype=torch.int32, device=self.device)
                        self._fill_status = torch.zeros(1, dtype=torch.int32, device=self.device)
                    rc = self.lib.vlfm_fill_small_holes(_lib.ptr(self._dev_depth), h, w, float(self._hole_area_thresh), _lib.ptr(self._fill),
                                                        _lib.ptr(self._fill_ws), _lib.ptr(self._fill_status), _lib.stream_ptr())
                    _lib.check(rc, "vlfm_fill_small_holes")
                    fill = self._fill
                rc = self.lib.vlfm_obstacle_update(ctypes.byref(p), 1, None, _lib.ptr(self._obst), _lib.ptr(self._nav),
                                                   _lib.ptr(self._dev_depth), _lib.ptr(self._dev_tf), _lib.ptr(fill),
                                                   _lib.ptr(self._status), _lib.stream_ptr())
                _lib.check(rc, "vlfm_obstacle_update")
                self._nav_valid = True
                # numpy raises IndexError synchronously (:101); the caller turns it into STOP
                # (base_objectnav_policy.py:157-162).  Only pay for the sync near the border.
                cx, cy = tf_camera_to_episodic[0, 3], tf_camera_to_episodic[1, 3]
                margin = (half + 2) / self.pixels_per_meter
                lim = self.size / 2 / self.pixels_per_meter
                if abs(cx) + margin >= lim or abs(cy) + margin >= lim:
                    if int(self._status[0].item()) & _lib.ST_SCATTER_OOB:
                        self._status.zero_()
                        raise IndexError("obstacle cell index out of bounds for the map")
            elif explore:
                self._upload(None, tf_camera_to_episodic)
            if not explore:
                return
            if self._explore_impl is None:
                from .explore import ExploreEngine

                self._explore_impl = ExploreEngine(self)
            nav_half = 0
            if self._dirty is not None:                # square around the explore pose that covers the dirty rectangle
                col, row = (int(v) for v in self._xy_to_px(np.asarray(tf_camera_to_episodic[:2, 3], dtype=np.float64).reshape(1, 2))[0])
                d = self._dirty
                nav_half = max(col - d[0], d[2] - 1 - col, row - d[1], d[3] - 1 - row, 0)
                self._dirty = None
            self._explore_impl.update(tf_camera_to_episodic, max_depth, topdown_fov, nav_half)
            self._front_cache = None

    def visualize(self) -> np.ndarray:
        """obstacle_map.py:171-193 (trajectory overlay omitted)."""
        import cv2

        vis = np.ones((self.size, self.size, 3), dtype=np.uint8) * 255
        vis[self.explored_area == 1] = (200, 255, 200)
        vis[self._navigable_map == 0] = self.radius_padding_color
        vis[self._map == 1] = (0, 0, 0)
        for f in self._frontiers_px:
            cv2.circle(vis, tuple([int(i) for i in f]), 5, (200, 0, 0), 2)
        return cv2.flip(vis, 0)
