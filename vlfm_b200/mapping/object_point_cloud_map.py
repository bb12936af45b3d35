"""GPU object point-cloud map behind the reference's ``ObjectPointCloudMap`` surface.

Reference: vlfm/mapping/object_point_cloud_map.py (class :16, update_map :32-77, get_best_object :79-102,
update_explored :104-132, get_target_cloud :134-141, _extract_object_cloud :143-163, _get_closest_point :165-189).

Device work (csrc/object_cloud.cu): mask erosion, masked unprojection in ``np.where`` order and the DBSCAN largest-cluster filter
(the reference's only native dependency on this path, Open3D ``cluster_dbscan``).  Host work: what the reference does with
numpy on the <= 5000 surviving points -- rigid transform, range ids, closest point, cone test -- kept in numpy with the same
calls, because the clouds are host state the policy reads (``self.clouds[name]`` numpy arrays) and because the reference's
randomness is numpy's GLOBAL generator (``np.random.rand`` / ``np.random.choice``), which this class draws from in the same
order so that a seeded run is reproducible against the reference.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Union

import numpy as np
import torch

from .. import _lib

MAX_POINTS = 5000          # get_random_subarray(cloud, 5000) (:160)
DBSCAN_EPS, DBSCAN_MIN_POINTS = 0.2, 100   # open3d_dbscan_filtering defaults (:192)


def too_offset(mask: np.ndarray) -> bool:
    """object_point_cloud_map.py:269-298 (bounding rectangle of the mask entirely in an outer third and near the image edge)"""
    ys, xs = np.nonzero(mask)
    if len(xs) == 0:
        x = w = 0
    else:
        x, w = int(xs.min()), int(xs.max()) - int(xs.min()) + 1
    third = mask.shape[1] // 3
    if x + w <= third:
        return x <= int(0.05 * mask.shape[1])
    if x >= 2 * third:
        return x + w >= int(0.95 * mask.shape[1])
    return False


def _transform_points(tf: np.ndarray, pts: np.ndarray) -> np.ndarray:      # geometry_utils.py:205-213
    hom = np.hstack((pts, np.ones((pts.shape[0], 1))))
    out = np.dot(tf, hom.T).T
    return out[:, :3] / out[:, 3:]


def _within_fov_cone(origin: np.ndarray, angle: float, fov: float, rng: float, points: np.ndarray) -> np.ndarray:   # geometry_utils.py:91-116
    d = points[:, :3] - origin
    dist = np.linalg.norm(d, axis=1)
    ang = np.arctan2(d[:, 1], d[:, 0])
    diff = np.mod(ang - angle + np.pi, 2 * np.pi) - np.pi
    return points[np.logical_and(dist <= rng, np.abs(diff) <= fov / 2)]


class ObjectPointCloudMap:
    clouds: Dict[str, np.ndarray] = {}
    use_dbscan: bool = True

    def __init__(self, erosion_size: float, device: Union[str, torch.device, None] = None) -> None:
        if not torch.cuda.is_available():
            raise _lib.VlfmError("vlfm_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        self._erosion_size = erosion_size
        self.last_target_coord: Union[np.ndarray, None] = None
        self.clouds = {}
        self._shape = None

    def reset(self) -> None:
        self.clouds = {}
        self.last_target_coord = None

    def has_object(self, target_class: str) -> bool:
        return target_class in self.clouds and len(self.clouds[target_class]) > 0

    # ------------------------------------------------------------------ device part ----
    def _buffers(self, h: int, w: int) -> None:
        if self._shape == (h, w):
            return
        dev = self.device
        self._pin_d = torch.empty((h, w), dtype=torch.float32).pin_memory()
        self._pin_m = torch.empty((h, w), dtype=torch.uint8).pin_memory()
        self._dev_d = torch.empty((h, w), dtype=torch.float32, device=dev)
        self._dev_m = torch.empty((h, w), dtype=torch.uint8, device=dev)
        self._pts = torch.empty((h * w, 3), dtype=torch.float64, device=dev)
        self._count = torch.zeros(2, dtype=torch.int32, device=dev)
        self._scratch = torch.empty(((h * w + 255) // 256 * 256 + 8 * h + 256 + 3) // 4, dtype=torch.int32, device=dev)
        n = ctypes.c_size_t(0)
        _lib.check(self.lib.vlfm_dbscan_workspace_bytes(MAX_POINTS, ctypes.byref(n)), "vlfm_dbscan_workspace_bytes")
        self._db_ws = torch.empty((n.value + 3) // 4, dtype=torch.int32, device=dev)
        self._gather = torch.empty(MAX_POINTS, dtype=torch.int32, device=dev)
        self._gathered = torch.empty((MAX_POINTS, 3), dtype=torch.float64, device=dev)
        self._out = torch.empty((MAX_POINTS, 3), dtype=torch.float64, device=dev)
        self._shape = (h, w)

    def _extract_object_cloud(self, depth: np.ndarray, object_mask: np.ndarray, min_depth: float, max_depth: float, fx: float,
                              fy: float) -> np.ndarray:
        """:143-163 -- erode, unproject (np.where order), random subsample to 5000, DBSCAN largest cluster; [N,3] float64."""
        if depth.ndim == 3:
            depth = depth.squeeze(2)
        h, w = depth.shape
        self._buffers(h, w)
        st = _lib.stream_ptr()
        with torch.cuda.device(self.device):
            self._pin_d.numpy()[...] = depth
            self._pin_m.numpy()[...] = object_mask
            self._dev_d.copy_(self._pin_d, non_blocking=True)
            self._dev_m.copy_(self._pin_m, non_blocking=True)
            rc = self.lib.vlfm_object_cloud_extract(_lib.ptr(self._dev_d), _lib.ptr(self._dev_m), h, w, int(self._erosion_size),
                                                    float(np.float32(max_depth - min_depth)), float(np.float32(min_depth)), float(fx), float(fy),
                                                    _lib.ptr(self._pts), h * w, _lib.ptr(self._count), _lib.ptr(self._scratch),
                                                    self._scratch.numel() * 4, st)
            _lib.check(rc, "vlfm_object_cloud_extract")
            n = int(self._count[0].item())                # the subsample draws from numpy's generator on the host: the count is needed here
            if n == 0:
                return np.zeros((0, 3))
            gather = None
            if n > MAX_POINTS:                            # get_random_subarray (:246-266)
                idx = np.random.choice(n, MAX_POINTS, replace=False)
                self._gather.copy_(torch.from_numpy(idx.astype(np.int32)))
                gather, n = self._gather, MAX_POINTS
            if not self.use_dbscan:
                pts = self._pts[:n] if gather is None else self._pts[gather.long()]
                return pts.cpu().numpy()
            rc = self.lib.vlfm_dbscan_largest_cluster(_lib.ptr(self._pts), _lib.ptr(gather), n, DBSCAN_EPS, DBSCAN_MIN_POINTS,
                                                      _lib.ptr(self._gathered), _lib.ptr(self._out), _lib.ptr(self._count[1:]),
                                                      _lib.ptr(self._db_ws), self._db_ws.numel() * 4, st)
            _lib.check(rc, "vlfm_dbscan_largest_cluster")
            m = int(self._count[1].item())
            return self._out[:m].cpu().numpy() if m else np.array([])

    # -------------------------------------------------------------------- host part ----
    def update_map(self, object_name: str, depth_img: np.ndarray, object_mask: np.ndarray, tf_camera_to_episodic: np.ndarray,
                   min_depth: float, max_depth: float, fx: float, fy: float) -> None:
        """:32-77"""
        local_cloud = self._extract_object_cloud(depth_img, object_mask, min_depth, max_depth, fx, fy)
        if len(local_cloud) == 0:
            return
        if too_offset(object_mask):
            within_range = np.ones_like(local_cloud[:, 0]) * np.random.rand()
        else:
            within_range = (local_cloud[:, 0] <= max_depth * 0.95) * 1.0  # 5% margin
            within_range = within_range.astype(np.float32)
            within_range[within_range == 0] = np.random.rand()
        global_cloud = _transform_points(tf_camera_to_episodic, local_cloud)
        global_cloud = np.concatenate((global_cloud, within_range[:, None]), axis=1)
        curr_position = tf_camera_to_episodic[:3, 3]
        closest_point = self._get_closest_point(global_cloud, curr_position)
        if np.linalg.norm(closest_point[:3] - curr_position) < 1.0:
            return   # too close to trust
        if object_name in self.clouds:
            self.clouds[object_name] = np.concatenate((self.clouds[object_name], global_cloud), axis=0)
        else:
            self.clouds[object_name] = global_cloud

    def get_best_object(self, target_class: str, curr_position: np.ndarray) -> np.ndarray:
        """:79-102"""
        closest_point_2d = self._get_closest_point(self.get_target_cloud(target_class), curr_position)[:2]
        if self.last_target_coord is None:
            self.last_target_coord = closest_point_2d
        else:
            delta_dist = np.linalg.norm(closest_point_2d - self.last_target_coord)
            if delta_dist < 0.1:
                return self.last_target_coord
            if delta_dist < 0.5 and np.linalg.norm(curr_position - closest_point_2d) > 2.0:
                return self.last_target_coord
            self.last_target_coord = closest_point_2d
        return self.last_target_coord

    def update_explored(self, tf_camera_to_episodic: np.ndarray, max_depth: float, cone_fov: float) -> None:
        """:104-132 -- drop detections that were out of range when seen and are now inside the near half of the view cone."""
        camera_coordinates = tf_camera_to_episodic[:3, 3]
        camera_yaw = float(np.arctan2(tf_camera_to_episodic[1, 0], tf_camera_to_episodic[0, 0]))
        for obj in self.clouds:
            within_range = _within_fov_cone(camera_coordinates, camera_yaw, cone_fov, max_depth * 0.5, self.clouds[obj])
            for range_id in set(within_range[..., -1].tolist()):
                if range_id == 1:
                    continue
                self.clouds[obj] = self.clouds[obj][self.clouds[obj][..., -1] != range_id]

    def get_target_cloud(self, target_class: str) -> np.ndarray:
        """:134-141"""
        target_cloud = self.clouds[target_class].copy()
        if np.any(target_cloud[:, -1] == 1):
            target_cloud = target_cloud[target_cloud[:, -1] == 1]
        return target_cloud

    def _get_closest_point(self, cloud: np.ndarray, curr_position: np.ndarray) -> np.ndarray:
        """:165-189"""
        ndim = curr_position.shape[0]
        if self.use_dbscan:
            return cloud[np.argmin(np.linalg.norm(cloud[:, :ndim] - curr_position, axis=1))]
        ref_point = np.concatenate((curr_position, np.array([0.5]))) if ndim == 2 else curr_position
        sorted_indices = np.argsort(np.linalg.norm(cloud[:, :3] - ref_point, axis=1))
        top_percent = sorted_indices[: int(0.25 * len(cloud))]
        try:
            median_index = top_percent[int(len(top_percent) / 2)]
        except IndexError:
            median_index = 0
        return cloud[median_index]
