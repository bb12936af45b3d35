"""Seeded synthetic RGB-D + pose generator (numpy only).

Shared by tests, bench.py and the golden-fixture script so that every consumer sees
bit-identical inputs for a given (seed, step).  Shapes and camera constants follow the
reference's Habitat configuration: 640x480, hfov 79 deg,
min/max depth 0.5/5.0 m, camera height 0.88 m, 0.25 m / 30 deg action set.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

HFOV_DEG = 79.0
MIN_DEPTH = 0.5
MAX_DEPTH = 5.0
CAMERA_HEIGHT = 0.88


@dataclass
class Frame:
    depth: np.ndarray  # (H, W) float32 in [0, 1]
    rgb: np.ndarray  # (H, W, 3) uint8
    tf: np.ndarray  # (4, 4) float64 camera -> episodic
    xy: Tuple[float, float]
    yaw: float


def focal_from_hfov(width: int, hfov_deg: float = HFOV_DEG) -> float:
    """vlfm/policy/habitat_policies.py:89-91: fx = fy = W / (2 tan(hfov/2))."""
    return width / (2.0 * math.tan(math.radians(hfov_deg) / 2.0))


def tf_from_pose(x: float, y: float, z: float, yaw: float) -> np.ndarray:
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, -s, 0.0, x], [s, c, 0.0, y], [0.0, 0.0, 1.0, z], [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)


def make_depth(rng: np.random.Generator, h: int, w: int, holes: bool = True) -> np.ndarray:
    """Piecewise-smooth 'room': per-column wall distance with a few jumps, a floor ramp
    below the horizon, optional zero-valued hole blobs.  float32 in [0, 1]."""
    u = np.linspace(0.0, 1.0, w)
    wall = 0.55 + 0.25 * np.sin(2 * np.pi * (rng.uniform(0.5, 2.0) * u + rng.uniform()))
    for _ in range(int(rng.integers(1, 5))):
        a, b = sorted(rng.uniform(0, 1, 2))
        wall[(u >= a) & (u <= b)] += rng.uniform(-0.3, 0.3)
    wall = np.clip(wall + rng.normal(0, 0.004, w), 0.02, 1.0)
    v = np.arange(h, dtype=np.float64)[:, None]
    horizon = h * 0.5
    with np.errstate(divide="ignore"):
        floor = np.where(v > horizon, (0.18 * h) / np.maximum(v - horizon, 1e-6) / 5.0, np.inf)
    img = np.minimum(wall[None, :], floor)
    img = np.clip(img, 0.0, 1.0).astype(np.float32)
    if holes:
        for _ in range(int(rng.integers(2, 7))):
            cy, cx = int(rng.integers(0, h)), int(rng.integers(0, w))
            ry, rx = int(rng.integers(2, 10)), int(rng.integers(2, 14))
            img[max(cy - ry, 0) : cy + ry, max(cx - rx, 0) : cx + rx] = 0.0
    return img


def make_rgb(rng: np.random.Generator, h: int, w: int) -> np.ndarray:
    """Value-noise texture, uint8."""
    coarse = rng.integers(0, 256, (h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
    up = np.kron(coarse, np.ones((16, 16, 1), np.float32))[:h, :w]
    fine = rng.integers(-20, 21, (h, w, 3)).astype(np.float32)
    return np.clip(up + fine, 0, 255).astype(np.uint8)


def trajectory(seed: int, steps: int, h: int = 480, w: int = 640, holes: bool = True, bound_m: float = 15.0,
               with_rgb: bool = False, start_xy: Tuple[float, float] = (0.0, 0.0)) -> List[Frame]:
    """Random walk with the Habitat action set; stays within +-bound_m of the origin."""
    rng = np.random.default_rng(1234 + seed)
    x, y = start_xy
    yaw = float(rng.uniform(-math.pi, math.pi))
    frames: List[Frame] = []
    for _ in range(steps):
        act = int(rng.integers(0, 3))
        if act == 0:
            nx, ny = x + 0.25 * math.cos(yaw), y + 0.25 * math.sin(yaw)
            if abs(nx - start_xy[0]) < bound_m and abs(ny - start_xy[1]) < bound_m:
                x, y = nx, ny
        elif act == 1:
            yaw += math.radians(30.0)
        else:
            yaw -= math.radians(30.0)
        yaw = (yaw + math.pi) % (2 * math.pi) - math.pi
        depth = make_depth(rng, h, w, holes)
        rgb = make_rgb(rng, h, w) if with_rgb else np.zeros((0, 0, 3), np.uint8)
        frames.append(Frame(depth, rgb, tf_from_pose(x, y, CAMERA_HEIGHT, yaw), (x, y), yaw))
    return frames


def make_object_mask(rng: np.random.Generator, h: int, w: int, side: str = "any") -> np.ndarray:
    """uint8 0/1 mask of a blob-shaped detection (what MobileSAM hands to ObjectPointCloudMap.update_map,
    base_objectnav_policy.py:311-346): an ellipse with a ragged rim plus a few stray specks (removed by the erosion).
    ``side``: "left" / "right" put the whole blob in an outer third of the image (the too_offset case)."""
    cy = int(rng.integers(int(0.3 * h), int(0.8 * h)))
    ry, rx = int(rng.integers(max(6, h // 12), max(8, h // 4))), int(rng.integers(max(6, w // 16), max(8, w // 6)))
    if side == "left":
        rx = min(rx, w // 8); cx = int(rng.integers(rx // 2, max(rx // 2 + 1, w // 3 - rx - 1)))
    elif side == "right":
        rx = min(rx, w // 8); cx = int(rng.integers(2 * (w // 3) + rx + 1, w - 1))
    else:
        cx = int(rng.integers(int(0.3 * w), int(0.7 * w)))
    yy, xx = np.mgrid[0:h, 0:w]
    ang = np.arctan2(yy - cy, xx - cx)
    rim = 1.0 + 0.12 * np.sin(5 * ang + rng.uniform(0, 6)) + 0.06 * np.sin(11 * ang + rng.uniform(0, 6))
    m = (((xx - cx) / (rx * rim)) ** 2 + ((yy - cy) / (ry * rim)) ** 2 <= 1.0)
    for _ in range(int(rng.integers(2, 8))):
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        m[y:y + 2, x:x + 2] = True
    return m.astype(np.uint8)
