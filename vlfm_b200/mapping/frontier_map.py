"""Per-frontier cached ITM cosine (reference: vlfm/mapping/frontier_map.py:10-77).

Pure host bookkeeping around ONE ``encoder.cosine(image, text)`` call per update that
introduces a new frontier; the encoder is the in-process GPU BLIP-2 (no HTTP client).
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import numpy as np


class Frontier:
    def __init__(self, xyz: np.ndarray, cosine: float):
        self.xyz = xyz
        self.cosine = cosine


class FrontierMap:
    frontiers: List[Frontier] = []

    def __init__(self, encoding_type: str = "cosine", encoder: Optional[Any] = None):
        if encoder is None:
            from ..vlm.blip2itm import BLIP2ITMClient

            encoder = BLIP2ITMClient()
        self.encoder = encoder
        self.frontiers = []

    def reset(self) -> None:
        self.frontiers = []

    def update(self, frontier_locations: List[np.ndarray], curr_image: np.ndarray, text: str) -> None:
        """frontier_map.py:25-52: drop vanished frontiers, tag new ones with the current
        frame's cosine (computed at most once)."""
        def known(loc, pool):
            return any(np.array_equal(loc, other) for other in pool)

        self.frontiers = [f for f in self.frontiers if known(f.xyz, frontier_locations)]
        score = None
        for loc in frontier_locations:
            if not known(loc, [f.xyz for f in self.frontiers]):
                if score is None:
                    score = self._encode(curr_image, text)
                self.frontiers.append(Frontier(loc, score))

    def _encode(self, image: np.ndarray, text: str) -> float:
        return self.encoder.cosine(image, text)

    def sort_waypoints(self) -> Tuple[np.ndarray, List[float]]:
        """frontier_map.py:66-77: descending by cached cosine."""
        scores = [f.cosine for f in self.frontiers]
        order = np.argsort([-c for c in scores])
        return np.array([self.frontiers[i].xyz for i in order]), [scores[i] for i in order]
