"""Detection container (reference: vlfm/vlm/detections.py:15-126)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


def _cxcywh_to_xyxy(b: torch.Tensor) -> torch.Tensor:
    cx, cy, w, h = b.unbind(-1)
    return torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), dim=-1)


class ObjectDetections:
    """boxes (N,4) normalised xyxy, logits (N,), phrases list[str]."""

    def __init__(self, boxes: torch.Tensor, logits: torch.Tensor, phrases: List[str], image_source: Optional[np.ndarray],
                 fmt: str = "cxcywh"):
        self.image_source = image_source
        self.boxes = boxes if fmt == "xyxy" else _cxcywh_to_xyxy(boxes)  # detections.py:29-32 (torchvision box_convert)
        self.logits = logits
        self.phrases = phrases
        self._annotated_frame: Optional[np.ndarray] = None

    @property
    def num_detections(self) -> int:
        return len(self.phrases)

    @property
    def annotated_frame(self) -> Optional[np.ndarray]:
        if self._annotated_frame is None and self.image_source is not None:
            import cv2

            img = self.image_source.copy()
            h, w = img.shape[:2]
            for box, logit, phrase in zip(self.boxes, self.logits, self.phrases):
                x0, y0, x1, y1 = (box * torch.tensor([w, h, w, h])).int().tolist()
                cv2.rectangle(img, (x0, y0), (x1, y1), (255, 0, 0), 2)
                cv2.putText(img, f"{phrase} {float(logit):.2f}", (x0, max(y0 - 4, 10)), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 0, 0), 1)
            self._annotated_frame = img
        return self._annotated_frame

    def __repr__(self) -> str:
        rows = [f"{p} ({float(l):.2f}): {b.tolist()}" for b, l, p in zip(self.boxes, self.logits, self.phrases)]
        return "\n".join(rows) if rows else "No detections"

    def filter_by_conf(self, conf_thresh: float) -> None:  # detections.py:64-71
        self._filter(torch.ge(self.logits, conf_thresh))

    def filter_by_class(self, classes: List[str]) -> None:  # detections.py:73-80
        self._filter(torch.tensor([p in classes for p in self.phrases], dtype=torch.bool))

    def _filter(self, keep: torch.Tensor) -> None:
        if keep.all():
            return
        self.boxes = self.boxes[keep]
        self.logits = self.logits[keep]
        self.phrases = [p for i, p in enumerate(self.phrases) if keep[i]]
        self._annotated_frame = None

    def to_json(self) -> dict:
        return {"boxes": self.boxes.tolist(), "logits": self.logits.tolist(), "phrases": self.phrases}

    @classmethod
    def from_json(cls, d: dict, image_source: Optional[np.ndarray] = None) -> "ObjectDetections":
        return cls(image_source=image_source, boxes=torch.tensor(d["boxes"]), logits=torch.tensor(d["logits"]),
                   phrases=d["phrases"], fmt="xyxy")
