"""fp32 CPU oracle for GroundingDINO.predict (reference: vlfm/vlm/grounding_dino.py:38-74).

TEST INFRASTRUCTURE.  PARITY UNPINNED w.r.t. GroundingDINO@eeba084 (absent offline, no
checkpoint, no reference test): the oracle is the architecture-equivalent HF
``GroundingDinoForObjectDetection`` (Swin-T, 900 queries) in float32 eager mode, fed by the
reference's own preprocessing (to_tensor + ImageNet normalise, no resize)."""
from __future__ import annotations

import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def preprocess(image: np.ndarray) -> torch.Tensor:
    t = torch.from_numpy(image.copy()).permute(2, 0, 1).to(torch.float32).div(255)
    return (t - torch.tensor(MEAN).view(3, 1, 1)) / torch.tensor(STD).view(3, 1, 1)


class GdinoOracle:
    def __init__(self, seed: int = 0):
        from transformers import GroundingDinoConfig, GroundingDinoForObjectDetection

        torch.manual_seed(seed)
        self.model = GroundingDinoForObjectDetection(GroundingDinoConfig()).eval().float()

    def state_dict(self):
        return self.model.state_dict()

    @torch.inference_mode()
    def backbone_features(self, image: np.ndarray):
        bb = self.model.model.backbone.conv_encoder.model
        return [f[0] for f in bb(preprocess(image).unsqueeze(0), return_dict=True).feature_maps]

    @torch.inference_mode()
    def raw_outputs(self, image: np.ndarray, input_ids, input_noise: float = 0.0, noise_seed: int = 0):
        """``input_noise`` > 0: the normalised pixels are multiplied by (1 + input_noise * N(0,1)) -- the oracle's own conditioning
        (how far ITS outputs move under a half-precision-rounding-sized perturbation) is the yardstick of the decision tests."""
        ids = torch.tensor([list(input_ids)], dtype=torch.long)
        h, w = image.shape[:2]
        px = preprocess(image).unsqueeze(0)
        if input_noise > 0.0:
            px = px * (1.0 + input_noise * torch.randn(px.shape, generator=torch.Generator().manual_seed(noise_seed)))
        out = self.model(pixel_values=px, input_ids=ids, token_type_ids=torch.zeros_like(ids),
                         attention_mask=torch.ones_like(ids), pixel_mask=torch.ones(1, h, w, dtype=torch.long))
        return out.logits[0].sigmoid(), out.pred_boxes[0]
