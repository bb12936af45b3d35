"""fp32 CPU oracle for BLIP2ITM.cosine (reference: vlfm/vlm/blip2itm.py:37-54).

TEST INFRASTRUCTURE.  PARITY UNPINNED with respect to lavis==1.0.2 (absent offline; no
reference test or golden vector exists for this path): the oracle is the
architecture-equivalent HF ``Blip2ForImageTextRetrieval`` run eagerly in float32 with the
ITC head (normalize(vision_proj(Q)) . normalize(text_proj(CLS)), max over queries), fed
by the reference's own preprocessing chain restated from blip2itm.py:48-49 and lavis'
BlipImageEvalProcessor: PIL bicubic Resize((224,224)) -> ToTensor -> Normalize(CLIP).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def preprocess(image: np.ndarray, size: int) -> torch.Tensor:
    from PIL import Image

    pil = Image.fromarray(image).resize((size, size), Image.BICUBIC)
    t = torch.from_numpy(np.asarray(pil).copy()).permute(2, 0, 1).to(torch.float32).div(255)
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(3, 1, 1)
    return t.sub(mean).div(std)


class Blip2Oracle:
    def __init__(self, dims, state_dict: Dict[str, torch.Tensor]):
        from transformers import Blip2ForImageTextRetrieval

        from vlfm_b200.vlm.blip2_config import hf_config

        with torch.device("meta"):
            model = Blip2ForImageTextRetrieval(hf_config(dims))
        model = model.to_empty(device="cpu")
        missing, unexpected = model.load_state_dict({k: v.clone() for k, v in state_dict.items()}, strict=False, assign=True)
        assert not unexpected, unexpected
        assert all("position_ids" in m for m in missing), missing
        model.embeddings.position_ids = torch.arange(dims.max_pos).expand((1, -1))
        self.model = model.eval().float()
        self.dims = dims

    @torch.inference_mode()
    def cosine(self, image: np.ndarray, token_ids: Sequence[int]) -> float:
        px = preprocess(image, self.dims.image).unsqueeze(0)
        ids = torch.tensor([list(token_ids)], dtype=torch.long)
        out = self.model(pixel_values=px, input_ids=ids, attention_mask=torch.ones_like(ids), use_image_text_matching_head=False)
        return float(out.logits_per_image.reshape(-1)[0])

    # The same forward split at the image / text boundary (modeling_blip_2.py, ITC branch of Blip2ForImageTextRetrieval.forward):
    # many frames x many prompts cost one ViT pass per frame.  tests/test_oracle_blip2_split.py pins the split to cosine().
    @torch.inference_mode()
    def image_features(self, image: np.ndarray) -> torch.Tensor:
        """normalised vision_projection of the 32 query outputs, [queries, proj]"""
        m = self.model
        px = preprocess(image, self.dims.image).unsqueeze(0)
        img = m.vision_model(pixel_values=px)[0]
        att = torch.ones(img.shape[:-1], dtype=torch.long)
        q = m.qformer(query_embeds=m.query_tokens.expand(1, -1, -1), encoder_hidden_states=img, encoder_attention_mask=att, return_dict=True)[0]
        return torch.nn.functional.normalize(m.vision_projection(q), dim=-1)[0]

    @torch.inference_mode()
    def text_feature(self, token_ids: Sequence[int]) -> torch.Tensor:
        m = self.model
        ids = torch.tensor([list(token_ids)], dtype=torch.long)
        emb = m.embeddings(input_ids=ids)
        t = m.qformer(query_embeds=emb, query_length=0, attention_mask=torch.ones_like(ids), return_dict=True)[0]
        return torch.nn.functional.normalize(m.text_projection(t[:, 0, :]), dim=-1)[0]

    @staticmethod
    def cosine_from(img_feat: torch.Tensor, txt_feat: torch.Tensor) -> float:
        return float((img_feat @ txt_feat).max())

    @torch.inference_mode()
    def image_tokens(self, image: np.ndarray) -> torch.Tensor:
        px = preprocess(image, self.dims.image).unsqueeze(0)
        return self.model.vision_model(pixel_values=px)[0][0]
