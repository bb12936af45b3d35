"""GroundingDINO behind the reference's class surface (vlfm/vlm/grounding_dino.py:23-85).

The Swin-T backbone -- the dense contraction north_star names -- runs on the hand-written sm_100a kernels
(``SwinBackboneEngine``); every nn.Linear, the deformable encoder layers, the image<->text fusion layers and the decoder layers
run on the library too (``gdino_accel``: tcgen05 GEMM, fused multi-scale deformable sampling, bi-attention).  The module graph
that sequences them is HF's ``GroundingDinoForObjectDetection`` (architecture-equivalent to groundingdino@eeba084); what is
still PyTorch glue is listed in DESIGN.md section 7.  Post-processing restates groundingdino.util.inference.predict.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch

from .detections import ObjectDetections
from .gdino_accel import accelerate
from .swin_engine import SwinBackboneEngine

GROUNDING_DINO_CONFIG = "GroundingDINO/groundingdino/config/GroundingDINO_SwinT_OGC.py"
GROUNDING_DINO_WEIGHTS = "data/groundingdino_swint_ogc.pth"
CLASSES = "chair . person . dog ."  # grounding_dino.py:20
BACKBONE_PREFIX = "model.backbone.conv_encoder.model."


class _Features(torch.nn.Module):
    """Stands in for the HF backbone module: returns the feature maps our engine produced."""

    def __init__(self):
        super().__init__()
        self.maps: List[torch.Tensor] = []

    def forward(self, pixel_values=None, **kw):
        from transformers.modeling_outputs import BackboneOutput

        return BackboneOutput(feature_maps=tuple(self.maps))


class SimpleCaptionTokenizer:
    """SYNTHETIC stand-in for bert-base-uncased (no vocab offline): [CLS]=101, '.'=1012, [SEP]=102,
    words -> stable ids in [2000, vocab).  decode() inverts it for the words it has seen."""

    def __init__(self, vocab: int = 30522):
        self.vocab = vocab
        self.words: Dict[int, str] = {}

    def encode(self, caption: str) -> List[int]:
        import zlib

        ids = [101]
        for tok in caption.replace(".", " . ").split():
            if tok == ".":
                ids.append(1012)
            else:
                i = 2000 + zlib.crc32(tok.encode()) % (self.vocab - 2000)
                self.words[i] = tok
                ids.append(i)
        return ids + [102]

    def decode(self, ids: List[int]) -> str:
        return " ".join("." if i == 1012 else self.words.get(i, "[UNK]") for i in ids if i not in (101, 102))


class GroundingDINO:
    def __init__(self, config_path: str = GROUNDING_DINO_CONFIG, weights_path: str = GROUNDING_DINO_WEIGHTS,
                 caption: str = CLASSES, box_threshold: float = 0.35, text_threshold: float = 0.25,
                 device: torch.device = torch.device("cuda"), state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 tokenizer: Optional[Any] = None, seed: int = 0, synthetic: bool = False):
        """``weights_path`` (grounding_dino.py:33 ``load_model(config_path, weights_path)``) is honoured: a groundingdino
        ``.pth`` (original key names, converted by ``gdino_weights.convert_groundingdino_state_dict``) or an HF-layout
        state dict; ``VLFM_GDINO_WEIGHTS`` overrides it.  ``config_path`` selects nothing here: the only architecture built
        is GroundingDINO_SwinT_OGC (the reference's default and the one its weights file is for).  Without a checkpoint
        the constructor RAISES unless ``synthetic=True`` (seeded random weights: tests / benchmarks only) -- a detector
        that silently runs on random weights returns meaningless boxes."""
        from transformers import GroundingDinoConfig, GroundingDinoForObjectDetection

        from .gdino_weights import load_checkpoint

        cfg = GroundingDinoConfig()
        real = False          # weights read from a checkpoint file (then the real vocabulary is mandatory)
        if state_dict is None:
            path = os.environ.get("VLFM_GDINO_WEIGHTS", "") or (weights_path if weights_path and os.path.exists(weights_path) else "")
            if path:
                state_dict = load_checkpoint(path)
                real = True
            elif not synthetic:
                raise FileNotFoundError(
                    f"GroundingDINO: no checkpoint at weights_path={weights_path!r} and VLFM_GDINO_WEIGHTS is unset. "
                    "Pass synthetic=True to run on seeded random weights (tests / benchmarks only).")
        torch.manual_seed(seed)
        model = GroundingDinoForObjectDetection(cfg)
        if state_dict is not None:
            missing, unexpected = model.load_state_dict(state_dict, strict=False)
            missing = [k for k in missing if "position_ids" not in k]
            if missing or unexpected:      # a checkpoint that does not cover the model is an error, never a silent partial load
                raise KeyError(f"GroundingDINO checkpoint does not match the model: {len(missing)} missing keys (e.g. {missing[:4]}), "
                               f"{len(unexpected)} unexpected keys (e.g. {list(unexpected)[:4]})")
        sd = model.state_dict()
        self.device = device
        self.backbone = SwinBackboneEngine(sd, prefix=BACKBONE_PREFIX, embed_dim=cfg.backbone_config.embed_dim,
                                           depths=cfg.backbone_config.depths, heads=cfg.backbone_config.num_heads,
                                           out_stages=tuple(cfg.backbone_config.out_indices), eps=cfg.backbone_config.layer_norm_eps,
                                           device=device)
        self._features = _Features()
        model.model.backbone.conv_encoder.model = self._features
        self.model = model.to(device).eval()
        # feature enhancer / decoder: nn.Linear -> tcgen05 GEMM, deformable-attention sampling -> vlfm_msda_forward
        self.accel = accelerate(self.model) if os.environ.get("VLFM_GDINO_ACCEL", "1") != "0" else {}
        # model-level glue (neck, proposal scoring, top-900 selection, heads) on the library's kernels; VLFM_GDINO_OWN_FORWARD=0 keeps
        # HF's GroundingDinoModel.forward for A/B comparisons
        self.fwd = None
        if os.environ.get("VLFM_GDINO_OWN_FORWARD", "1") != "0":
            from .gdino_forward import GdinoForward
            from .gdino_ops import LibOps

            self.fwd = GdinoForward(self.model, LibOps(), self.backbone)
        self.caption = caption
        self.box_threshold = box_threshold
        self.text_threshold = text_threshold
        if tokenizer is None:
            vocab = os.environ.get("VLFM_BERT_VOCAB", "")
            if vocab:
                from .blip2itm import WordPieceCaptionTokenizer

                tokenizer = WordPieceCaptionTokenizer(vocab)
            elif real and not synthetic:
                raise FileNotFoundError("GroundingDINO: real weights need the bert-base-uncased vocabulary: set VLFM_BERT_VOCAB=<vocab.txt> "
                                        "or pass tokenizer= (the crc32 stand-in only makes sense with synthetic weights)")
            else:
                if not synthetic:
                    import warnings

                    warnings.warn("GroundingDINO: in-memory state_dict without tokenizer= / VLFM_BERT_VOCAB: using the crc32 stand-in "
                                  "tokenizer, which is only meaningful with synthetic weights")
                tokenizer = SimpleCaptionTokenizer(cfg.text_config.vocab_size)
        self.tokenizer = tokenizer
        self._pin: Optional[torch.Tensor] = None
        self._dev: Optional[torch.Tensor] = None
        self._static: Dict[Any, Dict[str, Any]] = {}
        self._graph_ok = os.environ.get("VLFM_GDINO_GRAPH", "1") != "0"
        self._graph_max_batch = int(os.environ.get("VLFM_GDINO_GRAPH_MAX_BATCH", "4"))
        self.graph_error: Optional[str] = None

    def _forward_static(self, st):
        """One forward on the static buffers of ``st`` (eager or under CUDA-graph capture)."""
        if self.fwd is not None:
            st["logits"], st["boxes"] = self.fwd.forward(st["img"], st["key_ids"])
            st["keep"] = self.fwd.last       # a captured graph reads the cached shape / caption constants by address: keep them alive
            return
        self._features.maps = self.backbone.forward(st["img"])
        out = self.model(pixel_values=st["dummy"], input_ids=st["ids"], token_type_ids=st["tt"], attention_mask=st["am"], pixel_mask=st["pm"])
        st["logits"], st["boxes"] = out.logits.sigmoid(), out.pred_boxes
        # a captured graph reads the cached text-tower output by ADDRESS: this entry owns a reference, so evicting the caption
        # from the text cache (FIFO, 8 entries) can never free memory a live graph replays from
        tb = self.model.model.text_backbone
        if hasattr(tb, "cache"):
            st["text_ref"] = tb.cache.get(tb.key)

    @torch.inference_mode()
    def raw_outputs_device(self, images: torch.Tensor, input_ids: List[int]):
        """images [B,H,W,3] uint8 on the device -> (sigmoid logits [B,900,256], boxes [B,900,4] cxcywh).

        Small batches (the per-step policy call is batch 1) replay a CUDA graph of the whole detector per
        (batch, image size, caption): the module graph is ~700 launches and launch-bound otherwise."""
        b, h, w = images.shape[:3]
        key = (int(b), int(h), int(w), tuple(int(i) for i in input_ids))
        tb = self.model.model.text_backbone
        if hasattr(tb, "key"):
            tb.key = (key[3], key[0])
        st = self._static.get(key)
        if st is None:
            ids = torch.tensor([list(key[3])], dtype=torch.long, device=self.device).expand(b, -1).contiguous()
            st = {"img": torch.empty_like(images), "ids": ids, "tt": torch.zeros_like(ids), "am": torch.ones_like(ids),
                  "dummy": torch.zeros(b, 3, h, w, device=self.device),   # only its shape is used (pixel mask); features come from our engine
                  "pm": torch.ones(b, h, w, dtype=torch.long, device=self.device), "calls": 0, "graph": None, "key_ids": list(key[3])}
            if len(self._static) >= 4:
                self._static.pop(next(iter(self._static)))
            self._static[key] = st
        st["img"].copy_(images, non_blocking=True)
        st["calls"] += 1
        use_graph = self._graph_ok and b <= self._graph_max_batch
        if use_graph and st["graph"] is None and st["calls"] >= 2:       # call 1 warmed every cache up eagerly
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._forward_static(st)
                st["graph"] = g
            except Exception as e:  # a sync point inside the module graph: stay eager (loudly, once)
                self._graph_ok = False
                self.graph_error = repr(e)
                print(f"[vlfm_b200] GroundingDINO CUDA-graph capture disabled: {e}", flush=True)
                torch.cuda.synchronize()
        if st["graph"] is not None:
            st["graph"].replay()
        else:
            self._forward_static(st)
        return st["logits"], st["boxes"]

    def raw_outputs(self, image: np.ndarray, input_ids: List[int]):
        """-> (sigmoid logits [900,256], boxes [900,4] cxcywh) on the device."""
        image = np.ascontiguousarray(image, dtype=np.uint8)
        if self._pin is None or self._pin.shape[1:] != image.shape:
            self._pin = torch.empty((1,) + image.shape, dtype=torch.uint8).pin_memory()
            self._dev = torch.empty((1,) + image.shape, dtype=torch.uint8, device=self.device)
        self._pin[0].numpy()[...] = image
        self._dev.copy_(self._pin, non_blocking=True)
        logits, boxes = self.raw_outputs_device(self._dev, input_ids)
        return logits[0], boxes[0]

    def predict(self, image: np.ndarray, caption: Optional[str] = None) -> ObjectDetections:
        """grounding_dino.py:38-74."""
        caption_to_use = self.caption if caption is None else caption
        text = caption_to_use.lower().strip()
        if not text.endswith("."):
            text = text + "."
        ids = self.tokenizer.encode(text)
        logits, boxes = self.raw_outputs(image, ids)
        logits, boxes = logits.cpu(), boxes.cpu()
        keep = logits.max(dim=1)[0] > self.box_threshold
        logits, boxes = logits[keep], boxes[keep]
        phrases = []
        for row in logits:
            pos = row > self.text_threshold
            pos[0] = False
            pos[len(ids) - 1 :] = False
            phrases.append(self.tokenizer.decode([ids[i] for i in pos.nonzero(as_tuple=True)[0].tolist()]).replace(".", "").strip())
        det = ObjectDetections(boxes, logits.max(dim=1)[0], phrases, image_source=image)
        classes = caption_to_use[: -len(" .")].split(" . ")
        det.filter_by_class(classes)
        return det


_SHARED: Dict[str, GroundingDINO] = {}


class GroundingDINOClient:
    """Same signature as the HTTP client (grounding_dino.py:77-85); in-process."""

    def __init__(self, port: int = 12181, model: Optional[GroundingDINO] = None):
        if model is None:
            if "default" not in _SHARED:
                _SHARED["default"] = GroundingDINO(synthetic=os.environ.get("VLFM_SYNTHETIC_WEIGHTS", "") == "1")
            model = _SHARED["default"]
        self.model = model

    def predict(self, image_numpy: np.ndarray, caption: Optional[str] = "") -> ObjectDetections:
        return self.model.predict(image_numpy, caption=caption if caption else None)
