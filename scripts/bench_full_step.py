"""BASELINE.json configs[2]: the FULL policy step for a batch of environments on one B200 --
GroundingDINO detect + BLIP-2 ITC + ObstacleMap (hole fill, scatter, dilate, fog-of-war, frontiers) + ValueMap fuse +
frontier scoring -- with per-component CUDA-event times.  Auxiliary measurement (bench.py keeps configs[1]).

    python scripts/bench_full_step.py --batch 32 --steps 6 --warmup 3 [--no-gdino] [--streams 8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlfm_b200.mapping.obstacle_map import ObstacleMap  # noqa: E402
from vlfm_b200.mapping.value_map import ValueMapBatch  # noqa: E402
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory  # noqa: E402
from vlfm_b200.vlm.blip2_config import Blip2Dims, random_state_dict  # noqa: E402
from vlfm_b200.vlm.blip2itm import BLIP2ITM  # noqa: E402

H, W, G = 480, 640, 1000
MIN_D, MAX_D, FOV = 0.5, 5.0, float(np.deg2rad(79))
PROMPT = "Seems like there is a chair ahead."


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--hole-thresh", type=int, default=100000)
    ap.add_argument("--no-gdino", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, NF = a.batch, a.steps + a.warmup
    dims = Blip2Dims()
    itm = BLIP2ITM(state_dict=random_state_dict(dims, 0), dims=dims, max_batch=B, device=dev)
    gd = None
    if not a.no_gdino:
        from vlfm_b200.vlm.grounding_dino import GroundingDINO

        gd = GroundingDINO(device=dev, synthetic=True)
        ids = gd.tokenizer.encode("chair . couch . potted plant . bed . toilet . tv .")
    vmb = ValueMapBatch(B, 1, size=G, use_max_confidence=False, device=dev)
    oms = [ObstacleMap(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=a.hole_thresh, size=G, device=dev) for _ in range(B)]
    streams = [torch.cuda.Stream(dev) for _ in range(max(1, a.streams))]
    fx = focal_from_hfov(W)
    frames = [trajectory(e, NF, h=H, w=W, bound_m=15.0, with_rgb=True) for e in range(B)]
    # every step's frames wait in page-locked host memory, batched per step (as a vectorised simulator bridge leaves them)
    rgb_pin = torch.empty((NF, B, H, W, 3), dtype=torch.uint8).pin_memory()
    depth_pin = torch.empty((NF, B, H, W), dtype=torch.float32).pin_memory()
    tf_pin = torch.empty((NF, B, 16), dtype=torch.float64).pin_memory()
    for i in range(NF):
        for e in range(B):
            f = frames[e][i]
            rgb_pin[i, e].numpy()[...] = f.rgb; depth_pin[i, e].numpy()[...] = f.depth; tf_pin[i, e].numpy()[...] = f.tf.reshape(16)
            f.depth = depth_pin[i, e].numpy()                    # ObstacleMap reads the same page-locked frame
    rgb_dev, depth_dev, tf_dev = (torch.empty_like(t[0], device=dev) for t in (rgb_pin, depth_pin, tf_pin))
    names = ["h2d", "gdino", "itc", "obstacle+explore", "value_fuse", "frontier_scoring"]
    acc = {k: 0.0 for k in names}
    n_front = 0

    def step(i, timed):
        nonlocal n_front
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        main = torch.cuda.current_stream()
        ev[0].record()
        # page-locked host frames -> HBM (inside the step)
        rgb_dev.copy_(rgb_pin[i], non_blocking=True); depth_dev.copy_(depth_pin[i], non_blocking=True); tf_dev.copy_(tf_pin[i], non_blocking=True)
        ev[1].record()
        if gd is not None:
            logits, boxes = gd.raw_outputs_device(rgb_dev, ids)
            keep = logits.max(dim=2)[0] > gd.box_threshold    # compaction mask stays on the device
            _ = keep.sum()
        ev[2].record()
        cos = itm.cosine_device(rgb_dev, PROMPT)
        ev[3].record()
        for s in streams:
            s.wait_stream(main)
        for e in range(B):                                    # independent envs: round-robin over streams
            with torch.cuda.stream(streams[e % len(streams)]):
                oms[e].update_map(frames[e][i].depth, frames[e][i].tf, MIN_D, MAX_D, fx, fx, FOV)
        for s in streams:
            main.wait_stream(s)
        ev[4].record()
        vmb.update(cos.double().view(B, 1), depth_dev, tf_dev.view(B, 4, 4), MIN_D, MAX_D, FOV)
        ev[5].record()
        for e in range(B):                                    # ITMPolicy._sort_frontiers_by_value: D2H of the frontier list + disc medians
            fr = oms[e].frontiers
            n_front += len(fr)
            if len(fr):
                px = oms[e]._xy_to_px(fr[:, :2])
                vmb.disc_median(e, np.stack([px[:, 1], px[:, 0]], axis=1), 10)
        ev[6].record()
        torch.cuda.synchronize()
        if timed:
            for k, nme in enumerate(names):
                acc[nme] += ev[k].elapsed_time(ev[k + 1])

    for i in range(a.warmup):
        step(i, False)
    torch.cuda.synchronize()
    n_front = 0
    t0 = time.perf_counter()
    for i in range(a.steps):
        step(a.warmup + i, True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out = {
        "metric": "full-step env-steps/s (GroundingDINO + BLIP-2 ITC + Obstacle/Value/Frontier update)", "value": B * a.steps / wall,
        "unit": "env-steps/s", "n_gpus": 1, "batch": B, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps,
        "timing": "host wall clock around whole steps incl. H2D of RGB-D from page-locked host frames + D2H of frontier lists; per-component CUDA events",
        "component_ms_per_step": {k: v / a.steps for k, v in acc.items()},
        "frontiers_per_env_step": n_front / (B * a.steps), "streams": len(streams), "hole_area_thresh": a.hole_thresh,
        "gdino": "Swin-T, linears, deformable / fusion / decoder layers on own kernels; neck, query selection and glue HF PyTorch" if gd is not None else "skipped",
        "config": {"workload": f"configs[2]: full step, batch={B} envs, 640x480 RGB-D, 1000^2 grid, 1xB200", "data": "synthetic"},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
