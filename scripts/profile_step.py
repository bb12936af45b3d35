"""One configs[1] step (ITM cosine + value-map update, batch 1) for ncu:
cudaProfilerStart/Stop bracket exactly `--steps` steps after warm-up.  Run with
VLFM_NO_GRAPH=1 so every kernel is an individual launch."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import FOV, MIN_D, MAX_D, G, PROMPT, make_frames
from vlfm_b200.mapping.value_map import ValueMapBatch
from vlfm_b200.vlm.blip2_config import Blip2Dims, random_state_dict
from vlfm_b200.vlm.blip2itm import BLIP2ITM

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=1); ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
dims = Blip2Dims(); B = a.batch
itm = BLIP2ITM(state_dict=random_state_dict(dims, 0), dims=dims, max_batch=B)
eng = ValueMapBatch(B, 1, size=G, use_max_confidence=False)
fr = make_frames(0)
rgb = [torch.from_numpy(np.stack([f.rgb] * B)).cuda() for f in fr[:4]]
dep = [torch.from_numpy(np.stack([f.depth] * B)).cuda() for f in fr[:4]]
tfs = [torch.from_numpy(np.stack([f.tf] * B)).cuda() for f in fr[:4]]
def step(i):
    c = itm.cosine_device(rgb[i % 4], PROMPT)
    eng.update(c.double().view(B, 1), dep[i % 4], tfs[i % 4], MIN_D, MAX_D, FOV)
for i in range(3): step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(a.steps): step(3 + i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.steps, "steps")
