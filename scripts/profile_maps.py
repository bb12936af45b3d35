"""Grid-kernel workloads for ncu: value-map update and obstacle update at batch 1 and 32
(640x480 depth, 1000^2 grid).  cudaProfilerStart/Stop bracket one launch of each."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vlfm_b200 import _lib
from vlfm_b200.mapping.value_map import ValueMapBatch
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory

lib = _lib.load()
FOV = float(np.deg2rad(79.0)); G = 1000; H, W = 480, 640
frames = [trajectory(s, 4) for s in range(32)]
timed = "--time" in sys.argv

def run(B):
    eng = ValueMapBatch(B, 1, size=G, use_max_confidence=False)
    obst = torch.zeros(B, G, G, dtype=torch.uint8, device="cuda"); nav = torch.zeros_like(obst)
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    fx = focal_from_hfov(W)
    half = int(math.ceil(5.0 * 20 * math.sqrt(1 + (W / 2 / fx) ** 2))) + 3 + 2
    def step(i, full):
        depth = torch.from_numpy(np.stack([frames[b][i % 4].depth for b in range(B)])).cuda()
        tf = torch.from_numpy(np.stack([frames[b][i % 4].tf for b in range(B)])).cuda()
        vals = torch.full((B, 1), 0.5, dtype=torch.float64, device="cuda")
        def go():
            eng.update(vals, depth, tf, 0.5, 5.0, FOV)
            p = _lib.ObstacleParams(H, W, G, 20, 4.5, 0.5, 5.0, fx, fx, 0.61, 0.88, 7, 1 if full else 0, half)
            _lib.check(lib.vlfm_obstacle_update(ctypes.byref(p), B, None, obst.data_ptr(), nav.data_ptr(), depth.data_ptr(),
                                                tf.data_ptr(), None, status.data_ptr(), _lib.stream_ptr()), "obstacle")
        return go
    step(0, True)(); step(1, False)(); torch.cuda.synchronize()
    go = step(2, False)
    if timed:
        for _ in range(5): go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): go()
        e1.record(); torch.cuda.synchronize()
        print(f"B={B}: value+obstacle update {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per batch-step")
    else:
        torch.cuda.profiler.start(); go(); torch.cuda.synchronize(); torch.cuda.profiler.stop()

for B in (1, 32):
    run(B)
print("done")
