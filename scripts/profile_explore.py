import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vlfm_b200.mapping.obstacle_map import ObstacleMap
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory
fx = focal_from_hfov(640)
g = ObstacleMap(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=int(os.environ.get("HOLE", "100000")), size=1000)
fr = trajectory(3, 12, bound_m=12.0)
for f in fr[:10]: g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
torch.cuda.synchronize(); torch.cuda.profiler.start()
g.update_map(fr[10].depth, fr[10].tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("explored", int(g.explored_area.sum()), "frontiers", len(g._frontiers_px))
