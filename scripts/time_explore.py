import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vlfm_b200.mapping.obstacle_map import ObstacleMap
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory
from oracle.obstacle_map_oracle import ObstacleMapOracle
fx = focal_from_hfov(640)
g = ObstacleMap(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=int(os.environ.get("HOLE", "100000")), size=1000)
o = ObstacleMapOracle(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=int(os.environ.get("HOLE", "100000")), size=1000)
fr = trajectory(5, 40, bound_m=12.0)
for f in fr[:8]: g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
torch.cuda.synchronize(); t0 = time.perf_counter()
for f in fr[8:]: g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"GPU ObstacleMap.update_map (obstacle + explore + frontiers): {(t1-t0)/32*1e3:.2f} ms/step, explored {int(g.explored_area.sum())}, frontiers {len(g._frontiers_px)}")
for f in fr[:8]: o.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
t0 = time.perf_counter()
for f in fr[8:]: o.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
t1 = time.perf_counter()
print(f"CPU oracle (numpy/cv2): {(t1-t0)/32*1e3:.2f} ms/step, explored {int(o.explored_area.sum())}")
print("match:", np.array_equal(g.explored_area, o.explored_area), np.array_equal(np.asarray(g._frontiers_px), np.asarray(o._frontiers_px)))
