"""Development aid: step the GPU ObstacleMap and the oracle side by side, and on the first mismatch compare the
fog-of-war intermediates (cone / blocked / visible-after-rays / new explored window) kept in the workspace."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle.explore_oracle as ex
from oracle import contours as ct, cv_draw as dr, cv_prims as pr
from oracle.obstacle_map_oracle import ObstacleMapOracle
from vlfm_b200.mapping.obstacle_map import ObstacleMap
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory

seed, hw, size, steps, bound = int(sys.argv[1]), (int(sys.argv[2]), int(sys.argv[3])), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
ex.PRIMS = "numpy"
fx = focal_from_hfov(hw[1])
o = ObstacleMapOracle(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=-1, size=size)
g = ObstacleMap(0.61, 0.88, 0.18, area_thresh=1.5, hole_area_thresh=-1, size=size)
WIN = 512 * 512
for i, f in enumerate(trajectory(seed, steps, h=hw[0], w=hw[1], bound_m=bound)):
    prev_explored = o.explored_area.copy()
    o.update_obstacles(f.depth, f.tf, 0.5, 5.0, fx, fx)
    nav = np.asarray(o._navigable_map).astype(np.uint8)
    # oracle fog intermediates
    agent = o.xy_to_px(f.tf[:2, 3].reshape(1, 2))[0]
    yaw = float(np.arctan2(f.tf[1, 0], f.tf[0, 0]))
    heading = np.rad2deg(ex.wrap_heading(yaw + np.pi / 2)); fov = np.rad2deg(np.deg2rad(79)); L = 100.0
    src = agent.astype(int)
    cone = dr.ellipse_sector(size, size, (int(src[0]), int(src[1])), int(L), heading - fov / 2, heading + fov / 2).astype(np.uint8)
    blocked = cone & (1 - nav)
    conts = ct.find_external_contours(blocked, True)
    ex.explore_step(o, f.tf, 5.0, np.deg2rad(79))
    g.update_map(f.depth, f.tf, 0.5, 5.0, fx, fx, np.deg2rad(79))
    ge = g.explored_area
    if np.array_equal(ge, o.explored_area) and np.asarray(o._frontiers_px).shape == np.asarray(g._frontiers_px).shape and np.array_equal(np.asarray(o._frontiers_px), np.asarray(g._frontiers_px)):
        print("step", i, "ok", int(ge.sum()), len(o._frontiers_px)); continue
    print("step", i, "MISMATCH explored", int(ge.sum()), int(o.explored_area.sum()), "frontiers", np.asarray(g._frontiers_px).tolist(), np.asarray(o._frontiers_px).tolist())
    ys, xs = np.nonzero(ge != o.explored_area); print(" diff cells (x,y):", list(zip(xs.tolist(), ys.tolist()))[:10], "agent", agent.tolist())
    assert np.array_equal(g._navigable_map.astype(np.uint8), nav), "nav differs"
    ws8 = g._explore_impl.ws.view(torch.uint8).cpu().numpy()
    W0 = 209; ox, oy = int(agent[0]) - 104, int(agent[1]) - 104
    def win(k): return ws8[k * WIN : k * WIN + W0 * W0].reshape(W0, W0)
    gcone = win(0)
    ocone = cone[oy:oy + W0, ox:ox + W0]
    print(" cone equal:", np.array_equal(gcone, ocone), int(gcone.sum()), int(ocone.sum()))
    # blocked/visible are overwritten by later stages? blocked (1) and visible (2) persist; cut (3); newexp (4)
    print(" blocked equal:", np.array_equal(win(1), blocked[oy:oy + W0, ox:ox + W0]), int(win(1).sum()), int(blocked.sum()), "n contours", len(conts))
    # recompute oracle visible-after-cut
    pts = []
    for c in conts:
        if ct.is_convex(c):
            a, b = ex._extreme_bearing_points(src, c, heading); pts.append(a.reshape(-1, 2)); pts.append(b.reshape(-1, 2))
        else:
            pts.append(c.reshape(-1, 2))
    if pts:
        pts = np.concatenate(pts, 0)
        segs = ex._ray_segments(src, pts, L * 1.05)
        cut = np.zeros((size, size), bool)
        for a, b in segs: dr.thick_line2(cut, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])))
        ocut = cut[oy:oy + W0, ox:ox + W0]
        gcut = win(3) > 0
        vis_o = ((cone & nav) > 0) & ~cut
        print(" rays oracle", len(segs), " cut equal inside cone:", np.array_equal(gcut & (ocone > 0), ocut & (ocone > 0)), int((gcut & (ocone > 0)).sum()), int((ocut & (ocone > 0)).sum()))
        print(" visible equal:", np.array_equal(win(2) > 0, vis_o[oy:oy + W0, ox:ox + W0]))
        d = (gcut & (ocone > 0)) ^ (ocut & (ocone > 0)); yy, xx = np.nonzero(d); print("  cut diff (win x,y):", list(zip(xx.tolist(), yy.tolist()))[:10])
    break
