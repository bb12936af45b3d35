"""ObjectPointCloudMap on the GPU vs oracle/object_map_oracle.py (itself pinned to the real reference class and, for DBSCAN,
to scikit-learn): clouds bit-identical (float64 points, same order), same best object / target cloud, with numpy's global
generator seeded identically on both sides (the reference's own source of randomness)."""
import numpy as np
import pytest

from oracle import object_map_oracle as om
from vlfm_b200.utils.synthetic import focal_from_hfov, make_object_mask, tf_from_pose, trajectory

pytestmark = pytest.mark.gpu


def _scenario(seed, steps, h, w):
    rng = np.random.default_rng(100 + seed)
    fx = focal_from_hfov(w)
    out = []
    for i, f in enumerate(trajectory(seed, steps, h=h, w=w, bound_m=6.0)):
        side = ["any", "left", "any", "right", "any"][i % 5]
        mask = make_object_mask(rng, h, w, side)
        if side != "any":                                   # drop the stray specks so that the bounding box is the blob's (too_offset)
            import cv2
            mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, np.ones((5, 5), np.uint8))
        depth = f.depth.copy()
        if i % 3 == 2:
            depth[mask > 0] = np.float32(0.98)              # beyond 95 % of max_depth: out-of-range ids
        out.append((depth, mask, f.tf, fx))
    return out


def test_device_primitives_vs_oracle():
    """erode + unproject (np.where order) and the DBSCAN filter, each against its numpy statement"""
    from vlfm_b200.mapping.object_point_cloud_map import ObjectPointCloudMap

    rng = np.random.default_rng(0)
    for k in (0, 1, 3):
        m = ObjectPointCloudMap(erosion_size=k)
        m.use_dbscan = False
        for depth, mask, tf, fx in _scenario(k, 3, 240, 320):
            mask[:, :2] = 1                                  # touches the image border
            ref = om.object_cloud(depth, om.erode_mask(mask, k), 0.5, 5.0, fx, fx)
            np.random.seed(3)
            ref = om.random_subarray(ref, 5000)
            np.random.seed(3)
            got = m._extract_object_cloud(depth, mask, 0.5, 5.0, fx, fx)
            assert got.shape == ref.shape and np.array_equal(got, ref)
    m = ObjectPointCloudMap(erosion_size=1)
    for t in range(6):
        n = int(rng.integers(200, 5000))
        pts = np.concatenate([rng.normal(0, 0.1, (n // 2, 3)) + [2, 0, 0], rng.normal(0, 0.3, (n // 3, 3)) + [3, 1, 0], rng.uniform(-4, 4, (n // 6, 3))])
        pts = pts[rng.permutation(len(pts))]
        import ctypes, torch
        from vlfm_b200 import _lib
        m._buffers(60, 100)
        dev = torch.from_numpy(pts).cuda()
        rc = m.lib.vlfm_dbscan_largest_cluster(dev.data_ptr(), None, len(pts), 0.2, 100, None, m._out.data_ptr(), m._count[1:].data_ptr(),
                                               m._db_ws.data_ptr(), m._db_ws.numel() * 4, _lib.stream_ptr())
        _lib.check(rc, "dbscan")
        cnt = int(m._count[1].item())
        ref = om.dbscan_filter(pts)
        assert cnt == len(ref)
        if cnt:
            assert np.array_equal(m._out[:cnt].cpu().numpy(), ref)


@pytest.mark.parametrize("use_dbscan,hw", [(True, (240, 320)), (False, (240, 320)), (True, (480, 640))])
def test_object_map_vs_oracle(use_dbscan, hw):
    from vlfm_b200.mapping.object_point_cloud_map import ObjectPointCloudMap

    for seed in range(2):
        g, o = ObjectPointCloudMap(erosion_size=2), om.ObjectPointCloudMapOracle(erosion_size=2)
        g.use_dbscan = o.use_dbscan = use_dbscan
        seen_offset = removed = False
        for depth, mask, tf, fx in _scenario(seed, 7, *hw):
            seen_offset |= om.too_offset(mask)
            for m in (o, g):
                np.random.seed(7 + seed)
                m.update_map("chair", depth, mask, tf, 0.5, 5.0, fx, fx)
                m.update_explored(tf, 5.0, np.deg2rad(79))
            assert o.has_object("chair") == g.has_object("chair")
            if o.has_object("chair"):
                assert np.array_equal(o.clouds["chair"], g.clouds["chair"])
                pos = tf[:2, 3] + 0.3
                assert np.array_equal(o.get_best_object("chair", pos), g.get_best_object("chair", pos))
                assert np.array_equal(o.get_target_cloud("chair"), g.get_target_cloud("chair"))
        assert seen_offset and o.has_object("chair")
        # walk up to an out-of-range detection: update_explored drops it on both sides
        c = o.clouds["chair"]
        far = c[c[:, 3] != 1]
        if len(far):
            p = far[0, :3]
            tf2 = tf_from_pose(p[0] - 1.0, p[1], 0.88, 0.0)
            n0 = len(o.clouds["chair"])
            for m in (o, g):
                m.update_explored(tf2, 5.0, np.deg2rad(79))
            assert len(o.clouds["chair"]) < n0 and np.array_equal(o.clouds["chair"], g.clouds["chair"])
            removed = True
        assert removed or seed > 0
    with pytest.raises(Exception):
        ObjectPointCloudMap(erosion_size=2)._extract_object_cloud(np.zeros((4, 4), np.float32), np.zeros((5, 5), np.uint8), 0.5, 5.0, 1.0, 1.0)
