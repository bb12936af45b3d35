"""oracle/object_map_oracle.py: DBSCAN restatement vs scikit-learn's implementation; the whole class vs the REAL reference
class (vlfm/mapping/object_point_cloud_map.py) imported with a stub open3d."""
import cv2
import numpy as np
import pytest

from conftest import has_reference
from oracle import object_map_oracle as om
from vlfm_b200.utils.synthetic import focal_from_hfov, make_object_mask, trajectory


def _clustered(rng, n):
    k = int(rng.integers(1, 5))
    parts = []
    for _ in range(k):
        c = rng.uniform(-2, 2, 3)
        parts.append(c + rng.normal(0, rng.uniform(0.03, 0.25), (int(n // k), 3)))
    parts.append(rng.uniform(-3, 3, (n // 10, 3)))
    p = np.concatenate(parts)
    return p[rng.permutation(len(p))]


def test_dbscan_labels_match_sklearn():
    from sklearn.cluster import DBSCAN

    rng = np.random.default_rng(0)
    for t in range(12):
        pts = _clustered(rng, int(rng.integers(300, 2500)))
        mp = [100, 40, 10][t % 3]
        ref = DBSCAN(eps=0.2, min_samples=mp, algorithm="brute").fit(pts).labels_
        got = om.dbscan_labels(pts, 0.2, mp)
        assert np.array_equal(ref, got), (t, (ref != got).sum())
    assert len(om.dbscan_filter(rng.uniform(-50, 50, (500, 3)))) == 0          # only noise


def test_erode_restatement():
    rng = np.random.default_rng(1)
    for k in (0, 1, 2, 3):
        m = make_object_mask(rng, 120, 160)
        m[:, :3] = 1                                                           # touches the image edge: the border does not erode
        assert np.array_equal(cv2.erode(m * 255, None, iterations=k), om.erode_mask_numpy(m, k))


def _scenario(seed, steps=6, h=240, w=320):
    rng = np.random.default_rng(100 + seed)
    fx = focal_from_hfov(w)
    out = []
    for i, f in enumerate(trajectory(seed, steps, h=h, w=w, bound_m=6.0)):
        side = ["any", "left", "any", "right", "any"][i % 5]
        mask = make_object_mask(rng, h, w, side)
        depth = f.depth.copy()
        if i % 3 == 2:
            depth[mask > 0] = np.float32(0.98)                                # a far detection: out-of-range ids
        out.append((depth, mask, f.tf, fx))
    return out


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
@pytest.mark.parametrize("use_dbscan", [True, False])
def test_oracle_class_matches_the_reference_class(use_dbscan):
    from oracle import ref_import

    R = ref_import.object_map_module().ObjectPointCloudMap
    for seed in range(3):
        r, o = R(erosion_size=2), om.ObjectPointCloudMapOracle(erosion_size=2)
        r.reset()
        r.use_dbscan = o.use_dbscan = use_dbscan
        for depth, mask, tf, fx in _scenario(seed):
            for m in (r, o):
                np.random.seed(7 + seed)
                m.update_map("chair", depth, mask, tf, 0.5, 5.0, fx, fx)
                m.update_explored(tf, 5.0, np.deg2rad(79))
            assert r.has_object("chair") == o.has_object("chair")
            if r.has_object("chair"):
                assert np.array_equal(r.clouds["chair"], o.clouds["chair"])
                pos = tf[:2, 3] + 0.3
                assert np.array_equal(r.get_best_object("chair", pos), o.get_best_object("chair", pos))
                assert np.array_equal(r.get_target_cloud("chair"), o.get_target_cloud("chair"))
