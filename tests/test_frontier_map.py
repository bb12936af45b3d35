"""FrontierMap (vlfm/mapping/frontier_map.py:10-77): host bookkeeping around one cosine call per update that introduces a
new frontier.  Checked against explicit expectations everywhere, and against the REAL reference class (its HTTP encoder
replaced by the same scripted one) where /root/reference exists."""
import sys
import types

import numpy as np
import pytest

from conftest import has_reference
from vlfm_b200.mapping.frontier_map import FrontierMap


class ScriptedEncoder:
    def __init__(self):
        self.calls = 0

    def cosine(self, image, text):
        self.calls += 1
        return 0.1 * self.calls + float(image.sum() % 7) * 1e-3


def _stream(seed, steps=30):
    rng = np.random.default_rng(seed)
    pool = [rng.uniform(-5, 5, 2).round(2) for _ in range(12)]
    out = []
    for _ in range(steps):
        k = int(rng.integers(0, 6))
        idx = rng.choice(len(pool), size=k, replace=False)
        out.append(([pool[i].copy() for i in idx], rng.integers(0, 255, (4, 4, 3), dtype=np.uint8)))
    return out


def test_update_sort_reset_semantics():
    enc = ScriptedEncoder()
    fm = FrontierMap(encoder=enc)
    img = np.zeros((2, 2, 3), np.uint8)
    a, b, c = np.array([1.0, 2.0]), np.array([3.0, 4.0]), np.array([5.0, 6.0])
    fm.update([a, b], img, "x")
    assert enc.calls == 1 and [f.cosine for f in fm.frontiers] == [0.1, 0.1]            # one encode for both new frontiers
    fm.update([b.copy(), c], img, "x")                                                  # a vanished, b kept (array_equal), c new
    assert enc.calls == 2 and len(fm.frontiers) == 2
    assert np.array_equal(fm.frontiers[0].xyz, b) and fm.frontiers[0].cosine == 0.1 and fm.frontiers[1].cosine == 0.2
    fm.update([b, c], img, "x")
    assert enc.calls == 2                                                               # nothing new: no encode
    pts, vals = fm.sort_waypoints()
    assert vals == [0.2, 0.1] and np.array_equal(pts, np.array([c, b]))
    fm.reset()
    assert fm.frontiers == []
    fm.update([], img, "x")
    assert enc.calls == 2 and fm.frontiers == []


@pytest.mark.skipif(not has_reference(), reason="/root/reference not present")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_matches_live_reference(seed):
    stub = types.ModuleType("vlfm.vlm.blip2itm")
    stub.BLIP2ITMClient = ScriptedEncoder
    saved = {k: sys.modules.get(k) for k in ("vlfm.vlm.blip2itm", "vlfm.mapping.frontier_map")}
    sys.modules["vlfm.vlm.blip2itm"] = stub
    sys.modules.pop("vlfm.mapping.frontier_map", None)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    try:
        from vlfm.mapping.frontier_map import FrontierMap as RefFrontierMap  # type: ignore

        ref, got = RefFrontierMap(), FrontierMap(encoder=ScriptedEncoder())
        ref.frontiers = []
        for locs, img in _stream(seed):
            ref.update(locs, img, "a chair")
            got.update(locs, img, "a chair")
            assert len(ref.frontiers) == len(got.frontiers)
            for r, g in zip(ref.frontiers, got.frontiers):
                assert np.array_equal(r.xyz, g.xyz) and r.cosine == g.cosine
            if ref.frontiers:
                (rp, rv), (gp, gv) = ref.sort_waypoints(), got.sort_waypoints()
                assert rv == gv and np.array_equal(rp, gp)
        assert ref.encoder.calls == got.encoder.calls
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
