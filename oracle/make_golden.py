"""Generate tests/golden/*.npz by running the REAL reference (build container only).

Usage:  PYTHONPATH=/root/repo python oracle/make_golden.py
Needs /root/reference.  Inputs are regenerated from vlfm_b200.utils.synthetic with the
recorded seeds (an input checksum is stored so generator drift is detected); outputs
are stored sparsely (flat indices + values of non-zero cells).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_import  # noqa: E402
from vlfm_b200.utils.synthetic import focal_from_hfov, trajectory  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
FOV = float(np.deg2rad(79.0))

VALUE_CASES = [
    # name, channels, use_max_conf, fusion, size, seed, steps, (H, W), bound
    ("vm_weighted", 1, False, "default", 480, 11, 6, (480, 640), 5.0),
    ("vm_maxconf_c2", 2, True, "default", 480, 12, 6, (480, 640), 5.0),
    ("vm_edge_clip", 1, False, "default", 260, 13, 5, (240, 320), 6.2),
    ("vm_replace", 1, False, "replace", 400, 14, 4, (120, 160), 3.0),
    ("vm_equal", 1, False, "equal_weighting", 400, 15, 4, (120, 160), 3.0),
]


def digest(frames) -> str:
    h = hashlib.sha256()
    for f in frames:
        h.update(np.ascontiguousarray(f.depth).tobytes())
        h.update(np.ascontiguousarray(f.tf).tobytes())
    return h.hexdigest()


def sparse(a: np.ndarray):
    flat = a.reshape(-1)
    idx = np.flatnonzero(flat)
    return idx.astype(np.int32), flat[idx]


def value_cases() -> None:
    RV = ref_import.value_map_class()
    for name, ch, maxc, fus, size, seed, steps, (h, w), bound in VALUE_CASES:
        RV._confidence_masks.clear()
        ref = RV(ch, size=size, use_max_confidence=maxc, fusion_type=fus)
        frames = trajectory(seed, steps, h=h, w=w, bound_m=bound)
        rng = np.random.default_rng(seed)
        vals = rng.random((steps, ch))
        for f, v in zip(frames, vals):
            ref.update_map(v, f.depth, f.tf, 0.5, 5.0, FOV)
        ci, cv = sparse(ref._map)
        vi, vv = sparse(ref._value_map)
        wps = np.array([[f.xy[0] + 0.4, f.xy[1] - 0.3] for f in frames])
        red = (lambda s: [max(t) for t in s]) if ch > 1 else None
        sw, sv = ref.sort_waypoints(wps, 0.5, reduce_fn=red)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            channels=ch, use_max_confidence=maxc, fusion=fus, size=size, seed=seed, steps=steps,
            hw=np.array([h, w]), bound=bound, values=vals, input_sha256=digest(frames),
            conf_idx=ci, conf_val=cv, value_idx=vi, value_val=vv.astype(np.float64),
            value_dtype=str(ref._value_map.dtype), waypoints=wps, sorted_wp=sw,
            sorted_val=np.asarray(sv, dtype=np.float64),
        )
        print(name, "conf nz", ci.size, "value nz", vi.size)


def obstacle_cases() -> None:
    try:
        from oracle.make_golden_obstacle import obstacle_cases as run
    except ImportError:
        return
    run(OUT)


if __name__ == "__main__":
    assert ref_import.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    value_cases()
    obstacle_cases()
