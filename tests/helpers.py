"""Shared workload builders for the tests (synthetic scene -> oracle submap -> scans)."""
import functools

import numpy as np

SEVEN = np.array([[-3, 2, 0], [-4, 2, 0], [-5, 2, 0], [-6, 2, 0], [-6, 3, 1], [-6, 4, 2], [-7, 3, 1]], np.float32)


def seven_point_grid(orc, res, shift=(-1, 0, 0)):
    g = orc.Grid(res)
    for p in SEVEN:
        g.set_probability(g.cell_index(p + np.array(shift, np.float32)), 1.0)
    return g


def apply_pose(p7, pts):
    w, x, y, z = p7[3:]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return pts @ R.T + p7[:3]


def pose_error(a, b):
    """(translation error [m], rotation error [rad]) between two 7-vectors."""
    dt = np.linalg.norm(a[:3] - b[:3])
    qa, qb = a[3:] / np.linalg.norm(a[3:]), b[3:] / np.linalg.norm(b[3:])
    d = abs(float(np.dot(qa, qb)))
    return dt, 2 * np.arccos(min(1.0, d))


@functools.lru_cache(maxsize=4)
def workload(beams=16, num_map_scans=8, num_scans=4, hi_res=0.1, lo_res=0.45, start=2.0):
    """Builds a submap with the oracle's range-data inserter from `num_map_scans` sweeps, then `num_scans` further
    sweeps to register. Returns a dict; everything is deterministic."""
    import orc
    import synth
    scene = synth.Scene(42)
    opts = orc.FrontEndOptions.defaults()
    hi, lo = orc.Grid(hi_res), orc.Grid(lo_res)
    origin = np.zeros((1, 3), np.float32)
    t = start
    for _ in range(num_map_scans):
        rows = synth.make_scan(scene, beams, t)
        prev, cur = synth.pose7(t - 0.1), synth.pose7(t)
        ing = orc.ingest_scan(opts, rows, origin, prev, cur)
        local = apply_pose(cur, ing["returns_tracking"].astype(np.float64)).astype(np.float32)
        o = cur[:3].astype(np.float32)
        near = local[np.linalg.norm(local - o, axis=1) <= 20.0]
        hi.insert_range_data(o, near)
        lo.insert_range_data(o, local)
        t += 0.1
    rng = np.random.RandomState(45)
    scans, prevs, curs, truths, times = [], [], [], [], []
    for _ in range(num_scans):
        times.append(t)
        scans.append(synth.make_scan(scene, beams, t))
        prevs.append(synth.pose7(t - 0.1))
        truths.append(synth.pose7(t))
        curs.append(synth.perturb_pose(synth.pose7(t), rng, 0.05, 0.5))
        t += 0.1
    return {"opts": opts, "hi": hi, "lo": lo, "origin": origin, "scans": scans, "prev": np.array(prevs),
            "cur": np.array(curs), "truth": np.array(truths), "times": times, "submap_pose": orc.IDENTITY_POSE.copy()}
