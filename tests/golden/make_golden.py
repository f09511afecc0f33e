"""Freezes the oracle's outputs on seeded inputs into tests/golden/frontend_16beam.json.

The reference cannot run here (C++ tree with un-vendored dependencies), so these vectors are not reference
outputs: they are the ORACLE's outputs at the commit that pinned it to the reference's own fixtures
(tests/test_oracle_golden.py, tests/test_fcsm_oracle.py). Their job is to stop silent drift — of the synthetic generator, of
the oracle, of the device path — between rounds: tests/test_golden_vectors.py recomputes them with the oracle (CPU) and with
the CUDA path (GPU) and compares. Large arrays are stored as SHA-256 of their bytes, small ones in full.

    python tests/golden/make_golden.py        # rewrites the JSON; review the diff before committing
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute(orc):
    """Everything the golden file holds, from the oracle. Returns (record, inputs) — inputs feed the device side of the test."""
    from helpers import workload
    from test_decode import LAYOUTS, POSE, message
    from test_fcsm_oracle import CLOUD, TEST_OPTS, fixture_grid
    w = workload(beams=16, num_map_scans=6, num_scans=2)
    o = w["opts"]
    rec = {"workload": "synth.Scene(42), 16-beam, 6 map sweeps from t=2.0, 2 sweeps to register",
           "grid_hi_cells": sha(np.stack([c.astype(np.int64) for c in w["hi"].export()])),
           "grid_lo_cells": sha(np.stack([c.astype(np.int64) for c in w["lo"].export()])), "scans": []}
    for s in range(2):
        rows = w["scans"][s]
        ing = orc.ingest_scan(o, rows, w["origin"], w["prev"][s], w["cur"][s])
        pts = ing["returns_tracking"]
        hk, hp = orc.adaptive_voxel_filter(pts, o.hi_max_length, o.hi_min_num_points, o.hi_max_range)
        lk, lp = orc.adaptive_voxel_filter(pts, o.lo_max_length, o.lo_min_num_points, o.lo_max_range)
        m = orc.match_scan(o, pts, ing["current_pose"].astype(np.float64), w["submap_pose"], w["hi"], w["lo"])
        rec["scans"].append({
            "input_rows": sha(rows), "num_points": int(len(rows)),
            "first_keep": sha(ing["first_keep"].astype(np.int64)), "num_first": int(len(ing["first_keep"])),
            "returns_local": sha(ing["returns_local"]), "returns_tracking": sha(pts), "num_returns": int(len(pts)),
            "misses_tracking": sha(ing["misses_tracking"]), "current_pose": [float(v) for v in ing["current_pose"]],
            "adaptive_high": {"keep": sha(hk.astype(np.int64)), "count": int(len(hk)), "passes": [float(v) for v in hp]},
            "adaptive_low": {"keep": sha(lk.astype(np.int64)), "count": int(len(lk)), "passes": [float(v) for v in lp]},
            "pose_estimate_local": [float(v) for v in m["pose_estimate_local"]],
            "num_iterations": int(m["summary"]["num_iterations"]), "final_cost": float(m["summary"]["final_cost"])})
    g = fixture_grid(orc, (0.25, -0.1, 0.05))
    f = orc.fcsm_match_3dof(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **TEST_OPTS)
    rec["loop_closure_fixture"] = {"score": float(f.score), "offset": [int(v) for v in f.offset],
                                   "low_resolution_score": float(f.low_resolution_score), "pose": [float(v) for v in f.pose]}
    full, scan_index, num_scans = orc.fcsm_match_full(g, g, CLOUD, CLOUD, [0, 0, 0, 0.9987502603949663, 0, 0, 0.04997916927067833],
                                                      orc.IDENTITY_POSE, 0.1, xy_window=0.8, z_window=0.8, angular_window=0.3,
                                                      min_low_resolution_score=0.15, min_rotational_score=0.1, depth=6, full_depth=6)
    rec["loop_closure_full_match"] = {"score": float(full.score), "scan_index": int(scan_index), "num_scans": int(num_scans),
                                      "pose": [float(v) for v in full.pose]}
    # sparse pose adjustment: 2 submaps, 6 nodes on a circle, constraints from both submaps with fixed pseudo-noise
    ang = np.linspace(0, 2 * np.pi, 6, endpoint=False)
    nodes = [np.array([3 * np.cos(a), 3 * np.sin(a), 0.1 * k, np.cos(a / 2), 0, 0, np.sin(a / 2)]) for k, a in enumerate(ang)]
    submaps = [np.array([0, 0, 0, 1.0, 0, 0, 0]), np.array([1.0, 0.5, 0, np.cos(0.2), 0, 0, np.sin(0.2)])]
    cons = [(sid, k, nodes[k] + np.array([0.01 * ((k + sid) % 3 - 1), -0.02 * ((k * 2 + sid) % 3 - 1), 0.005 * k, 0, 0, 0, 0]), 1.0 + sid, 2.0)
            for sid in range(2) for k in range(6)]
    so, no, summ = orc.pose_graph_solve(submaps, [n + np.array([0.05, -0.05, 0.02, 0, 0, 0, 0]) for n in nodes], cons)
    rec["pose_graph"] = {"final_cost": float(summ["final_cost"]), "num_iterations": int(summ["num_iterations"]),
                         "node0": [float(v) for v in no[0]], "submap1": [float(v) for v in so[1]]}
    rec["decode"] = {}
    for name in LAYOUTS:
        data, step, offs, tt, _ = message(name, 4097, 11, last_is_bad=(name == "ouster48"))
        rows, off = orc.decode_point_cloud2(data, step, offs, tt, POSE)
        rec["decode"][name] = {"message": sha(data), "rows": sha(rows), "count": int(len(rows)), "stamp_offset": float(off)}
    return rec, w


if __name__ == "__main__":
    import orc
    record, _ = compute(orc)
    path = os.path.join(HERE, "frontend_16beam.json")
    json.dump(record, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")
