#!/usr/bin/env python
"""One correlative search on the configs[2] shape (0.05 m grid, 0.15 m / 1 deg window -> 456 533 candidates) through the C-ABI,
for ncu and for a CUDA-event timing of rtcsm_score_kernel alone (prints one JSON line)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "d-liom_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def main():
    import dliom
    import orc
    import synth
    from bench import apply_pose
    ctx = dliom.Context(0)
    scene = synth.Scene(42)
    opts = orc.FrontEndOptions.defaults()
    origin = np.zeros((1, 3), np.float32)
    hi, lo = ctx.grid(0.05), ctx.grid(0.45)
    ident = orc.IDENTITY_POSE.copy()
    for k in range(6):
        t = 2.0 + 0.1 * k
        rows = synth.make_scan(scene, 128, t)
        cur = synth.pose7(t)
        ing = orc.ingest_scan(opts, rows, origin, synth.pose7(t - 0.1), cur)
        local = apply_pose(cur, ing["returns_tracking"].astype(np.float64)).astype(np.float32)
        ctx.submap_insert_range_data(hi, lo, ident, cur[:3].astype(np.float32), local, high_resolution_max_range=20)
    t = 2.35
    rows = synth.make_scan(scene, 128, t)
    ing = orc.ingest_scan(opts, rows, origin, synth.pose7(t - 0.1), synth.pose7(t))
    pts = ing["returns_tracking"]
    hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
    cloud = pts[hk]
    init = synth.perturb_pose(synth.pose7(t), np.random.RandomState(3), 0.1, 0.5)
    ctx.set_profiling(False)
    got = ctx.rtcsm_match(hi, cloud, init, 0.15, np.deg2rad(1.0), 1e-1, 1e-1)
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        got = ctx.rtcsm_match(hi, cloud, init, 0.15, np.deg2rad(1.0), 1e-1, 1e-1)
    wall = (time.perf_counter() - t0) / n
    R, L = got["num_candidates"] // 343, 343
    print(json.dumps({"points": int(len(cloud)), "candidates": int(got["num_candidates"]), "score": float(got["score"]),
                      "wall_ms_per_call": wall * 1e3,
                      "algorithmic_bytes": float(R * (12.0 * len(cloud) + 2.0 * len(cloud) * L))}))


if __name__ == "__main__":
    main()
