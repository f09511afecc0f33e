"""Committed golden vectors (tests/golden/frontend_16beam.json, written by tests/golden/make_golden.py): the oracle must still
produce them (CPU), and the CUDA path must produce them too (GPU) — bit for bit where the work is integer / index / float32
selection, within the stated tolerance for the double-precision solve."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "frontend_16beam.json")))


def test_oracle_reproduces_the_golden_vectors(orc):
    from make_golden import compute
    now, _ = compute(orc)
    assert now["grid_hi_cells"] == GOLDEN["grid_hi_cells"] and now["grid_lo_cells"] == GOLDEN["grid_lo_cells"]
    assert now["decode"] == GOLDEN["decode"]
    assert now["loop_closure_fixture"] == GOLDEN["loop_closure_fixture"]
    assert now["loop_closure_full_match"] == GOLDEN["loop_closure_full_match"]
    assert now["pose_graph"]["num_iterations"] == GOLDEN["pose_graph"]["num_iterations"]
    for k in ("node0", "submap1"):
        assert np.allclose(now["pose_graph"][k], GOLDEN["pose_graph"][k], rtol=0, atol=1e-12)
    assert abs(now["pose_graph"]["final_cost"] - GOLDEN["pose_graph"]["final_cost"]) <= 1e-15
    for a, b in zip(now["scans"], GOLDEN["scans"]):
        for k in ("input_rows", "num_points", "first_keep", "num_first", "returns_local", "returns_tracking", "num_returns",
                  "misses_tracking", "current_pose", "adaptive_high", "adaptive_low", "num_iterations"):
            assert a[k] == b[k], k
        assert np.allclose(a["pose_estimate_local"], b["pose_estimate_local"], rtol=0, atol=1e-12)
        assert abs(a["final_cost"] - b["final_cost"]) <= 1e-12 * max(1.0, b["final_cost"])


@pytest.mark.gpu
def test_device_reproduces_the_golden_vectors(orc):
    import dliom
    from helpers import pose_error, workload
    from make_golden import sha
    from test_decode import LAYOUTS, POSE, message
    from test_fcsm_oracle import CLOUD, TEST_OPTS, fixture_grid
    ctx = dliom.Context(0)
    w = workload(beams=16, num_map_scans=6, num_scans=2)
    o = w["opts"]
    fo = dliom.FrontendOptions.from_oracle(o)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    assert sha(np.stack([c.astype(np.int64) for c in hi.export()])) == GOLDEN["grid_hi_cells"]
    res = ctx.frontend_match_batch(fo, w["scans"], w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    for s, g in enumerate(GOLDEN["scans"]):
        assert sha(w["scans"][s]) == g["input_rows"]
        ing = ctx.ingest_scan(fo, w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        assert sha(ing["first_keep"].astype(np.int64)) == g["first_keep"]
        for k in ("returns_local", "returns_tracking", "misses_tracking"):
            assert sha(ing[k]) == g[k], k
        assert [float(v) for v in ing["current_pose"]] == g["current_pose"]
        pts = ing["returns_tracking"]
        hk, hp = ctx.adaptive_voxel_filter(pts, o.hi_max_length, o.hi_min_num_points, o.hi_max_range)
        lk, lp = ctx.adaptive_voxel_filter(pts, o.lo_max_length, o.lo_min_num_points, o.lo_max_range)
        assert sha(hk.astype(np.int64)) == g["adaptive_high"]["keep"] and [float(v) for v in hp] == g["adaptive_high"]["passes"]
        assert sha(lk.astype(np.int64)) == g["adaptive_low"]["keep"] and [float(v) for v in lp] == g["adaptive_low"]["passes"]
        r = res[s]
        assert (r.num_first_filter, r.num_returns, r.num_high_resolution, r.num_low_resolution) == \
               (g["num_first"], g["num_returns"], g["adaptive_high"]["count"], g["adaptive_low"]["count"])
        dt, dr = pose_error(np.array(r.pose_estimate_local[:]), np.array(g["pose_estimate_local"]))
        assert dt < 1e-7 and dr < 1e-7            # BASELINE tolerance is 1e-4 m / 1e-5 rad
        assert r.summary.num_iterations == g["num_iterations"]
        assert abs(r.summary.final_cost - g["final_cost"]) <= 1e-6 * max(1.0, g["final_cost"])
    og = fixture_grid(orc, (0.25, -0.1, 0.05))
    dg = dliom.Grid.from_oracle(ctx, og)
    f = ctx.fcsm_match_3dof(dg, dg, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **TEST_OPTS)
    gf = GOLDEN["loop_closure_fixture"]
    assert float(f.score) == gf["score"] and float(f.low_resolution_score) == gf["low_resolution_score"]
    for name, gd in GOLDEN["decode"].items():
        data, step, offs, tt, _ = message(name, 4097, 11, last_is_bad=(name == "ouster48"))
        assert sha(data) == gd["message"]
        rows, off = ctx.decode_point_cloud2(data, step, offs, tt, POSE)
        assert sha(rows) == gd["rows"] and len(rows) == gd["count"] and off == gd["stamp_offset"]
    ctx.close()
