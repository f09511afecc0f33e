"""dl_frontend_submit / dl_frontend_collect (the streaming form of the host-buffer front end): same results as the blocking
call, bit for bit; two contexts in flight at once; misuse is refused instead of corrupting the batch in flight."""
import numpy as np
import pytest

from helpers import workload

pytestmark = pytest.mark.gpu


def rows16(scans):
    return [np.ascontiguousarray(s.view(np.uint8).reshape(-1, 32)[:, :16]).view(np.float32).reshape(-1, 4) for s in scans]


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert list(x.pose_estimate_local) == list(y.pose_estimate_local)
        assert (x.ok, x.num_first_filter, x.num_returns, x.num_high_resolution, x.num_low_resolution) == \
               (y.ok, y.num_first_filter, y.num_returns, y.num_high_resolution, y.num_low_resolution)
        assert x.summary.num_iterations == y.summary.num_iterations and x.summary.final_cost == y.summary.final_cost


def test_submit_collect_equals_blocking_call(orc):
    import dliom
    w = workload(beams=16, num_map_scans=8, num_scans=12)
    a, b = dliom.Context(0), dliom.Context(0)
    hi, lo = dliom.Grid.from_oracle(a, w["hi"]), dliom.Grid.from_oracle(a, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    fo.range_row_floats = 4
    scans = rows16(w["scans"])
    args = (w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    want = a.frontend_match_batch(fo, scans, *args)
    assert all(r.ok for r in want)
    # one context, submit then collect
    a.frontend_submit(fo, scans, *args)
    same(a.frontend_collect(), want)
    # two contexts in flight at once (grids shared read-only), collected in either order, repeatedly
    first = (scans[:7], w["prev"][:7], w["cur"][:7])
    second = (scans[7:], w["prev"][7:], w["cur"][7:])
    for _ in range(3):
        a.frontend_submit(fo, first[0], w["origin"], first[1], first[2], w["submap_pose"], hi, lo)
        b.frontend_submit(fo, second[0], w["origin"], second[1], second[2], w["submap_pose"], hi, lo)
        rb = b.frontend_collect()
        ra = a.frontend_collect()
        same(list(ra) + list(rb), want)
    a.close(); b.close()


def test_submit_misuse_is_refused(orc):
    import dliom
    w = workload(beams=16, num_map_scans=8, num_scans=12)
    c = dliom.Context(0)
    hi, lo = dliom.Grid.from_oracle(c, w["hi"]), dliom.Grid.from_oracle(c, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    fo.range_row_floats = 4
    scans = rows16(w["scans"])
    args = (w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    with pytest.raises(dliom.DlError):
        c._submitted = (None, 12)
        c.frontend_collect()                       # nothing submitted
    c.frontend_submit(fo, scans, *args)
    with pytest.raises(dliom.DlError):
        c.frontend_submit(fo, scans, *args)        # second batch on the same context
    with pytest.raises(dliom.DlError):
        c.voxel_filter(np.zeros((10, 3), np.float32), 0.5)   # scratch is owned by the batch in flight
    assert all(r.ok for r in c.frontend_collect())
    keep = c.voxel_filter(np.zeros((10, 3), np.float32), 0.5)  # and usable again afterwards
    assert len(keep) == 1
    c.close()


def test_strided_host_layout_single_copy(orc):
    """host_scan_stride_rows: scans in one allocation at a constant stride are uploaded with one strided copy per
    sub-batch; same results as the per-scan copies, and a stride that does not describe the pointers is refused."""
    import dliom
    w = workload(beams=16, num_map_scans=8, num_scans=12)
    c = dliom.Context(0)
    hi, lo = dliom.Grid.from_oracle(c, w["hi"]), dliom.Grid.from_oracle(c, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    fo.range_row_floats = 4
    scans = rows16(w["scans"])
    args = (w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    want = c.frontend_match_batch(fo, scans, *args)
    stride = max(len(s) for s in scans) + 5
    block = np.full((len(scans), stride, 4), np.nan, np.float32)      # NaN padding must never be looked at
    for b, s in enumerate(scans):
        block[b, :len(s)] = s
    views = [block[b, :len(s)] for b, s in enumerate(scans)]
    fo.host_scan_stride_rows = stride
    same(c.frontend_match_batch(fo, views, *args), want)
    c.frontend_submit(fo, views, *args)
    same(c.frontend_collect(), want)
    with pytest.raises(dliom.DlError):
        c.frontend_match_batch(fo, scans, *args)                      # independent buffers, but a stride was promised
    fo.host_scan_stride_rows = 3
    with pytest.raises(dliom.DlError):
        c.frontend_match_batch(fo, views, *args)                      # smaller than a scan
    c.close()


def test_twelve_byte_rows_with_time_runs_are_bit_identical(orc):
    """range_row_floats = 3: bare x y z rows + the per-point times as runs (a spinning LiDAR stamps a firing column with one time).
    Same floats reach the deskew, so every result equals the 16-byte-row call bit for bit; a quarter fewer bytes cross PCIe."""
    import dliom
    from helpers import workload
    w = workload()
    ctx = dliom.Context(0)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    fo4 = dliom.FrontendOptions.from_oracle(w["opts"])
    fo4.range_row_floats = 4
    rows4 = [np.ascontiguousarray(np.stack([s["x"], s["y"], s["z"], s["t"]], 1)) for s in w["scans"]]
    want = ctx.frontend_match_batch(fo4, rows4, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    fo3 = dliom.FrontendOptions.from_oracle(w["opts"])
    runs = dliom.TimeRuns([s["t"] for s in w["scans"]]).attach(fo3)
    rows3 = [np.ascontiguousarray(r[:, :3]) for r in rows4]
    got = ctx.frontend_match_batch(fo3, rows3, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    tr = fo3._time_runs
    assert tr.offsets[-1] < 0.1 * sum(len(r) for r in rows4)          # ~1 800 columns per 28 800-point sweep
    for a, b in zip(got, want):
        assert a.ok == b.ok == 1 and list(a.pose_estimate_local) == list(b.pose_estimate_local)
        assert (a.num_first_filter, a.num_returns, a.num_misses, a.num_high_resolution) == (b.num_first_filter, b.num_returns, b.num_misses, b.num_high_resolution)
    # a scan without per-point times (|t_0| < 1e-3 -> no deskew, LTB:430-433) as a single run
    flat = dliom.FrontendOptions.from_oracle(w["opts"])
    dliom.TimeRuns([np.zeros(len(r), np.float32) for r in rows3]).attach(flat)
    z4 = [np.ascontiguousarray(np.concatenate([r, np.zeros((len(r), 1), np.float32)], 1)) for r in rows3]
    a = ctx.frontend_match_batch(flat, rows3, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    b = ctx.frontend_match_batch(fo4, z4, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    assert all(list(x.pose_estimate_local) == list(y.pose_estimate_local) for x, y in zip(a, b))
    # argument checks: runs missing, not starting at row 0
    bad = dliom.FrontendOptions.from_oracle(w["opts"])
    bad.range_row_floats = 3
    with pytest.raises(dliom.DlError):
        ctx.frontend_match_batch(bad, rows3, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    broken = dliom.TimeRuns([s["t"] for s in w["scans"]])
    broken.first_row[0] = 1          # a scan's first run must start at its row 0
    broken.attach(bad)
    with pytest.raises(dliom.DlError):
        ctx.frontend_match_batch(bad, rows3, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    ctx.close()


def test_key_range_flag_is_per_scan(orc):
    """The fused front half keys its second voxel filter on voxel indices RELATIVE to the scan's pose (dl_frontend.cu): with a
    0.4 mm voxel the 16-bit axis span of a ~29 k-point scan is +-13 m, so a scan with farther points is flagged (ok = -1) while a
    scan of the same batch whose points all lie within 5 m is registered normally."""
    import dliom
    from helpers import workload
    w = workload()
    ctx = dliom.Context(0)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    fo.range_row_floats = 4
    fo.voxel_filter_size = 4e-4
    full = np.ascontiguousarray(np.stack([w["scans"][0][k] for k in "xyzt"], 1))
    near = np.ascontiguousarray(full[np.linalg.norm(full[:, :3], axis=1) < 5.0])
    near[-1, 3] = 0.0
    assert len(near) > 500 and np.linalg.norm(full[:, :3], axis=1).max() > 20.0
    res = ctx.frontend_match_batch(fo, [near, full], w["origin"], w["prev"][[0, 0]], w["cur"][[0, 0]], w["submap_pose"], hi, lo)
    assert res[0].ok == 1 and res[1].ok == -1
    # the stock voxel size is nowhere near the limit
    fo.voxel_filter_size = w["opts"].voxel_filter_size
    res = ctx.frontend_match_batch(fo, [near, full], w["origin"], w["prev"][[0, 0]], w["cur"][[0, 0]], w["submap_pose"], hi, lo)
    assert res[0].ok == 1 and res[1].ok == 1
    ctx.close()
