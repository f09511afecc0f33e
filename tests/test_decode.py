"""Wire format -> TimedPointCloud rows: SensorBridge::HandlePointCloud2Message + HandleRangefinder restated on the
raw sensor_msgs/PointCloud2 bytes. CPU part: the oracle against an independent numpy reading of the same loops; GPU part:
the device decode against the oracle, bit for bit, for the four sensor types, packed and aligned layouts, NaN / Inf points."""
import numpy as np
import pytest

LAYOUTS = {
    # name: (numpy dtype of a point, (x, y, z, time) offsets, time type)
    "velodyne_packed22": (np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                                    "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22}), (0, 4, 8, 18), 1),
    "velodyne_pcl32": (np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                                 "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32}), (0, 4, 8, 24), 1),
    "ouster48": (np.dtype({"names": ["x", "y", "z", "intensity", "time", "reflectivity", "ring", "noise", "range"],
                           "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "u1", "<u2", "<u4"],
                           "offsets": [0, 4, 8, 16, 20, 24, 26, 28, 32], "itemsize": 48}), (0, 4, 8, 20), 2),
    "robosense32": (np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "u1", "<u2", "<f8"],
                              "offsets": [0, 4, 8, 16, 18, 24], "itemsize": 32}), (0, 4, 8, 24), 3),
    "xyzi16": (np.dtype({"names": ["x", "y", "z", "intensity"], "formats": ["<f4"] * 4, "offsets": [0, 4, 8, 12], "itemsize": 16}),
               (0, 4, 8, 0), 0),
}
POSE = np.array([0.3, -0.2, 1.7, 0.9987502603949663, 0.0, 0.0, 0.04997916927067833])   # 0.1 rad about z


def message(name, n, seed, last_is_bad=False):
    dt, offs, tt = LAYOUTS[name]
    rng = np.random.default_rng(seed)
    pts = np.zeros(n, dt)
    xyz = rng.uniform(-60, 60, (n, 3)).astype(np.float32)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    bad = rng.choice(n, max(n // 50, 1), replace=False) if n else np.array([], int)
    for k, i in enumerate(bad):
        pts[["x", "y", "z"][k % 3]][i] = [np.nan, np.inf, -np.inf][k % 3]
    if last_is_bad and n:
        pts["y"][-1] = np.nan
    if tt == 1:
        pts["time"] = np.sort(rng.uniform(0, 0.1, n)).astype(np.float32)
    elif tt == 2:
        pts["time"] = np.sort(rng.integers(0, 100_000_000, n)).astype(np.uint32)
    elif tt == 3:
        pts["time"] = 1.6e9 + np.sort(rng.uniform(0, 0.1, n))
    return pts.view(np.uint8).reshape(-1), dt.itemsize, offs, tt, pts


def numpy_reading(pts, tt, pose):
    """The reference's loops with numpy scalars (independent of the oracle's C++)."""
    from helpers import apply_pose
    f = np.float32
    keep = np.isfinite(pts["x"]) & np.isfinite(pts["y"]) & np.isfinite(pts["z"])
    if tt == 1:
        last = np.float64(pts["time"][-1])
        t = (pts["time"].astype(np.float64) - last).astype(f)
    elif tt == 2:
        scaled = (pts["time"].astype(f) * f(1e-9)).astype(f)
        last = np.float64(scaled[-1])
        t = (scaled.astype(np.float64) - last).astype(f)
    elif tt == 3:
        last = pts["time"][-1]
        t = (pts["time"] - last).astype(f)
    else:
        last, t = 0.0, np.zeros(len(pts), f)
    return keep, t, (float(last) if tt in (1, 2) else 0.0)


@pytest.mark.parametrize("name", list(LAYOUTS))
def test_oracle_decode_semantics(orc, name):
    data, step, offs, tt, pts = message(name, 5000, 3)
    rows, off = orc.decode_point_cloud2(data, step, offs, tt, POSE)
    keep, t, want_off = numpy_reading(pts, tt, POSE)
    assert len(rows) == int(keep.sum()) and off == want_off
    assert np.array_equal(rows[:, 3], t[keep])                         # per-point times, order kept, NaN / Inf dropped
    ident, _ = orc.decode_point_cloud2(data, step, offs, tt, orc.IDENTITY_POSE)
    assert np.array_equal(ident[:, :3], np.stack([pts["x"], pts["y"], pts["z"]], 1)[keep])
    assert np.abs(rows[:, :3] - (ident[:, :3].astype(np.float64) @ rot(POSE).T + POSE[:3])).max() < 2e-5
    if tt:
        assert rows[-1, 3] == 0.0 and np.all(rows[:, 3] <= 0)          # CHECK_LE(ranges.back().time, 0) downstream


def test_transform_timed_point_cloud_reference_fixture(orc):
    """sensor/point_cloud_test.cc:28-51 (TransformPointCloud / TransformTimedPointCloud): a quarter turn about z maps
    (0.5, 0.5, 1) to (-0.5, 0.5, 1) and (3.5, 0.5, 42) to (-0.5, 3.5, 42) within 1e-6; times pass through."""
    dt, offs, tt = LAYOUTS["xyzi16"]
    pts = np.zeros(2, dt)
    pts["x"], pts["y"], pts["z"] = [0.5, 3.5], [0.5, 0.5], [1.0, 42.0]
    quarter = np.array([0, 0, 0, np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4)])     # Embed3D(Rigid2f::Rotation(M_PI_2))
    rows, _ = orc.decode_point_cloud2(pts.view(np.uint8).reshape(-1), dt.itemsize, offs, tt, quarter)
    assert np.abs(rows[:, :3] - [[-0.5, 0.5, 1.0], [-0.5, 3.5, 42.0]]).max() < 1e-6 and np.all(rows[:, 3] == 0)


def rot(p):
    w, x, y, z = p[3:]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_oracle_decode_edge_cases(orc):
    data, step, offs, tt, pts = message("velodyne_packed22", 300, 5, last_is_bad=True)
    rows, off = orc.decode_point_cloud2(data, step, offs, tt, POSE)
    assert off == float(pts["time"][-1])           # the LAST point of the message sets the reference time even when it is dropped
    assert len(rows) < 300 and rows[-1, 3] <= 0
    empty, off0 = orc.decode_point_cloud2(np.zeros(0, np.uint8), step, offs, tt, POSE)
    assert len(empty) == 0 and off0 == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(LAYOUTS))
def test_device_decode_bit_exact(orc, name):
    import dliom
    ctx = dliom.Context(0)
    for n, seed, bad_last in ((1, 1, False), (257, 2, True), (130656, 3, False)):
        data, step, offs, tt, _ = message(name, n, seed, bad_last)
        want, woff = orc.decode_point_cloud2(data, step, offs, tt, POSE)
        got, goff = ctx.decode_point_cloud2(data, step, offs, tt, POSE)
        assert goff == woff and got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    got, goff = ctx.decode_point_cloud2(np.zeros(0, np.uint8), LAYOUTS[name][0].itemsize, LAYOUTS[name][1], LAYOUTS[name][2], POSE)
    assert len(got) == 0 and goff == 0.0
    ctx.close()


@pytest.mark.gpu
def test_device_decode_feeds_the_front_end(orc):
    """Message uploaded as it arrived, decoded on the device straight into the buffer dl_frontend_match_batch_dev reads:
    same poses as the host path that receives the already-decoded TimedPointCloud."""
    import ctypes as C
    import dliom
    from helpers import workload
    w = workload(beams=16, num_map_scans=8, num_scans=2)
    ctx = dliom.Context(0)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    fo.range_row_floats = 4
    dt, offs, tt = LAYOUTS["velodyne_packed22"]
    rows_host, sizes = [], []
    cap = max(len(s) for s in w["scans"])
    d_rows = ctx.device_alloc(2 * cap * 16)
    for k, s in enumerate(w["scans"]):
        msg = np.zeros(len(s), dt)
        msg["x"], msg["y"], msg["z"] = s["x"], s["y"], s["z"]
        msg["time"] = (s["t"].astype(np.float64) + 0.1).astype(np.float32)      # the driver stamps from the FIRST point
        raw = msg.view(np.uint8).reshape(-1)
        d_msg = ctx.device_alloc(len(raw))
        ctx.copy_to_device(d_msg, raw)
        kept, off = ctx.decode_point_cloud2_dev(d_msg, len(s), dt.itemsize, offs, tt, orc.IDENTITY_POSE,
                                                C.c_void_p(d_rows.value + k * cap * 16))
        ctx.device_free(d_msg)
        rows, _ = orc.decode_point_cloud2(raw, dt.itemsize, offs, tt, orc.IDENTITY_POSE)
        assert kept == len(rows) == len(s)
        rows_host.append(rows); sizes.append(kept)
    want = ctx.frontend_match_batch(fo, rows_host, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    d_res = ctx.device_alloc(2 * C.sizeof(dliom.ScanResult))
    ctx.frontend_match_batch_dev(fo, d_rows, cap, np.array(sizes, np.int64), w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo, d_res)
    got = ctx.fetch_results(d_res, 2)
    for a, b in zip(got, want):
        assert a.ok == 1 and list(a.pose_estimate_local) == list(b.pose_estimate_local)
    ctx.close()
