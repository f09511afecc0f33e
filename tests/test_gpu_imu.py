"""GPU parity of the IMU row: device pre-integration and the fused scan-match + IMU solve against the
oracle (oracle/orc_imu.h). fp64 throughout: compared to ~1e-12 relative (pre-integration) / 1e-7 m (solve)."""
import numpy as np
import pytest

import imu_synth
from helpers import pose_error, workload

pytestmark = pytest.mark.gpu
NOISE = [3.99e-2, 1.56e-2, 6.4e-5, 3.6e-5]


@pytest.fixture(scope="module")
def ctx():
    import dliom
    c = dliom.Context(0)
    yield c
    c.close()


def test_preintegration_matches_oracle(ctx, orc):
    rng = np.random.RandomState(1)
    intervals, biases = [], []
    for k in range(9):
        t0 = 2.0 + 0.37 * k
        n = [0.1, 0.1, 0.05, 0.2, 0.1, 0.005, 0.1, 0.1, 0.1][k]
        intervals.append(imu_synth.samples(t0, t0 + n, noise=(3.99e-2, 1.56e-2), seed=k))
        biases.append(rng.normal(0, [2e-2] * 3 + [2e-3] * 3))
    got = ctx.imu_preintegrate(NOISE, intervals, np.array(biases))
    for k, (dt, acc, gyr) in enumerate(intervals):
        want = orc.imu_preintegrate(NOISE, biases[k][:3], biases[k][3:], dt, acc, gyr)
        g = got[k]
        assert abs(g.sum_dt - want.sum_dt) < 1e-15
        for name in ("delta_p", "delta_q", "delta_v"):
            assert np.allclose(np.array(getattr(g, name)), np.array(getattr(want, name)), rtol=1e-13, atol=1e-15), name
        assert np.allclose(np.array(g.jacobian), np.array(want.jacobian), rtol=1e-12, atol=1e-16)
        assert np.allclose(np.array(g.covariance), np.array(want.covariance), rtol=1e-11, atol=1e-24)
        si = imu_synth.state(2.0 + 0.37 * k, biases[k][:3], biases[k][3:])
        assert np.allclose(ctx.imu_predict(si, g), orc.imu_predict(si, want), rtol=1e-13, atol=1e-13)


def test_preintegration_single_sample_interval(ctx, orc):
    dt, acc, gyr = imu_synth.samples(3.0, 3.1)
    got = ctx.imu_preintegrate(NOISE, [(dt[:1], acc[:1], gyr[:1])], np.zeros((1, 6)))[0]
    assert got.sum_dt == 0.0 and list(got.delta_q) == [1, 0, 0, 0]
    assert np.array_equal(np.array(got.jacobian).reshape(15, 15), np.eye(15))


def test_fused_match_matches_oracle(ctx, orc):
    import dliom
    w = workload()
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    problems, si_list, init_list, preints, wants = [], [], [], [], []
    for s in range(len(w["scans"])):
        t1 = w["times"][s]
        dt, acc, gyr = imu_synth.samples(t1 - 0.1, t1, noise=(3.99e-2, 1.56e-2), seed=10 + s)
        m = orc.imu_preintegrate(NOISE, [0, 0, 0], [0, 0, 0], dt, acc, gyr)
        si = imu_synth.state(t1 - 0.1)
        pred = orc.imu_predict(si, m)
        ing = orc.ingest_scan(w["opts"], w["scans"][s], w["origin"], si[:7], pred[:7])
        pts = ing["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        init = pred.copy()
        init[:3] += [0.03, -0.02, 0.01]
        for tw, rw, iw in ((0.0, 0.0, 1.0), (5.0, 4e2, 0.5)):
            want, ws = orc.fused_match([pts[hk], pts[lk]], [w["hi"], w["lo"]], [1.0, 6.0], tw, rw, init[:3], si, init, m,
                                       imu_weight=iw)
            got, gs = ctx.fused_match_batch([[pts[hk], pts[lk]]], [[hi, lo]], [1.0, 6.0], tw, rw, [w["submap_pose"]], [si],
                                            [init], [ctx.imu_preintegrate(NOISE, [(dt, acc, gyr)], np.zeros((1, 6)))[0]],
                                            imu_weight=iw)
            dtn, drn = pose_error(got[0][:7], want[:7])
            assert dtn < 1e-7 and drn < 1e-8, (s, dtn, drn)
            assert np.allclose(got[0][7:], want[7:], atol=1e-7)
            assert abs(gs[0]["final_cost"] - ws["final_cost"]) <= 1e-6 * max(1.0, ws["final_cost"])  # QR (oracle) vs normal equations
            assert gs[0]["num_iterations"] == ws["num_iterations"]


def test_fused_match_in_a_rotated_submap_frame(ctx, orc):
    """The solve is frame-invariant: the same problem posed against a submap whose local pose is not identity."""
    import dliom
    w = workload()
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    t1 = w["times"][0]
    dt, acc, gyr = imu_synth.samples(t1 - 0.1, t1)
    m = ctx.imu_preintegrate(NOISE, [(dt, acc, gyr)], np.zeros((1, 6)))[0]
    si = imu_synth.state(t1 - 0.1)
    pred = ctx.imu_predict(si, m)
    ing = orc.ingest_scan(w["opts"], w["scans"][0], w["origin"], si[:7], pred[:7])
    pts = ing["returns_tracking"]
    hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
    lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
    base, _ = ctx.fused_match_batch([[pts[hk], pts[lk]]], [[hi, lo]], [1.0, 6.0], 0.0, 0.0, [w["submap_pose"]], [si], [pred], [m])
    # move the world: submap local pose S, states mapped by S as well -> the answer must map by S too
    S = orc.angle_axis_pose((3.0, -2.0, 0.5), 0.3, (0, 0, 1))

    def move(x):
        from helpers import apply_pose
        y = x.copy()
        y[:3] = apply_pose(S, x[:3][None])[0]
        a, b = S[3:], x[3:7]
        y[3:7] = [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                  a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]
        y[7:10] = apply_pose(np.concatenate([[0, 0, 0], S[3:]]), x[7:10][None])[0]
        return y
    g = apply_pose_vec = None
    from helpers import apply_pose
    Gm = apply_pose(np.concatenate([[0, 0, 0], S[3:]]), np.array([[0, 0, 9.8]]))[0]
    moved, _ = ctx.fused_match_batch([[pts[hk], pts[lk]]], [[hi, lo]], [1.0, 6.0], 0.0, 0.0, [S], [move(si)], [move(pred)], [m],
                                     gravity=Gm)
    want = move(base[0])
    dtn, drn = pose_error(moved[0][:7], want[:7])
    assert dtn < 1e-6 and drn < 1e-7
    assert np.allclose(moved[0][7:10], want[7:10], atol=1e-6)


def test_frontend_batch_with_fused_imu_solve(ctx, orc):
    """dl_frontend_match_batch_imu = the batched front half (ingest, filters) + the 15-parameter fused solve. Checked against
    the oracle's own chain: ingest -> adaptive filters -> fused_scan_match seeded with the IMU prediction, per scan."""
    import dliom
    w = workload()
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    o = w["opts"]
    scans, states_i, preds, preints, wants = [], [], [], [], []
    for s in range(len(w["scans"])):
        t1 = w["times"][s]
        dt, acc, gyr = imu_synth.samples(t1 - 0.1, t1, noise=(3.99e-2, 1.56e-2), seed=30 + s)
        m = orc.imu_preintegrate(NOISE, [0, 0, 0], [0, 0, 0], dt, acc, gyr)
        si = imu_synth.state(t1 - 0.1)
        pred = orc.imu_predict(si, m)
        ing = orc.ingest_scan(o, w["scans"][s], w["origin"], si[:7], pred[:7])
        pts = ing["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, o.hi_max_length, o.hi_min_num_points, o.hi_max_range)
        lk, _ = orc.adaptive_voxel_filter(pts, o.lo_max_length, o.lo_min_num_points, o.lo_max_range)
        init = pred.copy()
        init[:7] = np.concatenate([ing["current_pose"][:3].astype(np.float64), ing["current_pose"][3:].astype(np.float64)])
        want, ws = orc.fused_match([pts[hk], pts[lk]], [w["hi"], w["lo"]], [o.occ_w0, o.occ_w1], o.trans_w, o.rot_w, init[:3], si,
                                   init, m, imu_weight=0.7, max_iter=o.max_iter)
        scans.append(w["scans"][s]); states_i.append(si); preds.append(pred); wants.append((want, ws))
        preints.append(ctx.imu_preintegrate(NOISE, [(dt, acc, gyr)], np.zeros((1, 6)))[0])
    res, states = ctx.frontend_match_batch_imu(fo, scans, w["origin"], states_i, preds, preints, w["submap_pose"], hi, lo,
                                               imu_weight=0.7)
    for r, x, (want, ws) in zip(res, states, wants):
        assert r.ok == 1
        dtn, drn = pose_error(x[:7], want[:7])
        assert dtn < 1e-6 and drn < 1e-7, (dtn, drn)
        assert np.allclose(x[7:], want[7:], atol=1e-6)
        assert np.allclose(np.array(r.pose_estimate_local[:]), x[:7], atol=1e-12)
        assert r.summary.num_iterations == ws["num_iterations"]


def _imu_batch(w, orc):
    intervals, states_i = [], []
    for s in range(len(w["scans"])):
        t1 = w["times"][s]
        intervals.append(imu_synth.samples(t1 - 0.1, t1, noise=(3.99e-2, 1.56e-2), seed=30 + s))
        states_i.append(imu_synth.state(t1 - 0.1, ba=(0.01, -0.02, 0.005), bg=(1e-3, -2e-3, 5e-4)))
    return intervals, states_i


def test_frontend_imu_samples_device_chain(ctx, orc):
    """dl_frontend_match_batch_imu_samples: raw samples in, pre-integration + prediction + deskew constants + information
    matrix + fused solve on the device. Against (a) the oracle chain and (b) the host-prepared chain of the same library."""
    import dliom
    w = workload()
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    o = w["opts"]
    intervals, states_i = _imu_batch(w, orc)
    imu = dliom.ImuSamples(NOISE, intervals, states_i, imu_weight=0.7)
    res, states, predicted = ctx.frontend_match_batch_imu_samples(fo, w["scans"], w["origin"], imu, w["submap_pose"], hi, lo)
    preds, preints = [], []
    for s in range(len(w["scans"])):
        dt, acc, gyr = intervals[s]
        si = states_i[s]
        m = orc.imu_preintegrate(NOISE, si[10:13], si[13:16], dt, acc, gyr)
        pred = orc.imu_predict(si, m)
        # prediction: no transcendental function involved -> the device reproduces the oracle to rounding
        assert np.allclose(predicted[s], pred, rtol=0, atol=1e-11)
        ing = orc.ingest_scan(o, w["scans"][s], w["origin"], si[:7], pred[:7])
        pts = ing["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, o.hi_max_length, o.hi_min_num_points, o.hi_max_range)
        lk, _ = orc.adaptive_voxel_filter(pts, o.lo_max_length, o.lo_min_num_points, o.lo_max_range)
        init = pred.copy()
        init[:7] = np.concatenate([ing["current_pose"][:3].astype(np.float64), ing["current_pose"][3:].astype(np.float64)])
        want, ws = orc.fused_match([pts[hk], pts[lk]], [w["hi"], w["lo"]], [o.occ_w0, o.occ_w1], o.trans_w, o.rot_w, init[:3], si,
                                   init, m, imu_weight=0.7, max_iter=o.max_iter)
        r, x = res[s], states[s]
        assert r.ok == 1
        assert r.num_returns == len(pts) and r.num_high_resolution == len(hk) and r.num_low_resolution == len(lk)
        dtn, drn = pose_error(x[:7], want[:7])
        assert dtn < 1e-6 and drn < 1e-7, (dtn, drn)
        assert np.allclose(x[7:], want[7:], atol=1e-6)
        assert np.allclose(np.array(r.pose_estimate_local[:]), x[:7], atol=1e-12)
        assert r.summary.num_iterations == ws["num_iterations"]
        preds.append(pred)
        preints.append(ctx.imu_preintegrate(NOISE, [(dt, acc, gyr)], np.array([si[10:16]]))[0])
    # (b) same library, factors prepared on the host from finished pre-integrations: the two chains share every operation
    res_h, states_h = ctx.frontend_match_batch_imu(fo, w["scans"], w["origin"], states_i, preds, preints, w["submap_pose"], hi, lo,
                                                   imu_weight=0.7)
    for s in range(len(w["scans"])):
        assert res_h[s].summary.num_iterations == res[s].summary.num_iterations
        assert np.allclose(states_h[s], states[s], rtol=0, atol=1e-9)


def test_frontend_imu_samples_streaming_and_device_resident(ctx, orc):
    import ctypes as C
    import dliom
    w = workload()
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    intervals, states_i = _imu_batch(w, orc)
    imu = dliom.ImuSamples(NOISE, intervals, states_i, imu_weight=0.7)
    res, states, _ = ctx.frontend_match_batch_imu_samples(fo, w["scans"], w["origin"], imu, w["submap_pose"], hi, lo)
    # streaming
    ctx.frontend_submit_imu_samples(fo, w["scans"], w["origin"], imu, w["submap_pose"], hi, lo)
    with pytest.raises(dliom.DlError):   # one batch in flight per context
        ctx.frontend_submit_imu_samples(fo, w["scans"], w["origin"], imu, w["submap_pose"], hi, lo)
    res_s, states_s = ctx.frontend_collect_imu()
    assert np.array_equal(states_s, states)
    assert all(list(a.pose_estimate_local) == list(b.pose_estimate_local) and a.ok == b.ok for a, b in zip(res_s, res))
    # device-resident scans, results and states
    n = len(w["scans"])
    sizes = np.array([len(s) for s in w["scans"]], np.int64)
    cap = int(sizes.max())
    rows = np.zeros((n, cap, 8), np.float32)
    for b, sc in enumerate(w["scans"]):
        rows[b, :len(sc)] = sc.view(np.float32).reshape(-1, 8)
    d_rows = ctx.device_alloc(rows.nbytes)
    d_res = ctx.device_alloc(n * C.sizeof(dliom.ScanResult))
    d_states = ctx.device_alloc(n * 16 * 8)
    ctx.copy_to_device(d_rows, rows)
    ctx.frontend_match_batch_imu_samples_dev(fo, imu, d_rows, cap, sizes, w["origin"], w["submap_pose"], hi, lo, d_res, d_states)
    res_d = ctx.fetch_results(d_res, n)
    states_d = np.zeros((n, 16))
    ctx.check(ctx.L.dl_copy_to_host(ctx.h, states_d.ctypes.data, d_states, states_d.nbytes))
    assert np.array_equal(states_d, states)
    assert all(list(a.pose_estimate_local) == list(b.pose_estimate_local) for a, b in zip(res_d, res))
    for p in (d_rows, d_res, d_states):
        ctx.device_free(p)


def test_frontend_imu_samples_without_samples_marks_scan(ctx, orc):
    """An interval without usable samples has no factor: ok = -2, the other scans of the batch are unaffected."""
    import dliom
    w = workload()
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    intervals, states_i = _imu_batch(w, orc)
    full = dliom.ImuSamples(NOISE, intervals, states_i, imu_weight=0.7)
    res, states, _ = ctx.frontend_match_batch_imu_samples(fo, w["scans"], w["origin"], full, w["submap_pose"], hi, lo)
    holed = list(intervals)
    holed[1] = (np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3)))
    imu = dliom.ImuSamples(NOISE, holed, states_i, imu_weight=0.7)
    res2, states2, _ = ctx.frontend_match_batch_imu_samples(fo, w["scans"], w["origin"], imu, w["submap_pose"], hi, lo)
    assert res2[1].ok == -2
    for s in (0, 2, 3):
        assert res2[s].ok == 1 and np.array_equal(states2[s], states[s])
    bad = dliom.ImuSamples(NOISE, intervals, states_i, imu_weight=0.7)
    bad.offsets[0] = 1
    with pytest.raises(dliom.DlError):
        ctx.frontend_match_batch_imu_samples(fo, w["scans"], w["origin"], bad, w["submap_pose"], hi, lo)
