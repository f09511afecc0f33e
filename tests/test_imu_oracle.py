"""CPU checks of the IMU oracle (oracle/orc_imu.h): the pre-integrator restated from integration_base.h, the 15-dim
residual of integration_base.h:267-301 and its Jacobian, and the fused scan-match + IMU solve. The reference has no
runnable test for this row, so these are consistency / known-physics checks."""
import numpy as np

import imu_synth
from helpers import pose_error, workload

NOISE = [3.99e-2, 1.56e-2, 6.4e-5, 3.6e-5]   # D/config/kaist.lua:38-43


def test_preintegration_predicts_the_analytic_trajectory(orc):
    t0, t1 = 3.0, 3.1
    dt, acc, gyr = imu_synth.samples(t0, t1)
    m = orc.imu_preintegrate(NOISE, [0, 0, 0], [0, 0, 0], dt, acc, gyr)
    assert abs(m.sum_dt - 0.1) < 1e-12
    pred = orc.imu_predict(imu_synth.state(t0), m)
    truth = imu_synth.state(t1)
    assert np.linalg.norm(pred[:3] - truth[:3]) < 1e-5       # mid-point rule on a smooth 0.1 s arc
    assert np.linalg.norm(pred[7:10] - truth[7:10]) < 1e-5
    assert pose_error(pred[:7], truth[:7])[1] < 1e-7
    r, _ = orc.imu_residual(imu_synth.state(t0), truth, m)
    assert np.max(np.abs(r)) < 1e-5


def test_preintegration_first_sample_only_latches(orc):
    dt, acc, gyr = imu_synth.samples(3.0, 3.1)
    m1 = orc.imu_preintegrate(NOISE, [0, 0, 0], [0, 0, 0], dt[:1], acc[:1], gyr[:1])
    assert m1.sum_dt == 0.0 and list(m1.delta_q) == [1, 0, 0, 0]
    J = np.array(m1.jacobian).reshape(15, 15)
    assert np.array_equal(J, np.eye(15)) and not np.any(np.array(m1.covariance))


def test_covariance_symmetric_positive_definite_and_bias_jacobians(orc):
    dt, acc, gyr = imu_synth.samples(3.0, 3.1)
    m = orc.imu_preintegrate(NOISE, [0.02, -0.01, 0.03], [1e-3, 2e-3, -1e-3], dt, acc, gyr)
    P = np.array(m.covariance).reshape(15, 15)
    assert np.allclose(P, P.T, atol=1e-18) and np.linalg.eigvalsh(P).min() > 0
    # d(delta_p, delta_v)/d(ba), d(.)/d(bg) from the propagated Jacobian vs re-integration with a perturbed bias
    J = np.array(m.jacobian).reshape(15, 15)
    for col, (dba, dbg) in enumerate([((1e-4, 0, 0), (0, 0, 0)), ((0, 0, 0), (0, 1e-5, 0))]):
        m2 = orc.imu_preintegrate(NOISE, np.array([0.02, -0.01, 0.03]) + dba, np.array([1e-3, 2e-3, -1e-3]) + dbg,
                                  dt, acc, gyr)
        d = np.concatenate([dba, dbg])
        lin_p = J[0:3, 9:15] @ d
        lin_v = J[6:9, 9:15] @ d
        assert np.allclose(np.array(m2.delta_p) - np.array(m.delta_p), lin_p, rtol=2e-2, atol=1e-10)
        assert np.allclose(np.array(m2.delta_v) - np.array(m.delta_v), lin_v, rtol=2e-2, atol=1e-10)


def test_residual_jacobian_matches_finite_differences(orc):
    dt, acc, gyr = imu_synth.samples(3.0, 3.1)
    m = orc.imu_preintegrate(NOISE, [0, 0, 0], [0, 0, 0], dt, acc, gyr)
    si = imu_synth.state(3.0)
    sj = imu_synth.state(3.1)
    sj[:3] += [0.02, -0.01, 0.005]
    sj[7:10] += [0.05, 0.02, -0.01]
    sj[10:] = [1e-3, -2e-3, 5e-4, 1e-4, 2e-4, -1e-4]
    r0, J = orc.imu_residual(si, sj, m)

    def plus(x, d):
        y = x.copy()
        y[:3] += d[:3]
        n = np.linalg.norm(d[3:6])
        dq = np.array([1.0, 0, 0, 0]) if n == 0 else np.concatenate([[np.cos(n)], np.sin(n) / n * d[3:6]])
        a, b = dq, x[3:7]
        y[3:7] = [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                  a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]
        y[7:] += d[6:]
        return y
    h = 1e-6
    for c in range(15):
        e = np.zeros(15)
        e[c] = h
        rp, _ = orc.imu_residual(si, plus(sj, e), m, jacobian=False)
        rm, _ = orc.imu_residual(si, plus(sj, -e), m, jacobian=False)
        assert np.allclose((rp - rm) / (2 * h), J[:, c], atol=1e-6), c


def fused_inputs(orc, w, s):
    """State i = truth one scan earlier, pre-integration over the scan, state j initial = IMU prediction."""
    import synth
    t1 = w["times"][s]
    t0 = t1 - 0.1
    dt, acc, gyr = imu_synth.samples(t0, t1)
    m = orc.imu_preintegrate(NOISE, [0, 0, 0], [0, 0, 0], dt, acc, gyr)
    si = imu_synth.state(t0)
    pred = orc.imu_predict(si, m)
    ing = orc.ingest_scan(w["opts"], w["scans"][s], w["origin"], si[:7], pred[:7])
    pts = ing["returns_tracking"]
    hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
    lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
    return si, pred, m, [pts[hk], pts[lk]]


def test_fused_solve_stays_consistent_with_imu_and_scan(orc):
    w = workload()
    si, pred, m, clouds = fused_inputs(orc, w, 0)
    init = pred.copy()
    init[:3] += [0.03, -0.02, 0.01]       # a wrong initial guess the two information sources must pull back
    out, s = orc.fused_match(clouds, [w["hi"], w["lo"]], [1.0, 6.0], 0.0, 0.0, init[:3], si, init, m, imu_weight=1.0)
    assert s["final_cost"] < s["initial_cost"]
    truth = imu_synth.state(w["times"][0])
    assert np.linalg.norm(out[:3] - truth[:3]) < np.linalg.norm(init[:3] - truth[:3])
    # with an overwhelming IMU weight the estimate collapses onto the IMU prediction
    out2, _ = orc.fused_match(clouds, [w["hi"], w["lo"]], [1.0, 6.0], 0.0, 0.0, init[:3], si, init, m, imu_weight=1e3,
                              max_iter=50)
    assert np.linalg.norm(out2[:3] - pred[:3]) < 1e-3 and np.linalg.norm(out2[7:10] - pred[7:10]) < 1e-2


def test_frontend_batch_imu_equals_the_stepwise_chain():
    """orc_frontend_batch_imu (the pooled CPU baseline of the IMU-coupled front end) = pre-integrate -> predict -> ingest ->
    adaptive filters -> fused match done call by call; identical for 1 and many threads (work queue, no shared state)."""
    import orc
    from helpers import workload
    import imu_synth
    w = workload()
    o = w["opts"]
    noise = [3.99e-2, 1.56e-2, 6.4e-5, 3.6e-5]
    intervals, states_i = [], []
    for s in range(len(w["scans"])):
        t1 = w["times"][s]
        intervals.append(imu_synth.samples(t1 - 0.1, t1, noise=(3.99e-2, 1.56e-2), seed=30 + s))
        states_i.append(imu_synth.state(t1 - 0.1, ba=(0.01, -0.02, 0.005), bg=(1e-3, -2e-3, 5e-4)))
    secs, states, pred, ok, iters = orc.frontend_batch_imu(o, w["scans"], w["origin"], noise, states_i, intervals, w["submap_pose"],
                                                           w["hi"], w["lo"], 3, imu_weight=0.7)
    _, states1, _, ok1, _ = orc.frontend_batch_imu(o, w["scans"], w["origin"], noise, states_i, intervals, w["submap_pose"],
                                                   w["hi"], w["lo"], 1, imu_weight=0.7)
    assert np.array_equal(states, states1) and np.array_equal(ok, ok1) and all(ok == 1) and secs > 0
    for s in range(len(w["scans"])):
        dt, acc, gyr = intervals[s]
        si = states_i[s]
        m = orc.imu_preintegrate(noise, si[10:13], si[13:16], dt, acc, gyr)
        p = orc.imu_predict(si, m)
        assert np.array_equal(p, pred[s])
        ing = orc.ingest_scan(o, w["scans"][s], w["origin"], si[:7], p[:7])
        pts = ing["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, o.hi_max_length, o.hi_min_num_points, o.hi_max_range)
        lk, _ = orc.adaptive_voxel_filter(pts, o.lo_max_length, o.lo_min_num_points, o.lo_max_range)
        init = p.copy()
        init[:7] = ing["current_pose"].astype(np.float64)
        want, ws = orc.fused_match([pts[hk], pts[lk]], [w["hi"], w["lo"]], [o.occ_w0, o.occ_w1], o.trans_w, o.rot_w, init[:3], si,
                                   init, m, imu_weight=0.7, max_iter=o.max_iter)
        # the batch chain renormalises the float-cast pose when it composes it with the submap pose (as the device does)
        from helpers import pose_error
        dtn, drn = pose_error(states[s][:7], want[:7])
        assert dtn < 1e-9 and drn < 1e-8 and np.allclose(states[s][7:], want[7:], rtol=0, atol=1e-9)
        assert iters[s] == ws["num_iterations"]
    # an empty interval has no factor
    holed = list(intervals)
    holed[2] = (np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3)))
    _, _, _, ok2, _ = orc.frontend_batch_imu(o, w["scans"], w["origin"], noise, states_i, holed, w["submap_pose"], w["hi"], w["lo"], 2)
    assert ok2.tolist() == [1, 1, -2, 1]
