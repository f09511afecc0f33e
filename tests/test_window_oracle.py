"""Oracle of the two-stage mode's second stage (oracle/orc_window.h; reference LocalTrajectoryBuilder3D::WindowOptimize,
LTB:693-863): the lag-one smoother = prior on the previous key + IMU factor with bias correction + matched pose prior
[+ gravity direction]. Properties any correct fusion must have; the device version is compared to it in tests/test_gpu_window.py.
GTSAM is not available to pin it further (parity unpinned for this row)."""
import numpy as np
import pytest

import imu_synth
from helpers import pose_error

NOISE = [3.99e-2, 1.56e-2, 6.4e-5, 3.6e-5]


def tight_prior(sp=1e-3, sr=1e-3, sv=1e-2, sb=1e-3):
    return np.diag([1 / sp ** 2] * 3 + [1 / sr ** 2] * 3 + [1 / sv ** 2] * 3 + [1 / sb ** 2] * 6)


def interval(orc, t0, t1, ba=(0, 0, 0), bg=(0, 0, 0), lin_ba=None, lin_bg=None):
    dt, acc, gyr = imu_synth.samples(t0, t1, ba=ba, bg=bg)
    return orc.imu_preintegrate(NOISE, ba if lin_ba is None else lin_ba, bg if lin_bg is None else lin_bg, dt, acc, gyr)


def test_consistent_inputs_reproduce_the_truth(orc):
    t0, t1 = 3.0, 3.1
    m = interval(orc, t0, t1)
    si, sj = imu_synth.state(t0), imu_synth.state(t1)
    xi, xj, info, s = orc.window_optimize(si, tight_prior(), m, sj[:7])
    dt, dr = pose_error(xj[:7], sj[:7])
    assert dt < 2e-4 and dr < 1e-5            # mid-point integration error of the synthetic interval, nothing else
    assert np.abs(xj[7:10] - sj[7:10]).max() < 2e-3 and np.abs(xj[10:]).max() < 1e-4
    assert s["termination"] == 0 and s["final_cost"] < 1e-2
    assert np.allclose(info, info.T, rtol=1e-9, atol=1e-6) and np.linalg.eigvalsh(0.5 * (info + info.T)).min() > 0
    # stationary: starting from the answer, Gauss-Newton has nothing to do
    _, xj2, _, s2 = orc.window_optimize(si, tight_prior(), m, sj[:7], initial_j=xj)
    assert s2["num_iterations"] <= 2 and pose_error(xj2[:7], xj[:7])[0] < 1e-9


def test_fusion_weighs_the_matched_pose_against_the_imu(orc):
    t0, t1 = 3.0, 3.1
    m = interval(orc, t0, t1)
    si, sj = imu_synth.state(t0), imu_synth.state(t1)
    z = sj[:7].copy()
    z[0] += 0.10                               # the scan matcher says 10 cm further along x than the IMU prediction
    pulls = []
    prior = tight_prior(sp=0.05, sv=0.5)       # previous key known to 5 cm / 0.5 m/s: the prediction is worth ~7 cm
    for sigma in (1.0, 0.05, 0.005):
        _, xj, info, _ = orc.window_optimize(si, prior, m, z, sigma_t=sigma)
        pulls.append(xj[0] - sj[0])
    assert 0 < pulls[0] < pulls[1] < pulls[2] < 0.10 + 1e-6      # between the two, closer to the matcher as its sigma shrinks
    assert pulls[0] < 0.01 and pulls[2] > 0.08
    # the carried information grows with a more certain matcher
    i_loose = orc.window_optimize(si, prior, m, z, sigma_t=1.0)[2]
    i_tight = orc.window_optimize(si, prior, m, z, sigma_t=0.005)[2]
    assert np.trace(i_tight[:3, :3]) > 10 * np.trace(i_loose[:3, :3])


def test_bias_correction_and_bias_observability(orc):
    """The interval is pre-integrated at WRONG linearisation biases; the first-order correction (integration_base.h:283-290)
    must absorb the difference when the prior knows the true biases."""
    t0, t1 = 3.0, 3.1
    ba, bg = np.array([0.05, -0.03, 0.02]), np.array([2e-3, -1e-3, 3e-3])
    m = interval(orc, t0, t1, ba=ba, bg=bg, lin_ba=(0, 0, 0), lin_bg=(0, 0, 0))     # measurements carry the biases
    si, sj = imu_synth.state(t0, ba=ba, bg=bg), imu_synth.state(t1, ba=ba, bg=bg)
    _, xj, _, s = orc.window_optimize(si, tight_prior(), m, sj[:7])
    dt, dr = pose_error(xj[:7], sj[:7])
    assert dt < 5e-4 and dr < 5e-5 and s["final_cost"] < 1e-9
    # without the prior's knowledge of the biases (zero-mean prior) the same data leave a visible residual
    s0 = imu_synth.state(t0)
    _, xj0, _, s_bad = orc.window_optimize(s0, tight_prior(), m, sj[:7])
    assert s_bad["final_cost"] > 1e-5


def test_gravity_factor_pulls_roll_and_pitch(orc):
    t0, t1 = 3.0, 3.1
    m = interval(orc, t0, t1)
    si, sj = imu_synth.state(t0), imu_synth.state(t1)
    tilt = np.array([np.cos(0.02), np.sin(0.02), 0, 0])            # matched pose with 0.04 rad of roll error
    z = sj[:7].copy()
    q = sj[3:7]
    z[3:] = [q[0] * tilt[0] - q[1] * tilt[1] - q[2] * tilt[2] - q[3] * tilt[3], q[0] * tilt[1] + q[1] * tilt[0] + q[2] * tilt[3] - q[3] * tilt[2],
             q[0] * tilt[2] - q[1] * tilt[3] + q[2] * tilt[0] + q[3] * tilt[1], q[0] * tilt[3] + q[1] * tilt[2] - q[2] * tilt[1] + q[3] * tilt[0]]
    loose = np.diag([1e6] * 3 + [1e-2] * 3 + [1e4] * 3 + [1e6] * 6)    # the prior says nothing about the attitude
    roll = lambda x: np.arctan2(2 * (x[3] * x[4] + x[5] * x[6]), 1 - 2 * (x[4] ** 2 + x[5] ** 2))
    _, without, _, _ = orc.window_optimize(si, loose, m, z, sigma_r=0.05)
    _, with_g, _, _ = orc.window_optimize(si, loose, m, z, sigma_r=0.05, gravity_factor=(1e-3, (0, 0, 1), (0, 0, 1)))
    assert abs(roll(with_g)) < 0.25 * abs(roll(without)) and abs(roll(without)) > 0.01


def test_not_positive_definite_inputs_are_refused(orc):
    m = interval(orc, 3.0, 3.1)
    with pytest.raises(RuntimeError):
        orc.window_optimize(imu_synth.state(3.0), -tight_prior(), m, imu_synth.state(3.1)[:7])
