"""Synthetic 200 Hz IMU for the analytic trajectory of tools/synth.py (specific force = R^T (a + G), G = +9.8 z,
the convention of the reference's in-repo integrator, integration_base.h:292-297). Noise-free unless asked."""
import numpy as np

G = np.array([0.0, 0.0, 9.8])


def _pos(t):
    return np.array([10.0 * t, 0.6 * np.sin(0.3 * t), 0.0])


def _vel(t):
    return np.array([10.0, 0.18 * np.cos(0.3 * t), 0.0])


def _acc(t):
    return np.array([0.0, -0.054 * np.sin(0.3 * t), 0.0])


def _yaw(t):
    return 0.16 * np.sin(0.5 * t)


def _yaw_rate(t):
    return 0.08 * np.cos(0.5 * t)


def state(t, ba=(0, 0, 0), bg=(0, 0, 0)):
    """16-vector nav state (p, q wxyz, v, ba, bg) of the tracking frame at time t."""
    y = _yaw(t)
    return np.concatenate([_pos(t), [np.cos(y / 2), 0, 0, np.sin(y / 2)], _vel(t), ba, bg])


def samples(t0, t1, rate=200.0, ba=(0, 0, 0), bg=(0, 0, 0), noise=None, seed=44):
    """IMU samples covering [t0, t1]: returns (dt[n], acc[n,3], gyr[n,3]); sample 0 is at t0 and only latches the
    integrator's acc_0 / gyr_0, sample k integrates over [t_{k-1}, t_k]."""
    n = int(round((t1 - t0) * rate))
    ts = t0 + np.arange(n + 1) / rate
    dt = np.full(n + 1, 1.0 / rate)
    acc, gyr = np.zeros((n + 1, 3)), np.zeros((n + 1, 3))
    for k, t in enumerate(ts):
        y = _yaw(t)
        c, s = np.cos(y), np.sin(y)
        Rt = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])
        acc[k] = Rt @ (_acc(t) + G) + np.asarray(ba)
        gyr[k] = np.array([0, 0, _yaw_rate(t)]) + np.asarray(bg)
    if noise is not None:
        rng = np.random.RandomState(seed)
        acc += rng.normal(0, noise[0], acc.shape)
        gyr += rng.normal(0, noise[1], gyr.shape)
    return dt, acc, gyr
