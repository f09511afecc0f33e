"""Device fixed-lag smoother (csrc/dl_window.cu, dl_window_optimize_batch) against its oracle (oracle/orc_window.h): the device
differentiates with forward-mode duals, the oracle with central differences, so agreement to ~1e-7 checks both; plus a chained
run where the carried information of one call is the prior of the next, on both sides."""
import numpy as np
import pytest

import imu_synth
from helpers import pose_error

pytestmark = pytest.mark.gpu
NOISE = [3.99e-2, 1.56e-2, 6.4e-5, 3.6e-5]


@pytest.fixture(scope="module")
def ctx():
    import dliom
    c = dliom.Context(0)
    yield c
    c.close()


def prior(sp=0.05, sr=0.01, sv=0.5, sb=1e-2):
    return np.diag([1 / sp ** 2] * 3 + [1 / sr ** 2] * 3 + [1 / sv ** 2] * 3 + [1 / sb ** 2] * 6)


def problems(orc, ctx, n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for k in range(n):
        t0 = 2.0 + 0.37 * k
        ba, bg = rng.normal(0, 2e-2, 3), rng.normal(0, 2e-3, 3)
        lin_ba, lin_bg = ba + rng.normal(0, 5e-3, 3), bg + rng.normal(0, 5e-4, 3)     # linearised a little off the truth
        dt, acc, gyr = imu_synth.samples(t0, t0 + 0.1, ba=ba, bg=bg, noise=(3.99e-2, 1.56e-2), seed=seed + k)
        m_o = orc.imu_preintegrate(NOISE, lin_ba, lin_bg, dt, acc, gyr)
        m_d = ctx.imu_preintegrate(NOISE, [(dt, acc, gyr)], np.array([np.concatenate([lin_ba, lin_bg])]))[0]
        si = imu_synth.state(t0, ba=lin_ba, bg=lin_bg)
        si[:3] += rng.uniform(-0.03, 0.03, 3)
        si[7:10] += rng.uniform(-0.2, 0.2, 3)
        z = imu_synth.state(t0 + 0.1)[:7]
        z[:3] += rng.uniform(-0.03, 0.03, 3)
        A = rng.normal(0, 1, (15, 15))
        info = prior() + 0.05 * (A @ A.T) * np.outer(np.sqrt(np.diag(prior())), np.sqrt(np.diag(prior()))) / 15   # correlated, SPD
        out.append((si, info, m_o, m_d, z))
    return out


def test_device_matches_oracle(ctx, orc):
    ps = problems(orc, ctx, 9, 3)
    for gf in (None, (5e-2, (0.01, -0.02, 1.0), (0, 0, 1))):
        wi, wj, winfo, ws = zip(*[orc.window_optimize(si, info, m_o, z, gravity_factor=gf) for si, info, m_o, _, z in ps])
        gi, gj, ginfo, gs = ctx.window_optimize_batch([p[0] for p in ps], [p[1] for p in ps], [p[3] for p in ps], [p[4] for p in ps],
                                                      gravity_factor=gf)
        for k in range(len(ps)):
            assert gs[k]["termination"] == 0 and ws[k]["termination"] == 0
            assert gs[k]["num_iterations"] == ws[k]["num_iterations"]
            for a, b in ((gj[k], wj[k]), (gi[k], wi[k])):
                dt, dr = pose_error(a[:7], b[:7])
                assert dt < 1e-7 and dr < 1e-7 and np.abs(a[7:] - b[7:]).max() < 1e-7
            assert abs(gs[k]["final_cost"] - ws[k]["final_cost"]) <= 1e-6 * max(ws[k]["final_cost"], 1e-6)
            scale = np.sqrt(np.outer(np.diag(winfo[k]), np.diag(winfo[k])))
            assert np.abs(ginfo[k] - winfo[k]).max() <= 1e-5 * scale.max() and (np.abs(ginfo[k] - winfo[k]) / scale).max() < 1e-4
            assert np.allclose(ginfo[k], ginfo[k].T, rtol=1e-9, atol=1e-7 * scale.max())


def test_chained_steps_carry_the_information(ctx, orc):
    """Ten scans of one trajectory: estimate and information of step k are the prior of step k + 1 (what replaces iSAM2's marginal
    between the reference's re-seeds, LTB:750-797). Device and oracle chains run independently and must stay together."""
    rng = np.random.RandomState(11)
    t = 2.0
    so = sd = imu_synth.state(t)
    io = idv = prior(sp=0.01, sr=0.005, sv=0.1, sb=1e-2)
    for k in range(10):
        dt, acc, gyr = imu_synth.samples(t, t + 0.1, noise=(3.99e-2, 1.56e-2), seed=50 + k)
        z = imu_synth.state(t + 0.1)[:7]
        z[:3] += rng.normal(0, 0.01, 3)
        m_o = orc.imu_preintegrate(NOISE, so[10:13], so[13:16], dt, acc, gyr)
        m_d = ctx.imu_preintegrate(NOISE, [(dt, acc, gyr)], np.array([sd[10:16]]))[0]
        _, so, io, s_o = orc.window_optimize(so, io, m_o, z, sigma_t=0.02, sigma_r=0.005)
        _, sdj, idj, s_d = ctx.window_optimize_batch([sd], [idv], [m_d], [z], sigma_t=0.02, sigma_r=0.005)
        sd, idv = sdj[0], idj[0]
        assert s_d[0]["termination"] == 0
        dtn, drn = pose_error(sd[:7], so[:7])
        assert dtn < 1e-6 and drn < 1e-6 and np.abs(sd[7:] - so[7:]).max() < 1e-5
        truth = imu_synth.state(t + 0.1)
        assert pose_error(sd[:7], truth[:7])[0] < 0.05
        t += 0.1
    assert np.linalg.eigvalsh(0.5 * (idv + idv.T)).min() > 0


def test_argument_and_definiteness_errors(ctx, orc):
    import dliom
    ps = problems(orc, ctx, 2, 5)
    bad = [-ps[0][1], ps[1][1]]                     # first trajectory: prior information not positive definite
    gi, gj, ginfo, gs = ctx.window_optimize_batch([p[0] for p in ps], bad, [p[3] for p in ps], [p[4] for p in ps])
    assert gs[0]["termination"] == 2 and not gj[0].any() and gs[1]["termination"] == 0
    with pytest.raises(dliom.DlError):
        ctx.window_optimize_batch([ps[0][0]], [ps[0][1]], [ps[0][3]], [ps[0][4]], sigma_t=0.0)
