"""Oracle for the sparse pose adjustment this fork's OptimizationProblem3D::Solve reduces to (oracle/orc_posegraph.h):
SPA residuals against finite differences and hand-computed cases, the reference's own test of Solve (a statistical property:
optimization_problem_3d_test.cc:106-196), exact recovery on consistent constraints, and the gauge the first submap fixes.
CPU only: the device side of this row is not built yet."""
import numpy as np
import pytest


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def qrot(q, v):
    return qmul(qmul(q, np.array([0.0, *v])), q * [1, -1, -1, -1])[1:]


def aa_to_q(v):
    """transform::AngleAxisVectorToRotationQuaternion."""
    v = np.asarray(v, np.float64)
    n = np.linalg.norm(v)
    return np.array([np.cos(n / 2), *(np.sin(n / 2) / n * v)]) if n * n > 1e-8 else np.array([1.0, *(0.5 * v)])


def compose(a, b):
    return np.array([*(qrot(a[3:], b[:3]) + a[:3]), *qmul(a[3:], b[3:])])


def inverse(a):
    qi = a[3:] * [1, -1, -1, -1]
    return np.array([*(-qrot(qi, a[:3])), *qi])


def angle(p):
    return 2 * np.arctan2(np.linalg.norm(p[4:7]), abs(p[3]))


def test_spa_residual_values_and_jacobian(orc):
    rng = np.random.default_rng(1)
    for _ in range(20):
        pi = np.array([*rng.uniform(-5, 5, 3), *aa_to_q(rng.uniform(-1, 1, 3))])
        pj = np.array([*rng.uniform(-5, 5, 3), *aa_to_q(rng.uniform(-1, 1, 3))])
        rel = compose(inverse(pi), pj)
        e, _ = orc.spa_residual(pi, pj, rel, 2.0, 3.0)
        assert np.abs(e).max() < 1e-12                                   # a consistent constraint has no error
        noise = np.array([*rng.uniform(-0.3, 0.3, 3), *aa_to_q(rng.uniform(-0.2, 0.2, 3))])
        z = compose(rel, noise)
        e, jac = orc.spa_residual(pi, pj, z, 2.0, 3.0)
        # by hand: translation error = (z.t - R_i^-1 (t_j - t_i)) * w_t; rotation error = angle-axis of (q_j^-1 q_i z.q) * w_r
        assert np.allclose(e[:3], (z[:3] - rel[:3]) * 2.0, atol=1e-12)
        d = qmul(qmul(pj[3:] * [1, -1, -1, -1], pi[3:]), z[3:])
        d = d if d[0] >= 0 else -d
        ang = 2 * np.arctan2(np.linalg.norm(d[1:]), d[0])
        assert np.allclose(e[3:], ang / np.sin(ang / 2) * d[1:] * 3.0, atol=1e-12)
        # ambient Jacobian (d / d [q_i t_i q_j t_j]) against central differences
        x0 = np.concatenate([pi[3:], pi[:3], pj[3:], pj[:3]])
        num = np.zeros((6, 14))
        for k in range(14):
            h = 1e-6
            for sgn in (1, -1):
                x = x0.copy(); x[k] += sgn * h
                ee, _ = orc.spa_residual(np.concatenate([x[4:7], x[0:4]]), np.concatenate([x[11:14], x[7:11]]), z, 2.0, 3.0)
                num[:, k] += sgn * ee / (2 * h)
        assert np.abs(jac - num).max() < 1e-6


def test_exact_recovery_from_consistent_constraints(orc):
    """Noise-free constraints from every submap to every node: the optimum reproduces the ground-truth geometry RELATIVE TO THE
    FIRST SUBMAP. The first submap keeps its translation and its yaw (ConstantYawQuaternionPlus) but may tilt: roll and pitch
    of the whole map are a gauge freedom here, because the IMU terms that pin gravity are commented out in this fork
    (optimization_problem_3d.cc:350-489)."""
    rng = np.random.default_rng(2)
    submaps = [np.array([0, 0, 0, 1.0, 0, 0, 0]), np.array([4.0, 1.0, 0.2, *aa_to_q([0, 0, 0.5])])]
    truth = [np.array([*rng.uniform(-8, 8, 3), *aa_to_q(rng.uniform(-0.6, 0.6, 3))]) for _ in range(12)]
    cons = [(s, n, compose(inverse(submaps[s]), truth[n]), 1.0, 1.0) for s in range(2) for n in range(12)]
    start_nodes = [compose(t, np.array([*rng.uniform(-0.5, 0.5, 3), *aa_to_q(rng.uniform(-0.2, 0.2, 3))])) for t in truth]
    start_submaps = [submaps[0], compose(submaps[1], np.array([0.3, -0.2, 0.1, *aa_to_q([0.02, -0.03, 0.1])]))]
    s_out, n_out, summary = orc.pose_graph_solve(start_submaps, start_nodes, cons)
    assert summary["final_cost"] < 1e-12 * max(summary["initial_cost"], 1.0) and summary["termination"] == 0
    assert np.allclose(s_out[0][:3], 0)                                  # constant block
    yaw = lambda q: np.arctan2(*qrot(q, [1.0, 0, 0])[[1, 0]])
    tilt_only = qmul(start_submaps[0][3:] * [1, -1, -1, -1], s_out[0][3:])   # q0^-1 * q0' = product of xy-plane rotations:
    assert abs(tilt_only[3]) < 1e-4 and abs(yaw(s_out[0][3:])) < 1e-4      # yaw-free to first order (second order: tilt^2)
    to_first = inverse(s_out[0])
    for got, want in zip(n_out, truth):                                  # submaps[0] is the identity: truth IS relative to it
        rel = compose(inverse(want), compose(to_first, got))
        assert np.linalg.norm(rel[:3]) < 1e-5 and angle(rel) < 1e-5
    rel = compose(inverse(submaps[1]), compose(to_first, s_out[1]))
    assert np.linalg.norm(rel[:3]) < 1e-5 and angle(rel) < 1e-5


def test_reference_reduces_noise(orc):
    """optimization_problem_3d_test.cc:106-196 (ReducesNoise): 100 noisy nodes, two equally noisy observations per node from
    submaps 0 and 1, one wildly wrong observation with weight 1e-9 from submap 2 (rotated by pi); after Solve the summed
    translation and rotation errors must be below 80 % of what they were. Same construction, numpy RNG."""
    rng = np.random.default_rng(0)

    def random_transform(ts, rs):
        return np.array([*rng.uniform(-ts, ts, 3), *aa_to_q(rng.uniform(-rs, rs, 3))])

    def random_yaw_only(ts, rs):
        return np.array([*rng.uniform(-ts, ts, 3), *aa_to_q([0, 0, rng.uniform(-rs, rs)])])

    def add_noise(t, noise):      # AddNoise: rotation noise.q * t.q, translation t.t + noise.t
        return np.array([*(t[:3] + noise[:3]), *qmul(noise[3:], t[3:])])

    n = 100
    truth = [random_transform(10.0, 3.0) for _ in range(n)]
    noise = [random_yaw_only(0.2, 0.3) for _ in range(n)]
    nodes = [add_noise(t, z) for t, z in zip(truth, noise)]
    submap2 = np.array([0, 0, 0, *aa_to_q([0, 0, np.pi])])
    cons = []
    for j in range(n):
        cons.append((0, j, add_noise(truth[j], noise[j]), 1.0, 1.0))
        cons.append((1, j, add_noise(truth[j], random_yaw_only(0.2, 0.3)), 1.0, 1.0))
        cons.append((2, j, compose(compose(inverse(submap2), truth[j]), random_transform(1e3, 3.0)), 1e-9, 1e-9))
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])

    def errors(ps):
        return (sum(np.linalg.norm(t[:3] - p[:3]) for t, p in zip(truth, ps)),
                sum(angle(compose(inverse(t), p)) for t, p in zip(truth, ps)))
    t_before, r_before = errors(nodes)
    _, n_out, summary = orc.pose_graph_solve([ident, ident, submap2], nodes, cons, max_iter=50)
    t_after, r_after = errors(n_out)
    assert 0.8 * t_before > t_after and 0.8 * r_before > r_after
    assert summary["final_cost"] < summary["initial_cost"]


def test_fix_z_keeps_heights(orc):
    """options.fix_z_in_3d: SubsetParameterization(3, {2}) on every free translation."""
    rng = np.random.default_rng(4)
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
    truth = [np.array([*rng.uniform(-5, 5, 3), *aa_to_q(rng.uniform(-0.3, 0.3, 3))]) for _ in range(6)]
    start = [t + np.array([0.2, -0.1, 0.0, 0, 0, 0, 0]) for t in truth]
    cons = [(0, k, truth[k], 1.0, 1.0) for k in range(6)]
    _, n_out, summary = orc.pose_graph_solve([ident], start, cons, fix_z=True)
    for a, b, t in zip(n_out, start, truth):
        assert a[2] == b[2] and np.allclose(a[:2], t[:2], atol=1e-6)
    # heights that disagree with the constraints stay where they are; the rest of the error is minimised around them
    lifted = [t + np.array([0.2, -0.1, 0.4, 0, 0, 0, 0]) for t in truth]
    _, n_out, summary = orc.pose_graph_solve([ident], lifted, cons, fix_z=True)
    assert all(a[2] == b[2] for a, b in zip(n_out, lifted)) and summary["final_cost"] < summary["initial_cost"]


def test_solution_is_a_local_minimum(orc):
    """Independent of the solver: at the returned poses no small perturbation of any node (translation or rotation) lowers
    the total SPA cost, and the cost equals the summary's final_cost."""
    rng = np.random.default_rng(8)
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
    submaps = [ident, np.array([2.0, -1.0, 0.3, *aa_to_q([0.05, -0.02, 0.7])])]
    truth = [np.array([*rng.uniform(-6, 6, 3), *aa_to_q(rng.uniform(-0.8, 0.8, 3))]) for _ in range(8)]
    cons = []
    for s in range(2):
        for k in range(8):
            z = compose(compose(inverse(submaps[s]), truth[k]), np.array([*rng.normal(0, 0.05, 3), *aa_to_q(rng.normal(0, 0.03, 3))]))
            cons.append((s, k, z, 1.0 + s, 3.0))
    s_out, n_out, summary = orc.pose_graph_solve(submaps, truth, cons)

    def cost(sub, nodes):
        return 0.5 * sum(float(np.sum(orc.spa_residual(sub[c[0]], nodes[c[1]], c[2], c[3], c[4])[0] ** 2)) for c in cons)
    base = cost(s_out, n_out)
    assert abs(base - summary["final_cost"]) <= 1e-12 * max(1.0, base) and summary["termination"] == 0
    for _ in range(60):
        k = rng.integers(8)
        moved = [p.copy() for p in n_out]
        if rng.random() < 0.5:
            moved[k][:3] += rng.normal(0, 1e-3, 3)
        else:
            moved[k][3:] = qmul(aa_to_q(rng.normal(0, 1e-3, 3)), moved[k][3:])
        assert cost(s_out, moved) >= base - 1e-12


def test_normal_cholesky_and_dense_qr_agree(orc):
    """The two linear solvers behind the same trust-region loop (the reference's SPARSE_NORMAL_CHOLESKY forms and factors the
    normal equations; the scan matcher's DENSE_QR factors the augmented Jacobian) take the same steps: same iteration
    counts, poses equal to 1e-9."""
    rng = np.random.default_rng(12)
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
    submaps = [ident, np.array([3.0, 2.0, -0.1, *aa_to_q([0.01, 0.02, -0.4])]), np.array([-2.0, 1.0, 0.2, *aa_to_q([0, 0, 2.0])])]
    truth = [np.array([*rng.uniform(-8, 8, 3), *aa_to_q(rng.uniform(-1, 1, 3))]) for _ in range(15)]
    cons = [(s, k, compose(compose(inverse(submaps[s]), truth[k]), np.array([*rng.normal(0, 0.1, 3), *aa_to_q(rng.normal(0, 0.05, 3))])),
             1.0, 2.0) for s in range(3) for k in range(15) if (s + k) % 2 == 0 or s == 0]
    start = [compose(t, np.array([*rng.normal(0, 0.3, 3), *aa_to_q(rng.normal(0, 0.1, 3))])) for t in truth]
    sa, na, qa = orc.pose_graph_solve(submaps, start, cons, linear_solver="dense_qr")
    sb, nb, qb = orc.pose_graph_solve(submaps, start, cons, linear_solver="normal_cholesky")
    assert qa["num_iterations"] == qb["num_iterations"] and qa["termination"] == qb["termination"] == 0
    assert abs(qa["final_cost"] - qb["final_cost"]) <= 1e-12 * max(1.0, qa["final_cost"])
    assert np.abs(na - nb).max() < 1e-9 and np.abs(sa - sb).max() < 1e-9
