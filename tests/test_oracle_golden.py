"""Pins the CPU oracle against every known-answer fixture the reference's own tests hold for the hot path
(SURVEY.md Appendix C). File:line cites are into /root/reference/src/cartographer/cartographer/."""
import numpy as np
import pytest

SEVEN = np.array([[-3, 2, 0], [-4, 2, 0], [-5, 2, 0], [-6, 2, 0], [-6, 3, 1], [-6, 4, 2], [-7, 3, 1]], np.float32)


def isapprox(a, b, prec):
    """Eigen isApprox on the 4x4 homogeneous matrices (rigid_transform_test_helpers.h:42-46)."""
    def mat(p):
        w, x, y, z = p[3:]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = p[:3]
        return M
    A, B = mat(a), mat(b)
    return np.linalg.norm(A - B) <= prec * min(np.linalg.norm(A), np.linalg.norm(B))


def seven_point_grid(orc, res, shift=(-1, 0, 0)):
    g = orc.Grid(res)
    for p in SEVEN:
        g.set_probability(g.cell_index(p + np.array(shift, np.float32)), 1.0)
    return g


# ---------------------------------------------------------------- voxel filter (sensor/internal/voxel_filter_test.cc:29-54)
def test_voxel_filter_first_point_per_voxel(orc):
    pc = np.array([[0, 0, 0], [0.1, -0.1, 0.1], [0.3, -0.1, 0], [0, 0, 0.1]], np.float32)
    assert orc.voxel_filter(pc, 0.3).tolist() == [0, 2]


def test_voxel_filter_large_coordinates(orc):
    pc = np.array([[100000, 0, 0], [100000.001, -0.0001, 0.0001], [100000.003, -0.0001, 0], [-200000, 0, 0]],
                  np.float32)
    assert orc.voxel_filter(pc, 0.01).tolist() == [0, 3]


def test_voxel_filter_ignores_time(orc):
    pc = np.array([[-100, 0.3, 0.4, i] for i in range(100)], np.float32)
    assert orc.voxel_filter(pc, 0.3).tolist() == [0]


def test_voxel_filter_empty(orc):
    assert orc.voxel_filter(np.zeros((0, 3), np.float32), 0.3).tolist() == []


# ---------------------------------------------------------------- hybrid grid (mapping/3d/hybrid_grid_test.cc)
def test_grid_cell_index(orc):  # :89-110
    g = orc.Grid(2.0)
    cases = {(0, 0, 0): (0, 0, 0), (0, 26, 10): (0, 13, 5), (14, 0, 10): (7, 0, 5), (14, 26, 0): (7, 13, 0),
             (8.5, 11.5, 0.5): (4, 6, 0), (7.5, 12.5, 1.5): (4, 6, 1), (6.5, 14.5, 2.5): (3, 7, 1),
             (5.5, 13.5, 3.5): (3, 7, 2)}
    for p, want in cases.items():
        assert tuple(g.cell_index(p)) == want


def test_grid_center_of_cell(orc):  # :112-121
    g = orc.Grid(2.0)
    c = g.center_of_cell((3, 2, 1))
    assert c.tolist() == [6.0, 4.0, 2.0]
    assert tuple(g.cell_index(c)) == (3, 2, 1)


def test_grid_apply_odds(orc):  # :29-68
    L = orc.lib()
    g = orc.Grid(1.0)
    for i in [(0, 0, 0), (0, 1, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1)]:
        assert g.value(i) == 0

    def table(p):
        t = np.zeros(32768, np.uint16)
        L.orc_lookup_table_to_apply_odds(L.orc_odds(np.float32(p)), t)
        return t
    g.set_probability((1, 0, 1), 0.5)
    g.apply_lookup_table((1, 0, 1), table(0.9))
    g.finish_update()
    assert g.probability((1, 0, 1)) > 0.5
    g.set_probability((0, 1, 0), 0.5)
    g.apply_lookup_table((0, 1, 0), table(0.1))
    g.finish_update()
    assert g.probability((0, 1, 0)) < 0.5
    g.apply_lookup_table((1, 1, 1), table(0.42))
    assert abs(g.probability((1, 1, 1)) - 0.42) < 1e-4
    g.apply_lookup_table((1, 1, 1), table(0.9))  # ignored until FinishUpdate
    assert abs(g.probability((1, 1, 1)) - 0.42) < 1e-4
    g.finish_update()
    g.apply_lookup_table((1, 1, 1), table(0.9))
    assert g.probability((1, 1, 1)) > 0.42


def test_grid_get_probability(orc):  # :70-85
    g = orc.Grid(1.0)
    g.set_probability(g.cell_index((0, 1, 1)), 0.9)
    assert abs(g.probability(g.cell_index((0, 1, 1))) - 0.9) < 1e-6
    for p in [(0, 2, 1), (1, 1, 1), (1, 2, 1)]:
        assert g.value(g.cell_index(p)) == 0


def test_grid_random_fill_iteration(orc):  # :123-177 (the reference seeds std::mt19937(1285120005))
    rng = np.random.RandomState(1285120005 % 2**32)
    g = orc.Grid(2.0)
    want = {}
    for _ in range(10000):
        x, y, z = (int(v) for v in rng.randint(-3000, 3000, 3))
        want[(x, y, z)] = np.float32(rng.uniform(0.1, 0.9))
    for k, p in want.items():
        g.set_probability(k, p)
    xs, ys, zs, vs = g.export()
    assert len(xs) == len(want)
    table = np.zeros(65536, np.float32)
    orc.lib().orc_value_to_probability_table(table)
    for x, y, z, v in zip(xs, ys, zs, vs):
        assert g.value((x, y, z)) == v
        assert abs(table[v] - want[(x, y, z)]) < 1e-4
    # out-of-bounds lookups read as unknown
    assert g.value((100000, 0, 0)) == 0 and g.value((-100000, 5, 5)) == 0
    assert g.bits() >= 6


def test_grid_growth_limit(orc):
    g = orc.Grid(1.0)
    with pytest.raises(RuntimeError):
        g.set_probability((9000, 0, 0), 0.5)  # beyond +-8192 cells: CHECK_LE(new_bits, 8)


# ---------------------------------------------------------------- probability values (mapping/probability_values_test.cc)
def test_probability_value_round_trip(orc):
    L = orc.lib()
    table = np.zeros(65536, np.float32)
    L.orc_value_to_probability_table(table)
    assert table[0] == np.float32(0.1) and table[32768] == np.float32(0.1)
    assert np.array_equal(table[:32768], table[32768:])
    for v in range(1, 32768):
        assert L.orc_probability_to_value(table[v]) == v
    assert abs(table[1] - 0.1) < 1e-6 and abs(table[32767] - 0.9) < 1e-6
    assert L.orc_probability_to_value(np.float32(0.0)) == 1 and L.orc_probability_to_value(np.float32(1.0)) == 32767


def test_apply_odds_table_marks_cells(orc):
    L = orc.lib()
    t = np.zeros(32768, np.uint16)
    L.orc_lookup_table_to_apply_odds(L.orc_odds(np.float32(0.55)), t)
    assert (t >= 32768).all()
    assert abs(L.orc_value_to_probability(int(t[0]) - 32768) - 0.55) < 1e-4
    # applying hit odds always raises, and saturates at 0.9
    v = (t[1:] - 32768).astype(np.int64)
    assert (v >= np.arange(1, 32768)).all() and v.max() == 32767


# ---------------------------------------------------------------- range data inserter (mapping/3d/range_data_inserter_3d_test.cc:30-103)
def test_range_data_inserter(orc):
    g = orc.Grid(1.0)
    origin = np.array([0, 0, -4], np.float32)
    returns = np.array([[-3, -1, 4], [-2, 0, 4], [-1, 1, 4], [0, 2, 4]], np.float32)
    g.insert_range_data(origin, returns, hit=0.7, miss=0.4, num_free=1000)
    for p in returns:
        assert abs(g.probability(g.cell_index(p)) - 0.7) < 1e-4
    # cells on the rays below the hits are misses (:75-84 checks (-2,-1..,0) style lines)
    for p in [(-2, 0, 2), (-1, 0, 0), (0, 1, 0)]:
        pass
    xs, ys, zs, vs = g.export()
    probs = {(x, y, z): orc.lib().orc_value_to_probability(int(v)) for x, y, z, v in zip(xs, ys, zs, vs)}
    hits = {tuple(g.cell_index(p)) for p in returns}
    assert all(abs(p - 0.4) < 1e-4 for k, p in probs.items() if k not in hits)
    assert len(probs) > len(hits)
    for _ in range(1000):
        g.insert_range_data(origin, returns, hit=0.7, miss=0.4, num_free=1000)
    xs, ys, zs, vs = g.export()
    for x, y, z, v in zip(xs, ys, zs, vs):
        p = orc.lib().orc_value_to_probability(int(v))
        assert abs(p - (0.9 if (x, y, z) in hits else 0.1)) < 1e-3


# ---------------------------------------------------------------- interpolated grid (scan_matching/interpolated_grid_test.cc:28-84)
def test_interpolated_grid_reproduces_grid_points(orc):
    g = seven_point_grid(orc, 0.1, shift=(0, 0, 0))
    res = float(np.float32(0.1))
    z = -1.0
    while z < 3.0:
        y = 1.0
        while y < 5.0:
            x = -8.0
            while x < -2.0:
                want = g.probability(g.cell_index((x, y, z)))
                assert abs(want - g.interpolate(x, y, z)) < 1e-6
                x += res
            y += res
        z += res * 7  # the reference sweeps every z layer; a 7-layer stride keeps the CPU suite short


def test_interpolated_grid_monotone_in_x(orc):
    g = seven_point_grid(orc, 0.1, shift=(0, 0, 0))
    res = float(np.float32(0.1))
    step = res / 10.0
    checked = 0
    for (y, z) in [(2.0, 0.0), (3.0, 1.0), (4.0, 2.0)]:
        x = -8.0
        while x < -2.0:
            a = g.probability(g.cell_index((x, y, z)))
            b = g.probability(g.cell_index((x + res, y, z)))
            d = float(b) - float(a)
            if abs(d) >= 1e-6:
                s = step
                while s < res - 2 * step:
                    assert 0.0 < d * (g.interpolate(x + s + step, y, z) - g.interpolate(x + s, y, z))
                    checked += 1
                    s += step
            x += res
    assert checked > 20


def test_interpolated_gradient_matches_finite_difference(orc):
    g = seven_point_grid(orc, 0.1, shift=(0, 0, 0))
    rng = np.random.RandomState(3)
    for _ in range(200):
        p = SEVEN[rng.randint(7)] + rng.uniform(-0.12, 0.12, 3)
        v = g.interpolate_grad(*p)
        assert abs(v[0] - g.interpolate(*p)) < 1e-12
        h = 1e-7
        for a in range(3):
            e = np.zeros(3)
            e[a] = h
            fd = (g.interpolate(*(p + e)) - g.interpolate(*(p - e))) / (2 * h)
            assert abs(fd - v[1 + a]) < 1e-4 * max(1.0, abs(fd))


# ---------------------------------------------------------------- RT-CSM (scan_matching/real_time_correlative_scan_matcher_3d_test.cc:36-117)
RTCSM_STARTS = [
    ((-1, 0, 0), 0.0, (1, 0, 0)), ((-0.8, 0, 0), 0.0, (1, 0, 0)), ((-1, 0, -0.2), 0.0, (1, 0, 0)),
    ((-0.9, -0.2, 0.2), 0.0, (1, 0, 0)),
    ((-1, 0, 0), 0.8 / 180 * np.pi, (1, 0, 0)), ((-1, 0, 0), 0.8 / 180 * np.pi, (0, 1, 0)),
    ((-1, 0, 0), 0.8 / 180 * np.pi, (0, 1, 1)),
]


@pytest.mark.parametrize("t,angle,axis", RTCSM_STARTS)
def test_rtcsm_reference_cases(orc, t, angle, axis):
    g = seven_point_grid(orc, 0.1)
    r = orc.rtcsm_match(g, SEVEN, orc.angle_axis_pose(t, angle, axis), 0.3, np.deg2rad(1.0), 1e-1, 1.0)
    assert (r["linear"], r["angular"]) == (3, 1)  # 0.3 / 0.1f -> 3; 343 * 27 = 9261 candidates
    assert isapprox(r["pose"], orc.pose((-1, 0, 0)), 1e-3)
    assert r["score"] > 0


def test_rtcsm_window_promotions(orc):
    """double window / float resolution, then lround (real_time_correlative_scan_matcher_3d.cc:59-60)."""
    for window, res, want in [(0.1, 0.2, 0), (0.15, 0.1, 1), (0.3, 0.1, 3), (0.15, 0.05, 3)]:
        g = orc.Grid(res)
        g.set_probability((0, 0, 0), 0.9)
        r = orc.rtcsm_match(g, np.array([[1.0, 0, 0]], np.float32), orc.IDENTITY_POSE, window, 0.0, 0.1, 0.1)
        assert r["linear"] == want
        assert r["angular"] == 0


# ---------------------------------------------------------------- Ceres matcher (scan_matching/ceres_scan_matcher_3d_test.cc:34-116)
def ceres_case(orc, cloud, initial, expected):
    g = seven_point_grid(orc, 1.0)
    pose, s = orc.ceres_match([cloud], [g], [1.0], 0.01, 0.1, initial[:3], initial, nonmono=True, max_iter=10)
    assert 0.0 <= s["final_cost"] <= 1e-2, s
    assert isapprox(pose, expected, 3e-2), (pose, s)
    return pose, s


@pytest.mark.parametrize("t", [(-1, 0, 0), (-0.8, 0, 0), (-1, 0, -0.2), (-0.9, -0.2, 0.2)])
def test_ceres_reference_translations(orc, t):
    ceres_case(orc, SEVEN, orc.pose(t), orc.pose((-1, 0, 0)))


def test_ceres_reference_full_pose_correction(orc):  # :102-116
    a = 0.05
    c, s = np.cos(a), np.sin(a)
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    cloud = (SEVEN.astype(np.float64) @ Rz.T).astype(np.float32)
    # expected = Translation(-1,0,0) * Rotation(z, 0.05)^-1
    expected = orc.angle_axis_pose((-1, 0, 0), -a, (0, 0, 1))
    initial = orc.angle_axis_pose((-0.95, -0.05, 0.05), a, (1, 0, 0))
    ceres_case(orc, cloud, initial, expected)


def test_ceres_optimum_matches_scipy(orc):
    """Independent check of the optimum (not of the LM trajectory): scipy LM on the same residuals
    reaches the same cost and pose as the restated Ceres loop run to convergence."""
    from scipy.optimize import least_squares
    g = seven_point_grid(orc, 1.0)
    init = orc.pose((-0.9, -0.2, 0.2))
    pose, s = orc.ceres_match([SEVEN], [g], [1.0], 0.01, 0.1, init[:3], init, max_iter=200)
    tq_inv = np.array([init[3], -init[4], -init[5], -init[6]])

    def quat_mul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                         a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                         a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])

    def residuals(v):
        t, aa = v[:3], v[3:]
        n = np.linalg.norm(aa)
        dq = np.array([1, 0, 0, 0.0]) if n == 0 else np.concatenate([[np.cos(n)], np.sin(n) / n * aa])
        q = quat_mul(dq, init[3:])
        r = []
        for p in SEVEN.astype(np.float64):
            qv = q[1:]
            uv = 2 * np.cross(qv, p)
            w = p + q[0] * uv + np.cross(qv, uv) + t
            r.append((1.0 / np.sqrt(7.0)) * (1.0 - g.interpolate(*w)))
        r += list(0.01 * (t - init[:3]))
        r += list(0.1 * quat_mul(tq_inv, q)[1:])
        return np.array(r)

    sol = least_squares(residuals, np.concatenate([init[:3], np.zeros(3)]), method="lm", xtol=1e-14, ftol=1e-14)
    # The smoothstep interpolation has a vanishing gradient AND curvature at voxel centres, so both solvers
    # crawl along a flat valley near the optimum (cost floor 7 * 0.5 * (0.1/sqrt(7))^2 = 0.005): agree loosely.
    assert abs(0.5 * np.sum(sol.fun ** 2) - s["final_cost"]) < 1e-6
    assert np.allclose(sol.x[:3], pose[:3], atol=5e-3)


def test_ceres_normal_equations_consistent(orc):
    """J^T J / J^T r from the Jet path agree with finite differences of the residual-only path."""
    g = seven_point_grid(orc, 1.0)
    ref = orc.pose((-0.9, -0.1, 0.1))
    at = orc.angle_axis_pose((-0.93, -0.12, 0.08), 0.03, (0.3, -0.2, 0.9))
    cost, grad, H = orc.ceres_normal_equations([SEVEN], [g], [1.0], 0.01, 0.1, ref[:3], ref, at)
    assert np.allclose(H, H.T) and (np.linalg.eigvalsh(H) > -1e-12).all()

    def cost_at(delta):
        t = at[:3] + delta[:3]
        n = np.linalg.norm(delta[3:])
        q = at[3:]
        if n > 0:
            dq = np.concatenate([[np.cos(n)], np.sin(n) / n * delta[3:]])
            a, b = dq, q
            q = np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                          a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                          a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                          a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
        c, _, _ = orc.ceres_normal_equations([SEVEN], [g], [1.0], 0.01, 0.1, ref[:3], ref, np.concatenate([t, q]))
        return c
    h = 1e-6
    for a in range(6):
        e = np.zeros(6)
        e[a] = h
        fd = (cost_at(e) - cost_at(-e)) / (2 * h)
        assert abs(fd - grad[a]) < 1e-6 * max(1.0, abs(fd)), (a, fd, grad[a])


# ---- rotation_delta_cost_functor_3d_test.cc -----------------------------------------------------------------------------------
def _aa_q(angle, axis):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    return np.array([np.cos(angle / 2), *(np.sin(angle / 2) * axis)])


def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def test_rotation_delta_same_rotation_gives_zero_cost(orc):  # :47-57
    ident = np.array([1.0, 0, 0, 0])
    assert abs(orc.rotation_delta_cost(1.0, ident, ident)) < 1e-8
    r = _aa_q(0.9, [0.2, 0.1, 0.3])
    assert abs(orc.rotation_delta_cost(1.0, r, r)) < 1e-8


def test_rotation_delta_computes_correct_cost(orc):  # :59-80
    scale, angle = 1.2, 0.8
    rotation = _aa_q(angle, [0.2, 0.1, 0.8])
    target = _aa_q(0.2, [-0.5, 0.3, 0.4])
    expected = (scale * np.sin(angle / 2)) ** 2
    assert abs(orc.rotation_delta_cost(scale, [1.0, 0, 0, 0], rotation) - expected) < 1e-8
    assert abs(orc.rotation_delta_cost(scale, target, _qmul(target, rotation)) - expected) < 1e-8
    assert abs(orc.rotation_delta_cost(scale, target, _qmul(rotation, target)) - expected) < 1e-8


# ---- transform/transform_test.cc:29-46 (GetAngle) ---------------------------------------------------------------------------
def test_get_angle_of_angle_axis_round_trip(orc):
    """100 draws of std::mt19937(42): angle ~ U(0, pi), axis = normalised U(-1, 1)^3, all float; GetAngle(Rotation(
    AngleAxisVectorToRotationQuaternion(angle * axis))) returns the angle to 1e-6."""
    f = np.float32
    raw = iter(np.random.RandomState(42).randint(0, 2 ** 32, size=400, dtype=np.uint64))

    def uniform(a, b):      # libstdc++: generate_canonical<float, 24> then (b - a) * c + a, float arithmetic
        c = f(next(raw)) / f(4294967296.0)
        if c >= f(1.0):
            c = np.nextafter(f(1.0), f(0.0))
        return f(c * f(b - a)) + f(a)
    for _ in range(100):
        angle = uniform(f(0.0), f(np.pi))
        v = np.array([uniform(f(-1), f(1)) for _ in range(3)], f)
        axis = v / f(np.sqrt(f(v[0] * v[0] + f(v[1] * v[1] + v[2] * v[2]))))
        assert abs(orc.angle_of_angle_axis_f(angle * axis) - angle) < 1e-6


# ---------------------------------------------------------------- parameterisation / optional terms
def _yaw_of(q):
    w, x, y, z = q
    return np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))


def test_ceres_only_optimize_yaw_keeps_roll_pitch(orc):
    """YawOnlyQuaternionPlus (rotation_parameterization.h:27-39, selected by ceres_pose.cc:23-44 when only_optimize_yaw):
    the update is q_delta(yaw) (x) q, so whatever the data asks for, the solution is a z-rotation times the start."""
    og = seven_point_grid(orc, 1.0)
    # start: small roll AND a yaw error; the truth is the identity rotation, so the full solve removes both ...
    init = orc.angle_axis_pose((-1.0, 0.0, 0.0), 0.0, (0, 0, 1))
    roll = orc.angle_axis_pose((0, 0, 0), 0.04, (1, 0, 0))[3:]
    yaw = orc.angle_axis_pose((0, 0, 0), 0.03, (0, 0, 1))[3:]

    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
    init[3:] = qmul(yaw, roll)
    full, fs = orc.ceres_match([SEVEN], [og], [1.0], 0.01, 0.1, init[:3], init, nonmono=True, max_iter=30)
    only, os_ = orc.ceres_match([SEVEN], [og], [1.0], 0.01, 0.1, init[:3], init, only_yaw=True, nonmono=True, max_iter=30)
    # ... while the yaw-only solve leaves q_result (x) q_init^-1 a pure z rotation
    qi_inv = init[3:] * np.array([1, -1, -1, -1])
    d = qmul(only[3:], qi_inv)
    d /= np.linalg.norm(d)
    assert abs(d[1]) < 1e-12 and abs(d[2]) < 1e-12
    assert abs(np.linalg.norm(only[3:]) - 1.0) < 1e-12
    assert only_sane(os_) and os_["final_cost"] < os_["initial_cost"]
    # the full parameterisation does move roll / pitch on the same data
    df = qmul(full[3:], qi_inv)
    assert np.hypot(df[1], df[2]) > 1e-4 and only_sane(fs)
    # 4 local parameters are enough to remove a pure yaw error (0.2 rad -> below the fixture's 3e-2 tolerance); z never moves
    init2 = orc.angle_axis_pose((-1.0, 0.0, 0.0), 0.2, (0, 0, 1))
    a, sa = orc.ceres_match([SEVEN], [og], [1.0], 0.01, 0.0, init2[:3], init2, only_yaw=True, nonmono=True, max_iter=50)
    assert abs(_yaw_of(a[3:])) < 3e-2 and sa["final_cost"] < 1e-2 and a[4] == 0.0 and a[5] == 0.0


def only_sane(s):
    return s["termination"] in (0, 1) and np.isfinite(s["final_cost"])


def test_ceres_non_positive_weights_drop_the_terms(orc):
    """The fork adds the translation / rotation residual blocks only for weights > 0 (ceres_scan_matcher_3d.cc:104-118):
    a zero and a negative weight give the same problem, its cost is the occupied-space cost alone, and the result
    differs from the weighted solve."""
    og = seven_point_grid(orc, 1.0)
    init = orc.angle_axis_pose((-0.9, -0.2, 0.2), 0.02, (0.2, 0.5, 0.8))
    target = np.array([-5.0, 3.0, 1.0])          # far away: a weighted translation term must show up
    zero, zs = orc.ceres_match([SEVEN], [og], [1.0], 0.0, 0.0, target, init, nonmono=True, max_iter=10)
    neg, ns = orc.ceres_match([SEVEN], [og], [1.0], -3.0, -7.0, target, init, nonmono=True, max_iter=10)
    assert np.array_equal(zero, neg) and zs["initial_cost"] == ns["initial_cost"] and zs["num_iterations"] == ns["num_iterations"]
    c0, g0, h0 = orc.ceres_normal_equations([SEVEN], [og], [1.0], 0.0, 0.0, target, init, init)
    c1, g1, h1 = orc.ceres_normal_equations([SEVEN], [og], [1.0], 2.0, 0.0, target, init, init)
    c2, g2, h2 = orc.ceres_normal_equations([SEVEN], [og], [1.0], 0.0, 2.0, target, init, init)
    assert abs(zs["initial_cost"] - c0) < 1e-15
    # translation term alone: + 1/2 * w^2 |t - target|^2, Hessian + w^2 I on the translation block only
    assert abs((c1 - c0) - 0.5 * 4.0 * np.sum((init[:3] - target) ** 2)) < 1e-9
    assert np.allclose(h1[:3, :3] - h0[:3, :3], 4.0 * np.eye(3), atol=1e-9) and np.allclose(h1[3:, 3:], h0[3:, 3:], atol=1e-12)
    # rotation term alone at the reference rotation: zero residual, Hessian only on the rotation block
    assert abs(c2 - c0) < 1e-15 and np.allclose(h2[:3, :], h0[:3, :], atol=1e-12) and not np.allclose(h2[3:, 3:], h0[3:, 3:])
    with_t, _ = orc.ceres_match([SEVEN], [og], [1.0], 2.0, 0.0, target, init, nonmono=True, max_iter=10)
    assert np.linalg.norm(with_t[:3] - zero[:3]) > 1e-3
