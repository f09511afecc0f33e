"""Oracle pinning for the loop-closure coarse matcher: the branch-and-bound restatement against the
reference's own fixture (SM/fast_correlative_scan_matcher_3d_test.cc:35-150, reduced to the translation-only
MatchWith3DofInitial form this fork calls) and against a brute-force numpy scoring of every leaf."""
import numpy as np

CLOUD = np.array([[4, 0, 0], [4.5, 0, 0], [5, 0, 0], [5.5, 0, 0], [0, 4, 0], [0, 4.5, 0], [0, 5, 0], [0, 5.5, 0],
                  [0, 0, 4], [0, 0, 4.5], [0, 0, 5], [0, 0, 5.5]], np.float32)
TEST_OPTS = dict(xy_window=0.8, z_window=0.8, min_low_resolution_score=0.15, min_rotational_score=0.1, depth=6, full_depth=6)


def fixture_grid(orc, shift):
    g = orc.Grid(0.05)
    g.insert_range_data(np.asarray(shift, np.float32), (CLOUD + np.asarray(shift, np.float32)).astype(np.float32),
                        hit=0.7, miss=0.4, num_free=5)
    return g


def brute_force(orc, grid, pts, pose7, wxy, wz):
    """score of every leaf, computed independently of the oracle's grid stack: V8 from the exported cells."""
    xs, ys, zs, vs = grid.export()
    f = np.float32
    prob = np.array([orc.lib().orc_value_to_probability(int(v)) for v in vs], np.float32)
    scaled = (prob - f(0.1)) * (f(255.0) / (f(0.9) - f(0.1)))
    v8 = np.where(scaled > 0, np.floor(scaled + f(0.5)), np.ceil(scaled - f(0.5))).astype(np.int64)   # lround
    table = {(int(x), int(y), int(z)): int(v) for x, y, z, v in zip(xs, ys, zs, v8)}
    from helpers import apply_pose
    cells = [tuple(int(c) for c in grid.cell_index(p)) for p in apply_pose(pose7, pts.astype(np.float64)).astype(np.float32)]
    out = {}
    for oz in range(-wz, wz + 1):
        for oy in range(-wxy, wxy + 1):
            for ox in range(-wxy, wxy + 1):
                s = sum(table.get((c[0] + ox, c[1] + oy, c[2] + oz), 0) for c in cells)
                out[(ox, oy, oz)] = f(0.1) + (f(s) / f(len(cells))) * ((f(0.9) - f(0.1)) / f(255.0))
    return out


def test_reference_fixture_translation_only(orc):
    rng = np.random.default_rng(42)
    for _ in range(10):
        shift = (0.7 * rng.uniform(-1, 1, 3)).astype(np.float32)
        g = fixture_grid(orc, shift)
        r = orc.fcsm_match_3dof(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **TEST_OPTS)
        assert r.found and r.score > 0.1 and r.low_resolution_score > 0.14 and r.rotational_score > 0.09
        assert np.abs(np.array(r.pose[:3]) - shift).max() < 0.05            # IsNearly(expected_pose, 0.05)
        assert np.allclose(np.array(r.pose[3:]), [1, 0, 0, 0])
        far = np.array([[42, 42, 42]], np.float32)                          # low-resolution gate rejects everything
        assert not orc.fcsm_match_3dof(g, g, CLOUD, far, orc.IDENTITY_POSE, 0.1, **TEST_OPTS).found


def test_branch_and_bound_equals_brute_force(orc):
    """The bounds are admissible: the B&B winner's score is the maximum over all leaves, for full-resolution-only stacks
    and for stacks with half-resolution levels."""
    rng = np.random.default_rng(3)
    for depth, full in ((6, 6), (5, 2), (4, 1), (1, 1)):
        shift = (0.3 * rng.uniform(-1, 1, 3)).astype(np.float32)
        g = fixture_grid(orc, shift)
        guess = np.array([0.05, -0.02, 0.01, 1, 0, 0, 0], np.float64)
        opts = dict(TEST_OPTS, xy_window=0.4, z_window=0.3, depth=depth, full_depth=full)
        r = orc.fcsm_match_3dof(g, g, CLOUD, CLOUD, guess, 0.1, **opts)
        scores = brute_force(orc, g, CLOUD, guess, 8, 6)
        best = max(scores.values())
        assert r.found and np.float32(r.score) == best
        assert scores[tuple(r.offset)] == best
        assert r.leaves_scored <= len(scores)
        want_t = np.float32(0.05) * np.array(r.offset, np.float32) + guess[:3].astype(np.float32)
        assert np.array_equal(np.array(r.pose[:3], np.float32), want_t)


def test_min_score_above_best_returns_nothing(orc):
    g = fixture_grid(orc, (0.1, 0.0, -0.1))
    assert not orc.fcsm_match_3dof(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.95, **TEST_OPTS).found


def test_matcher_object_is_reentrant(orc):
    """One per-submap matcher (stack built once) queried from several threads gives what fresh matchers give."""
    from concurrent.futures import ThreadPoolExecutor
    g = fixture_grid(orc, (0.25, -0.1, 0.05))
    m = orc.FastCorrelativeScanMatcher(g, g, **TEST_OPTS)
    rng = np.random.default_rng(9)
    guesses = [np.array([*(0.2 * rng.uniform(-1, 1, 3)), 1, 0, 0, 0]) for _ in range(16)]
    with ThreadPoolExecutor(8) as ex:
        got = list(ex.map(lambda q: m.match(CLOUD, CLOUD, q, 0.1), guesses))
    for q, r in zip(guesses, got):
        want = orc.fcsm_match_3dof(g, g, CLOUD, CLOUD, q, 0.1, **TEST_OPTS)
        assert (r.found, r.score, list(r.offset), list(r.pose)) == (want.found, want.score, list(want.offset), list(want.pose))


# ---- the full matcher (yaw search) against the reference's own tests -------------------------------------------------------
def test_rotational_matcher_only_same_histogram_is_score_one(orc):
    """SM/rotational_scan_matcher_test.cc:28-36."""
    h = np.array([1, 43, 0.5, 0.3123, 23, 42, 0], np.float32)
    s = orc.rotational_match(h, h, 0.0, [0.0, 1.0])
    assert abs(s[0] - 1.0) < 1e-6 and s[1] < 1.0


def test_rotational_matcher_interpolates_as_expected(orc):
    """SM/rotational_scan_matcher_test.cc:38-66: rotating the t-th fraction of a bucket in / out, both directions."""
    n, per = 10, np.float32(np.pi / 10)
    unit = lambda k: np.eye(n, dtype=np.float32)[k]
    t = np.float32(0.0)
    while t < 1.0:
        want = 0.0 if t == 0 else float(t / np.hypot(t, 1 - t))
        assert abs(orc.rotational_match(unit(3), unit(2), 0.0, [t * per])[0] - want) < 1e-6
        assert abs(orc.rotational_match(unit(3), unit(2), 0.0, [(2 - t) * per])[0] - want) < 1e-6
        s = orc.rotational_match(unit(3), unit(4), 0.0, [-t * per, (t - 2) * per])
        assert abs(s[0] - want) < 1e-6 and abs(s[1] - want) < 1e-6
        t = np.float32(t + np.float32(0.1))


def _mt19937_uniform_float(raw):
    """std::uniform_real_distribution<float>(-1, 1) over std::mt19937 draws, as libstdc++ computes it:
    generate_canonical<float, 24> = float(draw) / 2^32 (float arithmetic, clamped below 1), then * (b - a) + a."""
    f = np.float32
    c = f(raw) / f(4294967296.0)
    if c >= f(1.0):
        c = np.nextafter(f(1.0), f(0.0))
    return f(c * f(2.0)) + f(-1.0)


def _affine(pose7):
    from test_decode import rot
    m = np.eye(4)
    m[:3, :3] = rot(np.asarray(pose7, np.float64))
    m[:3, 3] = pose7[:3]
    return m


def test_reference_fixture_full_match(orc):
    """SM/fast_correlative_scan_matcher_3d_test.cc:140-172 (CorrectPoseForMatch), including its random poses: std::mt19937(42)
    through uniform_real_distribution<float>, x/y/z = 0.7 u, yaw = 0.2 u; zero rotational histograms (every yaw step passes);
    IsNearly = Eigen isApprox on the 4x4 matrices with epsilon 0.05."""
    raw = iter(np.random.RandomState(42).randint(0, 2 ** 32, size=80, dtype=np.uint64))
    f = np.float32
    opts = dict(xy_window=0.8, z_window=0.8, angular_window=0.3, min_low_resolution_score=0.15, min_rotational_score=0.1, depth=6,
                full_depth=6)
    for _ in range(20):
        x, y, z = (f(0.7) * _mt19937_uniform_float(next(raw)) for _ in range(3))
        theta = f(0.2) * _mt19937_uniform_float(next(raw))
        expected = np.array([x, y, z, np.cos(theta / 2), 0, 0, np.sin(theta / 2)], np.float64)
        from helpers import apply_pose
        g = orc.Grid(0.05)
        g.insert_range_data(np.array([x, y, z], f), apply_pose(expected, CLOUD.astype(np.float64)).astype(f), hit=0.7, miss=0.4, num_free=5)
        r, scan_index, num_scans = orc.fcsm_match_full(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, orc.IDENTITY_POSE, 0.1, **opts)
        assert r.found and r.score > 0.1 and r.rotational_score > 0.09 and r.low_resolution_score > 0.14
        assert num_scans > 20          # ~ 2 * 0.3 / acos(1 - res^2 / (2 range^2)) yaw steps
        a, b = _affine(np.array(r.pose[:])), _affine(expected)
        assert np.linalg.norm(a - b) <= 0.05 * min(np.linalg.norm(a), np.linalg.norm(b))
        far = np.array([[42, 42, 42]], f)
        assert not orc.fcsm_match_full(g, g, CLOUD, far, orc.IDENTITY_POSE, orc.IDENTITY_POSE, 0.1, **opts)[0].found


def test_compute_histogram_of_a_room(orc):
    """ComputeHistogram on a square room seen from its middle: the weight lands in the buckets of the two wall directions
    (0 = pi and pi/2); rotating the room by 0.5 rad moves the peaks by one bucket (pi/10 per bucket). No reference fixture
    exists for this function: this pins the restated semantics, not the reference."""
    t = np.arange(-4, 4, 0.3, dtype=np.float32)
    walls = np.concatenate([np.stack([t, np.full_like(t, 4)], 1), np.stack([t, np.full_like(t, -4)], 1),
                            np.stack([np.full_like(t, 4), t], 1), np.stack([np.full_like(t, -4), t], 1)])
    for yaw, peaks in ((0.0, {0, 9, 5}), (0.5, {1, 6})):
        c, s_ = np.cos(yaw), np.sin(yaw)
        xy = walls @ np.array([[c, s_], [-s_, c]], np.float32)
        pts = np.concatenate([xy, np.full((len(xy), 1), 0.05, np.float32)], 1)
        h = orc.compute_histogram(pts, 10)
        assert h.sum() > 10
        assert sum(h[k] for k in peaks) > 0.85 * h.sum()


def test_precomputation_grid_against_naive_maximum(orc):
    """SM/precomputation_grid_3d_test.cc:31-78 (TestAgainstNaiveAlgorithm): 1 000 random cells of a 2 m grid set to random
    probabilities; for depths 0..3 the precomputed value at a random cell equals (to 1e-2, the 8-bit quantisation) the largest
    probability inside the 2^depth cube that starts there."""
    rng = np.random.default_rng(23847)
    g = orc.Grid(2.0)
    for _ in range(1000):
        g.set_probability(tuple(int(v) for v in rng.integers(-50, 50, 3)), float(rng.uniform(0.1, 0.9)))
    for depth in range(4):
        width = 1 << depth
        cells = rng.integers(-50, 50, (100, 3))
        values = orc.precomputation_values(g, depth, cells)
        for c, v in zip(cells, values):
            naive = max(g.probability((int(c[0]) + dx, int(c[1]) + dy, int(c[2]) + dz))
                        for dx in range(width) for dy in range(width) for dz in range(width))
            naive = max(naive, 0.0)
            assert abs(naive - (0.1 + v * (0.8 / 255.0))) < 1e-2
