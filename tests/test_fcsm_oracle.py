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
