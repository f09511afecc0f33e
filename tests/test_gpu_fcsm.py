"""GPU parity of the loop-closure coarse matcher: the device scores the whole translation window;
the oracle runs the reference's precomputation stack + branch and bound. Scores are integer sums -> bit-exact."""
import numpy as np
import pytest

from helpers import pose_error, workload
from test_fcsm_oracle import CLOUD, TEST_OPTS, fixture_grid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import dliom
    c = dliom.Context(0)
    yield c
    c.close()


def same(got, want, unique=True):
    assert bool(got.found) == bool(want.found)
    if not want.found:
        return
    assert np.float32(got.score) == np.float32(want.score)
    assert np.float32(got.rotational_score) == np.float32(want.rotational_score)
    if unique:
        assert tuple(got.offset) == tuple(want.offset)
        assert np.array_equal(np.array(got.pose_estimate[:]), np.array(want.pose[:]))
        assert np.float32(got.low_resolution_score) == np.float32(want.low_resolution_score)


def test_reference_fixture(ctx, orc):
    import dliom
    rng = np.random.default_rng(42)
    for _ in range(6):
        shift = (0.7 * rng.uniform(-1, 1, 3)).astype(np.float32)
        og = fixture_grid(orc, shift)
        g = dliom.Grid.from_oracle(ctx, og)
        want = orc.fcsm_match_3dof(og, og, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **TEST_OPTS)
        got = ctx.fcsm_match_3dof(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **TEST_OPTS)
        same(got, want, unique=False)
        assert np.abs(np.array(got.pose_estimate[:3]) - shift).max() < 0.05
        assert got.num_candidates == 33 ** 3
        far = np.array([[42, 42, 42]], np.float32)
        assert not ctx.fcsm_match_3dof(g, g, CLOUD, far, orc.IDENTITY_POSE, 0.1, **TEST_OPTS).found
        assert not ctx.fcsm_match_3dof(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.95, **TEST_OPTS).found


def test_scene_submap_pose_graph_options(ctx, orc):
    """pose_graph.lua's constraint_builder options (5 m x 5 m x 1 m window at 0.1 m: 101 x 101 x 21 leaves, depth 8 /
    full-resolution depth 3) on a synthetic-scene submap, node displaced by metres."""
    import dliom
    w = workload(beams=16, num_map_scans=40, num_scans=3)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    rng = np.random.default_rng(5)
    for k in range(3):
        pts = orc.ingest_scan(w["opts"], w["scans"][k], w["origin"], w["prev"][k], w["truth"][k])["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        guess = np.array(w["truth"][k], np.float64)
        guess[:3] += rng.uniform(-1, 1, 3) * [2.5, 2.5, 0.5]
        kw = dict(min_low_resolution_score=0.3)
        want = orc.fcsm_match_3dof(w["hi"], w["lo"], pts[hk], pts[lk], guess, 0.15, **kw)
        got = ctx.fcsm_match_3dof(hi, lo, pts[hk], pts[lk], guess, 0.15, **kw)
        assert want.found
        same(got, want)
        assert got.num_candidates == 101 * 101 * 21 and want.leaves_scored < got.num_candidates
        assert np.abs(np.array(got.pose_estimate[:3]) - w["truth"][k][:3]).max() < 0.3
        # the stock thresholds (min_score 0.55 is far above what a 16-beam sweep reaches here): both say "no constraint"
        assert not orc.fcsm_match_3dof(w["hi"], w["lo"], pts[hk], pts[lk], guess, 0.55).found
        assert not ctx.fcsm_match_3dof(hi, lo, pts[hk], pts[lk], guess, 0.55).found


def test_low_resolution_gate_skips_better_scoring_leaves(ctx, orc):
    """The low-resolution grid disagrees with the high-resolution one by two cells: every leaf around the high-resolution
    optimum fails the gate and the matcher has to walk down the score order (the reference does this inside the branch
    and bound, the device by rejecting the arg-max and re-reducing). Equal-score leaves abound here, so only the score
    and the gate outcome are compared, not which of the tied leaves was reported."""
    import dliom
    shift = np.array([0.2, -0.1, 0.15], np.float32)
    ohi, olo = fixture_grid(orc, shift), fixture_grid(orc, shift - np.array([0.1, 0, 0], np.float32))
    hi, lo = dliom.Grid.from_oracle(ctx, ohi), dliom.Grid.from_oracle(ctx, olo)
    opts = dict(TEST_OPTS, min_low_resolution_score=0.5)
    want = orc.fcsm_match_3dof(ohi, olo, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **opts)
    got = ctx.fcsm_match_3dof(hi, lo, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **opts)
    best = orc.fcsm_match_3dof(ohi, ohi, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, **TEST_OPTS)
    assert want.found and want.score < best.score
    same(got, want, unique=False)
    assert got.low_resolution_score >= 0.5


def test_argument_errors(ctx, orc):
    import dliom
    og = fixture_grid(orc, (0, 0, 0))
    g = dliom.Grid.from_oracle(ctx, og)
    with pytest.raises(dliom.DlError):
        ctx.fcsm_match_3dof(g, g, CLOUD[:0], CLOUD, orc.IDENTITY_POSE, 0.1)
    with pytest.raises(dliom.DlError):
        ctx.fcsm_match_3dof(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, 0.1, depth=0)


def test_constraint_search_batch(ctx, orc):
    """ConstraintBuilder3D::ComputeConstraint (constraint_builder_3d.cc:261-333) for a batch of (node, submap) pairs in one
    call: coarse search, min_score prune, refinement seeded with the coarse pose on the device. Pairs mix two submaps,
    several nodes, a node whose guess is far outside the window (pruned) — the oracle chains its own two matchers."""
    import dliom
    w = workload(beams=16, num_map_scans=40, num_scans=3)
    w2 = workload(beams=16, num_map_scans=12, num_scans=3)
    subs = [(w["hi"], w["lo"]), (w2["hi"], w2["lo"])]
    dev = [(dliom.Grid.from_oracle(ctx, h), dliom.Grid.from_oracle(ctx, l)) for h, l in subs]
    rng = np.random.default_rng(11)
    guesses, his, los, hg, lg, og = [], [], [], [], [], []
    for k in range(3):
        pts = orc.ingest_scan(w["opts"], w["scans"][k], w["origin"], w["prev"][k], w["truth"][k])["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        for s in range(2):
            for far in (False, True):
                g = np.array(w["truth"][k], np.float64)
                g[:3] += rng.uniform(-1, 1, 3) * [2.0, 2.0, 0.4] + (np.array([60.0, 0, 0]) if far else 0)
                guesses.append(g); his.append(pts[hk]); los.append(pts[lk])
                hg.append(dev[s][0]); lg.append(dev[s][1]); og.append(subs[s])
    opt = dliom.ConstraintOptions.defaults(min_score=0.15, min_low_resolution_score=0.3)
    got = ctx.constraint_search_batch(opt, guesses, his, los, hg, lg)
    assert len(got) == 12
    found = 0
    for c, g, h, l, (oh, ol) in zip(got, guesses, his, los, og):
        coarse = orc.fcsm_match_3dof(oh, ol, h, l, g, 0.15, min_low_resolution_score=0.3)
        assert bool(c.found) == bool(coarse.found)
        if not coarse.found:
            assert c.translation_weight == 0 and not any(c.pose[:])
            continue
        found += 1
        assert np.float32(c.score) == np.float32(coarse.score)
        assert np.float32(c.low_resolution_score) == np.float32(coarse.low_resolution_score)
        assert np.array_equal(np.array(c.coarse_pose[:]), np.array(coarse.pose[:]))
        cp = np.array(coarse.pose[:])
        want, summary = orc.ceres_match([h, l], [oh, ol], [5.0, 30.0], 10.0, 1.0, cp[:3], cp, max_iter=10)
        dt, dr = pose_error(np.array(c.pose[:]), want)
        assert dt < 1e-6 and dr < 1e-7
        assert c.summary.num_iterations == summary["num_iterations"]
        assert abs(c.summary.final_cost - summary["final_cost"]) <= 1e-6 * max(1.0, summary["final_cost"])
        assert (c.translation_weight, c.rotation_weight) == (1.1e4, 1e5)
    assert 1 <= found <= 6     # the six "far" guesses can never match


def test_constraint_search_batch_edge_cases(ctx, orc):
    import dliom
    og = fixture_grid(orc, (0, 0, 0))
    g = dliom.Grid.from_oracle(ctx, og)
    opt = dliom.ConstraintOptions.defaults()
    assert ctx.constraint_search_batch(opt, [], [], [], [], []) == []
    with pytest.raises(dliom.DlError):   # an empty cloud in the batch
        ctx.constraint_search_batch(opt, [orc.IDENTITY_POSE] * 2, [CLOUD, CLOUD[:0]], [CLOUD, CLOUD], [g, g], [g, g])


def test_pruned_search_equals_exhaustive(ctx, orc):
    """The default search opens only the 8^3 blocks whose exact bound (sliding maximum of the 8-bit grid) can still reach the
    best leaf; DLIOM_FCSM_EXHAUSTIVE=1 scores every leaf. Same answer, bit for bit — also after the grid changed (the search
    index is rebuilt from the grid's version) and for windows that are not multiples of the block size."""
    import os
    import dliom
    w = workload(beams=16, num_map_scans=40, num_scans=3)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    rng = np.random.default_rng(21)

    def both(pts_hi, pts_lo, guess, min_score, **kw):
        got = ctx.fcsm_match_3dof(hi, lo, pts_hi, pts_lo, guess, min_score, **kw)
        os.environ["DLIOM_FCSM_EXHAUSTIVE"] = "1"
        try:
            want = ctx.fcsm_match_3dof(hi, lo, pts_hi, pts_lo, guess, min_score, **kw)
        finally:
            del os.environ["DLIOM_FCSM_EXHAUSTIVE"]
        assert (got.found, got.score, got.low_resolution_score, list(got.offset), list(got.pose_estimate)) == \
               (want.found, want.score, want.low_resolution_score, list(want.offset), list(want.pose_estimate))
        return got

    clouds = []
    for k in range(3):
        pts = orc.ingest_scan(w["opts"], w["scans"][k], w["origin"], w["prev"][k], w["truth"][k])["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        clouds.append((pts[hk], pts[lk]))
        for window in ((5.0, 1.0), (3.3, 0.7), (0.4, 0.4)):
            guess = np.array(w["truth"][k], np.float64)
            guess[:3] += rng.uniform(-1, 1, 3) * [window[0] * 0.5, window[0] * 0.5, window[1] * 0.5]
            for ms, ml in ((0.15, 0.3), (0.05, 0.05), (0.6, 0.55)):
                both(pts[hk], pts[lk], guess, ms, xy_window=window[0], z_window=window[1], min_low_resolution_score=ml)
    # the grid changes (host cells, then a device-side insert): the index follows
    before = both(*clouds[0], w["truth"][0], 0.15, min_low_resolution_score=0.3)
    cell = w["hi"].cell_index(w["truth"][0][:3].astype(np.float32) + np.float32([3.0, 0.5, 0.2]))
    xs = np.arange(cell[0], cell[0] + 40, dtype=np.int32)
    hi.set_cells(xs, np.full(40, cell[1], np.int32), np.full(40, cell[2], np.int32), np.full(40, 32767, np.uint16))
    both(*clouds[0], w["truth"][0], 0.15, min_low_resolution_score=0.3)
    ctx.submap_insert_range_data(hi, lo, orc.IDENTITY_POSE, w["truth"][1][:3].astype(np.float32),
                                 clouds[1][0] + np.float32([0.3, 0.0, 0.0]), high_resolution_max_range=20)
    after = both(*clouds[0], w["truth"][0], 0.15, min_low_resolution_score=0.3)
    assert before.found and after.found


def test_full_match_with_yaw_search(ctx, orc):
    """dl_fcsm_match = FastCorrelativeScanMatcher3D::Match: yaw steps and rotational scores on the host, one translation search
    per passing step in a single device batch. Against the oracle's full matcher: the reference's fixture (zero histograms,
    rotated poses) and scene scans with real histograms where min_rotational_score removes most steps."""
    import dliom
    from helpers import apply_pose
    opts = dict(xy_window=0.8, z_window=0.8, angular_window=0.3, min_low_resolution_score=0.15, min_rotational_score=0.1, depth=6,
                full_depth=6)
    rng = np.random.default_rng(4)
    for _ in range(3):
        t = (0.7 * rng.uniform(-1, 1, 3)).astype(np.float32)
        theta = np.float32(0.2 * rng.uniform(-1, 1))
        expected = np.array([*t, np.cos(theta / 2), 0, 0, np.sin(theta / 2)], np.float64)
        og = orc.Grid(0.05)
        og.insert_range_data(t, apply_pose(expected, CLOUD.astype(np.float64)).astype(np.float32), hit=0.7, miss=0.4, num_free=5)
        g = dliom.Grid.from_oracle(ctx, og)
        want, ws, wn = orc.fcsm_match_full(og, og, CLOUD, CLOUD, orc.IDENTITY_POSE, orc.IDENTITY_POSE, 0.1, **opts)
        got = ctx.fcsm_match(g, g, CLOUD, CLOUD, orc.IDENTITY_POSE, orc.IDENTITY_POSE, 0.1, **opts)
        assert want.found and got.found
        assert np.float32(got.score) == np.float32(want.score) and np.float32(got.rotational_score) == np.float32(want.rotational_score)
        if got.scan_index == ws:
            assert list(got.pose_estimate) == list(want.pose) and np.float32(got.low_resolution_score) == np.float32(want.low_resolution_score)
        assert np.abs(np.array(got.pose_estimate[:3]) - t).max() < 0.05
    w = workload(beams=16, num_map_scans=40, num_scans=3)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    for k in range(2):
        pts = orc.ingest_scan(w["opts"], w["scans"][k], w["origin"], w["prev"][k], w["truth"][k])["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        scan_hist = orc.compute_histogram(pts, 120)
        submap_hist = orc.compute_histogram(apply_pose(np.concatenate([[0, 0, 0], w["truth"][k][3:]]), pts.astype(np.float64)).astype(np.float32), 120)
        node = np.array(w["truth"][k], np.float64)
        node[:3] += [0.4, -0.3, 0.05]
        half = 0.04 * (1 if k == 0 else -1)     # a yaw error of 0.08 rad for the search to undo
        dq = np.array([np.cos(half), 0, 0, np.sin(half)])
        a, b = dq, node[3:].copy()
        node[3:] = [a[0] * b[0] - a[3] * b[3], a[0] * b[1] - a[3] * b[2], a[0] * b[2] + a[3] * b[1], a[0] * b[3] + a[3] * b[0]]
        kw = dict(xy_window=1.0, z_window=0.3, angular_window=0.2, min_low_resolution_score=0.3, min_rotational_score=0.77)
        want, ws, wn = orc.fcsm_match_full(w["hi"], w["lo"], pts[hk], pts[lk], node, orc.IDENTITY_POSE, 0.15, histogram=scan_hist,
                                           submap_histogram=submap_hist, **kw)
        got = ctx.fcsm_match(hi, lo, pts[hk], pts[lk], node, orc.IDENTITY_POSE, 0.15, submap_histogram=submap_hist,
                             scan_histogram=scan_hist, **kw)
        assert bool(got.found) == bool(want.found)
        if want.found:
            assert np.float32(got.score) == np.float32(want.score)
            if got.scan_index == ws:
                assert list(got.pose_estimate) == list(want.pose) and np.float32(got.rotational_score) == np.float32(want.rotational_score)
