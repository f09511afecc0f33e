"""GPU parity: the CUDA path, called through the C-ABI (libdliom_b200.so), against the CPU oracle on the same inputs.
Integer / index / float-score work is compared bit for bit; the fp64 least-squares solve within the north-star
tolerance (1e-4 m, 1e-5 rad) — in practice ~1e-9."""
import numpy as np
import pytest

from helpers import SEVEN, pose_error, seven_point_grid, workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import dliom
    c = dliom.Context(0)
    yield c
    c.close()


def dev_grid(ctx, og):
    import dliom
    return dliom.Grid.from_oracle(ctx, og)


# ---------------------------------------------------------------- grid
def test_grid_lookup_matches_tree(ctx, orc):
    rng = np.random.RandomState(7)
    og = orc.Grid(2.0)
    cells = rng.randint(-3000, 3000, (5000, 3))
    for c in cells:
        og.set_probability(c, np.float32(rng.uniform(0.1, 0.9)))
    g = dev_grid(ctx, og)
    q = np.concatenate([cells, cells + rng.randint(-2, 3, cells.shape), rng.randint(-9000, 9000, (2000, 3)),
                        np.array([[2 ** 30, 0, 0], [-2 ** 30, 5, 5], [0, 0, 2 ** 31 - 1]])]).astype(np.int32)
    want = np.array([og.value(c) for c in q], np.uint16)
    assert np.array_equal(g.lookup(q), want)


def test_grid_incremental_update(ctx, orc):
    og = seven_point_grid(orc, 0.1)
    g = dev_grid(ctx, og)
    og.set_probability((500, -300, 20), 0.7)       # forces growth + a new brick
    og.set_probability(og.cell_index(SEVEN[0] + np.float32([-1, 0, 0])), 0.3)
    g.set_cells(*og.export())
    q = np.array([[500, -300, 20], og.cell_index(SEVEN[0] + np.float32([-1, 0, 0])), [0, 0, 0]], np.int32)
    assert np.array_equal(g.lookup(q), np.array([og.value(c) for c in q], np.uint16))


def test_grid_range_error(ctx):
    import dliom
    g = ctx.grid(1.0)
    with pytest.raises(dliom.DlError) as e:
        g.set_cells([9000], [0], [0], [5])
    assert e.value.status == -3


def test_interpolation_value_and_gradient(ctx, orc):
    og = seven_point_grid(orc, 0.1, shift=(0, 0, 0))
    g = dev_grid(ctx, og)
    rng = np.random.RandomState(3)
    pts = np.concatenate([SEVEN[rng.randint(7, size=4000)] + rng.uniform(-0.25, 0.25, (4000, 3)),
                          rng.uniform(-10, 10, (500, 3))])
    got = g.interpolate(pts)
    want = np.array([og.interpolate_grad(*p) for p in pts])
    assert np.max(np.abs(got[:, 0] - want[:, 0])) < 1e-14
    assert np.max(np.abs(got[:, 1:] - want[:, 1:])) < 1e-9


# ---------------------------------------------------------------- voxel filter
@pytest.mark.parametrize("stride", [3, 4, 8])
def test_voxel_filter_bit_exact(ctx, orc, stride):
    rng = np.random.RandomState(11 + stride)
    pts = np.zeros((60000, stride), np.float32)
    pts[:, :3] = rng.normal(0, 12, (60000, 3))
    pts[::7, :3] = np.round(pts[::7, :3] / 0.15) * 0.15 + 0.075   # many points exactly on voxel faces
    for res in (0.075, 0.15, 0.3, 2.0):
        assert np.array_equal(ctx.voxel_indices(pts, res), orc.voxel_indices(pts, res))
        assert np.array_equal(ctx.voxel_filter(pts, res), orc.voxel_filter(pts, res))


def test_voxel_indices_reciprocal_path_adversarial(ctx, orc):
    """dl_voxel_indices runs the multiply-by-reciprocal fast path with its exact-division fallback (round_div); the
    oracle divides. Hammer the rounding boundaries: points at k + 0.5 voxels +- a few ulps, huge and tiny values."""
    rng = np.random.RandomState(99)
    for res in (np.float32(0.075), np.float32(0.15), np.float32(0.1), np.float32(0.45), np.float32(2.0), np.float32(3.25)):
        k = rng.randint(-200000, 200000, 200000).astype(np.float32)
        base = ((k + np.float32(0.5)) * res).astype(np.float32)
        ulps = rng.randint(-4, 5, len(base))
        x = base.copy()
        for _ in range(4):
            x = np.where(ulps > 0, np.nextafter(x, np.float32(np.inf)), np.where(ulps < 0, np.nextafter(x, np.float32(-np.inf)), x))
            ulps = ulps - np.sign(ulps)
        pts = np.stack([x, rng.uniform(-1e5, 1e5, len(x)).astype(np.float32),
                        (rng.standard_cauchy(len(x)) * 1e-3).astype(np.float32)], 1).astype(np.float32)
        pts[:10, 2] = [0.0, -0.0, 1e-30, -1e-30, 1e-45, 3e6, -3e6, 1e7 * float(res), 0.5 * float(res), -0.5 * float(res)]
        assert np.array_equal(ctx.voxel_indices(pts, res), orc.voxel_indices(pts, res)), float(res)


def test_voxel_filter_reference_fixtures(ctx):
    pc = np.array([[0, 0, 0], [0.1, -0.1, 0.1], [0.3, -0.1, 0], [0, 0, 0.1]], np.float32)
    assert ctx.voxel_filter(pc, 0.3).tolist() == [0, 2]
    pc = np.array([[100000, 0, 0], [100000.001, -0.0001, 0.0001], [100000.003, -0.0001, 0], [-200000, 0, 0]], np.float32)
    assert ctx.voxel_filter(pc, 0.01).tolist() == [0, 3]
    pc = np.array([[-100, 0.3, 0.4, i] for i in range(100)], np.float32)
    assert ctx.voxel_filter(pc, 0.3).tolist() == [0]
    assert ctx.voxel_filter(np.zeros((0, 3), np.float32), 0.3).tolist() == []


def test_voxel_filter_full_scan_size(ctx, orc):
    w = workload(beams=64, num_map_scans=1, num_scans=1)
    rows = w["scans"][0]
    pts = rows.view(np.float32).reshape(-1, 8)
    assert len(pts) > 100000
    assert np.array_equal(ctx.voxel_filter(pts, 0.075), orc.voxel_filter(pts, 0.075))


@pytest.mark.parametrize("opts", [(2.0, 150, 15.0), (4.0, 200, 60.0), (0.5, 3000, 40.0), (2.0, 1e9, 50.0), (2.0, 5, 0.01)])
def test_adaptive_voxel_filter_same_passes_and_survivors(ctx, orc, opts):
    rng = np.random.RandomState(5)
    pts = (rng.normal(0, 1, (20000, 3)) * np.array([15, 15, 2])).astype(np.float32)
    want_keep, want_passes = orc.adaptive_voxel_filter(pts, *opts)
    keep, passes = ctx.adaptive_voxel_filter(pts, *opts)
    assert np.array_equal(passes, want_passes)
    assert np.array_equal(keep, want_keep)


@pytest.mark.parametrize("opts", [(4.0, 200, 60.0), (0.5, 3000, 40.0), (2.0, 150, 15.0)])
def test_adaptive_voxel_filter_large_cloud(ctx, orc, opts):
    """40 000 points: more than the 22 528 the shared-memory fast mode holds when the range crop keeps most of them (generic
    global-memory search, two-sweep crop and compaction), and the fast mode again when max_range = 15 m crops the cloud down."""
    rng = np.random.RandomState(6)
    pts = (rng.normal(0, 1, (40000, 3)) * np.array([15, 15, 2])).astype(np.float32)
    want_keep, want_passes = orc.adaptive_voxel_filter(pts, *opts)
    keep, passes = ctx.adaptive_voxel_filter(pts, *opts)
    assert np.array_equal(passes, want_passes)
    assert np.array_equal(keep, want_keep)


def test_adaptive_voxel_filter_small_inputs(ctx, orc):
    for n in (0, 1, 100, 150, 151):
        pts = np.random.RandomState(n).normal(0, 5, (n, 3)).astype(np.float32)
        want, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0) if n else (np.zeros(0, np.int64), None)
        keep, _ = ctx.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        assert np.array_equal(keep, want)


# ---------------------------------------------------------------- correlative matcher
RTCSM_STARTS = [((-1, 0, 0), 0.0, (1, 0, 0)), ((-0.8, 0, 0), 0.0, (1, 0, 0)), ((-1, 0, -0.2), 0.0, (1, 0, 0)),
                ((-0.9, -0.2, 0.2), 0.0, (1, 0, 0)), ((-1, 0, 0), 0.8 / 180 * np.pi, (1, 0, 0)),
                ((-1, 0, 0), 0.8 / 180 * np.pi, (0, 1, 0)), ((-1, 0, 0), 0.8 / 180 * np.pi, (0, 1, 1))]


@pytest.mark.parametrize("t,angle,axis", RTCSM_STARTS)
def test_rtcsm_reference_cases_bit_exact(ctx, orc, t, angle, axis):
    og = seven_point_grid(orc, 0.1)
    g = dev_grid(ctx, og)
    init = orc.angle_axis_pose(t, angle, axis)
    want = orc.rtcsm_match(og, SEVEN, init, 0.3, np.deg2rad(1.0), 1e-1, 1.0, want_scores=True)
    got = ctx.rtcsm_match(g, SEVEN, init, 0.3, np.deg2rad(1.0), 1e-1, 1.0, want_scores=True)
    assert (got["linear"], got["angular"]) == (want["linear"], want["angular"]) == (3, 1)
    assert got["angular_step"] == want["angular_step"]
    assert np.array_equal(got["scores"].view(np.uint32), want["scores"].view(np.uint32))
    assert got["best_index"] == want["best_index"]
    assert got["score"] == want["score"]
    assert np.array_equal(got["pose"], want["pose"])


def test_rtcsm_scene_cloud_bit_exact(ctx, orc):
    w = workload()
    ing = orc.ingest_scan(w["opts"], w["scans"][0], w["origin"], w["prev"][0], w["cur"][0])
    keep, _ = orc.adaptive_voxel_filter(ing["returns_tracking"], 2.0, 150, 15.0)
    cloud = ing["returns_tracking"][keep]
    g = dev_grid(ctx, w["hi"])
    want = orc.rtcsm_match(w["hi"], cloud, w["cur"][0], 0.15, np.deg2rad(1.0), 1e-1, 1e-1, want_scores=True)
    got = ctx.rtcsm_match(g, cloud, w["cur"][0], 0.15, np.deg2rad(1.0), 1e-1, 1e-1, want_scores=True)
    assert got["num_candidates"] == len(want["scores"])
    assert np.array_equal(got["scores"].view(np.uint32), want["scores"].view(np.uint32))
    assert got["best_index"] == want["best_index"] and np.array_equal(got["pose"], want["pose"])


# ---------------------------------------------------------------- least-squares matcher
def test_normal_equations_match_jet_path(ctx, orc):
    og = seven_point_grid(orc, 1.0)
    g = dev_grid(ctx, og)
    ref = orc.pose((-0.9, -0.1, 0.1))
    at = orc.angle_axis_pose((-0.93, -0.12, 0.08), 0.03, (0.3, -0.2, 0.9))
    wc, wg, wh = orc.ceres_normal_equations([SEVEN], [og], [1.0], 0.01, 0.1, ref[:3], ref, at)
    c, gr, h = ctx.ceres_normal_equations([SEVEN], [g], [1.0], 0.01, 0.1, ref[:3], ref, at)
    assert abs(c - wc) < 1e-15
    assert np.max(np.abs(gr - wg)) < 1e-13 and np.max(np.abs(h - wh)) < 1e-12


CERES_STARTS = [(-1, 0, 0), (-0.8, 0, 0), (-1, 0, -0.2), (-0.9, -0.2, 0.2)]


@pytest.mark.parametrize("t", CERES_STARTS)
def test_ceres_reference_cases(ctx, orc, t):
    og = seven_point_grid(orc, 1.0)
    g = dev_grid(ctx, og)
    init = orc.pose(t)
    want, ws = orc.ceres_match([SEVEN], [og], [1.0], 0.01, 0.1, init[:3], init, nonmono=True, max_iter=10)
    got, gs = ctx.ceres_match([SEVEN], [g], [1.0], 0.01, 0.1, init[:3], init, nonmono=True, max_iter=10)
    dt, dr = pose_error(got, want)
    assert dt < 1e-7 and dr < 1e-7, (got, want)
    assert gs["final_cost"] <= 1e-2 and abs(gs["final_cost"] - ws["final_cost"]) < 1e-10
    assert gs["num_iterations"] == ws["num_iterations"] and gs["termination"] == ws["termination"]
    assert gs["num_successful_steps"] == ws["num_successful_steps"]


def test_ceres_scene_two_grids_and_batch(ctx, orc):
    w = workload()
    hi, lo = dev_grid(ctx, w["hi"]), dev_grid(ctx, w["lo"])
    problems, wants, inits = [], [], []
    for s in range(len(w["scans"])):
        ing = orc.ingest_scan(w["opts"], w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        pts = ing["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        init = w["cur"][s]
        want, ws = orc.ceres_match([pts[hk], pts[lk]], [w["hi"], w["lo"]], [1.0, 6.0], 5.0, 4e2, init[:3], init)
        got, gs = ctx.ceres_match([pts[hk], pts[lk]], [hi, lo], [1.0, 6.0], 5.0, 4e2, init[:3], init)
        dt, dr = pose_error(got, want)
        assert dt < 1e-4 and dr < 1e-5, (s, dt, dr)          # north-star tolerance
        assert dt < 1e-7 and dr < 1e-8, (s, dt, dr)          # what the fp64 path actually achieves
        assert abs(gs["final_cost"] - ws["final_cost"]) <= 1e-6 * ws["final_cost"]
        assert gs["num_iterations"] == ws["num_iterations"]
        problems.append([pts[hk], pts[lk]])
        wants.append(want)
        inits.append(init)
    poses, sums = ctx.ceres_match_batch(problems, [[hi, lo]] * len(problems), [1.0, 6.0], 5.0, 4e2,
                                        np.array(inits)[:, :3].copy(), np.array(inits))
    for s in range(len(problems)):
        dt, dr = pose_error(poses[s], wants[s])
        assert dt < 1e-7 and dr < 1e-8


@pytest.mark.parametrize("angle,axis,rot_w", [(0.2, (0, 0, 1), 0.0), (0.05, (0.3, 0.1, 0.9), 0.1), (0.03, (1, 0, 0), 0.1)])
def test_ceres_only_optimize_yaw(ctx, orc, angle, axis, rot_w):
    """YawOnlyQuaternionPlus (rotation_parameterization.h:27-39; ceres_pose.cc:23-44): 4 local parameters on the device as in
    the oracle — same iterates, and the result is a z rotation times the start."""
    og = seven_point_grid(orc, 1.0)
    g = dev_grid(ctx, og)
    init = orc.angle_axis_pose((-0.95, 0.05, 0.1), angle, axis)
    want, ws = orc.ceres_match([SEVEN], [og], [1.0], 0.01, rot_w, init[:3], init, only_yaw=True, nonmono=True, max_iter=25)
    got, gs = ctx.ceres_match([SEVEN], [g], [1.0], 0.01, rot_w, init[:3], init, only_yaw=True, nonmono=True, max_iter=25)
    dt, dr = pose_error(got, want)
    # the 7-point fixture is nearly flat in yaw (the reference pins it to 3e-2): 25 iterations amplify last-bit differences
    # of the two arithmetic orders to ~1e-6 along that valley; iteration counts and costs still agree
    assert dt < 5e-6 and dr < 5e-6, (got, want)
    assert gs["num_iterations"] == ws["num_iterations"] and gs["termination"] == ws["termination"]
    assert gs["num_successful_steps"] == ws["num_successful_steps"]
    assert abs(gs["final_cost"] - ws["final_cost"]) < 1e-6 * ws["final_cost"]
    qi, q = init[3:], got[3:]
    d = np.array([q[0] * qi[0] + q[1] * qi[1] + q[2] * qi[2] + q[3] * qi[3],          # q (x) qi^-1
                  -q[0] * qi[1] + q[1] * qi[0] - q[2] * qi[3] + q[3] * qi[2],
                  -q[0] * qi[2] + q[1] * qi[3] + q[2] * qi[0] - q[3] * qi[1],
                  -q[0] * qi[3] - q[1] * qi[2] + q[2] * qi[1] + q[3] * qi[0]])
    assert abs(d[1]) < 1e-12 and abs(d[2]) < 1e-12


def test_ceres_only_optimize_yaw_scene(ctx, orc):
    w = workload()
    hi, lo = dev_grid(ctx, w["hi"]), dev_grid(ctx, w["lo"])
    ing = orc.ingest_scan(w["opts"], w["scans"][1], w["origin"], w["prev"][1], w["cur"][1])
    pts = ing["returns_tracking"]
    hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
    lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
    init = w["cur"][1]
    want, ws = orc.ceres_match([pts[hk], pts[lk]], [w["hi"], w["lo"]], [1.0, 6.0], 5.0, 4e2, init[:3], init, only_yaw=True)
    got, gs = ctx.ceres_match([pts[hk], pts[lk]], [hi, lo], [1.0, 6.0], 5.0, 4e2, init[:3], init, only_yaw=True)
    dt, dr = pose_error(got, want)
    assert dt < 1e-7 and dr < 1e-8 and gs["num_iterations"] == ws["num_iterations"]


@pytest.mark.parametrize("tw,rw", [(0.0, 0.0), (-3.0, -7.0), (2.0, 0.0), (0.0, 2.0), (-1.0, 5.0)])
def test_ceres_non_positive_weights_drop_the_terms(ctx, orc, tw, rw):
    """ceres_scan_matcher_3d.cc:104-118 (fork): the delta residual blocks exist only for weights > 0."""
    og = seven_point_grid(orc, 1.0)
    g = dev_grid(ctx, og)
    init = orc.angle_axis_pose((-0.9, -0.2, 0.2), 0.02, (0.2, 0.5, 0.8))
    target = np.array([-5.0, 3.0, 1.0])
    want, ws = orc.ceres_match([SEVEN], [og], [1.0], tw, rw, target, init, nonmono=True, max_iter=10)
    got, gs = ctx.ceres_match([SEVEN], [g], [1.0], tw, rw, target, init, nonmono=True, max_iter=10)
    dt, dr = pose_error(got, want)
    assert dt < 1e-7 and dr < 1e-7
    assert abs(gs["initial_cost"] - ws["initial_cost"]) < 1e-13 and gs["num_iterations"] == ws["num_iterations"]
    wc, wg, wh = orc.ceres_normal_equations([SEVEN], [og], [1.0], tw, rw, target, init, init)
    c, gr, h = ctx.ceres_normal_equations([SEVEN], [g], [1.0], tw, rw, target, init, init)
    assert abs(c - wc) < 1e-13 and np.max(np.abs(gr - wg)) < 1e-12 and np.max(np.abs(h - wh)) < 1e-11
    if tw <= 0 and rw <= 0:   # identical to the problem without the two blocks
        z, zs = ctx.ceres_match([SEVEN], [g], [1.0], 0.0, 0.0, target, init, nonmono=True, max_iter=10)
        assert np.array_equal(z, got) and zs["initial_cost"] == gs["initial_cost"]


def test_ceres_argument_checks(ctx, orc):
    import dliom
    og = seven_point_grid(orc, 1.0)
    g = dev_grid(ctx, og)
    init = orc.pose((-1, 0, 0))
    with pytest.raises(dliom.DlError):   # CHECK_EQ(occupied_space_weight_size, pairs)
        ctx.ceres_match([SEVEN], [g], [1.0, 2.0], 0.01, 0.1, init[:3], init)
    with pytest.raises(dliom.DlError):   # CHECK_GT(weight, 0)
        ctx.ceres_match([SEVEN], [g], [0.0], 0.01, 0.1, init[:3], init)
    with pytest.raises(dliom.DlError) as e:
        ctx.ceres_match([np.zeros((0, 3), np.float32)], [g], [1.0], 0.01, 0.1, init[:3], init)
    assert e.value.status == -4


# ---------------------------------------------------------------- scan ingest + whole front end
def test_ingest_scan_bit_exact(ctx, orc):
    import dliom
    w = workload()
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    for s in range(2):
        want = orc.ingest_scan(w["opts"], w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        got = ctx.ingest_scan(fo, w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        assert np.array_equal(got["first_keep"], want["first_keep"])
        for k in ("returns_local", "returns_tracking", "misses_tracking", "current_pose"):
            assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k


def test_ingest_scan_without_point_times(ctx, orc):
    import dliom
    w = workload()
    rows = w["scans"][0].copy()
    rows["t"] = 0.0           # LTB:430-433: |t_first| < 1e-3 -> no deskew
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    want = orc.ingest_scan(w["opts"], rows, w["origin"], w["prev"][0], w["cur"][0])
    got = ctx.ingest_scan(fo, rows, w["origin"], w["prev"][0], w["cur"][0])
    assert np.array_equal(got["returns_tracking"].view(np.uint32), want["returns_tracking"].view(np.uint32))


def test_ingest_scan_any_time_order(ctx, orc):
    """The deskew pose is shared by runs of equal point times inside 32-point tiles. Orders that defeat the grouping must
    give the same bits: rows shuffled (hardly any two neighbours share a time), every point with its own time (one
    pose per point, the table's worst case), and a scan whose size is not a multiple of 32."""
    import dliom
    w = workload()
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    rng = np.random.default_rng(12)
    base = w["scans"][1]
    shuffled = base[rng.permutation(len(base))].copy()
    shuffled["t"][-1] = 0.0
    unique = base[:5003].copy()
    unique["t"] = np.linspace(-0.1, 0.0, len(unique)).astype(np.float32)
    for rows in (shuffled, unique, base[:4097].copy()):
        want = orc.ingest_scan(w["opts"], rows, w["origin"], w["prev"][1], w["cur"][1])
        got = ctx.ingest_scan(fo, rows, w["origin"], w["prev"][1], w["cur"][1])
        assert np.array_equal(got["first_keep"], want["first_keep"])
        for k in ("returns_local", "returns_tracking", "misses_tracking", "current_pose"):
            assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k


def test_frontend_batch_timed_point_cloud_rows(ctx, orc):
    """16-byte TimedPointCloud rows (x y z t, single sensor) give exactly the results of the 32-byte RangeMeasurement rows."""
    import dliom
    w = workload()
    hi, lo = dev_grid(ctx, w["hi"]), dev_grid(ctx, w["lo"])
    fo8 = dliom.FrontendOptions.from_oracle(w["opts"])
    fo4 = dliom.FrontendOptions.from_oracle(w["opts"])
    fo4.range_row_floats = 4
    rows4 = [np.ascontiguousarray(s.view(np.float32).reshape(-1, 8)[:, :4]) for s in w["scans"]]
    r8 = ctx.frontend_match_batch(fo8, w["scans"], w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    r4 = ctx.frontend_match_batch(fo4, rows4, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    for a, b in zip(r8, r4):
        assert list(a.pose_estimate_local) == list(b.pose_estimate_local)
        assert (a.num_first_filter, a.num_returns, a.num_misses, a.num_high_resolution, a.num_low_resolution) == \
               (b.num_first_filter, b.num_returns, b.num_misses, b.num_high_resolution, b.num_low_resolution)


def test_frontend_batch_full_size_scans(ctx, orc):
    """BASELINE configs[1] size (64-beam, ~130k points): every count and the pose against the oracle."""
    import dliom
    w = workload(beams=64, num_map_scans=6, num_scans=3)
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    hi, lo = dev_grid(ctx, w["hi"]), dev_grid(ctx, w["lo"])
    res = ctx.frontend_match_batch(fo, w["scans"], w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    for s, r in enumerate(res):
        ing = orc.ingest_scan(w["opts"], w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        want = orc.match_scan(w["opts"], ing["returns_tracking"], ing["current_pose"].astype(np.float64), w["submap_pose"],
                              w["hi"], w["lo"])
        assert (r.num_first_filter, r.num_returns, r.num_misses) == (len(ing["first_keep"]), len(ing["returns_tracking"]),
                                                                      len(ing["misses_tracking"]))
        assert (r.num_high_resolution, r.num_low_resolution) == (len(want["hi_keep"]), len(want["lo_keep"]))
        dt, dr = pose_error(np.array(r.pose_estimate_local), want["pose_estimate_local"])
        assert r.ok == 1 and dt < 1e-7 and dr < 1e-8


@pytest.mark.parametrize("use_rtcsm", [0, 1])
def test_frontend_batch_matches_oracle(ctx, orc, use_rtcsm):
    import dliom
    w = workload()
    opts = orc.FrontEndOptions.defaults(use_rtcsm=use_rtcsm)
    fo = dliom.FrontendOptions.from_oracle(opts)
    hi, lo = dev_grid(ctx, w["hi"]), dev_grid(ctx, w["lo"])
    res = ctx.frontend_match_batch(fo, w["scans"], w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    for s, r in enumerate(res):
        ing = orc.ingest_scan(opts, w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        pred = ing["current_pose"].astype(np.float64)
        want = orc.match_scan(opts, ing["returns_tracking"], pred, w["submap_pose"], w["hi"], w["lo"])
        assert r.ok == 1 and want["ok"]
        assert (r.num_high_resolution, r.num_low_resolution) == (len(want["hi_keep"]), len(want["lo_keep"]))
        assert r.num_returns == len(ing["returns_tracking"]) and r.num_first_filter == len(ing["first_keep"])
        if use_rtcsm:
            assert np.float32(r.rtcsm_score) == np.float32(want["rtcsm_score"])
        dt, dr = pose_error(np.array(r.pose_estimate_local), want["pose_estimate_local"])
        assert dt < 1e-7 and dr < 1e-8, (s, dt, dr)
        assert r.summary.final_cost < r.summary.initial_cost
        # sanity against the synthetic ground truth (the reference's prior weights keep the solve near its start)
        dt, dr = pose_error(np.array(r.pose_estimate_local), w["truth"][s])
        assert dt < 0.2 and dr < 0.02, (s, dt, dr)


def test_full_cloud_solve_on_a_cluster(ctx, orc, monkeypatch):
    """SURVEY 8d's mode F: adaptive filters opened up, the matcher sees every point of the second voxel filter (thousands per
    solve). The batched front end then spreads each problem over a thread-block cluster (partial normal equations through
    distributed shared memory): same result as one CTA per problem and as the oracle, same iteration counts."""
    import dliom
    w = workload()
    opts = orc.FrontEndOptions.defaults()
    opts.hi_min_num_points = 1e9
    opts.lo_min_num_points = 1e9
    fo = dliom.FrontendOptions.from_oracle(opts)
    hi, lo = dev_grid(ctx, w["hi"]), dev_grid(ctx, w["lo"])
    args = (fo, w["scans"], w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
    runs = {}
    for cs in ("1", "2", "8", None):        # None: the library picks (4 problems -> 8 CTAs each)
        if cs is None:
            monkeypatch.delenv("DLIOM_NLS_CLUSTER", raising=False)
        else:
            monkeypatch.setenv("DLIOM_NLS_CLUSTER", cs)
        runs[cs] = ctx.frontend_match_batch(*args)
    for s, one in enumerate(runs["1"]):
        assert one.ok == 1 and one.num_high_resolution + one.num_low_resolution > 3000
        for cs in ("2", "8", None):
            r = runs[cs][s]
            dt, dr = pose_error(np.array(r.pose_estimate_local), np.array(one.pose_estimate_local))
            assert r.ok == 1 and dt < 1e-9 and dr < 1e-10, (cs, s, dt, dr)
            assert r.summary.num_iterations == one.summary.num_iterations
            assert abs(r.summary.final_cost - one.summary.final_cost) <= 1e-9 * abs(one.summary.final_cost)
        ing = orc.ingest_scan(opts, w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        want = orc.match_scan(opts, ing["returns_tracking"], ing["current_pose"].astype(np.float64), w["submap_pose"], w["hi"], w["lo"])
        assert (one.num_high_resolution, one.num_low_resolution) == (len(want["hi_keep"]), len(want["lo_keep"]))
        dt, dr = pose_error(np.array(runs[None][s].pose_estimate_local), want["pose_estimate_local"])
        assert dt < 1e-7 and dr < 1e-8, (s, dt, dr)


# ---------------------------------------------------------------- BASELINE.json configs[2] and configs[3] as parity cases
def test_config2_128_beam_fine_grid_correlative_then_ceres(ctx, orc):
    """configs[2]: 128-beam (~260k points) scan, 0.05 m HybridGrid, correlative + Ceres refine — the whole front end
    with use_online_correlative_scan_matching on, against the oracle (456 533 candidates: 7^3 x 11^3)."""
    import dliom
    w = workload(beams=128, num_map_scans=6, num_scans=1, hi_res=0.05)
    opts = orc.FrontEndOptions.defaults(use_rtcsm=1)
    fo = dliom.FrontendOptions.from_oracle(opts)
    hi, lo = dev_grid(ctx, w["hi"]), dev_grid(ctx, w["lo"])
    assert len(w["scans"][0]) > 250000
    r = ctx.frontend_match_batch(fo, w["scans"], w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)[0]
    ing = orc.ingest_scan(opts, w["scans"][0], w["origin"], w["prev"][0], w["cur"][0])
    want = orc.match_scan(opts, ing["returns_tracking"], ing["current_pose"].astype(np.float64), w["submap_pose"],
                          w["hi"], w["lo"])
    assert (r.num_first_filter, r.num_returns) == (len(ing["first_keep"]), len(ing["returns_tracking"]))
    assert (r.num_high_resolution, r.num_low_resolution) == (len(want["hi_keep"]), len(want["lo_keep"]))
    assert np.float32(r.rtcsm_score) == np.float32(want["rtcsm_score"])      # correlative stage: bit-exact
    dt, dr = pose_error(np.array(r.pose_estimate_local), want["pose_estimate_local"])
    assert r.ok == 1 and dt < 1e-7 and dr < 1e-8


def test_config3_many_submaps_in_one_batch(ctx, orc):
    """configs[3] shape: one scan registered against MANY active submaps in a single launch (what ConstraintBuilder3D
    farms out to its thread pool): 8 different (hi, lo) grid pairs, one problem each, checked against the oracle."""
    w = workload()
    ing = orc.ingest_scan(w["opts"], w["scans"][0], w["origin"], w["prev"][0], w["cur"][0])
    pts = ing["returns_tracking"]
    hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
    lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
    rng = np.random.RandomState(8)
    ogrids, dgrids, inits, targets = [], [], [], []
    for k in range(8):   # submaps = the map grids shifted by whole voxels (different content per problem)
        shift_hi, shift_lo = rng.randint(-3, 4, 3), rng.randint(-1, 2, 3)
        pair = []
        for src, res, sh in ((w["hi"], 0.1, shift_hi), (w["lo"], 0.45, shift_lo)):
            g = orc.Grid(res)
            xs, ys, zs, vs = src.export()
            keep = rng.rand(len(xs)) < 0.8
            for x, y, z, v in zip(xs[keep], ys[keep], zs[keep], vs[keep]):
                g.set_value((x + sh[0], y + sh[1], z + sh[2]), v)
            pair.append(g)
        ogrids.append(pair)
        dgrids.append([dev_grid(ctx, pair[0]), dev_grid(ctx, pair[1])])
        init = w["cur"][0].copy()
        init[:3] += shift_hi * 0.1
        inits.append(init)
        targets.append(init[:3].copy())
    poses, sums = ctx.ceres_match_batch([[pts[hk], pts[lk]]] * 8, dgrids, [1.0, 6.0], 5.0, 4e2, np.array(targets), np.array(inits))
    for k in range(8):
        want, ws = orc.ceres_match([pts[hk], pts[lk]], ogrids[k], [1.0, 6.0], 5.0, 4e2, targets[k], inits[k])
        dt, dr = pose_error(poses[k], want)
        assert dt < 1e-7 and dr < 1e-8, k
        assert sums[k]["num_iterations"] == ws["num_iterations"]
