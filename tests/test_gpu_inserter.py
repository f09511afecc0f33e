"""GPU parity of the grid WRITE side: RangeDataInserter3D::Insert / Submap3D::InsertRangeData executed on
the device grid, compared cell for cell (uint16 values, bit-exact) with the oracle's HybridGrid."""
import numpy as np
import pytest

from helpers import SEVEN, apply_pose, pose_error, workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import dliom
    c = dliom.Context(0)
    yield c
    c.close()


def cells(export):
    xs, ys, zs, vs = export
    return {(int(x), int(y), int(z)): int(v) for x, y, z, v in zip(xs, ys, zs, vs)}


def test_reference_inserter_fixture(ctx, orc):
    """mapping/3d/range_data_inserter_3d_test.cc:30-103: hit 0.7 / miss 0.4 / 1000 free voxels."""
    og, g = orc.Grid(1.0), ctx.grid(1.0)
    origin = np.array([0, 0, -4], np.float32)
    returns = np.array([[-3, -1, 4], [-2, 0, 4], [-1, 1, 4], [0, 2, 4]], np.float32)
    for k in range(40):
        og.insert_range_data(origin, returns, hit=0.7, miss=0.4, num_free=1000)
        g.insert_range_data(origin, returns, hit=0.7, miss=0.4, num_free=1000)
        if k in (0, 1, 39):
            assert cells(g.export()) == cells(og.export())
    for p in returns:   # the reference's own assertions after many inserts: hits -> 0.9, ray cells -> 0.1
        assert abs(orc.lib().orc_value_to_probability(int(g.lookup([og.cell_index(p)])[0])) - 0.9) < 1e-3


def test_scene_submap_built_on_device_equals_oracle(ctx, orc):
    """Several sweeps inserted into EMPTY device grids (growth from 2^3 top cells, lock-free brick allocation, hits
    before misses, update markers cleared) -> exactly the cells the oracle builds; then the matcher reads them."""
    import synth
    scene = synth.Scene(42)
    opts = orc.FrontEndOptions.defaults()
    ohi, olo = orc.Grid(0.1), orc.Grid(0.45)
    dhi, dlo = ctx.grid(0.1), ctx.grid(0.45)
    origin0 = np.zeros((1, 3), np.float32)
    t = 2.0
    for k in range(5):
        rows = synth.make_scan(scene, 16, t)
        cur = synth.pose7(t)
        ing = orc.ingest_scan(opts, rows, origin0, synth.pose7(t - 0.1), cur)
        local = apply_pose(cur, ing["returns_tracking"].astype(np.float64)).astype(np.float32)
        o = cur[:3].astype(np.float32)
        d = local - o
        rng2 = (d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])).astype(np.float32)   # Eigen order, float32
        near = local[np.sqrt(rng2).astype(np.float32) <= np.float32(20.0)]
        ohi.insert_range_data(o, near)
        olo.insert_range_data(o, local)
        ctx.submap_insert_range_data(dhi, dlo, orc.IDENTITY_POSE, o, local, high_resolution_max_range=20)
        t += 0.1
    assert cells(dhi.export()) == cells(ohi.export())
    assert cells(dlo.export()) == cells(olo.export())
    assert ohi.bits() >= 3     # the grids really grew
    # the freshly written device grids serve the matcher without any upload
    rows = synth.make_scan(scene, 16, t)
    ing = orc.ingest_scan(opts, rows, origin0, synth.pose7(t - 0.1), synth.pose7(t))
    pts = ing["returns_tracking"]
    hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
    lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
    init = synth.pose7(t)
    want, _ = orc.ceres_match([pts[hk], pts[lk]], [ohi, olo], [1.0, 6.0], 5.0, 4e2, init[:3], init)
    got, _ = ctx.ceres_match([pts[hk], pts[lk]], [dhi, dlo], [1.0, 6.0], 5.0, 4e2, init[:3], init)
    dt, dr = pose_error(got, want)
    assert dt < 1e-7 and dr < 1e-7   # (2 acos(|<q1,q2>|) has a ~3e-8 floor in double)


def test_host_cells_after_device_insert(ctx, orc):
    """dl_grid_set_cells after a device-side insert first pulls the device copy back into the host mirror."""
    og, g = orc.Grid(0.5), ctx.grid(0.5)
    origin = np.array([0.2, -0.1, 0.3], np.float32)
    pts = (np.random.RandomState(4).normal(0, 6, (3000, 3))).astype(np.float32)
    og.insert_range_data(origin, pts)
    g.insert_range_data(origin, pts)
    og.set_value((70, -3, 2), 1234)
    g.set_cells([70], [-3], [2], [1234])
    og.insert_range_data(origin, pts[:500])
    g.insert_range_data(origin, pts[:500])
    assert cells(g.export()) == cells(og.export())


def test_inserter_argument_checks(ctx):
    import dliom
    g = ctx.grid(1.0)
    with pytest.raises(dliom.DlError):     # CHECK_GT(hit_probability, 0.5)
        g.insert_range_data(np.zeros(3, np.float32), SEVEN, hit=0.4, miss=0.3)
    with pytest.raises(dliom.DlError) as e:  # beyond +-8192 cells: CHECK_LE(new_bits, 8)
        g.insert_range_data(np.zeros(3, np.float32), np.array([[9000.0, 0, 0]], np.float32))
    assert e.value.status == -3
    g.insert_range_data(np.zeros(3, np.float32), np.zeros((0, 3), np.float32))   # empty: no-op
