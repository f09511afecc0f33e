"""The NCCL exchange steps issued from the C-ABI (csrc/dl_comm.cu), on one GPU: a world of one rank exercises the same
ncclCommInitRank / ncclAllGather / ncclAllReduce / ncclBroadcast calls as the 8-GPU box (bench.py --gpus N runs them with N
ranks). The constraint exchange must return exactly what dl_constraint_search_batch computes for the same pairs."""
import ctypes as C

import numpy as np
import pytest

from helpers import workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import dliom
    c = dliom.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def comm(ctx):
    import dliom
    c = dliom.Comm(ctx, dliom.comm_unique_id(), 0, 1)
    yield c
    c.close()


def test_row_layout():
    import dliom
    assert C.sizeof(dliom.ConstraintRow) == 96 and C.sizeof(dliom.ExchangeInfo) == 24


def test_device_collectives_world_of_one(ctx, comm):
    x = np.arange(1000, dtype=np.float64) * 0.5
    d_a, d_b = ctx.device_alloc(x.nbytes), ctx.device_alloc(x.nbytes)
    ctx.copy_to_device(d_a, x)
    comm.all_gather_dev(d_a, d_b, x.nbytes)
    comm.all_reduce_f64_dev(d_b, len(x))          # sum over one rank
    comm.broadcast_dev(d_b, x.nbytes, 0)
    ctx.synchronize()
    out = np.zeros_like(x)
    ctx.check(ctx.L.dl_copy_to_host(ctx.h, out.ctypes.data, d_b, out.nbytes))
    assert np.array_equal(out, x)
    ctx.device_free(d_a)
    ctx.device_free(d_b)


def test_constraint_exchange_equals_batch_search(ctx, comm, orc):
    import dliom
    w = workload(beams=16, num_map_scans=40, num_scans=3)
    hi, lo = dliom.Grid.from_oracle(ctx, w["hi"]), dliom.Grid.from_oracle(ctx, w["lo"])
    rng = np.random.default_rng(5)
    guesses, his, los, sub, node = [], [], [], [], []
    for k in range(3):
        pts = orc.ingest_scan(w["opts"], w["scans"][k], w["origin"], w["prev"][k], w["truth"][k])["returns_tracking"]
        hk, _ = orc.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = orc.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        for far in (False, True):
            g = np.array(w["truth"][k], np.float64)
            g[:3] += rng.uniform(-1, 1, 3) * [2.0, 2.0, 0.4] + (np.array([60.0, 0, 0]) if far else 0)
            guesses.append(g); his.append(pts[hk]); los.append(pts[lk]); sub.append(7); node.append(100 + 2 * k + int(far))
    opt = dliom.ConstraintOptions.defaults(min_score=0.15, min_low_resolution_score=0.3)
    want = ctx.constraint_search_batch(opt, guesses, his, los, [hi] * 6, [lo] * 6)
    table, info = ctx.constraint_search_exchange(comm, opt, 8, sub, node, guesses, his, los, [hi] * 6, [lo] * 6)
    assert len(table) == 8 and info.bytes_sent == 8 * 96 and info.bytes_received == 8 * 96 and info.collective_ms > 0
    assert info.found_total == sum(c.found for c in want) >= 1
    for k, (r, c) in enumerate(zip(table[:6], want)):
        assert (r.submap_id, r.node_id, r.rank) == (7, 100 + k, 0)
        assert r.found == c.found
        if c.found:
            assert r.score == c.score and r.low_resolution_score == c.low_resolution_score
            assert list(r.pose) == list(c.pose)
            assert (r.translation_weight, r.rotation_weight) == (c.translation_weight, c.rotation_weight)
    assert table[6].found == -1 and table[7].found == -1
    # an empty shard still takes part in the collective
    table0, info0 = ctx.constraint_search_exchange(comm, opt, 4, [], [], [], [], [], [], [])
    assert all(r.found == -1 for r in table0) and info0.found_total == 0
    with pytest.raises(dliom.DlError):   # more pairs than the agreed capacity
        ctx.constraint_search_exchange(comm, opt, 2, sub, node, guesses, his, los, [hi] * 6, [lo] * 6)
