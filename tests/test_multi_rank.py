"""world_size-2 gloo test (CPU) of the N>1 host logic: the problem list is partitioned across ranks with no overlap
and no gap, per-rank results gather back in rank order, and the step time is the max over ranks."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "d-liom_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"),
                    os.path.join(ROOT, "tools")]
    import torch.distributed as dist
    from dliom import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import orc
        from helpers import workload
        w = workload(beams=16, num_map_scans=4, num_scans=5)
        mine = shard.shard_range(len(w["scans"]), rank, world)
        rows = []
        for s in mine:   # each rank registers only its own scans (with the CPU oracle here: no GPU in this test)
            ing = orc.ingest_scan(w["opts"], w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
            m = orc.match_scan(w["opts"], ing["returns_tracking"], ing["current_pose"].astype(np.float64),
                               w["submap_pose"], w["hi"], w["lo"])
            rows.append(np.concatenate([[s], m["pose_estimate_local"]]))
        gathered = shard.gather_results(dist, np.array(rows).reshape(-1, 8))
        slowest = shard.max_over_ranks(dist, 10.0 + rank)
        if rank == 0:
            np.save(os.path.join(tmp, "gathered.npy"), np.concatenate(gathered))
            np.save(os.path.join(tmp, "slowest.npy"), np.array([slowest]))
    finally:
        dist.destroy_process_group()


def test_partition_is_exact():
    sys.path.insert(0, os.path.join(ROOT, "d-liom_b200"))
    from dliom import shard
    for n in (0, 1, 5, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            parts = [list(shard.shard_range(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_ranks_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = np.load(tmp_path / "gathered.npy")
    assert g[:, 0].astype(int).tolist() == [0, 1, 2, 3, 4]     # rank order == scan order, nothing lost or duplicated
    assert np.load(tmp_path / "slowest.npy")[0] == 11.0          # max over ranks
    # same poses as a single-process run
    sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
    import orc
    from helpers import workload
    w = workload(beams=16, num_map_scans=4, num_scans=5)
    for s in range(5):
        ing = orc.ingest_scan(w["opts"], w["scans"][s], w["origin"], w["prev"][s], w["cur"][s])
        m = orc.match_scan(w["opts"], ing["returns_tracking"], ing["current_pose"].astype(np.float64), w["submap_pose"],
                           w["hi"], w["lo"])
        assert np.array_equal(g[s, 1:], m["pose_estimate_local"])


class _Hit:
    def __init__(self, found, score, pose):
        self.found, self.score, self.low_resolution_score, self.pose = found, score, 0.5, list(pose)
        self.translation_weight, self.rotation_weight = 1.1e4, 1e5


def _constraint_worker(rank, world, port, tmp):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "d-liom_b200")]
    import torch.distributed as dist
    from dliom import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        submaps, nodes, hits = _constraint_problem()
        mine = shard.shard_by_owner(submaps, rank, world)
        assert all(shard.owner_of_submap(submaps[i], world) == rank for i in mine)
        rows = shard.constraint_rows([submaps[i] for i in mine], [nodes[i] for i in mine], [hits[i] for i in mine])
        table = shard.all_gather_constraints(dist, rows)
        biggest = max(len(shard.shard_by_owner(submaps, r, world)) for r in range(world))
        one_call = shard.all_gather_constraints(dist, rows, max_rows=biggest)   # single fixed-size collective
        assert np.array_equal(table, one_call)
        np.save(os.path.join(tmp, f"table{rank}.npy"), table)
    finally:
        dist.destroy_process_group()


def _constraint_problem():
    rng = np.random.RandomState(7)
    submaps = [int(s) for s in rng.randint(0, 5, 23)]
    nodes = list(range(100, 123))
    hits = [_Hit(bool(rng.rand() < 0.6), float(rng.rand()), rng.randn(7)) for _ in submaps]
    return submaps, nodes, hits


def test_constraint_all_gather_two_ranks_gloo(tmp_path):
    """Loop-closure sharding: pairs go to the rank that owns the submap, pruned pairs vanish, and after the one exchange
    step every rank holds the same table a single process would have produced."""
    sys.path.insert(0, os.path.join(ROOT, "d-liom_b200"))
    from dliom import shard
    port = 31500 + os.getpid() % 2000
    mp.spawn(_constraint_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "table0.npy"), np.load(tmp_path / "table1.npy")
    submaps, nodes, hits = _constraint_problem()
    single = shard.all_gather_constraints(None, shard.constraint_rows(submaps, nodes, hits))
    assert np.array_equal(t0, t1) and np.array_equal(t0, single)
    assert len(single) == sum(h.found for h in hits) and t0.shape[1] == len(shard.CONSTRAINT_COLUMNS)
    assert np.all(np.diff(t0[:, 0]) >= 0)


def _broadcast_worker(rank, world, port, tmp):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "d-liom_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"),
                    os.path.join(ROOT, "tools")]
    import torch.distributed as dist
    from dliom import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import orc
        cells = None
        if rank == 1:   # rank 1 owns the finished submap
            from helpers import workload
            cells = workload(beams=16, num_map_scans=3, num_scans=1)["hi"].export()
        xs, ys, zs, vs = shard.broadcast_cells(dist, cells, src=1)
        g = orc.Grid(0.1)   # every rank rebuilds the grid from the broadcast cells
        for x, y, z, v in zip(xs[:2000], ys[:2000], zs[:2000], vs[:2000]):
            g.set_value((int(x), int(y), int(z)), int(v))
        ragged = shard.all_gather_ragged(dist, np.full((3 + rank, 2), float(rank), np.float32))
        np.savez(os.path.join(tmp, f"cells{rank}.npz"), xs=xs, ys=ys, zs=zs, vs=vs,
                 probe=np.array([g.value((int(xs[i]), int(ys[i]), int(zs[i]))) for i in range(0, 2000, 97)]),
                 ragged0=ragged[0], ragged1=ragged[1])
    finally:
        dist.destroy_process_group()


def test_submap_broadcast_two_ranks_gloo(tmp_path):
    """A finished submap crosses ranks once, as cells in the ToProto layout; ragged node data is all-gathered."""
    port = 33500 + os.getpid() % 2000
    mp.spawn(_broadcast_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "cells0.npz"), np.load(tmp_path / "cells1.npz")
    for k in ("xs", "ys", "zs", "vs", "probe", "ragged0", "ragged1"):
        assert np.array_equal(a[k], b[k]), k
    assert len(a["xs"]) > 2000 and a["vs"].dtype == np.uint16 and a["vs"].max() < 32768
    assert np.array_equal(a["probe"], a["vs"][0:2000:97])
    assert a["ragged0"].shape == (3, 2) and a["ragged1"].shape == (4, 2) and a["ragged1"][0, 0] == 1.0
