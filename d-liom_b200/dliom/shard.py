"""Multi-GPU plumbing for the batched front end: one process per GPU, registration problems are independent, so the
work list is PARTITIONED across ranks and nothing crosses NVLink on the data path. torch.distributed is
used for the rendezvous, the barrier around timed regions and the max-over-ranks of the device time only."""


def shard_range(num_items, rank, world):
    """Contiguous, balanced partition of range(num_items): the first (num_items % world) ranks get one extra item."""
    base, extra = divmod(num_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def max_over_ranks(dist, value, device=None):
    """MAX all-reduce of a python float (the contract times a multi-GPU step as the slowest rank's device time)."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_results(dist, local_rows, device=None):
    """All-gather of per-rank result rows (numpy 2-D float64, same column count) -> list ordered by rank."""
    import numpy as np
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [np.asarray(local_rows)]
    world = dist.get_world_size()
    rows = torch.as_tensor(np.ascontiguousarray(local_rows, np.float64), device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([rows.shape[0]], dtype=torch.int64, device=device))
    width = rows.shape[1]
    biggest = int(max(int(c[0]) for c in counts))
    padded = torch.zeros((biggest, width), dtype=torch.float64, device=device)
    padded[:rows.shape[0]] = rows
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return [o[:int(c[0])].cpu().numpy() for o, c in zip(out, counts)]


# ---- loop closure (BASELINE configs[3]/[4]): (node, submap) searches are sharded by SUBMAP OWNER so that every
# grid is resident on exactly one GPU; the only exchange step is the all-gather of the constraint records every rank
# needs for the (replicated, host-side) pose graph — ncclAllGather over NVLink on the GPU box, gloo in the CPU tests.
CONSTRAINT_COLUMNS = ("submap_id", "node_id", "score", "low_resolution_score", "x", "y", "z", "qw", "qx", "qy", "qz",
                      "translation_weight", "rotation_weight")


def owner_of_submap(submap_id, world):
    return int(submap_id) % world


def shard_by_owner(submap_ids, rank, world):
    """Indices of the pairs this rank searches: those whose submap it owns. Order preserved."""
    return [i for i, s in enumerate(submap_ids) if owner_of_submap(s, world) == rank]


def constraint_rows(submap_ids, node_ids, constraints):
    """Found constraints (dliom.Constraint or anything with the same attributes) -> float64 rows in CONSTRAINT_COLUMNS
    order; pruned pairs produce no row, exactly like ComputeConstraint leaving *constraint null."""
    import numpy as np
    rows = [[s, n, c.score, c.low_resolution_score, *c.pose[:], c.translation_weight, c.rotation_weight]
            for s, n, c in zip(submap_ids, node_ids, constraints) if c.found]
    return np.array(rows, np.float64).reshape(-1, len(CONSTRAINT_COLUMNS))


def all_gather_constraints(dist, local_rows, device=None):
    """Every rank ends up with the same, canonically ordered (submap_id, node_id) constraint table."""
    import numpy as np
    rows = np.concatenate(gather_results(dist, local_rows, device))
    order = np.lexsort((rows[:, 1], rows[:, 0]))
    return rows[order]
