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


def all_gather_constraints(dist, local_rows, device=None, max_rows=None):
    """Every rank ends up with the same, canonically ordered (submap_id, node_id) constraint table.

    max_rows: an upper bound on any rank's row count that every rank can compute without talking (the size of the
    largest shard of the pair list). With it the exchange is ONE all-gather of a fixed-size block whose first row carries
    the count; without it the counts are exchanged first (two collectives)."""
    import numpy as np
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1 or max_rows is None:
        rows = np.concatenate(gather_results(dist, local_rows, device))
    else:
        world = dist.get_world_size()
        local = np.asarray(local_rows, np.float64).reshape(-1, len(CONSTRAINT_COLUMNS))
        if len(local) > max_rows:
            raise ValueError("max_rows is smaller than this rank's row count")
        block = np.zeros((max_rows + 1, local.shape[1]), np.float64)
        block[0, 0] = len(local)
        block[1:1 + len(local)] = local
        mine = torch.as_tensor(block, device=device)
        out = torch.empty((world * mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(out, mine)
        got = out.cpu().numpy().reshape(world, mine.shape[0], mine.shape[1])
        rows = np.concatenate([got[r, 1:1 + int(got[r, 0, 0])] for r in range(world)])
    order = np.lexsort((rows[:, 1], rows[:, 0]))
    return rows[order]


def broadcast_cells(dist, cells, src, device=None):
    """A finished submap moves between ranks once: its cells in the HybridGrid::ToProto layout
    (x, y, z int32 + uint16 value, what Grid.export() returns and Grid.set_cells() takes) are broadcast from `src`.
    `cells` is ignored on the other ranks. One size broadcast + one payload broadcast (int32 x 4 per cell)."""
    import numpy as np
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tuple(np.asarray(c) for c in cells)
    rank = dist.get_rank()
    count = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        xs, ys, zs, vs = cells
        count[0] = len(xs)
    dist.broadcast(count, src)
    n = int(count[0])
    payload = torch.zeros((4, max(n, 1)), dtype=torch.int32, device=device)
    if rank == src and n:
        payload[:, :n] = torch.as_tensor(np.stack([np.asarray(xs, np.int32), np.asarray(ys, np.int32), np.asarray(zs, np.int32),
                                                   np.asarray(vs, np.uint16).astype(np.int32)]), device=device)
    dist.broadcast(payload, src)
    got = payload.cpu().numpy()[:, :n]
    return got[0].copy(), got[1].copy(), got[2].copy(), got[3].astype(np.uint16)


def all_gather_ragged(dist, local_rows, device=None):
    """All-gather of per-rank float32 row blocks of different lengths (node clouds, pose guesses) -> list by rank."""
    import numpy as np
    import torch
    local = np.ascontiguousarray(local_rows, np.float32)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world = dist.get_world_size()
    flat = torch.as_tensor(local.reshape(-1), device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([flat.numel()], dtype=torch.int64, device=device))
    biggest = max(int(c[0]) for c in counts)
    padded = torch.zeros(max(biggest, 1), dtype=torch.float32, device=device)
    padded[:flat.numel()] = flat
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    width = local.shape[1] if local.ndim == 2 else 1
    return [o[:int(c[0])].cpu().numpy().reshape(-1, width) if local.ndim == 2 else o[:int(c[0])].cpu().numpy()
            for o, c in zip(out, counts)]
