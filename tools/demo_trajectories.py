"""BASELINE configs[4] in miniature: one trajectory per GPU, everything of the path on the device.

  per rank:   build the trajectory's submap on the device (front-end ingest + range-data insertion, dl_submap_insert_range_data),
              register the following sweeps against it (dl_frontend_match_batch) -> nodes
  exchange 1: all-gather of the nodes' filtered clouds and poses (a few KB per node)
  per rank:   search every node of every trajectory against the rank's own submap (dl_constraint_search_batch):
              searches are sharded by submap owner, no grid ever moves
  exchange 2: all-gather of the constraint records -> identical table on every rank

    python tools/demo_trajectories.py                                   # 1 GPU, 1 trajectory
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/demo_trajectories.py

Prints one JSON line (rank 0) and exits non-zero if, at the reference's min_score (0.55), a node's constraint against its OWN
trajectory's submap is further than 0.2 m from the synthetic truth or one against ANOTHER trajectory's submap further than 0.3 m.
The same searches at --min-score (0.3) are reported next to it: that is where round 1's 4.8 m "inter-trajectory error" came from.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "d-liom_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "tools")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--beams", type=int, default=64)
    ap.add_argument("--map-scans", type=int, default=20)
    ap.add_argument("--nodes", type=int, default=6)
    ap.add_argument("--spacing", type=float, default=0.6, help="seconds between trajectory starts (10 m/s)")
    ap.add_argument("--min-score", type=float, default=0.3)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = dist if world > 1 else None
    dev = torch.device("cuda", local)
    import dliom
    import orc
    import synth
    from dliom import shard
    from helpers import apply_pose

    ctx = dliom.Context(local)
    scene = synth.Scene(42)
    opts = orc.FrontEndOptions.defaults()          # parameter block only; nothing of the oracle computes here
    fo = dliom.FrontendOptions.from_oracle(opts)
    origin = np.zeros((1, 3), np.float32)
    identity = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)

    # ---- 1. the trajectory's submap, built on the device
    t = 2.0 + args.spacing * rank
    hi, lo = ctx.grid(0.1), ctx.grid(0.45)
    t0 = time.perf_counter()
    for _ in range(args.map_scans):
        rows = synth.make_scan(scene, args.beams, t)
        cur = synth.pose7(t)
        ing = ctx.ingest_scan(fo, rows, origin, synth.pose7(t - 0.1), cur)
        local_pts = apply_pose(cur, ing["returns_tracking"].astype(np.float64)).astype(np.float32)
        ctx.submap_insert_range_data(hi, lo, identity, cur[:3].astype(np.float32), local_pts, high_resolution_max_range=20)
        t += 0.1
    build_s = time.perf_counter() - t0

    # ---- 2. nodes: the next sweeps, registered against the submap; their filtered clouds are what loop closure needs
    rng = np.random.RandomState(100 + rank)
    scans, prev, guess, truth = [], [], [], []
    for _ in range(args.nodes):
        scans.append(synth.make_scan(scene, args.beams, t))
        prev.append(synth.pose7(t - 0.1)); truth.append(synth.pose7(t))
        guess.append(synth.perturb_pose(synth.pose7(t), rng, 0.05, 0.5))
        t += 0.1
    t0 = time.perf_counter()
    res = ctx.frontend_match_batch(fo, scans, origin, np.array(prev), np.array(guess), identity, hi, lo)
    match_s = time.perf_counter() - t0
    front_err = max(float(np.abs(np.array(r.pose_estimate_local[:3]) - tr[:3]).max()) for r, tr in zip(res, truth))
    node_hi, node_lo = [], []
    for s, p, tr in zip(scans, prev, truth):
        pts = ctx.ingest_scan(fo, s, origin, p, tr)["returns_tracking"]
        hk, _ = ctx.adaptive_voxel_filter(pts, 2.0, 150, 15.0)
        lk, _ = ctx.adaptive_voxel_filter(pts, 4.0, 200, 60.0)
        node_hi.append(pts[hk]); node_lo.append(pts[lk])

    # ---- exchange 1: every rank gets every node (cloud rows tagged with their node number, poses)
    t0 = time.perf_counter()
    tag = lambda clouds: np.concatenate([np.column_stack([np.full(len(c), k, np.float32), c]) for k, c in enumerate(clouds)])
    all_hi = shard.all_gather_ragged(d, tag(node_hi), dev)
    all_lo = shard.all_gather_ragged(d, tag(node_lo), dev)
    all_truth = shard.all_gather_ragged(d, np.array(truth, np.float32), dev)   # float32 is plenty for a pose GUESS
    exchange_s = time.perf_counter() - t0

    # ---- 3. this rank's shard of the loop-closure searches: every node against the submap it owns
    pair_nodes, g7, hs, ls = [], [], [], []
    grng = np.random.RandomState(7)
    for r in range(len(all_truth)):
        for k in range(len(all_truth[r])):
            node_id = 1000 * r + k
            gg = all_truth[r][k].astype(np.float64)
            gg[:3] += grng.uniform(-1, 1, 3) * [2.0, 2.0, 0.4]
            pair_nodes.append((node_id, all_truth[r][k].astype(np.float64)))
            g7.append(gg)
            hs.append(all_hi[r][all_hi[r][:, 0] == k][:, 1:]); ls.append(all_lo[r][all_lo[r][:, 0] == k][:, 1:])
    # The asserted table uses the reference's min_score (pose_graph.lua: 0.55); --min-score (default 0.3) is run as well and only
    # REPORTED: in this corridor-like street (facades parallel to the driving direction) a node that sees only the far end of a
    # neighbouring submap has a nearly flat score along the street, and below ~0.5 the best leaf can sit metres away along x —
    # round 1's "4.8 m inter-trajectory error" was exactly that (error along x, score 0.3-0.45), not a frame bug in the exchange.
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.tensor(list(dliom.comm_unique_id()), dtype=torch.uint8))
    if world > 1:
        dist.broadcast(idt, 0)
    comm = dliom.Comm(ctx, bytes(idt.cpu().numpy().tolist()), rank, world)
    node_ids = [n for n, _ in pair_nodes]
    truth_of = {n: tr for n, tr in pair_nodes}

    def search(min_score):
        opt = dliom.ConstraintOptions.defaults(min_score=min_score, min_low_resolution_score=min(0.55, max(0.3, min_score)))
        t0 = time.perf_counter()
        # searches sharded by submap owner (this rank owns submap `rank`) + ONE ncclAllGather of the rows, issued from the C-ABI
        table, info = ctx.constraint_search_exchange(comm, opt, len(g7), [rank] * len(g7), node_ids, g7, hs, ls, [hi] * len(g7),
                                                     [lo] * len(g7))
        return table, info, time.perf_counter() - t0
    search(0.55)                          # warm-up: NCCL sets its connections up lazily on the first collective
    table, info, search_s = search(0.55)
    loose, _, _ = search(args.min_score)

    def errors(tab):
        own, other = [], []
        for r in tab:
            if r.found != 1:
                continue
            tr = truth_of[r.node_id]
            e = np.array(r.pose[:3]) - tr[:3]
            (own if r.node_id // 1000 == r.submap_id else other).append((float(np.abs(e).max()), [float(v) for v in e], float(r.score)))
        return own, other
    own, other = errors(table)
    own_l, other_l = errors(loose)
    worst = max([e[0] for e in own], default=0.0)
    worst_other = max([e[0] for e in other], default=0.0)
    worst_loose = max(other_l, default=(0.0, [0, 0, 0], 0.0))
    stats = torch.tensor([worst, front_err, build_s, match_s, exchange_s, search_s, info.collective_ms, worst_other], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    worst, front_err, worst_other = float(stats[0]), float(stats[1]), float(stats[7])
    if rank == 0:
        print(json.dumps({"demo": "configs[4] shape: %d trajectories x (%d map sweeps + %d nodes), %d-beam" %
                                  (world, args.map_scans, args.nodes, args.beams), "n_gpus": world,
                          "searches": world * len(g7), "min_score": 0.55,
                          "constraints": len(own) + len(other), "intra_trajectory": len(own), "inter_trajectory": len(other),
                          "max_constraint_error_m": worst, "max_inter_trajectory_constraint_error_m": worst_other,
                          "max_front_end_error_m": front_err,
                          "loose_threshold": {"min_score": args.min_score, "inter_trajectory": len(other_l),
                                              "worst_inter_error_m": worst_loose[0], "worst_inter_error_xyz": worst_loose[1],
                                              "score_of_worst": worst_loose[2],
                                              "note": "reported only: low-score matches along the street's symmetry axis"},
                          "collective": {"name": "ncclAllGather (dl_constraint_search_exchange)", "bytes": int(info.bytes_received),
                                         "ms": float(stats[6])},
                          "seconds_max_over_ranks": {"build_submap": float(stats[2]), "front_end": float(stats[3]),
                                                     "node_exchange": float(stats[4]), "search_and_exchange": float(stats[5])}}))
    comm.close()
    if world > 1:
        dist.destroy_process_group()
    # every constraint that passes the reference's min_score must be right: own submap within 2 voxels, other trajectories' within 0.3 m
    return 0 if worst < 0.2 and worst_other < 0.3 and front_err < 0.1 else 1


if __name__ == "__main__":
    sys.exit(main())
