#!/usr/bin/env python
"""bench.py — scans/sec of the scan-to-submap registration hot path (BASELINE.json metric) on N B200s.

A "step" is one pass of the whole per-scan hot path (first voxel filter -> deskew/transform/range gate -> second
voxel filters -> adaptive voxel filters -> Levenberg-Marquardt point-to-grid match) over one batch of synthetic
64-beam scans (configs[1]: 64-beam ~130k pts/scan, single submap, 1 x B200) against one 0.1 m / 0.45 m submap.

  value  scans/s with the batch already resident in HBM (dl_frontend_match_batch_dev), device-timed with CUDA
         events on the library's stream, max over ranks.
  e2e    the same metric through the C-ABI call that takes HOST buffers (dl_frontend_match_batch): pinned host
         scans copied to the device and results copied back inside the timed region (wall clock around the call).
  roofline       achieved algorithmic GB/s of the dominant stage (per-stage CUDA events) vs the measured HBM peak.
  cpu_baseline   the oracle (CPU restatement of the reference path) on the host cores, bounded sample (N=1, rank 0).

`--impl reference` times that CPU path alone (all host threads) and prints the same line with "impl": "reference".
Multi-GPU: one process per GPU (torchrun), scans are independent -> sharded across ranks, no data-path collective,
"scaling": "weak". Inputs per step exceed L2 (148 scans x 2.1 MB = 309 MB > 126 MB), so no explicit L2 flush.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "d-liom_b200"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "scans/sec (64-beam, 10 Hz) per GPU; pose RMSE vs reference CPU"
UNIT = "scans/s"


def oracle():
    """The CPU checker / baseline. Imported lazily and only on the cpu_baseline and --impl reference legs."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    orc.lib()
    return orc


def apply_pose(p7, pts):
    w, x, y, z = p7[3:]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return pts @ R.T + p7[:3]


def build_workload(args, rank):
    """Submap cells (built with the oracle's range-data inserter, like the reference builds a submap: hit 0.55 /
    miss 0.49 / 2 free voxels, high-res max range 20 m) + the batch of scans to register. Deterministic."""
    import synth
    orc = oracle()
    scene = synth.Scene(42)
    opts = orc.FrontEndOptions.defaults()
    hi, lo = orc.Grid(0.1), orc.Grid(0.45)
    origin = np.zeros((1, 3), np.float32)
    t0 = 2.0
    for k in range(args.map_scans):
        t = t0 + 0.1 * k
        rows = synth.make_scan(scene, args.beams, t)
        cur = synth.pose7(t)
        ing = orc.ingest_scan(opts, rows, origin, synth.pose7(t - 0.1), cur)
        local = apply_pose(cur, ing["returns_tracking"].astype(np.float64)).astype(np.float32)
        o = cur[:3].astype(np.float32)
        hi.insert_range_data(o, local[np.linalg.norm(local - o, axis=1) <= 20.0])
        lo.insert_range_data(o, local)
    rng = np.random.RandomState(45 + rank)
    distinct = min(args.batch, args.distinct_scans)
    scans, prevs, curs, truths = [], [], [], []
    for j in range(distinct):
        # sweeps interleaved with the map sweeps (half a period later), per-rank offset -> different data per GPU
        t = t0 + 0.05 + 0.1 * ((j * 7 + rank * 3) % max(args.map_scans - 1, 1)) + 0.001 * rank
        scans.append(synth.make_scan(scene, args.beams, t))
        prevs.append(synth.pose7(t - 0.1))
        truths.append(synth.pose7(t))
        curs.append(synth.perturb_pose(synth.pose7(t), rng, 0.1, 1.0))
    idx = [j % distinct for j in range(args.batch)]
    return {"orc": orc, "opts": opts, "hi": hi, "lo": lo, "origin": origin,
            "scans": [scans[i] for i in idx], "prev": np.array([prevs[i] for i in idx]),
            "cur": np.array([curs[i] for i in idx]), "truth": np.array([truths[i] for i in idx]),
            "submap_pose": orc.IDENTITY_POSE.copy()}


def pose_errors(a, b):
    dt = np.linalg.norm(a[:, :3] - b[:, :3], axis=1)
    qa = a[:, 3:] / np.linalg.norm(a[:, 3:], axis=1, keepdims=True)
    qb = b[:, 3:] / np.linalg.norm(b[:, 3:], axis=1, keepdims=True)
    d = np.abs(np.sum(qa * qb, axis=1)).clip(max=1.0)
    return dt, 2 * np.arccos(d)


class ClockSampler(threading.Thread):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 9:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[1]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][2]), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle port) on all host threads."""
    if rank != 0:
        return
    w = build_workload(args, 0)
    orc = w["orc"]
    threads = os.cpu_count() or 1
    sample = args.batch  # the whole batch is a bounded sample already (~30 ms of CPU work per scan)
    sel = slice(0, sample)
    times = []
    for step in range(args.warmup + args.steps):
        secs, poses, ok = orc.frontend_batch(w["opts"], w["scans"][sel], w["origin"], w["prev"][sel], w["cur"][sel],
                                             w["submap_pose"], w["hi"], w["lo"], threads)
        if step >= args.warmup:
            times.append(secs)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, args.batch),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} scans per step, {args.steps} steps, one scan per host thread"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, batch):
    return {"workload": f"configs[1]: {args.beams}-beam scans (~130k pts) vs one submap (0.1 m / 0.45 m), whole front-end "
                        f"hot path, pipeline-faithful filters", "scans_per_step_per_gpu": batch, "beams": args.beams,
            "map_scans": args.map_scans, "row_bytes": 4 * args.row_floats,
            "l2_policy": f"inputs ({batch} x {130605 * 4 * args.row_floats / 1e6:.1f} MB) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"scans sharded over {args.gpus} gpu(s), no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=148, help="scans per step per GPU (default: one per SM)")
    ap.add_argument("--beams", type=int, default=64)
    ap.add_argument("--map-scans", type=int, default=40)
    ap.add_argument("--distinct-scans", type=int, default=16)
    ap.add_argument("--cpu-sample", type=int, default=0, help="scans in the cpu_baseline sample (0 = 2 x threads)")
    ap.add_argument("--row-floats", type=int, default=4, choices=[4, 8],
                    help="4: TimedPointCloud rows x y z t (what AddRangeData receives); 8: RangeMeasurement rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import dliom
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = build_workload(args, rank)
    B = args.batch
    ctx = dliom.Context(local_rank)
    hi, lo = ctx.grid(0.1), ctx.grid(0.45)
    hi.set_cells(*w["hi"].export())
    lo.set_cells(*w["lo"].export())
    fo = dliom.FrontendOptions.from_oracle(w["opts"])
    fo.range_row_floats = args.row_floats
    if not os.environ.get("DLIOM_BENCH_PER_SCAN_COPIES"):
        fo.host_scan_stride_rows = -1   # set below once the staging layout is known
    row_bytes = 4 * args.row_floats

    sizes = np.array([len(s) for s in w["scans"]], np.int64)
    cap = int(sizes.max())
    # pinned host staging (e2e) and the HBM-resident copy (value)
    host = torch.zeros((B, cap, row_bytes), dtype=torch.uint8).pin_memory()
    for b, s in enumerate(w["scans"]):
        host[b, :len(s)] = torch.from_numpy(s.view(np.uint8).reshape(-1, 32)[:, :row_bytes].copy())
    dev = host.to(f"cuda:{local_rank}")
    results_dev = torch.zeros(B * C.sizeof(dliom.ScanResult), dtype=torch.uint8, device=f"cuda:{local_rank}")
    if fo.host_scan_stride_rows:
        fo.host_scan_stride_rows = cap   # the pinned staging tensor is (B, cap, row_bytes): one allocation, constant stride
    host_rows = dliom.HostScanBatch([host[b, :int(sizes[b])].numpy() for b in range(B)])
    # second context + second pinned copy of the scans for the streaming e2e loop (double buffering)
    ctx2 = dliom.Context(local_rank)
    host2 = host.clone().pin_memory()
    host_rows2 = dliom.HostScanBatch([host2[b, :int(sizes[b])].numpy() for b in range(B)])
    stream = torch.cuda.ExternalStream(ctx.stream, device=f"cuda:{local_rank}")

    results_dev2 = torch.zeros_like(results_dev)
    dev_lanes = [(ctx, results_dev), (ctx2, results_dev2)]
    if os.environ.get("DLIOM_BENCH_ONE_CONTEXT"):
        dev_lanes = [dev_lanes[0]]
    for _ in range(int(os.environ.get("DLIOM_BENCH_CONTEXTS", "2")) - 2):   # experiments: deeper rotation of contexts
        dev_lanes.append((dliom.Context(local_rank), torch.zeros_like(results_dev)))
    extra_ctx = [c for c, _ in dev_lanes[2:]]

    def step_dev(i=0):
        """One pass of the hot path over the HBM-resident batch. Successive steps alternate between two contexts (own
        streams, own scratch), so the latency-bound back half of step i overlaps the front half of step i+1."""
        c, out = dev_lanes[i % len(dev_lanes)]
        c.frontend_match_batch_dev(fo, C.c_void_p(dev.data_ptr()), cap, sizes, w["origin"], w["prev"], w["cur"],
                                   w["submap_pose"], hi, lo, C.c_void_p(out.data_ptr()))

    def step_e2e():
        return ctx.frontend_match_batch(fo, host_rows, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)

    lanes = [(ctx, host_rows), (ctx2, host_rows2)]

    def run_streaming(steps):
        """K batches through dl_frontend_submit / dl_frontend_collect on two alternating contexts: batch i+1 is uploading
        while batch i computes; every batch's inputs cross PCIe and every batch's results are read back."""
        out = None
        for i in range(steps):
            c, rows = lanes[i & 1]
            if i >= 2:
                out = c.frontend_collect()
            c.frontend_submit(fo, rows, w["origin"], w["prev"], w["cur"], w["submap_pose"], hi, lo)
        for i in range(max(steps - 2, 0), steps):
            out = lanes[i & 1][0].frontend_collect()
        return out

    def barrier():
        torch.cuda.synchronize()
        ctx.synchronize()
        ctx2.synchronize()
        for c in extra_ctx:
            c.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (both paths), then parity of the batch against the oracle on a sample
    warm = max(args.warmup, 1)   # at least one untimed pass: scratch arenas are sized on first use and `res` below needs a result
    for k in range(2 * warm):
        step_dev(k)
    for _ in range(warm):
        step_e2e()
    run_streaming(max(warm, 2))
    ctx.synchronize()
    res = ctx.fetch_results(C.c_void_p(results_dev.data_ptr()), B)

    # ---- timed: device-resident
    ctx.set_profiling(True)
    ctx.read_profile()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches + ctx2.launches + sum(c.launches for c in extra_ctx)
    barrier()
    # CUDA events on the launching streams: the first context's stream opens the region; the closing event is recorded on
    # the same stream after it has been made to wait for the other context's stream (an event wait, no host sync).
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    others = [torch.cuda.ExternalStream(c.stream, device=f"cuda:{local_rank}") for c in [ctx2] + extra_ctx]
    gate = torch.cuda.Event()
    gate.record(stream)
    for s2 in others:
        s2.wait_event(gate)           # no context starts before e0
    e0.record(stream)
    for k in range(args.steps):
        step_dev(k)
    for s2 in others:
        tail = torch.cuda.Event()
        tail.record(s2)
        stream.wait_event(tail)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launches + ctx2.launches + sum(c.launches for c in extra_ctx) - launches0
    profile = ctx.read_profile()
    ctx.set_profiling(False)
    # ---- timed: end to end (host buffers in, results out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = step_e2e()
    barrier()
    e2e_sync_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    res_stream = run_streaming(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    stream_equal = all(list(a.pose_estimate_local) == list(c.pose_estimate_local) and a.ok == c.ok and
                       a.num_returns == c.num_returns for a, c in zip(res_stream, res_e2e))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # ---- latency of ONE scan through the blocking call (what a 10 Hz single-trajectory node sees)
    one = dliom.HostScanBatch([host[0, :int(sizes[0])].numpy()])
    lat = []
    for _ in range(25):
        t1 = time.perf_counter()
        ctx.frontend_match_batch(fo, one, w["origin"], w["prev"][:1], w["cur"][:1], w["submap_pose"], hi, lo)
        lat.append((time.perf_counter() - t1) * 1e3)
    single_scan_ms = float(np.median(lat[5:]))
    # ---- the PCIe ceiling of the e2e number: the same pinned bytes copied with nothing else running
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    c0.record()
    for _ in range(5):
        dev.copy_(host, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 5 * host.numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9

    t = torch.tensor([ms_total, e2e_s * 1e3, e2e_sync_s * 1e3], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total, e2e_sync_ms_total = float(t[0]), float(t[1]), float(t[2])
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)
    e2e_value = world * B / (e2e_ms_total / args.steps / 1e3)

    if rank == 0:
        # ---- roofline of the dominant stage from the per-stage CUDA-event times (algorithmic bytes)
        n_raw = float(sizes.sum())
        n1 = sum(r.num_first_filter for r in res)
        n_ret_local = n1  # classify touches every first-filter survivor
        n2 = sum(r.num_returns for r in res)
        evals = sum(r.summary.num_evaluations * (r.num_high_resolution + r.num_low_resolution) for r in res)
        adaptive_bytes = sum(12.0 * (r.num_cropped_high * r.num_passes_high + r.num_cropped_low * r.num_passes_low) +
                             12.0 * (r.num_high_resolution + r.num_low_resolution) for r in res)
        stage_bytes = {
            "voxel_filter_first": 16.0 * n_raw + 16.0 * n1,                          # 16 N_in + 16 N_out
            "ingest_second_filter": 28.0 * n_ret_local + 12.0 * n1 + 12.0 * n2,      # ingest 28 N + second pass 12 N_in + 12 N_out
            "adaptive_voxel_filter": adaptive_bytes,
            "nls_solve": 28.0 * evals,
        }
        stages = {k: {"ms_per_step": v[0] / max(v[1], 1), "bytes_per_step": stage_bytes.get(k)} for k, v in profile.items()}
        dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "of measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "of fallback (6.65 TB/s)"
        achieved = (stages[dom]["bytes_per_step"] or 0.0) / (stages[dom]["ms_per_step"] * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                    "stages": {k: {"ms_per_step": round(v["ms_per_step"], 4),
                                   "gbps": None if not v["bytes_per_step"] else round(v["bytes_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9, 2)}
                               for k, v in stages.items()}}

        # ---- CPU baseline (oracle) on a bounded sample, and pose parity of the GPU batch against it
        cpu = None
        parity = None
        if world == 1:
            orc = w["orc"]
            threads = os.cpu_count() or 1
            sample = args.cpu_sample or min(B, 2 * threads)
            secs, poses, ok = orc.frontend_batch(w["opts"], w["scans"][:sample], w["origin"], w["prev"][:sample],
                                                 w["cur"][:sample], w["submap_pose"], w["hi"], w["lo"], threads)
            secs1, _, _ = orc.frontend_batch(w["opts"], w["scans"][:max(sample // threads, 2)], w["origin"],
                                             w["prev"], w["cur"], w["submap_pose"], w["hi"], w["lo"], 1)
            cpu = {"value": sample / secs, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{sample} of the batch's scans, one scan per host thread ({secs:.2f} s)",
                   "single_thread_value": max(sample // threads, 2) / secs1}
            got = np.array([list(res[i].pose_estimate_local) for i in range(sample)])
            dt, dr = pose_errors(got, poses)
            parity = {"scans": sample, "rmse_m": float(np.sqrt(np.mean(dt ** 2))), "rmse_rad": float(np.sqrt(np.mean(dr ** 2))),
                      "max_m": float(dt.max()), "max_rad": float(dr.max()), "all_ok": bool(all(r.ok for r in res))}

        h2d = int(sizes.sum() * row_bytes)
        d2h = int(B * C.sizeof(dliom.ScanResult))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 (indices, scores) + f64 (least squares)", "data": "synthetic",
                "config": workload_config(args, B),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms_total / args.steps,
                        "timing": "wall clock around K x (dl_frontend_submit, dl_frontend_collect) on two alternating contexts: "
                                  "batch i+1 uploads while batch i computes; all K uploads, solves and result reads inside",
                        "sync_call": {"value": world * B / (e2e_sync_ms_total / args.steps / 1e3), "ms_per_step": e2e_sync_ms_total / args.steps,
                                      "note": "one blocking dl_frontend_match_batch per step, nothing overlaps across steps"},
                        "streaming_equals_sync_results": stream_equal,
                        "pcie_h2d_gbs": round(h2d_gbs, 2), "copy_only_ms_per_step": round(h2d / h2d_gbs / 1e6, 3),
                        "pcie_note": "plain pinned cudaMemcpyAsync of the same buffers, measured in this run: the floor of e2e"},
                "latency": {"single_scan_ms": round(single_scan_ms, 3),
                            "note": "median wall time of dl_frontend_match_batch on ONE 64-beam scan (2.1 MB upload, 13 launches, result "
                                    "download); the CPU path takes 1000 / cpu_baseline.single_thread_value ms"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity_vs_cpu": parity,
                "clocks": sampler.summary()}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def dliom_range_dtype():
    return np.dtype([("x", np.float32), ("y", np.float32), ("z", np.float32), ("t", np.float32),
                     ("origin_index", np.uint64), ("_pad", np.uint64)])


if __name__ == "__main__":
    main()
