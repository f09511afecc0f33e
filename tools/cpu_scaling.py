"""CPU-arm scaling probe for the box bench.py runs on: cgroup CPU quota, affinity, and the oracle chain's throughput at several
thread counts (pooled, >= 4 scans per thread). Usage: python tools/cpu_scaling.py > gpurun_out/cpu_scaling.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "d-liom_b200")]
import bench  # noqa: E402


def read(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


class Args:
    batch, beams, map_scans, pairs, row_floats, gpus = 32, 64, 8, 0, 3, 1


out = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "cgroup_cpu_max": read("/sys/fs/cgroup/cpu.max"),
       "cfs_quota_us": read("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"), "cfs_period_us": read("/sys/fs/cgroup/cpu/cpu.cfs_period_us"),
       "loadavg": read("/proc/loadavg"), "physical_cores": bench.physical_cores(), "scaling": []}
w = bench.build_workload(Args, 0)
for th in (1, 8, 16, 32, 64, 128):
    if th > (os.cpu_count() or 1):
        break
    sel = [i % Args.batch for i in range(max(16, 6 * th))]
    bench.cpu_chain(w, sel[:th], th)
    secs = min(bench.cpu_chain(w, sel, th)[0] for _ in range(2))
    out["scaling"].append({"threads": th, "scans": len(sel), "scans_per_s": len(sel) / secs, "per_thread": len(sel) / secs / th})
print(json.dumps(out, indent=1))
