"""dl::round_div (the cell index of every voxel kernel) is host + device code: its host build is checked here against
lround(x / resolution) on the CPU, adversarially around the rounding boundaries; the device build is checked against the oracle
in tests/test_gpu_parity.py::test_voxel_indices_reciprocal_path_adversarial."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round_div_equals_lround_of_the_ieee_quotient(tmp_path):
    exe = str(tmp_path / "round_div_check")
    # the reference's flags: no -march, no fast-math, no contraction (cmake/functions.cmake:75,92-95)
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-ffp-contract=off", "-x", "c++",
                           os.path.join(ROOT, "tests", "cpp", "round_div_check.cc"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad=0" in out.stdout


def test_cpu_arm_is_sized_from_the_cgroup_quota(monkeypatch, tmp_path):
    """bench.usable_cpus(): os.cpu_count() capped by the affinity mask and the container's CFS quota (cgroup v2 cpu.max)."""
    import builtins
    import sys
    sys.path.insert(0, ROOT)
    import bench
    real_open = builtins.open
    fake = tmp_path / "cpu.max"

    def fake_open(path, *a, **k):
        if path == "/sys/fs/cgroup/cpu.max":
            return real_open(fake, *a, **k)
        return real_open(path, *a, **k)
    monkeypatch.setattr(bench.os, "cpu_count", lambda: 128)
    monkeypatch.setattr(bench.os, "sched_getaffinity", lambda pid: set(range(128)), raising=False)
    monkeypatch.setattr(builtins, "open", fake_open)
    fake.write_text("1600000 100000\n")
    assert bench.usable_cpus()[0] == 16
    fake.write_text("max 100000\n")
    assert bench.usable_cpus()[0] == 128
    fake.write_text("250000 100000\n")
    assert bench.usable_cpus()[0] == 3
