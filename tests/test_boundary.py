"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/dliom_b200.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dliom_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import dliom
    L = ctypes.CDLL(dliom.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert sorted(dliom.EXPORTS) == names


def test_struct_layouts_match_header():
    import dliom
    assert ctypes.sizeof(dliom.SolveSummary) == 40
    assert ctypes.sizeof(dliom.CeresOptions) == 8 + 8 * 4 + 16 + 16
    assert ctypes.sizeof(dliom.ScanResult) == 56 * 2 + 40 + 12 * 4
    assert ctypes.sizeof(dliom.RtcsmInfo) == 32


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import dliom
    with pytest.raises(dliom.DlError) as e:
        dliom.Context(0)
    assert e.value.status == -1   # DL_ERR_CUDA


def test_status_strings():
    import dliom
    L = dliom.lib()
    assert L.dl_status_string(0) == b"ok"
    assert b"CUDA" in L.dl_status_string(-1)


def _build_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "example_match")
    host = os.path.join(ROOT, "d-liom_b200", "host")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(host, "example_match.cc"), "-o", exe,
                           "-L" + os.path.join(ROOT, "d-liom_b200"), "-ldliom_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "d-liom_b200")])
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    """The header-only C++ mirror of the reference classes (d-liom_b200/host) builds against the C-ABI."""
    import subprocess
    import torch
    exe = _build_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "dliom error -1" in r.stderr


@pytest.mark.gpu
def test_cpp_shim_reference_matcher_case(tmp_path):
    import subprocess
    r = subprocess.run([_build_example(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "voxel filter kept" in r.stdout
