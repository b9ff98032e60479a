"""The C ABI from plain C++: integration/examples/abi_smoke.cpp is compiled with g++ against include/cubeslam_b200.h and linked with the
in-tree shared library -- no Python in the call path.  Without a GPU it must stop at csb_create (exit code 77: no CPU fallback); on a
B200 it runs line detection -> cuboid proposals -> BA linearisation (both Jacobian modes) and exits 0."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, csb):
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(csb.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "integration", "examples", "abi_smoke.cpp"),
                           csb.LIB_PATH, "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_cpp_program_builds_and_refuses_without_gpu(tmp_path, csb):
    import torch
    exe = _build(tmp_path, csb)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77 and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_program_runs_the_pipeline(tmp_path, csb):
    exe = _build(tmp_path, csb)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke: ok" in r.stdout
    print(r.stdout)
