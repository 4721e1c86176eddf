"""The C++ host mirror (include/gogp_b200.hpp): builds against the C-ABI; without a GPU
it must fail loudly, on a B200 it must reproduce the reference's tables."""
import os
import subprocess

import pytest

from tests.conftest import ROOT


def _build(built_lib):
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "cpp_host_test")
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp_host_test.cc"),
                           "-L" + libdir, "-lgogp_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_mirror_builds_and_fails_loudly_without_device(built_lib):
    import torch
    exe = _build(built_lib)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, "nodevice"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "panic" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_tables():
    exe = os.path.join(ROOT, "tests", "_build", "cpp_host_test")
    if not os.path.exists(exe):  # built on the CPU box and shipped; rebuild only if g++ is here
        from gogp_b200 import _lib
        exe = _build(_lib.LIB_PATH)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all ok" in r.stdout


def test_optimiser_loops_on_analytic_objectives():
    """gogp_b200/csrc/optimize.hpp (Adam / L-BFGS drivers behind gogp_optimize) is host code
    templated on the objective: unit-tested here on a quadratic, Rosenbrock and a domain wall."""
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "cpp_opt_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp_opt_test.cc")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all ok" in r.stdout
