"""include/convex_mpc_b200.hpp: the C++ mirror of the reference interface over the C ABI.
The driver tests/cpp/test_mpc_b200.cpp is the reference's test/test_mpc.cpp restated."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_mpc_b200.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_mpc_b200")


def _build():
    deps = [SRC, os.path.join(ROOT, "include", "convex_mpc_b200.hpp"), os.path.join(ROOT, "include", "mpc_b200.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), SRC,
                               "-L" + os.path.join(ROOT, "go1_qp_mpc_controller_b200"), "-lmpc_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "go1_qp_mpc_controller_b200"), "-o", EXE])


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(pkg):
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_like_test_mpc_cpp(pkg):
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "status 1 iters 50" in r.stdout
