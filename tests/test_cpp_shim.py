"""include/convex_mpc_b200.hpp: the C++ mirror of the reference interface over the C ABI.
The driver tests/cpp/test_mpc_b200.cpp is the reference's test/test_mpc.cpp restated."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_mpc_b200.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_mpc_b200")


SRC_STREAM = os.path.join(ROOT, "tests", "cpp", "test_stream_b200.cpp")
EXE_STREAM = os.path.join(ROOT, "tests", "cpp", "test_stream_b200")


def _build(src=SRC, exe=EXE):
    deps = [src, os.path.join(ROOT, "include", "convex_mpc_b200.hpp"), os.path.join(ROOT, "include", "mpc_b200.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), src,
                               "-L" + os.path.join(ROOT, "go1_qp_mpc_controller_b200"), "-lmpc_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "go1_qp_mpc_controller_b200"), "-o", exe])


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


def test_cpp_stream_driver_compiles(pkg):
    _build(SRC_STREAM, EXE_STREAM)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["gazebo", "hardware"])
def test_cpp_compute_grf_keeps_one_warm_solver(pkg, ob, which, tmp_path):
    """24 consecutive ticks of one robot through the C++ mirror's compute_grf, across a trot swap:
    tick 0 is initSolver, every later tick an update + warm solve of the SAME solver, exactly like the
    member solver of A1RobotControl.h:67 (A1RobotControl.cpp:522-540).  Tick by tick against the
    oracle's MpcStream: status, iteration count, GRF."""
    import numpy as np
    _build(SRC_STREAM, EXE_STREAM)
    cfg = pkg.config_default() if which == "gazebo" else pkg.config_hardware()
    T, robot = 24, 7
    st = np.stack([pkg.generate_stream_states(1006, robot, 1, 36 + t) for t in range(T)])       # (T, 1)
    assert (st[11]["contacts"] != st[12]["contacts"]).any()                                       # swap at tick 48
    ref = ob.mpc_stream(cfg, st)
    path = tmp_path / "records.bin"
    path.write_bytes(st.tobytes())
    r = subprocess.run([EXE_STREAM, str(path)] + (["hardware"] if which == "hardware" else []),
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [ln.split() for ln in r.stdout.strip().splitlines()]
    assert len(rows) == T
    for t, row in enumerate(rows):
        assert int(row[1]) == ref["status"][t, 0] == 1
        assert int(row[2]) == ref["iters"][t, 0], (t, row[2], ref["iters"][t, 0])
        g = np.array([float(v) for v in row[3:15]])
        den = max(np.linalg.norm(ref["grf"][t, 0]), 1.0)
        assert np.linalg.norm(g - ref["grf"][t, 0]) / den <= 1e-3, t
    # warm ticks are cheaper than the cold first tick
    assert np.mean([int(row[2]) for row in rows[1:11]]) < 0.7 * int(rows[0][2])
