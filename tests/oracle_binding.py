"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from go1_qp_mpc_controller_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(_ROOT, "oracle", "liboracle.so")


class OracleResult(C.Structure):
    _fields_ = [
        ("grf", C.c_double * 12),
        ("status", C.c_int32),
        ("iters", C.c_int32),
        ("rho_updates", C.c_int32),
        ("pad", C.c_int32),
        ("pri_res", C.c_double),
        ("dua_res", C.c_double),
    ]


ORACLE_RESULT_DTYPE = np.dtype(
    [
        ("grf", "<f8", 12),
        ("status", "<i4"),
        ("iters", "<i4"),
        ("rho_updates", "<i4"),
        ("pad", "<i4"),
        ("pri_res", "<f8"),
        ("dua_res", "<f8"),
    ]
)
assert C.sizeof(OracleResult) == ORACLE_RESULT_DTYPE.itemsize

_lib = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build_oracle()
        _lib = C.CDLL(_LIB)
    return _lib


def use_native():
    """bench.py's CPU legs: rebuild the oracle with -march=native ON this box and use that build from now
    on (the portable x86-64-v3 build travels with the snapshot and forgoes AVX-512).  Returns the flag
    string actually in use."""
    global _lib
    native = os.path.join(_ROOT, "oracle", "liboracle_native.so")
    try:
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(_ROOT, "oracle"), "native"],   # -B: never reuse a build from another box
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        _lib = C.CDLL(native)
        return "-O3 -march=native (built on this box)"
    except Exception:
        lib()
        return "-O3 -march=x86-64-v3 (portable build; native rebuild failed)"


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def mpc_build_qp(cfg, state_rec):
    """state_rec: numpy record array (STATE_DTYPE) of length 1 or a scalar record."""
    H = cfg.horizon
    n, m = 12 * H, 20 * H
    st = np.ascontiguousarray(np.atleast_1d(state_rec)[:1])
    P = np.empty((n, n))
    q = np.empty(n)
    l = np.empty(m)
    u = np.empty(m)
    rc = lib().oracle_mpc_build_qp(C.byref(cfg), _vp(st), _p(P), _p(q), _p(l), _p(u))
    assert rc == 0
    return P, q, l, u


def mpc_build_intermediates(cfg, state_rec):
    H = cfg.horizon
    n, s = 12 * H, 13 * H
    st = np.ascontiguousarray(np.atleast_1d(state_rec)[:1])
    out = dict(
        A_d=np.empty((13, 13)),
        B_d=np.empty((13, 12)),
        A_qp=np.empty((s, 13)),
        B_qp=np.empty((s, n)),
        x0=np.empty(13),
        x_ref=np.empty(s),
    )
    rc = lib().oracle_mpc_build_intermediates(
        C.byref(cfg), _vp(st), _p(out["A_d"]), _p(out["B_d"]), _p(out["A_qp"]), _p(out["B_qp"]),
        _p(out["x0"]), _p(out["x_ref"]))
    assert rc == 0
    return out


def mpc_compute_grf(cfg, states, threads=0, want_solutions=False, f32=False):
    states = np.ascontiguousarray(states)
    N = len(states)
    out = np.zeros(N, dtype=ORACLE_RESULT_DTYPE)
    sol = np.empty((N, 12 * cfg.horizon)) if want_solutions else None
    fn = lib().oracle_mpc_compute_grf_f32 if f32 else lib().oracle_mpc_compute_grf
    rc = fn(C.byref(cfg), _vp(states), C.c_int32(N), _vp(out), _p(sol), C.c_int32(threads))
    assert rc == 0
    return (out, sol) if want_solutions else out


def qp_mats_from_model(cfg, A_d, B_d_list, x0, x_ref, contacts):
    H = cfg.horizon
    n, m = 12 * H, 20 * H
    A_d = np.ascontiguousarray(A_d, dtype=np.float64)
    B_d_list = np.ascontiguousarray(B_d_list, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    x_ref = np.ascontiguousarray(x_ref, dtype=np.float64)
    c = np.ascontiguousarray(contacts, dtype=np.int32)
    P = np.empty((n, n))
    q = np.empty(n)
    l = np.empty(m)
    u = np.empty(m)
    rc = lib().oracle_qp_mats_from_model(C.byref(cfg), _p(A_d), _p(B_d_list), _p(x0), _p(x_ref),
                                         _p(c, C.c_int32), _p(P), _p(q), _p(l), _p(u))
    assert rc == 0
    return P, q, l, u


def osqp_solve_mpc(cfg, P, q, l, u):
    H = cfg.horizon
    n, m = 12 * H, 20 * H
    P = np.ascontiguousarray(P, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    x = np.empty(n)
    y = np.empty(m)
    info = np.zeros(1, dtype=ORACLE_RESULT_DTYPE)
    rc = lib().oracle_osqp_solve_mpc(C.byref(cfg), _p(P), _p(q), _p(l), _p(u), _p(x), _p(y), _vp(info))
    assert rc == 0
    return x, y, info[0]


def balance_build_qp(cfg, state_rec):
    st = np.ascontiguousarray(np.atleast_1d(state_rec)[:1])
    P = np.empty((12, 12))
    q = np.empty(12)
    l = np.empty(20)
    u = np.empty(20)
    rc = lib().oracle_balance_build_qp(C.byref(cfg), _vp(st), _p(P), _p(q), _p(l), _p(u))
    assert rc == 0
    return P, q, l, u


def balance_compute_grf(cfg, states, threads=0):
    states = np.ascontiguousarray(states)
    N = len(states)
    out = np.zeros(N, dtype=ORACLE_RESULT_DTYPE)
    rc = lib().oracle_balance_compute_grf(C.byref(cfg), _vp(states), C.c_int32(N), _vp(out),
                                          C.c_int32(threads))
    assert rc == 0
    return out


def mpc_stream(cfg, states_tn, threads=0):
    """states_tn: (ticks, n) record array; returns (ticks, n) oracle results (warm-started solver per robot)."""
    st = np.ascontiguousarray(states_tn)
    ticks, n = st.shape
    out = np.zeros((ticks, n), dtype=ORACLE_RESULT_DTYPE)
    rc = lib().oracle_mpc_stream(C.byref(cfg), _vp(st), C.c_int32(n), C.c_int32(ticks), _vp(out), C.c_int32(threads))
    assert rc == 0
    return out


def torque_map(states, torque_in, grf, balance=False):
    """compute_joint_torques for len(states) robots from body-frame GRFs (n, 12) in double."""
    st = np.ascontiguousarray(states)
    tin = np.ascontiguousarray(torque_in)
    g = np.ascontiguousarray(grf, dtype=np.float64)
    n = len(st)
    tau = np.zeros((n, 12))
    mask = np.zeros(n, dtype=np.int32)
    stride, off = (64, 54) if balance else (48, 43)
    rc = lib().oracle_torque_map(_vp(st), C.c_int32(stride), C.c_int32(off), _vp(tin), _p(g), C.c_int32(n), _p(tau),
                                 _p(mask, C.c_int32))
    assert rc == 0
    return tau, mask


def prep_stream(prep_cfg, sensors_tn):
    """sensors_tn: (ticks, n) RobotSensorIn records -> (states, torque_in, extras), each (ticks, n)."""
    s = np.ascontiguousarray(sensors_tn)
    ticks, n = s.shape
    st = np.zeros((ticks, n), dtype=abi.STATE_DTYPE)
    tin = np.zeros((ticks, n), dtype=abi.TORQUE_IN_DTYPE)
    ex = np.zeros((ticks, n), dtype=abi.PREP_OUT_DTYPE)
    rc = lib().oracle_prep_stream(C.byref(prep_cfg), _vp(s), C.c_int32(n), C.c_int32(ticks), _vp(st), _vp(tin), _vp(ex))
    assert rc == 0
    return st, tin, ex


def leg_fk_jac(rho_fix, q):
    r = np.ascontiguousarray(rho_fix, np.float64)
    qq = np.ascontiguousarray(q, np.float64)
    p = np.zeros(3)
    J = np.zeros((3, 3))
    lib().oracle_leg_fk_jac(_p(r), _p(qq), _p(p), _p(J))
    return p, J


def mpc_build_qp_ext(cfg, state_rec, gait_rec=None):
    H = cfg.horizon
    n, m = 12 * H, 20 * H
    st = np.ascontiguousarray(np.atleast_1d(state_rec)[:1])
    g = None if gait_rec is None else np.ascontiguousarray(np.atleast_1d(gait_rec)[:1])
    P = np.empty((n, n)); q = np.empty(n); l = np.empty(m); u = np.empty(m)
    rc = lib().oracle_mpc_build_qp_ext(C.byref(cfg), _vp(st), None if g is None else _vp(g), _p(P), _p(q), _p(l), _p(u))
    assert rc == 0
    return P, q, l, u


def mpc_compute_grf_ext(cfg, states, gait=None, threads=0):
    states = np.ascontiguousarray(states)
    g = None if gait is None else np.ascontiguousarray(gait)
    out = np.zeros(len(states), dtype=ORACLE_RESULT_DTYPE)
    rc = lib().oracle_mpc_compute_grf_ext(C.byref(cfg), _vp(states), None if g is None else _vp(g), C.c_int32(len(states)),
                                          _vp(out), C.c_int32(threads))
    assert rc == 0
    return out


def discretize_exact(cfg, state_rec):
    st = np.ascontiguousarray(np.atleast_1d(state_rec)[:1])
    A = np.empty((13, 13)); B = np.empty((13, 12))
    rc = lib().oracle_discretize_exact(C.byref(cfg), _vp(st), _p(A), _p(B))
    assert rc == 0
    return A, B


def max_threads():
    return int(lib().oracle_max_threads())


def constraint_matrix(H, mu):
    """Dense linear_constraints of ConvexMpc.cpp:46-58 (numpy, for certificates)."""
    n, m = 12 * H, 20 * H
    A = np.zeros((m, n))
    for i in range(4 * H):
        A[5 * i + 0, 3 * i + 0] = 1
        A[5 * i + 1, 3 * i + 0] = 1
        A[5 * i + 2, 3 * i + 1] = 1
        A[5 * i + 3, 3 * i + 1] = 1
        A[5 * i + 4, 3 * i + 2] = 1
        A[5 * i + 0, 3 * i + 2] = mu
        A[5 * i + 1, 3 * i + 2] = -mu
        A[5 * i + 2, 3 * i + 2] = mu
        A[5 * i + 3, 3 * i + 2] = -mu
    return A
