"""CPU tests: the C-ABI library loads, exports every symbol include/mpc_b200.h declares, the
record layouts match the header, defaults restate the reference's config files, and compute
entry points FAIL LOUDLY without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(?:int|void|int32_t|int64_t|const char \*)\s*\*?\s*((?:mpc|balance|prep|a1)_[a-z_0-9]+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(pkg):
    lib = pkg.load_library()
    declared = _header_functions()
    assert len(declared) >= 30
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(pkg.engine.EXPORTED_SYMBOLS) == declared


def test_record_layouts(pkg):
    a = pkg.abi
    assert C.sizeof(a.MpcStateIn) == 192 and C.sizeof(a.BalanceStateIn) == 256 and C.sizeof(a.MpcResult) == 64
    assert a.MpcStateIn.rot_mat.offset == 22 * 4 and a.MpcStateIn.foot_pos_abs.offset == 31 * 4
    assert a.MpcStateIn.contacts.offset == 43 * 4
    assert a.BalanceStateIn.rot_mat_z.offset == 33 * 4 and a.BalanceStateIn.contacts.offset == 54 * 4
    assert a.MpcResult.status.offset == 48
    assert a.STATE_DTYPE.fields["contacts"][1] == 43 * 4
    assert a.BALANCE_DTYPE.fields["foot_pos_abs"][1] == 42 * 4


def test_defaults_restate_reference_config(pkg):
    cfg = pkg.config_default()           # config/gazebo_a1_mpc.yaml:6-13,40-72
    assert cfg.horizon == 10 and cfg.dt == 0.0025 and cfg.mu == 0.3 and cfg.fz_max == 180.0 and cfg.mass == 12.0
    assert list(cfg.q_weights) == [20.0, 10.0, 1.0, 0.0, 0.0, 420.0, 0.05, 0.05, 0.05, 30.0, 30.0, 10.0, 0.0]
    assert list(cfg.r_weights) == [1e-7] * 12
    assert [cfg.inertia[0], cfg.inertia[4], cfg.inertia[8]] == [0.0168352186, 0.0656071082, 0.0742720659]
    hw = pkg.config_hardware()           # config/hardware_a1_mpc.yaml
    assert hw.mass == 13.5 and list(hw.r_weights) == [0.01, 0.01, 0.001] * 4
    assert list(hw.q_weights)[:6] == [150.0, 150.0, 50.0, 0.0, 0.0, 80.0]
    s = pkg.settings_osqp_default()      # osqp 0.6.x library defaults
    assert (s.rho, s.sigma, s.alpha, s.eps_abs, s.eps_rel) == (0.1, 1e-6, 1.6, 1e-3, 1e-3)
    assert (s.max_iter, s.check_termination, s.scaling, s.adaptive_rho) == (4000, 25, 10, 1)
    assert cfg.osqp.eps_abs == 1e-5 and cfg.osqp.eps_rel == 1e-5 and cfg.osqp.adaptive_rho_interval == 50
    b = pkg.balance_config_default()     # A1RobotControl.cpp:11-15, config/gazebo_a1_qp.yaml:54-68
    assert list(b.Q) == [1.0, 1.0, 1.0, 400.0, 400.0, 100.0] and b.R == 1e-3 and b.mu == 0.7 and b.F_max == 180.0
    assert list(b.kp_linear) == [100.0, 100.0, 300.0] and list(b.kd_angular) == [4.5, 4.5, 30.0]


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the engine must refuse, with MPC_ERR_NO_DEVICE and a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    with pytest.raises(pkg.MpcError) as ei:
        pkg.MpcEngine(pkg.config_default(), 0)
    assert ei.value.code == pkg.abi.MPC_ERR_NO_DEVICE and "no CPU fallback" in str(ei.value)
    with pytest.raises(pkg.MpcError):
        pkg.MpcEngine(pkg.balance_config_default(), 0, balance=True)


def test_product_never_imports_oracle():
    """The product path may not include, link or import anything under oracle/."""
    pk = os.path.join(ROOT, "go1_qp_mpc_controller_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_binding" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace(
                    "under oracle/", "").replace("under\noracle/", ""), f
    out = os.popen(f"ldd {os.path.join(pk, 'libmpc_b200.so')}").read()
    assert "oracle" not in out


def test_convexmpc_host_setters_match_oracle(pkg, ob):
    """The cheap ConvexMpc setters (host side of the facade) against the oracle's intermediates."""
    from go1_qp_mpc_controller_b200.convex_mpc import ConvexMpc
    cfg = pkg.config_default()
    rec = pkg.generate_states(1002, 3, 1)[0]
    im = ob.mpc_build_intermediates(cfg, rec)

    class _NoEngine:  # setters never touch the device
        pass
    m = ConvexMpc(np.array(cfg.q_weights[:]), np.array(cfg.r_weights[:]), engine=_NoEngine())
    m.calculate_A_mat_c(rec["euler"].astype(np.float64))
    m.calculate_B_mat_c(cfg.mass, np.array(cfg.inertia[:]).reshape(3, 3),
                        rec["rot_mat"].astype(np.float64).reshape(3, 3),
                        rec["foot_pos_abs"].astype(np.float64).reshape(4, 3).T)
    m.state_space_discretization(cfg.dt)
    np.testing.assert_allclose(m.A_mat_d, im["A_d"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(m.B_mat_d, im["B_d"], rtol=1e-12, atol=1e-16)
    assert np.array_equal(m.linear_constraints, ob.constraint_matrix(10, 0.3))
    assert m.Q[0] == 2 * cfg.q_weights[0] and m.R[0] == 2 * cfg.r_weights[0]
