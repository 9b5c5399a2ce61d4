"""Host-side binding of the C ABI (include/mpc_b200.h) through ctypes.

This is plumbing only: every compute call goes to libmpc_b200.so (hand-written
sm_100a kernels).  There is no CPU fallback and no import of anything under
oracle/: if the library is missing or no CUDA device is present the calls
raise.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpc_b200.so")

# every symbol include/mpc_b200.h declares
EXPORTED_SYMBOLS = [
    "mpc_settings_osqp_default", "mpc_settings_benchmark", "mpc_config_default",
    "mpc_config_hardware", "balance_config_default", "mpc_generate_states",
    "balance_generate_states", "mpc_engine_create", "mpc_engine_destroy", "mpc_last_error",
    "mpc_set_stream", "mpc_synchronize", "mpc_kernel_launches", "mpc_debug_phase_cycles", "mpc_load_states",
    "mpc_set_states_device", "mpc_build_qp", "mpc_build_qp_async", "mpc_get_qp", "mpc_solve",
    "mpc_solve_async", "mpc_get_results", "mpc_results_device", "mpc_get_solution",
    "mpc_compute_grf_batch", "mpc_qp_mats_from_model", "mpc_solve_qp", "balance_engine_create",
    "balance_qp_solve", "balance_load_states", "balance_solve", "balance_get_qp",
    "mpc_generate_stream_states", "mpc_solve_warm", "mpc_solve_warm_async", "mpc_stream_reset",
    "mpc_stream_step", "mpc_set_torque_inputs", "mpc_get_torques", "mpc_generate_torque_inputs",
    "prep_config_default", "mpc_prepare_states", "mpc_get_prepared", "mpc_prepare_reset", "mpc_generate_sensors",
    "a1_leg_fk_jac", "mpc_set_gait_inputs", "mpc_generate_gait_inputs",
    "mpc_stream_reset_slots", "mpc_engine_update_model",
    "mpc_fleet_create", "mpc_fleet_destroy", "mpc_fleet_last_error", "mpc_fleet_size", "mpc_fleet_shard_range",
    "mpc_measure_fp64_peak", "mpc_fleet_compute_grf_batch", "mpc_fleet_stream_step", "mpc_fleet_stream_reset", "mpc_fleet_kernel_launches",
]


class MpcError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"mpc_b200 error {code}: {text}")
        self.code = code


_lib = None


def load_library():
    """Load libmpc_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpcError(abi.MPC_ERR_NO_DEVICE,
                       f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    lib.mpc_last_error.restype = C.c_char_p
    lib.mpc_last_error.argtypes = [vp]
    lib.mpc_kernel_launches.restype = C.c_int64
    lib.mpc_kernel_launches.argtypes = [vp]
    lib.mpc_engine_destroy.restype = None
    lib.mpc_engine_destroy.argtypes = [vp]
    lib.mpc_engine_create.argtypes = [C.POINTER(abi.MpcConfig), i32, C.POINTER(vp)]
    lib.balance_engine_create.argtypes = [C.POINTER(abi.BalanceConfig), i32, C.POINTER(vp)]
    lib.mpc_generate_states.argtypes = [u64, u64, i32, vp]
    lib.balance_generate_states.argtypes = [u64, u64, i32, vp]
    lib.mpc_set_stream.argtypes = [vp, vp]
    lib.mpc_synchronize.argtypes = [vp]
    lib.mpc_load_states.argtypes = [vp, vp, i32]
    lib.mpc_set_states_device.argtypes = [vp, vp, i32]
    lib.mpc_build_qp.argtypes = [vp]
    lib.mpc_build_qp_async.argtypes = [vp]
    lib.mpc_get_qp.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.mpc_solve.argtypes = [vp]
    lib.mpc_solve_async.argtypes = [vp]
    lib.mpc_get_results.argtypes = [vp, vp]
    lib.mpc_results_device.argtypes = [vp, C.POINTER(vp)]
    lib.mpc_get_solution.argtypes = [vp, i32, vp]
    lib.mpc_compute_grf_batch.argtypes = [vp, vp, vp, i32]
    lib.mpc_generate_stream_states.argtypes = [u64, u64, i32, C.c_int64, vp]
    lib.mpc_set_torque_inputs.argtypes = [vp, vp, i32]
    lib.mpc_get_torques.argtypes = [vp, vp]
    lib.mpc_generate_torque_inputs.argtypes = [u64, u64, i32, vp]
    lib.prep_config_default.argtypes = [C.POINTER(abi.PrepConfig)]
    lib.mpc_prepare_states.argtypes = [vp, C.POINTER(abi.PrepConfig), vp, i32]
    lib.mpc_get_prepared.argtypes = [vp, vp, vp, vp]
    lib.mpc_prepare_reset.argtypes = [vp]
    lib.mpc_generate_sensors.argtypes = [u64, u64, i32, C.c_int64, vp]
    lib.a1_leg_fk_jac.argtypes = [vp, vp, vp, vp]
    lib.mpc_set_gait_inputs.argtypes = [vp, vp, i32]
    lib.mpc_generate_gait_inputs.argtypes = [u64, u64, i32, C.c_int64, vp]
    lib.mpc_solve_warm.argtypes = [vp]
    lib.mpc_solve_warm_async.argtypes = [vp]
    lib.mpc_stream_reset.argtypes = [vp]
    lib.mpc_stream_step.argtypes = [vp, vp, vp, i32]
    lib.mpc_stream_reset_slots.argtypes = [vp, vp, i32]
    lib.mpc_engine_update_model.argtypes = [vp, C.POINTER(abi.MpcConfig)]
    lib.mpc_measure_fp64_peak.argtypes = [i32, C.c_double, C.POINTER(C.c_double)]
    lib.mpc_fleet_create.argtypes = [C.POINTER(abi.MpcConfig), vp, i32, C.POINTER(vp)]
    lib.mpc_fleet_destroy.restype = None
    lib.mpc_fleet_destroy.argtypes = [vp]
    lib.mpc_fleet_last_error.restype = C.c_char_p
    lib.mpc_fleet_last_error.argtypes = [vp]
    lib.mpc_fleet_size.argtypes = [vp]
    lib.mpc_fleet_shard_range.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.mpc_fleet_compute_grf_batch.argtypes = [vp, vp, vp, i32]
    lib.mpc_fleet_stream_step.argtypes = [vp, vp, vp, i32]
    lib.mpc_fleet_stream_reset.argtypes = [vp]
    lib.mpc_fleet_kernel_launches.restype = C.c_int64
    lib.mpc_fleet_kernel_launches.argtypes = [vp]
    lib.mpc_qp_mats_from_model.argtypes = [vp] + [vp] * 9
    lib.mpc_solve_qp.argtypes = [vp] + [vp] * 7
    lib.balance_qp_solve.argtypes = [vp, vp, vp, i32]
    lib.balance_load_states.argtypes = [vp, vp, i32]
    lib.balance_solve.argtypes = [vp]
    lib.balance_get_qp.argtypes = [vp, i32, vp, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- configuration ---------------------------------------------------------

def config_default():
    cfg = abi.MpcConfig()
    load_library().mpc_config_default(C.byref(cfg))
    return cfg


def config_hardware():
    cfg = abi.MpcConfig()
    load_library().mpc_config_hardware(C.byref(cfg))
    return cfg


def balance_config_default():
    cfg = abi.BalanceConfig()
    load_library().balance_config_default(C.byref(cfg))
    return cfg


def settings_osqp_default():
    s = abi.MpcSolverSettings()
    load_library().mpc_settings_osqp_default(C.byref(s))
    return s


def generate_states(seed, first_index, n):
    out = np.zeros(n, dtype=abi.STATE_DTYPE)
    rc = load_library().mpc_generate_states(seed, first_index, n, _ptr(out))
    if rc:
        raise MpcError(rc, "mpc_generate_states")
    return out


def generate_stream_states(seed, first_index, n, tick):
    """The robots of generate_states(seed, first_index, n), `tick` control periods later."""
    out = np.zeros(n, dtype=abi.STATE_DTYPE)
    rc = load_library().mpc_generate_stream_states(seed, first_index, n, tick, _ptr(out))
    if rc:
        raise MpcError(rc, "mpc_generate_stream_states")
    return out


def generate_torque_inputs(seed, first_index, n):
    """Synthetic compute_joint_torques inputs (leg Jacobians, PD forces) for the robots of generate_states."""
    out = np.zeros(n, dtype=abi.TORQUE_IN_DTYPE)
    rc = load_library().mpc_generate_torque_inputs(seed, first_index, n, _ptr(out))
    if rc:
        raise MpcError(rc, "mpc_generate_torque_inputs")
    return out


def generate_gait_inputs(seed, first_index, n, tick=0):
    out = np.zeros(n, dtype=abi.GAIT_DTYPE)
    rc = load_library().mpc_generate_gait_inputs(seed, first_index, n, tick, _ptr(out))
    if rc:
        raise MpcError(rc, "mpc_generate_gait_inputs")
    return out


def prep_config_default():
    cfg = abi.PrepConfig()
    rc = load_library().prep_config_default(C.byref(cfg))
    if rc:
        raise MpcError(rc, "prep_config_default")
    return cfg


def generate_sensors(seed, first_index, n, tick=0):
    """Synthetic RobotSensorIn records for the robots of generate_stream_states at `tick`."""
    out = np.zeros(n, dtype=abi.SENSOR_DTYPE)
    rc = load_library().mpc_generate_sensors(seed, first_index, n, tick, _ptr(out))
    if rc:
        raise MpcError(rc, "mpc_generate_sensors")
    return out


def a1_leg_fk_jac(rho_fix, q):
    """Foot position (3,) and Jacobian (3, 3) of one A1 leg, rho_fix = (ox, oy, d, lt, lc)."""
    r = np.ascontiguousarray(rho_fix, np.float64)
    qq = np.ascontiguousarray(q, np.float64)
    p = np.zeros(3)
    J = np.zeros((3, 3))
    rc = load_library().a1_leg_fk_jac(_ptr(r), _ptr(qq), _ptr(p), _ptr(J))
    if rc:
        raise MpcError(rc, "a1_leg_fk_jac")
    return p, J


def generate_balance_states(seed, first_index, n):
    out = np.zeros(n, dtype=abi.BALANCE_DTYPE)
    rc = load_library().balance_generate_states(seed, first_index, n, _ptr(out))
    if rc:
        raise MpcError(rc, "balance_generate_states")
    return out


# ---- engine ------------------------------------------------------------------

class MpcEngine:
    """One engine = one GPU + one stream.  Mirrors the C ABI one to one."""

    def __init__(self, cfg=None, device=0, balance=False):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.balance = balance
        if balance:
            self.cfg = cfg if cfg is not None else balance_config_default()
            rc = self._lib.balance_engine_create(C.byref(self.cfg), device, C.byref(self._h))
        else:
            self.cfg = cfg if cfg is not None else config_default()
            rc = self._lib.mpc_engine_create(C.byref(self.cfg), device, C.byref(self._h))
        if rc:
            raise MpcError(rc, self._lib.mpc_last_error(None).decode())
        self.n = 0
        self._keep = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mpc_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise MpcError(rc, self._lib.mpc_last_error(self._h).decode())

    @property
    def horizon(self):
        return self.cfg.horizon

    def set_stream(self, cuda_stream):
        self._check(self._lib.mpc_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self._lib.mpc_synchronize(self._h))

    def kernel_launches(self):
        return int(self._lib.mpc_kernel_launches(self._h))

    def phase_cycles(self, enable=True):
        """Developer aid: read and reset admm_solve_kernel's per-phase cycle counters."""
        out = (C.c_int64 * 22)()
        self._lib.mpc_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        self._check(self._lib.mpc_debug_phase_cycles(self._h, 1 if enable else 0, out))
        d = dict(zip(["load_scale", "factor", "iterations", "checks", "output", "problems"], list(out)[:6]))
        d["fine"] = list(out)[6:]
        return d

    def load_states(self, states):
        states = np.ascontiguousarray(states)
        self._keep = states  # async H2D reads this buffer until the next sync
        if self.balance:
            assert states.dtype == abi.BALANCE_DTYPE
            self._check(self._lib.balance_load_states(self._h, _ptr(states), len(states)))
        else:
            assert states.dtype == abi.STATE_DTYPE
            self._check(self._lib.mpc_load_states(self._h, _ptr(states), len(states)))
        self.n = len(states)

    def set_states_device(self, dev_ptr, n):
        self._check(self._lib.mpc_set_states_device(self._h, C.c_void_p(dev_ptr), n))
        self.n = n

    def build_qp(self, sync=True):
        self._check((self._lib.mpc_build_qp if sync else self._lib.mpc_build_qp_async)(self._h))

    def get_qp(self, idx):
        if self.balance:
            P = np.empty((12, 12), np.float32)
            q = np.empty(12, np.float32)
            l = np.empty(20, np.float32)
            u = np.empty(20, np.float32)
            self._check(self._lib.balance_get_qp(self._h, idx, _ptr(P), _ptr(q), _ptr(l), _ptr(u)))
            return P, q, l, u
        n, m = 12 * self.horizon, 20 * self.horizon
        P = np.empty((n, n), np.float32)
        q = np.empty(n, np.float32)
        l = np.empty(m, np.float32)
        u = np.empty(m, np.float32)
        self._check(self._lib.mpc_get_qp(self._h, idx, _ptr(P), _ptr(q), _ptr(l), _ptr(u)))
        return P, q, l, u

    def solve(self, sync=True):
        if self.balance:
            self._check(self._lib.balance_solve(self._h))
            if sync:
                self.synchronize()
        else:
            self._check((self._lib.mpc_solve if sync else self._lib.mpc_solve_async)(self._h))

    def get_results(self, out=None):
        if out is None:
            out = np.zeros(self.n, dtype=abi.RESULT_DTYPE)
        self._check(self._lib.mpc_get_results(self._h, _ptr(out)))
        return out

    def results_device_ptr(self):
        p = C.c_void_p()
        self._check(self._lib.mpc_results_device(self._h, C.byref(p)))
        return p.value

    def get_solution(self, idx):
        x = np.empty(12 * self.horizon, np.float32)
        self._check(self._lib.mpc_get_solution(self._h, idx, _ptr(x)))
        return x

    def compute_grf_batch(self, states, out=None):
        """compute_grf for every record: host in, host out (H2D + kernels + D2H)."""
        states = np.ascontiguousarray(states)
        if out is None:
            out = np.zeros(len(states), dtype=abi.RESULT_DTYPE)
        if self.balance:
            self._check(self._lib.balance_qp_solve(self._h, _ptr(states), _ptr(out), len(states)))
        else:
            self._check(self._lib.mpc_compute_grf_batch(self._h, _ptr(states), _ptr(out), len(states)))
        self.n = len(states)
        return out

    def set_gait_inputs(self, gait):
        """Gait scheduler records of the loaded states (gait_aware engines)."""
        if gait is None:
            self._check(self._lib.mpc_set_gait_inputs(self._h, None, 0))
            return
        g = np.ascontiguousarray(gait)
        assert g.dtype == abi.GAIT_DTYPE
        self._keep_g = g
        self._check(self._lib.mpc_set_gait_inputs(self._h, _ptr(g), len(g)))

    # upstream state preparation on the device (orientation, leg kinematics, EKF, terrain pitch)
    def prepare_states(self, sensors, prep_cfg=None):
        """Sensor records in; afterwards the engine holds the states and torque inputs on the device."""
        s = np.ascontiguousarray(sensors)
        assert s.dtype == abi.SENSOR_DTYPE
        self._keep = s
        self._prep_cfg = prep_cfg if prep_cfg is not None else getattr(self, "_prep_cfg", None) or prep_config_default()
        self._check(self._lib.mpc_prepare_states(self._h, C.byref(self._prep_cfg), _ptr(s), len(s)))
        self.n = len(s)

    def get_prepared(self):
        st = np.zeros(self.n, dtype=abi.STATE_DTYPE)
        tin = np.zeros(self.n, dtype=abi.TORQUE_IN_DTYPE)
        ex = np.zeros(self.n, dtype=abi.PREP_OUT_DTYPE)
        self._check(self._lib.mpc_get_prepared(self._h, _ptr(st), _ptr(tin), _ptr(ex)))
        return st, tin, ex

    def prepare_reset(self):
        self._check(self._lib.mpc_prepare_reset(self._h))

    # torque map fused into the result writer (compute_joint_torques, A1RobotControl.cpp:289-319)
    def set_torque_inputs(self, torque_in):
        """Give the torque-map inputs of the loaded states (None switches the map off)."""
        if torque_in is None:
            self._check(self._lib.mpc_set_torque_inputs(self._h, None, 0))
            return
        t = np.ascontiguousarray(torque_in)
        assert t.dtype == abi.TORQUE_IN_DTYPE
        self._keep_t = t  # async H2D reads this buffer until the next sync
        self._check(self._lib.mpc_set_torque_inputs(self._h, _ptr(t), len(t)))

    def get_torques(self):
        out = np.zeros(self.n, dtype=abi.TORQUE_OUT_DTYPE)
        self._check(self._lib.mpc_get_torques(self._h, _ptr(out)))
        return out

    # warm-started streaming: slot i keeps robot i's solver alive between ticks
    # (A1RobotControl.cpp:522-538)
    def solve_warm(self, sync=True):
        self._check((self._lib.mpc_solve_warm if sync else self._lib.mpc_solve_warm_async)(self._h))

    def stream_reset(self):
        self._check(self._lib.mpc_stream_reset(self._h))

    def stream_reset_slots(self, idx):
        """Forget the live solvers of the robot slots `idx` only (their next tick is an initSolver)."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._check(self._lib.mpc_stream_reset_slots(self._h, _ptr(idx), len(idx)))

    def update_model(self, cfg):
        """New model constants (mass, inertia, weights, dt, solver settings) on a live engine: the warm
        solvers stay alive, like the reference's member solver across weight changes."""
        self._check(self._lib.mpc_engine_update_model(self._h, C.byref(cfg)))
        self.cfg = cfg

    def stream_step(self, states, out=None):
        """One control tick for every robot: host in, host out, warm-started solve."""
        states = np.ascontiguousarray(states)
        if out is None:
            out = np.zeros(len(states), dtype=abi.RESULT_DTYPE)
        self._check(self._lib.mpc_stream_step(self._h, _ptr(states), _ptr(out), len(states)))
        self.n = len(states)
        return out

    def qp_mats_from_model(self, A_mat_d, B_mat_d_list, mpc_states, mpc_states_d, contacts):
        H = self.horizon
        n, m = 12 * H, 20 * H
        A = np.ascontiguousarray(A_mat_d, np.float64)
        B = np.ascontiguousarray(B_mat_d_list, np.float64)
        x0 = np.ascontiguousarray(mpc_states, np.float64)
        xr = np.ascontiguousarray(mpc_states_d, np.float64)
        c = np.ascontiguousarray(contacts, np.int32)
        assert A.shape == (13, 13) and B.shape == (13 * H, 12) and x0.shape == (13,) and xr.shape == (13 * H,)
        P = np.empty((n, n))
        q = np.empty(n)
        l = np.empty(m)
        u = np.empty(m)
        self._check(self._lib.mpc_qp_mats_from_model(self._h, _ptr(A), _ptr(B), _ptr(x0), _ptr(xr), _ptr(c),
                                                     _ptr(P), _ptr(q), _ptr(l), _ptr(u)))
        return P, q, l, u

    def solve_qp(self, hessian, gradient, lb, ub):
        H = self.horizon
        n = 12 * H
        P = np.ascontiguousarray(hessian, np.float64)
        q = np.ascontiguousarray(gradient, np.float64)
        l = np.ascontiguousarray(lb, np.float64)
        u = np.ascontiguousarray(ub, np.float64)
        x = np.empty(n)
        status = C.c_int32()
        iters = C.c_int32()
        self._check(self._lib.mpc_solve_qp(self._h, _ptr(P), _ptr(q), _ptr(l), _ptr(u), _ptr(x),
                                           C.cast(C.byref(status), C.c_void_p),
                                           C.cast(C.byref(iters), C.c_void_p)))
        return x, status.value, iters.value


def measure_fp64_peak(device=0, ms_target=3.0):
    """This GPU's FP64 FMA rate in TFLOP/s, measured now (register-only DFMA kernel on every SM)."""
    v = C.c_double()
    lib = load_library()
    rc = lib.mpc_measure_fp64_peak(device, ms_target, C.byref(v))
    if rc:
        raise MpcError(rc, lib.mpc_last_error(None).decode())
    return v.value


def fleet_shard_range(n, ndev, i):
    """[begin, end) of shard i of ndev: ceil(n i / G) .. ceil(n (i + 1) / G) (host only)."""
    b, e = C.c_int32(), C.c_int32()
    rc = load_library().mpc_fleet_shard_range(n, ndev, i, C.byref(b), C.byref(e))
    if rc:
        raise MpcError(rc, "mpc_fleet_shard_range")
    return b.value, e.value


class MpcFleet:
    """Several GPUs of one box behind one call (mpc_fleet_* of the C ABI): contiguous shards, one
    stream and one pinned staging pair per GPU, results into ONE host array.  No NCCL, no torch."""

    def __init__(self, cfg=None, devices=(0,)):
        self._lib = load_library()
        self.cfg = cfg if cfg is not None else config_default()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        self._h = C.c_void_p()
        rc = self._lib.mpc_fleet_create(C.byref(self.cfg), _ptr(dev), len(dev), C.byref(self._h))
        if rc:
            raise MpcError(rc, self._lib.mpc_fleet_last_error(None).decode())
        self.devices = list(int(d) for d in dev)

    def _check(self, rc):
        if rc:
            raise MpcError(rc, self._lib.mpc_fleet_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mpc_fleet_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.mpc_fleet_size(self._h))

    def compute_grf_batch(self, states, out=None):
        states = np.ascontiguousarray(states)
        if out is None:
            out = np.zeros(len(states), dtype=abi.RESULT_DTYPE)
        self._check(self._lib.mpc_fleet_compute_grf_batch(self._h, _ptr(states), _ptr(out), len(states)))
        return out

    def stream_step(self, states, out=None):
        states = np.ascontiguousarray(states)
        if out is None:
            out = np.zeros(len(states), dtype=abi.RESULT_DTYPE)
        self._check(self._lib.mpc_fleet_stream_step(self._h, _ptr(states), _ptr(out), len(states)))
        return out

    def stream_reset(self):
        self._check(self._lib.mpc_fleet_stream_reset(self._h))

    def kernel_launches(self):
        return int(self._lib.mpc_fleet_kernel_launches(self._h))
