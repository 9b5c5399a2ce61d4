"""ctypes mirror of include/mpc_b200.h (record and config layouts).

Field order and sizes must match the C header exactly; tests/test_abi.py checks
sizeof() of every record against the values the header documents.
"""
import ctypes as C

import numpy as np

MPC_STATE_DIM = 13
MPC_NUM_DOF = 12
MPC_CONSTRAINT_DIM = 20
MPC_NUM_LEG = 4
MPC_INFTY = 1e30

MPC_OK = 0
MPC_ERR_INVALID = -1
MPC_ERR_NO_DEVICE = -2
MPC_ERR_CUDA = -3
MPC_ERR_STATE = -4
MPC_ERR_UNSUPPORTED = -5

STATUS_SOLVED = 1
STATUS_MAX_ITER_REACHED = -2
STATUS_PRIMAL_INFEASIBLE = -3
STATUS_DUAL_INFEASIBLE = -4
STATUS_UNSOLVED = -10


class MpcStateIn(C.Structure):
    _fields_ = [
        ("euler", C.c_float * 3),
        ("pos", C.c_float * 3),
        ("ang_vel", C.c_float * 3),
        ("lin_vel", C.c_float * 3),
        ("euler_d", C.c_float * 3),
        ("pos_d_z", C.c_float),
        ("lin_vel_d", C.c_float * 3),
        ("ang_vel_d", C.c_float * 3),
        ("rot_mat", C.c_float * 9),
        ("foot_pos_abs", C.c_float * 12),
        ("contacts", C.c_float * 4),
        ("pad", C.c_float),
    ]


class BalanceStateIn(C.Structure):
    _fields_ = [
        ("euler", C.c_float * 3),
        ("pos", C.c_float * 3),
        ("ang_vel", C.c_float * 3),
        ("lin_vel", C.c_float * 3),
        ("euler_d", C.c_float * 3),
        ("pos_d", C.c_float * 3),
        ("lin_vel_d", C.c_float * 3),
        ("ang_vel_d", C.c_float * 3),
        ("rot_mat", C.c_float * 9),
        ("rot_mat_z", C.c_float * 9),
        ("foot_pos_abs", C.c_float * 12),
        ("contacts", C.c_float * 4),
        ("pad", C.c_float * 6),
    ]


class MpcResult(C.Structure):
    _fields_ = [
        ("grf", C.c_float * 12),
        ("status", C.c_int32),
        ("iters", C.c_int32),
        ("rho_updates", C.c_int32),
        ("pri_res", C.c_float),
    ]


class MpcSolverSettings(C.Structure):
    _fields_ = [
        ("rho", C.c_double),
        ("sigma", C.c_double),
        ("alpha", C.c_double),
        ("eps_abs", C.c_double),
        ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double),
        ("eps_dual_inf", C.c_double),
        ("max_iter", C.c_int32),
        ("check_termination", C.c_int32),
        ("scaling", C.c_int32),
        ("adaptive_rho", C.c_int32),
        ("adaptive_rho_interval", C.c_int32),
        ("adaptive_rho_tolerance", C.c_double),
    ]


class MpcConfig(C.Structure):
    _fields_ = [
        ("horizon", C.c_int32),
        ("reserved0", C.c_int32),
        ("dt", C.c_double),
        ("mu", C.c_double),
        ("fz_min", C.c_double),
        ("fz_max", C.c_double),
        ("mass", C.c_double),
        ("inertia", C.c_double * 9),
        ("q_weights", C.c_double * 13),
        ("r_weights", C.c_double * 12),
        ("osqp", MpcSolverSettings),
        ("exact_discretization", C.c_int32),
        ("foot_drift", C.c_int32),
        ("gait_aware", C.c_int32),
        ("structured_solver", C.c_int32),
    ]


class BalanceConfig(C.Structure):
    _fields_ = [
        ("Q", C.c_double * 6),
        ("R", C.c_double),
        ("mu", C.c_double),
        ("F_min", C.c_double),
        ("F_max", C.c_double),
        ("mass", C.c_double),
        ("kp_linear", C.c_double * 3),
        ("kd_linear", C.c_double * 3),
        ("kp_angular", C.c_double * 3),
        ("kd_angular", C.c_double * 3),
        ("osqp", MpcSolverSettings),
    ]


# numpy views of the records, for slicing batches without per-element ctypes.
STATE_DTYPE = np.dtype(
    [
        ("euler", "<f4", 3),
        ("pos", "<f4", 3),
        ("ang_vel", "<f4", 3),
        ("lin_vel", "<f4", 3),
        ("euler_d", "<f4", 3),
        ("pos_d_z", "<f4"),
        ("lin_vel_d", "<f4", 3),
        ("ang_vel_d", "<f4", 3),
        ("rot_mat", "<f4", 9),
        ("foot_pos_abs", "<f4", 12),
        ("contacts", "<f4", 4),
        ("pad", "<f4"),
    ]
)
BALANCE_DTYPE = np.dtype(
    [
        ("euler", "<f4", 3),
        ("pos", "<f4", 3),
        ("ang_vel", "<f4", 3),
        ("lin_vel", "<f4", 3),
        ("euler_d", "<f4", 3),
        ("pos_d", "<f4", 3),
        ("lin_vel_d", "<f4", 3),
        ("ang_vel_d", "<f4", 3),
        ("rot_mat", "<f4", 9),
        ("rot_mat_z", "<f4", 9),
        ("foot_pos_abs", "<f4", 12),
        ("contacts", "<f4", 4),
        ("pad", "<f4", 6),
    ]
)
RESULT_DTYPE = np.dtype(
    [
        ("grf", "<f4", 12),
        ("status", "<i4"),
        ("iters", "<i4"),
        ("rho_updates", "<i4"),
        ("pri_res", "<f4"),
    ]
)

TORQUE_IN_DTYPE = np.dtype(
    [
        ("j_foot", "<f4", 36),
        ("foot_forces_kin", "<f4", 12),
        ("km_foot", "<f4", 3),
        ("torques_gravity", "<f4", 12),
        ("pad", "<f4"),
    ]
)
TORQUE_OUT_DTYPE = np.dtype([("joint_torques", "<f4", 12), ("nan_mask", "<i4"), ("pad", "<i4", 3)])
assert TORQUE_IN_DTYPE.itemsize == 256 and TORQUE_OUT_DTYPE.itemsize == 64

GAIT_DTYPE = np.dtype([("gait_counter", "<f4", 4), ("gait_counter_speed", "<f4", 4), ("counter_per_gait", "<f4"),
                       ("counter_per_swing", "<f4"), ("ticks_per_step", "<f4"), ("pad", "<f4")])
assert GAIT_DTYPE.itemsize == 48
SENSOR_DTYPE = np.dtype(
    [
        ("joint_pos", "<f4", 12), ("joint_vel", "<f4", 12), ("root_quat", "<f4", 4), ("imu_acc", "<f4", 3),
        ("imu_ang_vel", "<f4", 3), ("foot_force", "<f4", 4), ("root_pos", "<f4", 3), ("root_lin_vel", "<f4", 3),
        ("root_euler_d", "<f4", 3), ("root_pos_d_z", "<f4"), ("root_lin_vel_d", "<f4", 3), ("root_ang_vel_d", "<f4", 3),
        ("contacts", "<f4", 4), ("foot_pos_recent_contact", "<f4", 12), ("foot_forces_kin", "<f4", 12),
        ("movement_mode", "<f4"), ("dt", "<f4"), ("pad", "<f4", 12),
    ]
)
PREP_OUT_DTYPE = np.dtype(
    [
        ("root_euler", "<f4", 3), ("root_rot_mat", "<f4", 9), ("root_ang_vel", "<f4", 3), ("foot_pos_rel", "<f4", 12),
        ("foot_vel_rel", "<f4", 12), ("foot_pos_abs", "<f4", 12), ("foot_vel_abs", "<f4", 12),
        ("foot_pos_world", "<f4", 12), ("foot_vel_world", "<f4", 12), ("estimated_root_pos", "<f4", 3),
        ("estimated_root_vel", "<f4", 3), ("estimated_contacts", "<f4", 4), ("terrain_pitch_angle", "<f4"),
        ("root_euler_d_pitch", "<f4"), ("pad", "<f4", 29),
    ]
)
assert SENSOR_DTYPE.itemsize == 384 and PREP_OUT_DTYPE.itemsize == 512


class PrepConfig(C.Structure):
    _fields_ = [
        ("rho_fix", C.c_double * 20),
        ("km_foot", C.c_double * 3),
        ("torques_gravity", C.c_double * 12),
        ("use_estimator", C.c_int32),
        ("assume_flat_ground", C.c_int32),
        ("use_terrain_adapt", C.c_int32),
        ("pad", C.c_int32),
    ]


assert C.sizeof(MpcStateIn) == 192 == STATE_DTYPE.itemsize
assert C.sizeof(BalanceStateIn) == 256 == BALANCE_DTYPE.itemsize
assert C.sizeof(MpcResult) == 64 == RESULT_DTYPE.itemsize
