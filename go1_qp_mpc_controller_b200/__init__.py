"""B200-native batched convex-MPC ground-reaction-force engine.

Drop-in for ONE hot path of zerenluo123/Go1-QP-MPC-Controller:
A1RobotControl::compute_grf -> ConvexMpc -> OSQP.  The product is
csrc/ (hand-written sm_100a kernels behind the C ABI of include/mpc_b200.h);
this package is the host-side mirror of the reference interface.
"""
from . import abi  # noqa: F401
from .engine import (MpcEngine, MpcFleet, MpcError, fleet_shard_range, measure_fp64_peak, balance_config_default, config_default,  # noqa: F401
                     config_hardware, generate_balance_states, generate_states, generate_stream_states, generate_torque_inputs, generate_sensors, generate_gait_inputs, prep_config_default,
                     a1_leg_fk_jac, load_library,
                     settings_osqp_default)
from .convex_mpc import A1CtrlStates, A1RobotControl, ConvexMpc  # noqa: F401
