"""Host-side mirror of the reference interface for the hot path.

Same names, argument meaning and error behaviour as
  class ConvexMpc            src/a1_cpp/src/ConvexMpc.h:22-92
  A1CtrlStates (subset)      src/a1_cpp/src/A1CtrlStates.h:330-447
  A1RobotControl::compute_grf src/a1_cpp/src/A1RobotControl.h:44
so parity tests read like the reference's own test/test_mpc.cpp.  The cheap
setters run on the host (they fill a 13x13 and a 13x12 matrix); the heavy
calls -- calculate_qp_mats and the OSQP solve -- run on the GPU through the C
ABI.  Nothing here falls back to a CPU solve.
"""
import numpy as np

from . import abi
from .engine import MpcEngine, config_default

PLAN_HORIZON = 10
MPC_STATE_DIM = 13
MPC_CONSTRAINT_DIM = 20
NUM_LEG = 4
NUM_DOF = 12


def skew(v):
    """Utils::skew, utils/Utils.cpp:35-41."""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


class A1CtrlStates:
    """The fields compute_grf reads and writes (A1CtrlStates.h:330-447), defaults of reset()."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.stance_leg_control_type = 1
        self.use_terrain_adapt = 1
        self.robot_mass = 15.0
        self.a1_trunk_inertia = np.diag([0.0168352186, 0.0656071082, 0.0742720659])
        self.q_weights = np.array([80.0, 80.0, 1.0, 0.0, 0.0, 270.0, 1.0, 1.0, 20.0, 20.0, 20.0, 20.0, 0.0])
        self.r_weights = np.array([1e-5, 1e-5, 1e-6] * 4)
        self.root_pos = np.zeros(3)
        self.root_euler = np.zeros(3)
        self.root_rot_mat = np.zeros((3, 3))
        self.root_rot_mat_z = np.zeros((3, 3))
        self.root_lin_vel = np.zeros(3)
        self.root_ang_vel = np.zeros(3)
        self.root_pos_d = np.zeros(3)
        self.root_euler_d = np.zeros(3)
        self.root_lin_vel_d = np.zeros(3)
        self.root_lin_vel_d_world = np.zeros(3)
        self.root_ang_vel_d = np.zeros(3)
        self.foot_pos_abs = np.zeros((3, NUM_LEG))  # column per leg
        self.foot_pos_rel = np.zeros((3, NUM_LEG))
        self.contacts = [False] * NUM_LEG
        self.mpc_states = np.zeros(MPC_STATE_DIM)
        self.mpc_states_d = np.zeros(MPC_STATE_DIM * PLAN_HORIZON)
        self.foot_forces_grf = np.zeros((3, NUM_LEG))
        self.kp_linear = np.array([1000.0, 1000.0, 1000.0])
        self.kd_linear = np.array([200.0, 70.0, 120.0])
        self.kp_angular = np.array([650.0, 35.0, 1.0])
        self.kd_angular = np.array([4.5, 4.5, 30.0])
        # compute_joint_torques inputs/outputs (A1CtrlStates.h:95,122,129)
        self.j_foot = np.eye(NUM_DOF)
        self.foot_forces_kin = np.zeros((3, NUM_LEG))
        self.km_foot = np.array([0.1, 0.1, 0.1])
        self.torques_gravity = np.array([0.80, 0, 0, -0.80, 0, 0, 0.80, 0, 0, -0.80, 0, 0])
        self.joint_torques = np.zeros(NUM_DOF)

    def to_record(self):
        """Pack into the 192 B MpcStateIn record (include/mpc_b200.h)."""
        r = np.zeros(1, dtype=abi.STATE_DTYPE)
        r["euler"][0] = self.root_euler
        r["pos"][0] = self.root_pos
        r["ang_vel"][0] = self.root_ang_vel
        r["lin_vel"][0] = self.root_lin_vel
        r["euler_d"][0] = self.root_euler_d
        r["pos_d_z"][0] = self.root_pos_d[2]
        r["lin_vel_d"][0] = self.root_lin_vel_d
        r["ang_vel_d"][0] = self.root_ang_vel_d
        r["rot_mat"][0] = np.asarray(self.root_rot_mat).reshape(9)
        r["foot_pos_abs"][0] = np.asarray(self.foot_pos_abs).T.reshape(12)
        r["contacts"][0] = [1.0 if c else 0.0 for c in self.contacts]
        return r

    def to_torque_record(self):
        """Pack the compute_joint_torques inputs into the 256 B MpcTorqueIn record."""
        r = np.zeros(1, dtype=abi.TORQUE_IN_DTYPE)
        J = np.asarray(self.j_foot)
        r["j_foot"][0] = np.concatenate([J[3 * i:3 * i + 3, 3 * i:3 * i + 3].reshape(9) for i in range(NUM_LEG)])
        r["foot_forces_kin"][0] = np.asarray(self.foot_forces_kin).T.reshape(12)
        r["km_foot"][0] = self.km_foot
        r["torques_gravity"][0] = self.torques_gravity
        return r

    def to_balance_record(self):
        r = np.zeros(1, dtype=abi.BALANCE_DTYPE)
        r["euler"][0] = self.root_euler
        r["pos"][0] = self.root_pos
        r["ang_vel"][0] = self.root_ang_vel
        r["lin_vel"][0] = self.root_lin_vel
        r["euler_d"][0] = self.root_euler_d
        r["pos_d"][0] = self.root_pos_d
        r["lin_vel_d"][0] = self.root_lin_vel_d
        r["ang_vel_d"][0] = self.root_ang_vel_d
        r["rot_mat"][0] = np.asarray(self.root_rot_mat).reshape(9)
        r["rot_mat_z"][0] = np.asarray(self.root_rot_mat_z).reshape(9)
        r["foot_pos_abs"][0] = np.asarray(self.foot_pos_abs).T.reshape(12)
        r["contacts"][0] = [1.0 if c else 0.0 for c in self.contacts]
        return r


class ConvexMpc:
    """ConvexMpc(q_weights, r_weights) with the reference's five methods and public members."""

    def __init__(self, q_weights_, r_weights_, engine=None, horizon=PLAN_HORIZON):
        self.horizon = horizon
        self.mu = 0.3
        self.fz_min = 0.0
        self.fz_max = 0.0
        q = np.asarray(q_weights_, dtype=np.float64)
        r = np.asarray(r_weights_, dtype=np.float64)
        if q.shape != (MPC_STATE_DIM,) or r.shape != (NUM_DOF,):
            raise ValueError("q_weights must have 13 entries and r_weights 12")
        self.q_weights_mpc = np.tile(q, horizon)
        self.r_weights_mpc = np.tile(r, horizon)
        self.Q = 2.0 * self.q_weights_mpc  # diagonal (ConvexMpc.cpp:20)
        self.R = 2.0 * self.r_weights_mpc  # diagonal (ConvexMpc.cpp:41)
        n, m = NUM_DOF * horizon, MPC_CONSTRAINT_DIM * horizon
        lc = np.zeros((m, n))
        for i in range(NUM_LEG * horizon):  # ConvexMpc.cpp:46-58
            lc[5 * i + 0, 3 * i] = 1
            lc[5 * i + 1, 3 * i] = 1
            lc[5 * i + 2, 3 * i + 1] = 1
            lc[5 * i + 3, 3 * i + 1] = 1
            lc[5 * i + 4, 3 * i + 2] = 1
            lc[5 * i + 0, 3 * i + 2] = self.mu
            lc[5 * i + 1, 3 * i + 2] = -self.mu
            lc[5 * i + 2, 3 * i + 2] = self.mu
            lc[5 * i + 3, 3 * i + 2] = -self.mu
        self.linear_constraints = lc
        if engine is None:
            cfg = config_default()
            cfg.horizon = horizon
            for i in range(13):
                cfg.q_weights[i] = q[i]
            for i in range(12):
                cfg.r_weights[i] = r[i]
            engine = MpcEngine(cfg)
        self._engine = engine
        self.reset()

    def reset(self):
        """ConvexMpc.cpp:70-108."""
        H = self.horizon
        self.A_mat_c = np.zeros((13, 13))
        self.B_mat_c = np.zeros((13, 12))
        self.A_mat_d = np.zeros((13, 13))
        self.B_mat_d = np.zeros((13, 12))
        self.B_mat_d_list = np.zeros((13 * H, 12))
        self.A_qp = np.zeros((13 * H, 13))
        self.B_qp = np.zeros((13 * H, 12 * H))
        self.hessian = np.zeros((12 * H, 12 * H))
        self.gradient = np.zeros(12 * H)
        self.lb = np.zeros(20 * H)
        self.ub = np.zeros(20 * H)

    def calculate_A_mat_c(self, root_euler):
        """ConvexMpc.cpp:110-130 (yaw only)."""
        cy, sy = np.cos(root_euler[2]), np.sin(root_euler[2])
        self.A_mat_c[0:3, 6:9] = [[cy, sy, 0.0], [-sy, cy, 0.0], [0.0, 0.0, 1.0]]
        self.A_mat_c[3:6, 9:12] = np.eye(3)
        self.A_mat_c[11, NUM_DOF] = 1.0

    def calculate_B_mat_c(self, robot_mass, a1_trunk_inertia, root_rot_mat, foot_pos):
        """ConvexMpc.cpp:132-143; foot_pos is 3x4, column per leg."""
        Iw = root_rot_mat @ a1_trunk_inertia @ root_rot_mat.T
        Iw_inv = np.linalg.inv(Iw)
        for i in range(NUM_LEG):
            self.B_mat_c[6:9, 3 * i:3 * i + 3] = Iw_inv @ skew(foot_pos[:, i])
            self.B_mat_c[9:12, 3 * i:3 * i + 3] = np.eye(3) / robot_mass

    def state_space_discretization(self, dt):
        """ConvexMpc.cpp:145-156 (forward Euler)."""
        self.A_mat_d = np.eye(13) + self.A_mat_c * dt
        self.B_mat_d = self.B_mat_c * dt

    def calculate_qp_mats(self, state):
        """ConvexMpc.cpp:158-245 on the GPU (general dense build kernel)."""
        P, q, l, u = self._engine.qp_mats_from_model(self.A_mat_d, self.B_mat_d_list, state.mpc_states,
                                                     state.mpc_states_d,
                                                     [1 if c else 0 for c in state.contacts])
        self.fz_min, self.fz_max = 0.0, 180.0
        self.hessian, self.gradient, self.lb, self.ub = P, q, l, u


class A1RobotControl:
    """compute_grf (A1RobotControl.cpp:321-564) for one robot or a batch."""

    def __init__(self, cfg=None, device=0, use_sim_time="false"):
        self.use_sim_time = use_sim_time
        self._cfg = cfg if cfg is not None else config_default()
        self._device = device
        self._engine = None
        self._balance = None
        self._balance_key = None
        self.last_status = None
        self.last_iters = None
        self.mpc_init_counter = 0
        self._torques = None

    def _mpc_engine(self, state, mpc_dt):
        """The engine whose slot 0 is this controller's member solver (A1RobotControl.h:67).  The
        reference re-reads mass, inertia and weights from the state on every call into a fresh
        ConvexMpc (A1RobotControl.cpp:447) while the OsqpEigen solver lives on: a change of those is
        one more Hessian update, not a new solver -- so the engine's model is updated in place."""
        cfg = self._cfg
        inertia = [float(v) for v in np.asarray(state.a1_trunk_inertia).reshape(9)]
        same = (abs(cfg.dt - mpc_dt) == 0.0 and cfg.mass == state.robot_mass
                and list(cfg.q_weights) == [float(v) for v in state.q_weights]
                and list(cfg.r_weights) == [float(v) for v in state.r_weights]
                and list(cfg.inertia) == inertia)
        if not same:
            cfg.dt = mpc_dt
            cfg.mass = state.robot_mass
            for i in range(13):
                cfg.q_weights[i] = state.q_weights[i]
            for i in range(12):
                cfg.r_weights[i] = state.r_weights[i]
            for i in range(9):
                cfg.inertia[i] = inertia[i]
            if self._engine is not None:
                self._engine.update_model(cfg)
        if self._engine is None:
            self._engine = MpcEngine(cfg, self._device)
        return self._engine

    def _balance_engine(self, state):
        """Gains and mass are read from the state on every tick (A1RobotControl.cpp:380-391): a change
        rebuilds the engine (this branch's solver is a per-call local in the reference, :416-432)."""
        from .engine import balance_config_default
        key = (float(state.robot_mass),) + tuple(float(v) for f in ("kp_linear", "kd_linear", "kp_angular", "kd_angular")
                                                 for v in getattr(state, f))
        if self._balance is None or key != self._balance_key:
            bcfg = balance_config_default()
            bcfg.mass = state.robot_mass
            for i in range(3):
                bcfg.kp_linear[i] = state.kp_linear[i]
                bcfg.kd_linear[i] = state.kd_linear[i]
                bcfg.kp_angular[i] = state.kp_angular[i]
                bcfg.kd_angular[i] = state.kd_angular[i]
            if self._balance is not None:
                self._balance.close()
            self._balance = MpcEngine(bcfg, self._device, balance=True)
            self._balance_key = key
        return self._balance

    def compute_grf(self, state, dt):
        """Returns the 3x4 body-frame GRF like the reference; also writes mpc_states(_d).

        MPC branch: ONE persistent, warm-started solver like the reference's member
        (A1RobotControl.cpp:522-540): the first call is initSolver (cold), every later call is
        updateHessianMatrix / updateGradient / updateBounds + a warm solve()."""
        if state.stance_leg_control_type == 1:
            mpc_dt = dt if self.use_sim_time == "true" else 0.0025  # :462-467
            eng = self._mpc_engine(state, mpc_dt)
            state.mpc_states = np.concatenate([state.root_euler, state.root_pos, state.root_ang_vel,
                                               state.root_lin_vel, [-9.8]])
            state.root_lin_vel_d_world = state.root_rot_mat @ state.root_lin_vel_d
            eng.load_states(state.to_record())
            eng.set_torque_inputs(state.to_torque_record())
            eng.build_qp(sync=False)
            eng.solve_warm(sync=False)
            res = eng.get_results()
            self._torques = eng.get_torques()[0]
        else:
            bal = self._balance_engine(state)
            bal.load_states(state.to_balance_record())
            bal.set_torque_inputs(state.to_torque_record())
            bal.solve(sync=False)
            res = bal.get_results()
            self._torques = bal.get_torques()[0]
        self.last_status = int(res["status"][0])
        self.last_iters = int(res["iters"][0])
        return np.asarray(res["grf"][0], dtype=np.float64).reshape(4, 3).T

    def reset_solver(self):
        """Forget the member solver (the next MPC call is an initSolver again)."""
        if self._engine is not None:
            self._engine.stream_reset()

    def compute_joint_torques(self, state):
        """A1RobotControl.cpp:289-319 with the torques the device wrote next to the last GRF:
        zero for the first ten calls, NaN components keep their previous value."""
        self.mpc_init_counter += 1
        if self.mpc_init_counter < 10:
            state.joint_torques = np.zeros(NUM_DOF)
            return
        if self._torques is None:
            raise RuntimeError("compute_joint_torques before compute_grf")
        tau = np.asarray(self._torques["joint_torques"], dtype=np.float64)
        mask = int(self._torques["nan_mask"])
        for i in range(NUM_DOF):
            if not (mask >> i) & 1:
                state.joint_torques[i] = tau[i]

    def compute_grf_batch(self, records):
        """MPC branch for a batch of MpcStateIn records (engine-wide constants from cfg)."""
        if self._engine is None:
            self._engine = MpcEngine(self._cfg, self._device)
        return self._engine.compute_grf_batch(records)
