"""Batched host classes for the QP-based ZMP methods: LinearMpcZmp and IntrinsicallyStableMpc.

Each mirrors its reference class (constructor = QP structure, procOnce = per-tick coefficients +
post-processing) with a batch axis over initial states / reference sequences, and hands the QP itself
to a `qp_solve(QpProblemSet) -> QpResultArrays` callable: the CUDA engine (engine.QpEngine.solve) in
production, the oracle in the CPU-tier tests.
"""
import math

import numpy as np

from .linear_models import G, ComZmpModelJerkInput, InvariantSequentialExtension
from .qp import QpProblemSet


class LinearMpcZmp1d:
    """reference src/LinearMpcZmp.cpp:9-81 (Wieber 2006): jerk sequence with minimal norm keeping the ZMP
    inside its limits over the horizon."""

    def __init__(self, com_height, horizon_duration, horizon_dt):
        self.horizon_dt = horizon_dt
        self.horizon_steps = int(math.ceil(horizon_duration / horizon_dt))  # :13
        self.model = ComZmpModelJerkInput(com_height).calc_disc_matrix(horizon_dt)
        self.seq_ext = InvariantSequentialExtension(self.model, self.horizon_steps, True)
        N = self.horizon_steps
        self.Q = np.eye(N)                                             # obj_mat_ (:23)
        self.C = np.vstack([-self.seq_ext.B_seq, self.seq_ext.B_seq])  # ineq_mat_ (:25)

    def build_qp(self, initial_param, zmp_limits):
        """initial_param [B][3] (pos, vel, acc); zmp_limits [B][N][2] -> QpProblemSet (:54-60)."""
        # A_seq x0 as (a0 x0 + a1 x1) + a2 x2, the order of the device assembly (csrc/zmp_mpc.cu)
        A = self.seq_ext.A_seq
        ax0 = (A[None, :, 0] * initial_param[:, 0:1] + A[None, :, 1] * initial_param[:, 1:2]) + A[None, :, 2] * initial_param[:, 2:3]
        d = np.concatenate([ax0 - zmp_limits[:, :, 0], -ax0 + zmp_limits[:, :, 1]], axis=1)
        return QpProblemSet(self.Q, self.C, d)

    def proc_once(self, qp_solve, initial_param, zmp_limits, control_dt=-1.0):
        """-> planned ZMP [B] (:46-81)."""
        initial_param = np.atleast_2d(np.asarray(initial_param, dtype=np.float64))
        zmp_limits = np.asarray(zmp_limits, dtype=np.float64).reshape(len(initial_param), self.horizon_steps, 2)
        res = qp_solve(self.build_qp(initial_param, zmp_limits))
        com_jerk = res.x[:, 0]
        if control_dt < 0:
            control_dt = self.horizon_dt
        com_acc = initial_param[:, 2] + control_dt * com_jerk
        com_pos = (initial_param[:, 0] + control_dt * initial_param[:, 1]) + (0.5 * (control_dt * control_dt)) * initial_param[:, 2]
        zmp = np.clip(com_pos + self.model.C[0, 2] * com_acc, zmp_limits[:, 0, 0], zmp_limits[:, 0, 1])
        self.last_result = res
        return zmp


class LinearMpcZmp:
    """Two axes sharing one 1-D structure (reference src/LinearMpcZmp.cpp:83-112)."""

    def __init__(self, com_height, horizon_duration, horizon_dt):
        self.mpc_1d = LinearMpcZmp1d(com_height, horizon_duration, horizon_dt)

    def plan_batch(self, qp_solve, pos, vel, acc, zmp_limits_min, zmp_limits_max, control_dt=-1.0):
        """pos/vel/acc [B][2]; zmp_limits_min/max [B][N][2] (x, y) -> planned ZMP [B][2].
        Both axes go to the QP engine as one batch of 2B one-dimensional problems."""
        B = len(pos)
        ip = np.concatenate([np.stack([pos[:, a], vel[:, a], acc[:, a]], axis=1) for a in range(2)], axis=0)
        lim = np.concatenate([np.stack([zmp_limits_min[:, :, a], zmp_limits_max[:, :, a]], axis=2) for a in range(2)], axis=0)
        z = self.mpc_1d.proc_once(qp_solve, ip, lim, control_dt)
        return np.stack([z[:B], z[B:]], axis=1)


class IntrinsicallyStableMpc1d:
    """reference src/IntrinsicallyStableMpc.cpp:8-104 (Scianca et al. 2016)."""

    def __init__(self, com_height, horizon_duration, horizon_dt, weight_zmp=1.0, weight_zmp_vel=1e-3):
        self.horizon_dt = horizon_dt
        self.horizon_steps = N = int(math.ceil(horizon_duration / horizon_dt))
        self.weight_zmp, self.weight_zmp_vel = weight_zmp, weight_zmp_vel
        self.omega = math.sqrt(G / com_height)
        self.lam = math.exp(-1 * self.omega * horizon_dt)
        self.P = horizon_dt * np.tril(np.ones((N, N)))                      # :18-25
        self.Q = weight_zmp_vel * np.eye(N) + weight_zmp * self.P.T @ self.P  # :30-33
        a = np.zeros(N)
        a[0] = (1 - self.lam) / (self.omega * (1 - self.lam**N))           # :36
        for i in range(1, N):
            a[i] = self.lam * a[i - 1]
        self.A = a[None, :]
        self.C = np.vstack([-self.P, self.P])                               # :42

    def build_qp(self, capture_point, planned_zmp, ref_zmp, zmp_limits):
        """capture_point, planned_zmp [B]; ref_zmp [B][N]; zmp_limits [B][N][2] (:63-91)."""
        b = (capture_point - planned_zmp)[:, None]
        # w_zmp P' (z0 1 - z_ref): the sum over the stages runs sequentially (the order of the device assembly)
        v = planned_zmp[:, None] - ref_zmp
        acc = np.zeros_like(v)
        for i in range(self.horizon_steps):  # P is lower triangular: the skipped terms are exact zeros
            acc[:, :i + 1] = acc[:, :i + 1] + self.P[i][None, :i + 1] * v[:, i:i + 1]
        c = self.weight_zmp * acc
        d = np.concatenate([-zmp_limits[:, :, 0] + planned_zmp[:, None], zmp_limits[:, :, 1] - planned_zmp[:, None]], axis=1)
        return QpProblemSet(self.Q, self.C, d, self.A, b, c)

    def proc_once(self, qp_solve, capture_point, planned_zmp, ref_zmp, zmp_limits, control_dt=-1.0):
        capture_point, planned_zmp = np.atleast_1d(capture_point).astype(float), np.atleast_1d(planned_zmp).astype(float)
        B = len(capture_point)
        ref_zmp = np.asarray(ref_zmp, dtype=np.float64).reshape(B, self.horizon_steps)
        zmp_limits = np.asarray(zmp_limits, dtype=np.float64).reshape(B, self.horizon_steps, 2)
        res = qp_solve(self.build_qp(capture_point, planned_zmp, ref_zmp, zmp_limits))
        if control_dt < 0:
            control_dt = self.horizon_dt
        self.last_result = res
        return np.clip(planned_zmp + control_dt * res.x[:, 0], zmp_limits[:, 0, 0], zmp_limits[:, 0, 1])


class IntrinsicallyStableMpc:
    """reference src/IntrinsicallyStableMpc.cpp:106-139."""

    def __init__(self, com_height, horizon_duration, horizon_dt, weight_zmp=1.0, weight_zmp_vel=1e-3):
        self.mpc_1d = IntrinsicallyStableMpc1d(com_height, horizon_duration, horizon_dt, weight_zmp, weight_zmp_vel)

    def plan_batch(self, qp_solve, capture_point, planned_zmp, ref_zmp, zmp_limits_min, zmp_limits_max, control_dt=-1.0):
        """capture_point, planned_zmp [B][2]; ref_zmp, zmp_limits_min/max [B][N][2] -> planned ZMP [B][2]."""
        B = len(capture_point)
        cat = lambda f: np.concatenate([f(0), f(1)], axis=0)
        z = self.mpc_1d.proc_once(
            qp_solve, cat(lambda a: capture_point[:, a]), cat(lambda a: planned_zmp[:, a]), cat(lambda a: ref_zmp[:, :, a]),
            cat(lambda a: np.stack([zmp_limits_min[:, :, a], zmp_limits_max[:, :, a]], axis=2)), control_dt)
        return np.stack([z[:B], z[B:]], axis=1)
