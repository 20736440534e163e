"""Host-side containers that own the numpy buffers behind the C-ABI structs."""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import ptr


def default_alpha_list(n=11):
    """alpha_list = 10^linspace(0,-3,11) of nmpc_ddp, evaluated on the host."""
    return [10.0 ** (-3.0 * i / (n - 1)) for i in range(n)]


def ddp_config(with_input_constraint=False, **overrides):
    """nmpc_ddp::DDPSolver::Configuration defaults as a ccc_ddp_config_t."""
    c = _abi.DdpConfig()
    c.with_input_constraint = int(with_input_constraint)
    c.max_iter = 500
    c.reg_type = 1
    c.initial_lambda = 1e-4
    c.initial_dlambda = 1.0
    c.lambda_factor = 1.6
    c.lambda_min = 1e-6
    c.lambda_max = 1e10
    c.k_rel_norm_thre = 1e-4
    c.lambda_thre = 1e-5
    c.cost_update_ratio_thre = 0.0
    c.cost_update_thre = 1e-7
    al = default_alpha_list()
    c.n_alpha = len(al)
    for i, a in enumerate(al):
        c.alpha[i] = a
    c.boxqp_max_iter = 500
    c.boxqp_grad_thre = 1e-8
    c.boxqp_rel_improve_thre = 1e-8
    c.boxqp_step_factor = 0.6
    c.boxqp_min_step = 1e-22
    c.boxqp_armijo = 0.1
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def ddp_srb_config(**overrides):
    """Solver configuration of CCC::DdpSingleRigidBody's constructor (reference src/DdpSingleRigidBody.cpp:267-271)."""
    return ddp_centroidal_config(**overrides)


def ddp_centroidal_config(**overrides):
    """Solver configuration of CCC::DdpCentroidal's constructor (reference src/DdpCentroidal.cpp:197-201)."""
    kw = dict(initial_lambda=1e-6, lambda_min=1e-8, lambda_thre=1e-7)
    kw.update(overrides)
    return ddp_config(with_input_constraint=True, **kw)


class DdpResultArrays:
    """numpy buffers + ccc_ddp_result_t view."""

    def __init__(self, batch, horizon_steps, nx, m_max, trace_len=0):
        B, N = batch, horizon_steps
        self.x = np.zeros((B, N + 1, nx))
        self.u = np.zeros((B, N, m_max))
        self.cost = np.zeros(B)
        self.iters = np.zeros(B, dtype=np.int32)
        self.status = np.zeros(B, dtype=np.int32)
        self.trace_len = trace_len
        self.alpha_idx = np.full((B, max(trace_len, 1)), -4, dtype=np.int8)
        self.lambda_trace = np.zeros((B, max(trace_len, 1)))
        self.clamped = np.zeros((B, N), dtype=np.uint32)

    def as_struct(self):
        r = _abi.DdpResult()
        r.x, r.u, r.cost = ptr(self.x), ptr(self.u), ptr(self.cost)
        r.iters, r.status = ptr(self.iters), ptr(self.status)
        r.trace_len = self.trace_len
        r.alpha_idx = ptr(self.alpha_idx) if self.trace_len else None
        r.lambda_trace = ptr(self.lambda_trace) if self.trace_len else None
        r.clamped = ptr(self.clamped)
        return r


class DdpCentroidalProblemSet:
    """A batch of DdpCentroidal problems: shared schedules + per-problem initial states."""

    nx = 9

    def __init__(self, sched, sched_id, x0, mass, dt, w_run, w_term, u_lo=0.0, u_hi=1e6, u_init=None):
        self.sched = sched
        self.sched_id = np.ascontiguousarray(sched_id, dtype=np.int32)
        self.x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.mass, self.dt = float(mass), float(dt)
        self.w_run = np.asarray(w_run, dtype=np.float64)
        self.w_term = np.asarray(w_term, dtype=np.float64)
        self.u_lo, self.u_hi = float(u_lo), float(u_hi)
        self.u_init = None if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
        assert self.x0.shape == (len(self.sched_id), 9)
        assert self.w_run.shape == (10,) and self.w_term.shape == (9,)

    @classmethod
    def from_workload(cls, w):
        return cls(w["sched"], w["sched_id"], w["x0"], w["mass"], w["dt"], w["w_run"], w["w_term"], w["u_lo"], w["u_hi"])

    @property
    def batch(self):
        return len(self.sched_id)

    @property
    def N(self):
        return self.sched.N

    @property
    def m_max(self):
        return self.sched.m_max

    def subset(self, idx):
        idx = np.asarray(idx)
        return DdpCentroidalProblemSet(
            self.sched, self.sched_id[idx], self.x0[idx], self.mass, self.dt, self.w_run, self.w_term,
            self.u_lo, self.u_hi, None if self.u_init is None else self.u_init[idx])

    def as_struct(self):
        b = _abi.DdpCentroidalBatch()
        s = self.sched
        b.horizon_steps, b.batch, b.n_sched, b.m_max = s.N, self.batch, s.S, s.m_max
        b.dt, b.mass = self.dt, self.mass
        b.sched_id, b.m = ptr(self.sched_id), ptr(s.m)
        b.ridge, b.vertex, b.ref_pos = ptr(s.ridge), ptr(s.vertex), ptr(s.ref_pos)
        for i in range(10):
            b.w_run[i] = self.w_run[i]
        for i in range(9):
            b.w_term[i] = self.w_term[i]
        b.u_lo, b.u_hi = self.u_lo, self.u_hi
        b.x0 = ptr(self.x0)
        b.u_init = ptr(self.u_init)
        return b

    def new_result(self, trace_len=0):
        return DdpResultArrays(self.batch, self.N, 9, self.m_max, trace_len)


class DdpSrbProblemSet:
    """A batch of DdpSingleRigidBody problems: shared schedules + per-problem initial states (12)."""

    nx = 12

    def __init__(self, sched, sched_id, x0, mass, dt, w_run, w_term, u_lo=0.0, u_hi=1e6, u_init=None):
        self.sched = sched
        self.sched_id = np.ascontiguousarray(sched_id, dtype=np.int32)
        self.x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.mass, self.dt = float(mass), float(dt)
        self.w_run = np.asarray(w_run, dtype=np.float64)
        self.w_term = np.asarray(w_term, dtype=np.float64)
        self.u_lo, self.u_hi = float(u_lo), float(u_hi)
        self.u_init = None if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
        assert self.x0.shape == (len(self.sched_id), 12)
        assert self.w_run.shape == (13,) and self.w_term.shape == (12,)

    @classmethod
    def from_workload(cls, w):
        return cls(w["sched"], w["sched_id"], w["x0"], w["mass"], w["dt"], w["w_run"], w["w_term"], w["u_lo"], w["u_hi"])

    batch = property(lambda s: len(s.sched_id))
    N = property(lambda s: s.sched.N)
    m_max = property(lambda s: s.sched.m_max)

    def subset(self, idx):
        idx = np.asarray(idx)
        return DdpSrbProblemSet(self.sched, self.sched_id[idx], self.x0[idx], self.mass, self.dt, self.w_run, self.w_term,
                                self.u_lo, self.u_hi, None if self.u_init is None else self.u_init[idx])

    def as_struct(self):
        b = _abi.DdpSrbBatch()
        s = self.sched
        b.horizon_steps, b.batch, b.n_sched, b.m_max = s.N, self.batch, s.S, s.m_max
        b.dt, b.mass = self.dt, self.mass
        b.sched_id, b.m = ptr(self.sched_id), ptr(s.m)
        b.ridge, b.vertex, b.inertia, b.ref = ptr(s.ridge), ptr(s.vertex), ptr(s.inertia), ptr(s.ref)
        for i in range(13):
            b.w_run[i] = self.w_run[i]
        for i in range(12):
            b.w_term[i] = self.w_term[i]
        b.u_lo, b.u_hi = self.u_lo, self.u_hi
        b.x0 = ptr(self.x0)
        b.u_init = ptr(self.u_init)
        return b

    def new_result(self, trace_len=0):
        return DdpResultArrays(self.batch, self.N, 12, self.m_max, trace_len)


class DdpZmpProblemSet:
    """A batch of DdpZmp problems: shared reference schedules (ref ZMP, CoM height) + initial states (6)."""

    nx = 6
    m_max = 3

    def __init__(self, ref_zmp, com_z, sched_id, x0, mass, dt, weights=(1e2, 1e-1, 1e-4, 1.0, 1e2, 1.0), u_init=None):
        """ref_zmp [S][N+1][3], com_z [S][N+1]; weights = DdpZmp::WeightParam defaults (reference
        include/CCC/DdpZmp.h:71-76): running_com_pos_z, running_zmp, running_force_z, terminal_com_pos_xy,
        terminal_com_pos_z, terminal_com_vel."""
        self.ref_zmp = np.ascontiguousarray(ref_zmp, dtype=np.float64)
        self.com_z = np.ascontiguousarray(com_z, dtype=np.float64)
        self.sched_id = np.ascontiguousarray(sched_id, dtype=np.int32)
        self.x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.mass, self.dt = float(mass), float(dt)
        self.weights = np.asarray(weights, dtype=np.float64)
        self.u_init = None if u_init is None else np.ascontiguousarray(u_init, dtype=np.float64)
        self.S, self.N = self.ref_zmp.shape[0], self.ref_zmp.shape[1] - 1
        assert self.ref_zmp.shape == (self.S, self.N + 1, 3) and self.com_z.shape == (self.S, self.N + 1)
        assert self.x0.shape == (len(self.sched_id), 6)

    batch = property(lambda s: len(s.sched_id))

    def as_struct(self):
        b = _abi.DdpZmpBatch()
        b.horizon_steps, b.batch, b.n_sched = self.N, self.batch, self.S
        b.dt, b.mass = self.dt, self.mass
        b.sched_id, b.ref_zmp, b.com_z = ptr(self.sched_id), ptr(self.ref_zmp), ptr(self.com_z)
        for i in range(6):
            b.w[i] = self.weights[i]
        b.x0, b.u_init = ptr(self.x0), ptr(self.u_init)
        return b

    def new_result(self, trace_len=0):
        return DdpResultArrays(self.batch, self.N, 6, 3, trace_len)
