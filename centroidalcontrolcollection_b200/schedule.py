"""Schedule compiler: sample planOnce()'s callbacks into flat per-stage tables.

The reference passes time-varying problem data as std::function callbacks evaluated at
t_k = current_time + k * horizon_dt inside the solver (reference src/DdpCentroidal.cpp:21-30,
:36, :68, :225).  Callbacks cannot cross to the device, so they are sampled once per schedule
on the host into the tables of ccc_ddp_centroidal_batch_t (include/ccc_b200.h).
"""
import numpy as np

from . import _abi
from ._abi import CCC_DDP_M_MAX, ptr


class CentroidalSchedule:
    """Stage tables of one or more contact schedules (S of them) over N stages."""

    def __init__(self, n_sched, horizon_steps, m_max=CCC_DDP_M_MAX):
        self.S, self.N, self.m_max = n_sched, horizon_steps, m_max
        self.m = np.zeros((n_sched, horizon_steps), dtype=np.int32)
        self.ridge = np.zeros((n_sched, horizon_steps, m_max, 3))
        self.vertex = np.zeros((n_sched, horizon_steps, m_max, 3))
        self.ref_pos = np.zeros((n_sched, horizon_steps + 1, 3))

    def sample(self, s, motion_param_func, ref_data_func, current_time, dt):
        """motion_param_func(t) -> list of (vertex[mi,3], ridge[mi,3]); ref_data_func(t) -> pos[3]."""
        for k in range(self.N + 1):
            t = current_time + k * dt
            self.ref_pos[s, k] = ref_data_func(t)
            if k == self.N:
                break
            contacts = motion_param_func(t)
            j = 0
            for vtx, rdg in contacts:
                n = len(vtx)
                if j + n > self.m_max:
                    raise ValueError(f"stage {k}: input dimension {j + n} exceeds m_max={self.m_max}")
                self.vertex[s, k, j : j + n] = vtx
                self.ridge[s, k, j : j + n] = rdg
                j += n
            self.m[s, k] = j
            self.vertex[s, k, j:] = 0.0
            self.ridge[s, k, j:] = 0.0
        return self


class SrbSchedule(CentroidalSchedule):
    """Stage tables for CCC::DdpSingleRigidBody: contacts + inertia matrix per stage, pos/ori reference."""

    def __init__(self, n_sched, horizon_steps, m_max=CCC_DDP_M_MAX):
        super().__init__(n_sched, horizon_steps, m_max)
        self.inertia = np.tile(np.eye(3).reshape(1, 1, 9), (n_sched, horizon_steps, 1))
        self.ref = np.zeros((n_sched, horizon_steps + 1, 6))

    def sample(self, s, motion_param_func, ref_data_func, current_time, dt):
        """motion_param_func(t) -> (contacts, inertia[3,3]); ref_data_func(t) -> (pos[3], ori[3])."""
        inertias = {}

        def contacts_only(t):
            contacts, inertia = motion_param_func(t)
            inertias[t] = inertia
            return contacts

        refs = {}

        def pos_only(t):
            pos, ori = ref_data_func(t)
            refs[t] = (pos, ori)
            return pos

        super().sample(s, contacts_only, pos_only, current_time, dt)
        for k in range(self.N + 1):
            t = current_time + k * dt
            self.ref[s, k, 0:3], self.ref[s, k, 3:6] = refs[t]
            if k < self.N:
                self.inertia[s, k] = np.asarray(inertias[t], dtype=np.float64).reshape(9)
        return self


class FootstepPlans:
    """P footstep plans in the flat form of ccc_footstep_plans_t: what the reference tests' FootstepManager holds
    (tests/src/FootstepManager.h: initial Footstance :136-137, the appended Footsteps :36-77, horizon_duration_ :461,
    foot_size_ :464), one row per plan, compiled into reference-ZMP / ZMP-limit stage tables by ccc_footstep_compile."""

    def __init__(self, n_plans, max_steps, horizon_steps, horizon_dt, eps_reps=2, manager_horizon=10.0, foot_size=(0.1, 0.05)):
        self.P, self.F, self.N = int(n_plans), int(max_steps), int(horizon_steps)
        self.horizon_dt, self.eps_reps, self.manager_horizon = float(horizon_dt), int(eps_reps), float(manager_horizon)
        self.foot_size = tuple(float(v) for v in foot_size)
        self.current_time = np.zeros(self.P)
        self.stance0 = np.tile(np.array([[0.0, 0.1], [0.0, -0.1]]), (self.P, 1, 1))
        self.n_steps = np.zeros(self.P, dtype=np.int32)
        self.foot = np.zeros((self.P, self.F), dtype=np.int32)
        self.pos = np.zeros((self.P, self.F, 2))
        self.times = np.zeros((self.P, self.F, 4))

    def append_footstep(self, p, foot, pos, transit_start_time, transit_duration, swing_duration):
        """Footstep's constructor (:66-76) + FootstepManager::appendFootstep (:214-223) for plan p."""
        i = int(self.n_steps[p])
        if i >= self.F:
            raise ValueError("more footsteps than max_steps")
        if i > 0 and transit_start_time < self.times[p, i - 1, 3]:
            raise RuntimeError("transit_start_time of specified footstep must be after transit_end_time of last footstep")
        self.foot[p, i] = foot
        self.pos[p, i] = pos
        self.times[p, i] = (transit_start_time, transit_start_time + 0.5 * transit_duration,
                            transit_start_time + 0.5 * transit_duration + swing_duration,
                            transit_start_time + transit_duration + swing_duration)
        self.n_steps[p] = i + 1

    def as_struct(self):
        s = _abi.FootstepPlans()
        s.n_plans, s.max_steps, s.horizon_steps, s.eps_reps = self.P, self.F, self.N, self.eps_reps
        s.horizon_dt, s.manager_horizon = self.horizon_dt, self.manager_horizon
        s.foot_size[0], s.foot_size[1] = self.foot_size
        s.current_time, s.stance0, s.n_steps = ptr(self.current_time), ptr(self.stance0), ptr(self.n_steps)
        s.foot, s.pos, s.times = ptr(self.foot), ptr(self.pos), ptr(self.times)
        return s


class ZmpTables:
    """ccc_zmp_tables_t on the host: ref_zmp, lim_min, lim_max [P][N][2]."""

    def __init__(self, n_plans, horizon_steps):
        self.ref_zmp = np.zeros((n_plans, horizon_steps, 2))
        self.lim_min = np.zeros((n_plans, horizon_steps, 2))
        self.lim_max = np.zeros((n_plans, horizon_steps, 2))

    def as_struct(self):
        t = _abi.ZmpTables()
        t.ref_zmp, t.lim_min, t.lim_max = ptr(self.ref_zmp), ptr(self.lim_min), ptr(self.lim_max)
        return t
