"""Schedule compiler: sample planOnce()'s callbacks into flat per-stage tables.

The reference passes time-varying problem data as std::function callbacks evaluated at
t_k = current_time + k * horizon_dt inside the solver (reference src/DdpCentroidal.cpp:21-30,
:36, :68, :225).  Callbacks cannot cross to the device, so they are sampled once per schedule
on the host into the tables of ccc_ddp_centroidal_batch_t (include/ccc_b200.h).
"""
import numpy as np

from ._abi import CCC_DDP_M_MAX


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
