"""Host-side description of a batched receding-horizon run of CCC::DdpCentroidal (ccc_ddp_centroidal_closed_loop,
include/ccc_b200.h): the control loop of the reference's own test (tests/src/TestDdpCentroidal.cpp:94-150) for a
batch of plants.  The contact / reference callbacks are sampled once on the plant's time grid
(time = t0 + entry * sim_dt); stage k of control cycle t reads entry t + k * stride, stride = horizon_dt / sim_dt.
"""
import numpy as np

from . import _abi
from ._abi import ptr
from .schedule import CentroidalSchedule


class LoopResultArrays:
    def __init__(self, batch, ticks, m_max):
        self.plant = np.zeros((batch, ticks + 1, 9))  # position, velocity, angular momentum at the start of each cycle
        self.u0 = np.zeros((batch, ticks, m_max))      # force scales applied in each cycle
        self.iters = np.zeros((batch, ticks), dtype=np.int32)

    def as_struct(self):
        r = _abi.DdpCentroidalLoopResult()
        r.plant, r.u0, r.iters = ptr(self.plant), ptr(self.u0), ptr(self.iters)
        return r


class CentroidalLoop:
    def __init__(self, horizon_steps, horizon_dt, sim_dt, ticks, mass, w_run, w_term, n_sched=1, u_lo=0.0, u_hi=1e6,
                 max_iter_later=1, m_max=_abi.CCC_DDP_M_MAX):
        stride = int(round(horizon_dt / sim_dt))
        if abs(stride * sim_dt - horizon_dt) > 1e-12 * horizon_dt or stride < 1:
            raise ValueError("horizon_dt must be an integer multiple of sim_dt")
        self.N, self.dt, self.sim_dt, self.ticks, self.mass, self.stride = horizon_steps, horizon_dt, sim_dt, ticks, mass, stride
        self.grid_len = ticks - 1 + horizon_steps * stride + 1
        self.sched = CentroidalSchedule(n_sched, self.grid_len, m_max)
        self.w_run, self.w_term = np.asarray(w_run, dtype=np.float64), np.asarray(w_term, dtype=np.float64)
        self.u_lo, self.u_hi, self.max_iter_later = float(u_lo), float(u_hi), int(max_iter_later)
        self.sched_id = np.zeros(0, dtype=np.int32)
        self.plant0 = np.zeros((0, 9))
        self.disturb_tick, self.disturb_vel = -1, np.zeros(3)

    def sample(self, s, motion_param_func, ref_data_func, t0=0.0):
        """Sample the callbacks of schedule s at t0 + entry * sim_dt for every grid entry."""
        self.sched.sample(s, motion_param_func, ref_data_func, t0, self.sim_dt)
        return self

    def set_plants(self, sched_id, pos, vel, angular_momentum):
        self.sched_id = np.ascontiguousarray(sched_id, dtype=np.int32)
        self.plant0 = np.ascontiguousarray(np.concatenate([pos, vel, angular_momentum], axis=1), dtype=np.float64)
        return self

    def set_disturbance(self, tick, vel_impulse):
        """Velocity impulse added after the plant step of control cycle `tick`."""
        self.disturb_tick, self.disturb_vel = int(tick), np.asarray(vel_impulse, dtype=np.float64)
        return self

    batch = property(lambda s: len(s.sched_id))

    def new_result(self):
        return LoopResultArrays(self.batch, self.ticks, self.sched.m_max)

    def as_struct(self):
        ls = _abi.DdpCentroidalLoop()
        ls.horizon_steps, ls.batch, ls.n_sched, ls.m_max = self.N, self.batch, self.sched.S, self.sched.m_max
        ls.dt, ls.mass, ls.sim_dt = self.dt, self.mass, self.sim_dt
        ls.ticks, ls.stride, ls.grid_len, ls.max_iter_later = self.ticks, self.stride, self.grid_len, self.max_iter_later
        self._ref = np.ascontiguousarray(self.sched.ref_pos[:, : self.grid_len])
        ls.sched_id, ls.m, ls.ridge, ls.vertex, ls.ref_pos = ptr(self.sched_id), ptr(self.sched.m), ptr(self.sched.ridge), ptr(
            self.sched.vertex), ptr(self._ref)
        for i in range(10):
            ls.w_run[i] = self.w_run[i]
        for i in range(9):
            ls.w_term[i] = self.w_term[i]
        ls.u_lo, ls.u_hi = self.u_lo, self.u_hi
        ls.plant0 = ptr(self.plant0)
        ls.disturb_tick = self.disturb_tick
        for i in range(3):
            ls.disturb_vel[i] = self.disturb_vel[i]
        return ls
