"""Batched host class for LinearMpcXY (reference src/LinearMpcXY.cpp, include/CCC/LinearMpcXY.h):
linear MPC of the horizontal linear / angular momentum with the ridge force scales of every stage as
decision variables, for a predefined vertical motion and contact sequence.

The contact / reference schedule is shared by the batch (it is sampled on the horizon grid once per
call, like planOnce does, reference :96-115); the initial states differ per problem.  Condensing
(VariantSequentialExtension) and the QP matrices are built once per call on the host; the batch of QPs
(n = sum of ridge counts, one equality per contact stage, box bounds as 2n inequality rows) goes to
`qp_solve`: engine.QpEngine in production, the oracle in the CPU-tier tests.
"""
import numpy as np

from . import _abi
from ._abi import ptr
from .linear_models import G, StateSpaceModel, VariantSequentialExtension
from .qp import QpProblemSet


class MotionParam:
    """reference include/CCC/LinearMpcXY.h:38-51; the contact list is flattened to (vertex, ridge) tables
    [m][3] (contact.py)."""

    def __init__(self, com_z, total_force_z, vertex, ridge):
        self.com_z, self.total_force_z = float(com_z), float(total_force_z)
        self.vertex = np.asarray(vertex, dtype=np.float64).reshape(-1, 3)
        self.ridge = np.asarray(ridge, dtype=np.float64).reshape(-1, 3)


class WeightParam:
    """reference include/CCC/LinearMpcXY.h:104-143"""

    def __init__(self, linear_momentum_integral=(1.0, 1.0), linear_momentum=(0.0, 0.0), angular_momentum=(1.0, 1.0),
                 force=1e-5):
        self.linear_momentum_integral = np.asarray(linear_momentum_integral, dtype=np.float64)
        self.linear_momentum = np.asarray(linear_momentum, dtype=np.float64)
        self.angular_momentum = np.asarray(angular_momentum, dtype=np.float64)
        self.force = float(force)

    def output_weight(self, seq_len):  # src/LinearMpcXY.cpp:45-57
        one = np.array([self.linear_momentum_integral[0], self.linear_momentum[0], self.linear_momentum_integral[1],
                        self.linear_momentum[1], self.angular_momentum[0], self.angular_momentum[1]])
        return np.tile(one, seq_len)


class Model(StateSpaceModel):
    """state = (m c_x, m v_x, m c_y, m v_y, L_x, L_y), input = ridge force scales (src/LinearMpcXY.cpp:59-83)."""

    def __init__(self, mass, motion_param):
        m = len(motion_param.ridge)
        super().__init__(6, m, 0)
        self.motion_param = motion_param
        self.A[0, 1] = 1
        self.A[2, 3] = 1
        self.A[4, 2] = -1 * motion_param.total_force_z / mass
        self.A[5, 0] = motion_param.total_force_z / mass
        v, r, cz = motion_param.vertex, motion_param.ridge, motion_param.com_z
        for i in range(m):
            self.B[:, i] = [0, r[i, 0], 0, r[i, 1], -1 * (v[i, 2] - cz) * r[i, 1] + v[i, 1] * r[i, 2],
                            (v[i, 2] - cz) * r[i, 0] + -1 * v[i, 0] * r[i, 2]]


def to_state(mass, pos, vel, angular_momentum):
    """InitialParam::toState / RefData::toOutput (src/LinearMpcXY.cpp:26-38); inputs [..., 2] each."""
    pos, vel, am = (np.asarray(a, dtype=np.float64) for a in (pos, vel, angular_momentum))
    return np.stack([mass * pos[..., 0], mass * vel[..., 0], mass * pos[..., 1], mass * vel[..., 1], am[..., 0], am[..., 1]],
                    axis=-1)


class LinearMpcXY:
    def __init__(self, mass, horizon_dt, horizon_steps, weight_param=None):
        self.mass, self.horizon_dt, self.horizon_steps = mass, horizon_dt, horizon_steps
        self.weight_param = weight_param or WeightParam()
        self.force_range = (3.0, 3.0 * mass * G)  # :91

    def build_qp(self, motion_params, ref_output_seq, x0):
        """motion_params: one MotionParam per stage; ref_output_seq [6 N]; x0 [B][6] -> QpProblemSet
        (procOnce, src/LinearMpcXY.cpp:117-178)."""
        models = [Model(self.mass, mp).calc_disc_matrix(self.horizon_dt) for mp in motion_params]
        ext = VariantSequentialExtension(models, False)
        n = ext.total_input_dim
        w_out = self.weight_param.output_weight(len(models))
        BtW = ext.B_seq.T * w_out[None, :]
        Q = BtW @ ext.B_seq
        Q[np.diag_indices(n)] += self.weight_param.force
        x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
        resid = np.asarray(ref_output_seq, dtype=np.float64)[None, :] - x0 @ ext.A_seq.T - ext.E_seq[None, :]  # [B][6N]
        c = -1 * resid @ BtW.T
        stages = [mdl for mdl in models if mdl.input_dim > 0]  # no total-force constraint on flight stages (:124-131)
        A = np.zeros((len(stages), n))
        b = np.zeros(len(stages))
        acc = 0
        for e, mdl in enumerate(stages):
            m = mdl.input_dim
            A[e, acc:acc + m] = mdl.motion_param.ridge[:, 2]
            b[e] = mdl.motion_param.total_force_z
            acc += m
        # x_min <= x <= x_max (:176-177) as inequality rows: -x <= -x_min, x <= x_max
        C = np.vstack([-np.eye(n), np.eye(n)])
        d = np.concatenate([np.full(n, -self.force_range[0]), np.full(n, self.force_range[1])])
        B = len(x0)
        self.first_input_dim = models[0].input_dim
        return QpProblemSet(Q, C, np.tile(d, (B, 1)), A if len(stages) else None,
                            np.tile(b, (B, 1)) if len(stages) else None, c)

    def plan_batch(self, qp_solve, motion_param_func, ref_data_func, x0, current_time):
        """planOnce (src/LinearMpcXY.cpp:96-115) for a batch of initial states x0 [B][6] (to_state) ->
        force scales of the first stage [B][m_0].  ref_data_func(t) -> (pos, vel, angular_momentum)."""
        ts = [current_time + i * self.horizon_dt for i in range(self.horizon_steps)]
        motion_params = [motion_param_func(t) for t in ts]
        ref = np.concatenate([to_state(self.mass, *ref_data_func(t)) for t in ts])
        ps = self.build_qp(motion_params, ref, x0)
        res = qp_solve(ps)
        self.last_problem, self.last_result = ps, res
        return res.x[:, :self.first_input_dim]


class XySweepResultArrays:
    """ccc_linear_mpc_xy_result_t with every optional output allocated."""

    def __init__(self, batch, n_sched, horizon_steps, n, intermediates=True):
        self.u = np.zeros((batch, n))
        self.iters = np.zeros(batch, dtype=np.int32)
        self.status = np.zeros(batch, dtype=np.int32)
        self.n_active = np.zeros(batch, dtype=np.int32)
        self.active = np.full((batch, n), -1, dtype=np.int32)
        rows = 6 * horizon_steps
        self.A_seq = np.zeros((n_sched, rows, 6)) if intermediates else None
        self.B_seq = np.zeros((n_sched, rows, n)) if intermediates else None
        self.obj_mat = np.zeros((n_sched, n, n)) if intermediates else None
        self.obj_vec = np.zeros((batch, n)) if intermediates else None

    def as_struct(self):
        r = _abi.LinearMpcXyResult()
        r.u, r.iters, r.status, r.n_active, r.active = (ptr(self.u), ptr(self.iters), ptr(self.status), ptr(self.n_active),
                                                        ptr(self.active))
        r.A_seq, r.B_seq, r.obj_mat, r.obj_vec = ptr(self.A_seq), ptr(self.B_seq), ptr(self.obj_mat), ptr(self.obj_vec)
        return r

    def active_sets(self):
        return [tuple(sorted(int(v) for v in row[:k])) for row, k in zip(self.active, self.n_active)]


class XySweepProblemSet:
    """Flat form of LinearMpcXY::planOnce for B initial states over S sampled schedules
    (ccc_linear_mpc_xy_batch_t): everything after the callback sampling happens behind the C-ABI."""

    def __init__(self, mpc, horizon_steps, n_sched, m_max=32):
        self.mpc, self.N, self.S, self.m_max = mpc, int(horizon_steps), int(n_sched), int(m_max)
        N, S = self.N, self.S
        self.m = np.zeros((S, N), dtype=np.int32)
        self.ridge = np.zeros((S, N, m_max, 3))
        self.vertex = np.zeros((S, N, m_max, 3))
        self.com_z = np.zeros((S, N))
        self.total_force_z = np.zeros((S, N))
        self.ref_output = np.zeros((S, N, 6))
        self.sched_id = np.zeros(0, dtype=np.int32)
        self.x0 = np.zeros((0, 6))

    def sample(self, s, motion_param_func, ref_data_func, current_time):
        """What planOnce reads through its callbacks (src/LinearMpcXY.cpp:104-110) into schedule s."""
        for i in range(self.N):
            t = current_time + i * self.mpc.horizon_dt
            mp = motion_param_func(t)
            k = len(mp.ridge)
            assert k <= self.m_max
            self.m[s, i] = k
            self.ridge[s, i, :k] = mp.ridge
            self.vertex[s, i, :k] = mp.vertex
            self.com_z[s, i], self.total_force_z[s, i] = mp.com_z, mp.total_force_z
            self.ref_output[s, i] = to_state(self.mpc.mass, *ref_data_func(t))

    def set_initial_states(self, x0, sched_id):
        self.x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, 6)
        self.sched_id = np.ascontiguousarray(sched_id, dtype=np.int32)
        assert len(self.x0) == len(self.sched_id)

    @property
    def n(self):
        return int(self.m[0].sum())

    @property
    def n_eq(self):
        return int((self.m[0] > 0).sum())

    @property
    def batch(self):
        return len(self.x0)

    def as_struct(self):
        b = _abi.LinearMpcXyBatch()
        b.horizon_steps, b.batch, b.n_sched, b.m_max = self.N, self.batch, self.S, self.m_max
        b.dt, b.mass = self.mpc.horizon_dt, self.mpc.mass
        b.sched_id, b.m, b.ridge, b.vertex = ptr(self.sched_id), ptr(self.m), ptr(self.ridge), ptr(self.vertex)
        b.com_z, b.total_force_z, b.ref_output, b.x0 = ptr(self.com_z), ptr(self.total_force_z), ptr(self.ref_output), ptr(self.x0)
        w = self.mpc.weight_param.output_weight(1)
        for i in range(6):
            b.w_output[i] = w[i]
        b.w_force = self.mpc.weight_param.force
        b.force_lo, b.force_hi = self.mpc.force_range
        return b

    def new_result(self, intermediates=True):
        return XySweepResultArrays(self.batch, self.S, self.N, self.n, intermediates)

    def first_schedules(self, k):
        """The sub-sweep of schedules 0..k-1 and the initial states that use them."""
        sub = XySweepProblemSet(self.mpc, self.N, k, self.m_max)
        for f in ("m", "ridge", "vertex", "com_z", "total_force_z", "ref_output"):
            setattr(sub, f, np.ascontiguousarray(getattr(self, f)[:k]))
        keep = self.sched_id < k
        sub.set_initial_states(self.x0[keep], self.sched_id[keep])
        sub.index = np.where(keep)[0]
        return sub

    def host_problem(self, s, idx):
        """The QpProblemSet the host path (build_qp: scipy expm, numpy condensing) makes of schedule s and the
        initial states idx — the independent check of the closed-form discretisation."""
        mps = [MotionParam(self.com_z[s, i], self.total_force_z[s, i], self.vertex[s, i, :self.m[s, i]],
                           self.ridge[s, i, :self.m[s, i]]) for i in range(self.N)]
        return self.mpc.build_qp(mps, self.ref_output[s].reshape(-1), self.x0[idx])
