"""Batched host class for PreviewControlCentroidal (reference include/CCC/PreviewControlCentroidal.h,
src/PreviewControlCentroidal.cpp; Murooka et al. 2022): six one-dimensional preview controllers (three rotational,
three translational components; state (pos, vel, acc), input jerk, outputs (pos, inertia * acc)) followed by the
projection of the planned wrench onto the contacts' friction pyramids.

Setup (once per controller): six DARE / gain computations on the host (preview_control.PreviewControl with the
two-output CentroidalModel1d).  Online, for B initial states sharing the sampled reference (planOnce :91-130):
  * jerk = -K x + F ref_seq per component and problem: 6 B rows through `gemv` (ccc_preview_input on the GPU; the
    reference sequence interleaves position and wrench references, 2 N entries per row, :103-110);
  * control-step model: wrench = inertia * (acc + control_dt jerk) (:37-43), + m g on force z (:125);
  * wrench distribution (ForceColl::WrenchDistribution::run, an external dependency of the reference, restated
    from its published formulation — parity unpinned): min ||G lambda - w_des||^2_W + eps ||lambda||^2,
    lambda_min <= lambda <= lambda_max, G = grasp matrix of the ridges about the problem's CoM position.  G depends
    on the problem, so this is a QP batch with one matrix group per problem (`qp_solve_grouped`:
    engine.QpEngine.solve_grouped on the GPU, qp.QpGroupedProblemSet.solve_by_group with the oracle on the CPU).
Vectors follow SpaceVecAlg: MotionVecd / ForceVecd .vector() = (angular / moment, linear / force).
"""
import numpy as np

from .linear_models import G as GRAVITY
from .linear_models import StateSpaceModel
from .preview_control import PreviewControl
from .qp import QpGroupedProblemSet


def _centroidal_model_1d(inertia_param):
    """CentroidalModel1d (src/PreviewControlCentroidal.cpp:10-21)."""
    s = StateSpaceModel(3, 1, 2)
    s.A[0, 1] = 1
    s.A[1, 2] = 1
    s.B[2, 0] = 1
    s.C[0, 0] = 1
    s.C[1, 2] = inertia_param
    return s


def _numpy_gemv(K, F, x, ref_seq):
    return -(x @ K[0]) + ref_seq @ F[0]


class WrenchDistribution:
    """ForceColl::WrenchDistribution (configuration defaults as recalled: wrenchWeight 1, regularWeight 1e-8,
    ridgeForceMinMax (3, 1000)); `run` for a batch of desired wrenches and moment origins."""

    def __init__(self, vertex, ridge, wrench_weight=(1.0,) * 6, regular_weight=1e-8, ridge_force_min_max=(3.0, 1000.0)):
        self.vertex = np.asarray(vertex, dtype=np.float64).reshape(-1, 3)
        self.ridge = np.asarray(ridge, dtype=np.float64).reshape(-1, 3)
        self.w = np.asarray(wrench_weight, dtype=np.float64)  # (moment, force)
        self.reg, self.lim = regular_weight, ridge_force_min_max

    def build_qp(self, desired, origin):
        """desired [B][6] (moment, force), origin [B][3] -> (QpGroupedProblemSet with one group per problem, grasp [B][6][n])."""
        B, n = len(desired), len(self.ridge)
        grasp = np.zeros((B, 6, n))
        grasp[:, 0:3, :] = np.cross(self.vertex[None, :, :] - origin[:, None, :], self.ridge[None, :, :]).transpose(0, 2, 1)
        grasp[:, 3:6, :] = self.ridge.T[None, :, :]
        GtW = grasp.transpose(0, 2, 1) * self.w[None, None, :]
        Q = GtW @ grasp
        Q[:, np.arange(n), np.arange(n)] += self.reg
        c = -1 * np.einsum("bnk,bk->bn", GtW, desired)
        C = np.vstack([-np.eye(n), np.eye(n)])
        d = np.tile(np.concatenate([np.full(n, -self.lim[0]), np.full(n, self.lim[1])]), (B, 1))
        return QpGroupedProblemSet(Q, C, d, np.arange(B, dtype=np.int32), None, None, c), grasp

    def run(self, qp_solve_grouped, desired, origin):
        """-> (result wrench [B][6], ridge force scales [B][n])."""
        if len(self.ridge) == 0:
            return np.zeros_like(desired), np.zeros((len(desired), 0))
        gp, grasp = self.build_qp(np.asarray(desired, dtype=np.float64), np.asarray(origin, dtype=np.float64))
        res = qp_solve_grouped(gp)
        self.last_problem, self.last_result = gp, res
        return np.einsum("bkn,bn->bk", grasp, res.x), res.x


class PreviewControlCentroidal:
    def __init__(self, mass, moment_of_inertia, horizon_duration, horizon_dt, weight_pos=None, weight_wrench=None, weight_jerk=None):
        self.mass = mass
        wp = np.asarray(weight_pos if weight_pos is not None else [1e2] * 3 + [2e2] * 3, dtype=np.float64)           # :204-209
        ww = np.asarray(weight_wrench if weight_wrench is not None else [5e-3] * 3 + [5e-4] * 3, dtype=np.float64)
        wj = np.asarray(weight_jerk if weight_jerk is not None else [1e-8] * 6, dtype=np.float64)
        self.inertia = np.array(list(moment_of_inertia) + [mass] * 3, dtype=np.float64)
        self.pc_1d = [PreviewControl(_centroidal_model_1d(self.inertia[i]), horizon_duration, horizon_dt, [wp[i], ww[i]], [wj[i]])
                      for i in range(6)]
        self.horizon_steps, self.horizon_dt = self.pc_1d[0].horizon_steps, horizon_dt

    def planned_wrench(self, ref_data_func, pos, vel, acc, current_time, control_dt=-1.0, gemv=_numpy_gemv):
        """The six preview controllers (planOnce :97-125): pos / vel / acc [B][6] (angular, linear) -> wrench [B][6]
        (moment, force) before the projection.  ref_data_func(t) -> (pos[6], wrench[6])."""
        pos, vel, acc = (np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in (pos, vel, acc))
        B, N = len(pos), self.horizon_steps
        ref = np.zeros((6, 2 * N))
        for i in range(N):
            p, w = ref_data_func(current_time + (i + 1) * self.horizon_dt)
            ref[:, 2 * i], ref[:, 2 * i + 1] = p, w
        if control_dt < 0:
            control_dt = self.horizon_dt
        out = np.zeros((B, 6))
        for i in range(6):
            x = np.stack([pos[:, i], vel[:, i], acc[:, i]], axis=1)
            jerk = gemv(self.pc_1d[i].K, self.pc_1d[i].F, x, np.tile(ref[i], (B, 1)))
            a = acc[:, i] + control_dt * jerk
            out[:, i] = self.pc_1d[i].model.C[1, 2] * a
        out[:, 5] += self.mass * GRAVITY
        return out

    def plan_batch(self, qp_solve_grouped, vertex, ridge, ref_data_func, pos, vel, acc, current_time, control_dt=-1.0, gemv=_numpy_gemv):
        """planOnce (:91-130) for B initial states; (vertex, ridge) = flattened contact list at current_time."""
        desired = self.planned_wrench(ref_data_func, pos, vel, acc, current_time, control_dt, gemv)
        self.wrench_dist = WrenchDistribution(vertex, ridge)
        wrench, _ = self.wrench_dist.run(qp_solve_grouped, desired, np.atleast_2d(pos)[:, 3:6])
        return wrench
