"""Batched host class for LinearMpcZ (reference src/LinearMpcZ.cpp, include/CCC/LinearMpcZ.h): linear MPC of the
vertical CoM motion for a predefined contact sequence; the decision variables are the vertical forces of the contact
stages.

The contact / reference schedule is shared by the batch (sampled on the horizon grid once per call, reference
src/LinearMpcZ.cpp:59-67); the initial states differ per problem.  Condensing for the output (CoM height) with
per-stage input dimension 1 (contact) or 0 (flight) and the QP matrices are built once per call on the host; the
batch of QPs (n = contact stages, no equality, the force range as 2n bound rows) goes to `qp_solve`.
"""
import numpy as np

from .linear_models import G, StateSpaceModel, VariantSequentialExtension
from .qp import QpProblemSet


def _phase_model(mass, contact):
    """ModelContactPhase / ModelNoncontactPhase (src/LinearMpcZ.cpp:10-30): state (m z, m vz), output z."""
    s = StateSpaceModel(2, 1 if contact else 0, 1)
    s.A[0, 1] = 1
    if contact:
        s.B[1, 0] = 1
    s.C[0, 0] = 1 / mass
    s.E[1] = -1 * mass * G
    return s


class LinearMpcZ:
    def __init__(self, mass, horizon_dt, horizon_steps, weight_pos=1.0, weight_force=1e-7):
        self.mass, self.horizon_dt, self.horizon_steps = mass, horizon_dt, horizon_steps
        self.weight_pos, self.weight_force = weight_pos, weight_force  # WeightParam (include/CCC/LinearMpcZ.h:33-45)
        self.force_range = (10.0, 10.0 * mass * G)  # :39
        self.model_contact = _phase_model(mass, True).calc_disc_matrix(horizon_dt)
        self.model_noncontact = _phase_model(mass, False).calc_disc_matrix(horizon_dt)

    def build_qp(self, contacts, ref_pos_seq, x0):
        """contacts [N] bool, ref_pos_seq [N], x0 [B][2] (pos, vel) -> QpProblemSet (procOnce, :73-93)."""
        models = [self.model_contact if c else self.model_noncontact for c in contacts]
        ext = VariantSequentialExtension(models, True)
        n = ext.total_input_dim
        Q = self.weight_pos * ext.B_seq.T @ ext.B_seq
        Q[np.diag_indices(n)] += self.weight_force
        x = self.mass * np.atleast_2d(np.asarray(x0, dtype=np.float64))
        resid = np.asarray(ref_pos_seq, dtype=np.float64)[None, :] - x @ ext.A_seq.T - ext.E_seq[None, :]
        c = -1 * self.weight_pos * resid @ ext.B_seq
        C = np.vstack([-np.eye(n), np.eye(n)])
        d = np.concatenate([np.full(n, -self.force_range[0]), np.full(n, self.force_range[1])])
        return QpProblemSet(Q, C, np.tile(d, (len(x), 1)), None, None, c)

    def plan_batch(self, qp_solve, contact_func, ref_pos_func, x0, current_time):
        """planOnce (:48-71) for a batch of initial states x0 [B][2] -> planned vertical force [B]."""
        x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
        if not contact_func(current_time):  # planned force is always zero if there is no contact
            return np.zeros(len(x0))
        ts = [current_time + i * self.horizon_dt for i in range(self.horizon_steps)]
        ps = self.build_qp([contact_func(t) for t in ts], [ref_pos_func(t) for t in ts], x0)
        res = qp_solve(ps)
        self.last_problem, self.last_result = ps, res
        return res.x[:, 0]
