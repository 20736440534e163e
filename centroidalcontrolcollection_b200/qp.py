"""Host-side containers for the batched dense QP of include/ccc_b200.h (ccc_qp_batch_t / ccc_qp_result_t)."""
import numpy as np

from . import _abi
from ._abi import ptr


class QpResultArrays:
    def __init__(self, batch, n):
        self.x = np.zeros((batch, n))
        self.iters = np.zeros(batch, dtype=np.int32)
        self.status = np.zeros(batch, dtype=np.int32)
        self.n_active = np.zeros(batch, dtype=np.int32)
        self.active = np.full((batch, n), -1, dtype=np.int32)

    def as_struct(self):
        r = _abi.QpResult()
        r.x, r.iters, r.status, r.n_active, r.active = ptr(self.x), ptr(self.iters), ptr(self.status), ptr(self.n_active), ptr(self.active)
        return r

    def active_sets(self):
        """Final active sets as sorted tuples (order of activation is solver specific)."""
        return [tuple(sorted(int(v) for v in row[:k])) for row, k in zip(self.active, self.n_active)]


class QpProblemSet:
    """min 0.5 x'Qx + c'x  s.t.  A x = b, C x <= d  with Q, A, C shared by the batch."""

    def __init__(self, Q, C, d, A=None, b=None, c=None):
        self.Q = np.ascontiguousarray(Q, dtype=np.float64)
        self.C = np.ascontiguousarray(C, dtype=np.float64)
        self.d = np.ascontiguousarray(d, dtype=np.float64)
        self.n = self.Q.shape[0]
        self.n_ineq = self.C.shape[0]
        self.batch = self.d.shape[0]
        self.A = None if A is None else np.ascontiguousarray(A, dtype=np.float64).reshape(-1, self.n)
        self.n_eq = 0 if self.A is None else self.A.shape[0]
        self.b = None if b is None else np.ascontiguousarray(b, dtype=np.float64).reshape(self.batch, self.n_eq)
        self.c = None if c is None else np.ascontiguousarray(c, dtype=np.float64).reshape(self.batch, self.n)
        assert self.Q.shape == (self.n, self.n) and self.C.shape == (self.n_ineq, self.n) and self.d.shape == (self.batch, self.n_ineq)

    def subset(self, idx):
        idx = np.asarray(idx)
        return QpProblemSet(self.Q, self.C, self.d[idx], self.A, None if self.b is None else self.b[idx],
                            None if self.c is None else self.c[idx])

    def as_struct(self):
        s = _abi.QpBatch()
        s.n, s.n_eq, s.n_ineq, s.batch = self.n, self.n_eq, self.n_ineq, self.batch
        s.Q, s.A, s.C = ptr(self.Q), ptr(self.A), ptr(self.C)
        s.c, s.b, s.d = ptr(self.c), ptr(self.b), ptr(self.d)
        return s

    def new_result(self):
        return QpResultArrays(self.batch, self.n)

    def kkt_residuals(self, x, tol=1e-9):
        """(primal infeasibility, dual residual via non-negative least squares on the active rows)."""
        out = []
        for k in range(self.batch):
            xk = x[k]
            g = self.Q @ xk + (0 if self.c is None else self.c[k])
            viol = max(0.0, float((self.C @ xk - self.d[k]).max()))
            eq = 0.0 if self.A is None else float(np.abs(self.A @ xk - self.b[k]).max())
            act = np.where(self.C @ xk - self.d[k] > -tol)[0]
            rows = [self.C[act]] + ([self.A, -self.A] if self.A is not None else [])
            M = np.vstack(rows).T if sum(len(r) for r in rows) else np.zeros((self.n, 0))
            from scipy.optimize import nnls

            mu, rn = nnls(M, -g) if M.shape[1] else (np.zeros(0), float(np.linalg.norm(g)))
            out.append((max(viol, eq), rn))
        return out


class QpGroupedProblemSet:
    """Batch of QPs whose Q (and A) come in G groups, C shared: problem b uses group group_id[b]
    (ccc_qp_solve_grouped)."""

    def __init__(self, Q, C, d, group_id, A=None, b=None, c=None):
        self.Q = np.ascontiguousarray(Q, dtype=np.float64)
        self.G, self.n = self.Q.shape[0], self.Q.shape[1]
        self.C = np.ascontiguousarray(C, dtype=np.float64)
        self.d = np.ascontiguousarray(d, dtype=np.float64)
        self.n_ineq, self.batch = self.C.shape[0], self.d.shape[0]
        self.group_id = np.ascontiguousarray(group_id, dtype=np.int32)
        self.A = None if A is None else np.ascontiguousarray(A, dtype=np.float64).reshape(self.G, -1, self.n)
        self.n_eq = 0 if self.A is None else self.A.shape[1]
        self.b = None if b is None else np.ascontiguousarray(b, dtype=np.float64).reshape(self.batch, self.n_eq)
        self.c = None if c is None else np.ascontiguousarray(c, dtype=np.float64).reshape(self.batch, self.n)
        assert self.Q.shape == (self.G, self.n, self.n) and self.d.shape == (self.batch, self.n_ineq) and len(self.group_id) == self.batch

    def as_struct(self):
        s = _abi.QpBatch()
        s.n, s.n_eq, s.n_ineq, s.batch = self.n, self.n_eq, self.n_ineq, self.batch
        s.Q, s.A, s.C = ptr(self.Q), ptr(self.A), ptr(self.C)
        s.c, s.b, s.d = ptr(self.c), ptr(self.b), ptr(self.d)
        return s

    def new_result(self):
        return QpResultArrays(self.batch, self.n)

    def group(self, g):
        """(indices, QpProblemSet) of group g: what a one-group solver (the oracle) takes."""
        idx = np.where(self.group_id == g)[0]
        return idx, QpProblemSet(self.Q[g], self.C, self.d[idx], None if self.A is None else self.A[g],
                                 None if self.b is None else self.b[idx], None if self.c is None else self.c[idx])

    def solve_by_group(self, qp_solve):
        """Solve with a one-group solver, group by group (the oracle path of the tests)."""
        res = self.new_result()
        for g in range(self.G):
            idx, ps = self.group(g)
            if len(idx) == 0:
                continue
            r = qp_solve(ps)
            for f in ("x", "iters", "status", "n_active", "active"):
                getattr(res, f)[idx] = getattr(r, f)
        return res
