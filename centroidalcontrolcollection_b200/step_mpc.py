"""Host class of CCC::StepMpc (reference include/CCC/StepMpc.h, src/StepMpc.cpp; Xin et al. 2019: step-to-step MPC).

One unknown per element of the reference data (a support phase with its ZMP and end time): the ZMP of each phase.  The CoM
moves by the closed-form step model (cosh / sinh of omega x phase duration, src/StepMpc.cpp:8-25), condensed over the phases
by the variant sequential extension with outputs (position, velocity, capture point); the weighted least-squares system is
assembled term by term as the reference does (:27-178) and solved (the reference: Eigen colPivHouseholderQr, :181).
`plan_once` is the reference's planOnce for one problem in numpy (the check of the batched kernel); `plan_batch` flattens P
reference-data records and B initial parameters into the C-ABI struct and runs ccc_step_mpc_plan through `run`
(engine.step_mpc_plan).
"""
import math

import numpy as np

from . import _abi
from ._abi import ptr
from .linear_models import G

MAX_ELEMENTS = 16  # ccc_b200.h CCC_STEP_MPC_MAX_ELEMENTS


class StepMpc:
    def __init__(self, com_height, free_zmp=1e-2, fixed_zmp=1e0, double_support=1e0, pos=0.0, vel=0.0, capture_point_abs=1e1,
                 capture_point_rel=1e1):
        """WeightParam defaults of include/CCC/StepMpc.h:110-117."""
        self.com_height = com_height
        self.w = dict(free_zmp=free_zmp, fixed_zmp=fixed_zmp, double_support=double_support, pos=pos, vel=vel,
                      capture_point_abs=capture_point_abs, capture_point_rel=capture_point_rel)

    def _step_model(self, step_duration):
        """StepMpc1d::StepModel (src/StepMpc.cpp:8-25): Ad (2x2), Bd (2), C (3x2)."""
        omega = math.sqrt(G / self.com_height)
        e = math.exp(omega * step_duration)
        ei = 1.0 / e
        Ad = np.array([[0.5 * (e + ei), 0.5 * (e - ei) / omega], [0.5 * omega * (e - ei), 0.5 * (e + ei)]])
        Bd = np.array([1.0 - 0.5 * (e + ei), 0.5 * omega * (ei - e)])
        C = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0 / omega]])
        return Ad, Bd, C

    def system_1d(self, single, zmp, end_time, x0, current_time):
        """eq_mat, eq_vec of StepMpc1d::planOnce (:27-178) for one axis; single / zmp / end_time: the element list."""
        n = len(zmp)
        # variant sequential extension with outputs (include/CCC/VariantSequentialExtension.h:110-208)
        A_seq = np.zeros((3 * n, 2))
        B_seq = np.zeros((3 * n, n))
        Ax = np.zeros((n, 2, 2))
        Bx = np.zeros((n, 2, n))
        for i in range(n):
            dur = end_time[i] - (current_time if i == 0 else end_time[i - 1])
            Ad, Bd, C = self._step_model(dur)
            Ax[i] = Ad if i == 0 else Ad @ Ax[i - 1]
            if i > 0:
                Bx[i] = Ad @ Bx[i - 1]
            Bx[i][:, i] = Bd
            A_seq[3 * i:3 * i + 3] = C @ Ax[i]
            B_seq[3 * i:3 * i + 3] = C @ Bx[i]
        w = self.w
        eq_mat = np.zeros((n, n))
        eq_vec = np.zeros(n)
        diag = np.full(n, w["free_zmp"])
        diag[0] = w["fixed_zmp"]
        if n > 1 and not single[0]:
            diag[1] = w["fixed_zmp"]
        eq_mat[np.arange(n), np.arange(n)] += diag
        eq_vec += -1 * diag * zmp
        ds = w["double_support"] * np.array([[1.0, -2.0, 1.0], [-2.0, 4.0, -2.0], [1.0, -2.0, 1.0]])
        for i in range(1, n - 1):
            if single[i - 1] and not single[i] and single[i + 1]:
                eq_mat[i - 1:i + 2, i - 1:i + 2] += ds
        x0 = np.asarray(x0, dtype=np.float64)
        if w["pos"] > 0.0:
            ref = np.array([zmp[min(i + 1, n - 1)] for i in range(n)])
            S = np.zeros((n, 3 * n))
            S[np.arange(n), 3 * np.arange(n)] = 1.0
            sub = w["pos"] * B_seq.T @ S.T
            eq_mat += sub @ S @ B_seq
            eq_vec += sub @ (S @ A_seq @ x0 - ref)
        if w["vel"] > 0.0:
            S = np.zeros((n, 3 * n))
            S[np.arange(n), 3 * np.arange(n) + 1] = 1.0
            sub = w["vel"] * B_seq.T @ S.T @ S
            eq_mat += sub @ B_seq
            eq_vec += sub @ A_seq @ x0
        n_future_single = sum(1 for i in range(1, n) if single[i])
        if w["capture_point_abs"] > 0.0:
            ref = np.zeros(n)
            S = np.zeros((n, 3 * n))
            for i in range(n):
                if n_future_single == 0:
                    if i >= 1 or not single[i]:
                        ref[i] = zmp[i]
                        S[i, 3 * i + 2] = 1.0
                elif i >= 1 and single[i]:
                    ref[i] = zmp[i]
                    S[i, 3 * (i - 1) + 2] = 1.0
            sub = w["capture_point_abs"] * B_seq.T @ S.T
            eq_mat += sub @ S @ B_seq
            eq_vec += sub @ (S @ A_seq @ x0 - ref)
        if w["capture_point_rel"] > 0.0 and n_future_single >= 1:
            S = np.zeros((n, 3 * n))
            Z = np.zeros((n, n))
            for i in range(1, n):
                if single[i]:
                    S[i, 3 * (i - 1) + 2] = 1.0
                    Z[i, i] = 1.0
            sub = w["capture_point_rel"] * (S @ B_seq - Z).T
            eq_mat += sub @ (S @ B_seq - Z)
            eq_vec += sub @ S @ A_seq @ x0
        return eq_mat, eq_vec

    def plan_once(self, elements, pos, vel, current_time):
        """2-D planOnce (:195-247): elements = [(is_single_support, zmp[2], end_time), ...] -> (current_zmp[2],
        next_foot_zmp[2] or None)."""
        single = [bool(e[0]) for e in elements]
        end_time = [float(e[2]) for e in elements]
        cur, nxt = np.zeros(2), None
        for a in range(2):
            zmp = np.array([float(e[1][a]) for e in elements])
            M, v = self.system_1d(single, zmp, end_time, [pos[a], vel[a]], current_time)
            sol = np.linalg.solve(M, -1 * v)
            cur[a] = sol[0]
            for i in range(1, len(elements)):
                if single[i]:
                    nxt = np.zeros(2) if nxt is None else nxt
                    nxt[a] = sol[i]
                    break
        return cur, nxt

    def plan_batch(self, run, ref_data, current_times, pos, vel, plan_id):
        """ref_data: P element lists; current_times [P]; pos, vel [B][2]; plan_id [B] -> (current_zmp [B][2],
        next_foot_zmp [B][2], has_next [B])."""
        P, B = len(ref_data), len(plan_id)
        K = MAX_ELEMENTS
        keep = dict(plan_id=np.ascontiguousarray(plan_id, dtype=np.int32), pos=np.ascontiguousarray(pos, dtype=np.float64).reshape(B, 2),
                    vel=np.ascontiguousarray(vel, dtype=np.float64).reshape(B, 2), current_time=np.ascontiguousarray(current_times, dtype=np.float64),
                    n_elements=np.zeros(P, dtype=np.int32), single=np.zeros((P, K), dtype=np.int32), zmp=np.zeros((P, K, 2)),
                    end_time=np.zeros((P, K)))
        for p, elements in enumerate(ref_data):
            if not 1 <= len(elements) <= K:
                raise ValueError(f"StepMpc: 1..{K} elements per reference")
            keep["n_elements"][p] = len(elements)
            for i, (s, z, te) in enumerate(elements):
                keep["single"][p, i], keep["zmp"][p, i], keep["end_time"][p, i] = int(bool(s)), z, te
        bt = _abi.StepMpcBatch()
        bt.batch, bt.n_plans, bt.max_elements = B, P, K
        bt.com_height = self.com_height
        for k in ("free_zmp", "fixed_zmp", "double_support", "pos", "vel", "capture_point_abs", "capture_point_rel"):
            setattr(bt, "w_" + k, self.w[k])
        for k, v in keep.items():
            setattr(bt, "x_pos" if k == "pos" else "x_vel" if k == "vel" else k, ptr(v))
        return run(bt, B)
