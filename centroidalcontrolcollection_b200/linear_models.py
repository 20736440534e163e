"""Formulation helpers of the linear MPC methods (host side, setup time; numpy).

Restates reference include/CCC/StateSpaceModel.h (zero-order-hold discretisation, :164-216),
include/CCC/InvariantSequentialExtension.h (condensing, :103-181), src/CommonModels.cpp:8-17
(ComZmpModelJerkInput).  These run once per controller object, not per solve: the hot path is the QP.
"""
import numpy as np
from scipy.linalg import expm

G = 9.80665  # reference include/CCC/Constants.h:10


class StateSpaceModel:
    """x' = A x + B u + E,  y = C x + D u + F  (reference include/CCC/StateSpaceModel.h)."""

    def __init__(self, state_dim, input_dim, output_dim):
        if state_dim < 0 or input_dim < 0 or output_dim < 0:
            raise ValueError("dimensions must be non-negative")  # :46-83
        self.A = np.zeros((state_dim, state_dim))
        self.B = np.zeros((state_dim, input_dim))
        self.C = np.zeros((output_dim, state_dim))
        self.D = np.zeros((output_dim, input_dim))
        self.E = np.zeros(state_dim)
        self.F = np.zeros(output_dim)
        self.dt = 0.0
        self.Ad = self.Bd = self.Ed = None

    state_dim = property(lambda s: s.A.shape[0])
    input_dim = property(lambda s: s.B.shape[1])
    output_dim = property(lambda s: s.C.shape[0])

    def calc_disc_matrix(self, dt):
        """ZOH via exp([A B (E); 0] dt) (:164-216): the E column is appended only when E != 0."""
        n, m = self.state_dim, self.input_dim
        self.dt = dt
        if np.linalg.norm(self.E) == 0:
            M = np.zeros((n + m, n + m))
            M[:n, :n], M[:n, n:] = dt * self.A, dt * self.B
            X = expm(M)
            self.Ad, self.Bd, self.Ed = X[:n, :n], X[:n, n:n + m], np.zeros(n)
        else:
            M = np.zeros((n + m + 1, n + m + 1))
            M[:n, :n], M[:n, n:n + m], M[:n, n + m] = dt * self.A, dt * self.B, dt * self.E
            X = expm(M)
            self.Ad, self.Bd, self.Ed = X[:n, :n], X[:n, n:n + m], X[:n, n + m]
        return self

    def state_eq(self, x, u):
        return self.A @ x + self.B @ u + self.E

    def state_eq_disc(self, x, u):
        return self.Ad @ x + self.Bd @ u + self.Ed

    def observ_eq(self, x, u):
        return self.C @ x + self.D @ u + self.F


class ComZmpModelJerkInput(StateSpaceModel):
    """CoM-ZMP model with jerk input, ZMP output (reference src/CommonModels.cpp:8-17)."""

    def __init__(self, com_height):
        super().__init__(3, 1, 1)
        self.A[0, 1] = 1
        self.A[1, 2] = 1
        self.B[2, 0] = 1
        self.C[0, 0] = 1
        self.C[0, 2] = -1 * com_height / G


class InvariantSequentialExtension:
    """x_seq = A_seq x0 + B_seq u_seq + E_seq for an LTI model (reference
    include/CCC/InvariantSequentialExtension.h:103-181), optionally projected to outputs."""

    def __init__(self, model, seq_len, extend_for_output=False):
        n, m = model.state_dim, model.input_dim
        A_seq = np.zeros((seq_len * n, n))
        B_seq = np.zeros((seq_len * n, seq_len * m))
        E_seq = np.zeros(seq_len * n)
        for i in range(seq_len):
            A_seq[i * n:(i + 1) * n] = model.Ad if i == 0 else model.Ad @ A_seq[(i - 1) * n:i * n]
            for j in range(seq_len - i):
                if j == 0:
                    B_seq[i * n:(i + 1) * n, 0:m] = model.Bd if i == 0 else model.Ad @ B_seq[(i - 1) * n:i * n, 0:m]
                else:
                    B_seq[(i + j) * n:(i + j + 1) * n, j * m:(j + 1) * m] = B_seq[i * n:(i + 1) * n, 0:m]
            E_seq[i * n:(i + 1) * n] = model.Ed if i == 0 else model.Ad @ E_seq[(i - 1) * n:i * n] + model.Ed
        if extend_for_output:
            p = model.output_dim
            C_seq = np.zeros((seq_len * p, seq_len * n))
            for i in range(seq_len):
                C_seq[i * p:(i + 1) * p, i * n:(i + 1) * n] = model.C
            A_seq, B_seq, E_seq = C_seq @ A_seq, C_seq @ B_seq, C_seq @ E_seq
        self.A_seq, self.B_seq, self.E_seq, self.seq_len = A_seq, B_seq, E_seq, seq_len


class VariantSequentialExtension:
    """x_seq = A_seq x0 + B_seq u_seq + E_seq for a time-variant model list whose input dimension may
    change (and be zero) per stage (reference include/CCC/VariantSequentialExtension.h:110-208)."""

    def __init__(self, model_list, extend_for_output=False):
        n = model_list[0].state_dim
        L = len(model_list)
        self.model_list = model_list
        self.total_state_dim = L * n
        self.total_input_dim = sum(m.input_dim for m in model_list)
        self.total_output_dim = sum(m.output_dim for m in model_list)
        A_seq = np.zeros((L * n, n))
        B_seq = np.zeros((L * n, self.total_input_dim))
        E_seq = np.zeros(L * n)
        acc = 0
        for i, mi in enumerate(model_list):
            m = mi.input_dim
            A_seq[i * n:(i + 1) * n] = mi.Ad if i == 0 else mi.Ad @ A_seq[(i - 1) * n:i * n]
            for j in range(i, L):
                if j == i:
                    B_seq[j * n:(j + 1) * n, acc:acc + m] = mi.Bd
                else:
                    B_seq[j * n:(j + 1) * n, acc:acc + m] = model_list[j].Ad @ B_seq[(j - 1) * n:j * n, acc:acc + m]
            E_seq[i * n:(i + 1) * n] = mi.Ed if i == 0 else mi.Ad @ E_seq[(i - 1) * n:i * n] + mi.Ed
            acc += m
        if extend_for_output:
            C_seq = np.zeros((self.total_output_dim, L * n))
            acc_o = 0
            for i, mi in enumerate(model_list):
                p = mi.output_dim
                C_seq[acc_o:acc_o + p, i * n:(i + 1) * n] = mi.C
                acc_o += p
            # the reference's extension for outputs drops F (and requires D = 0), :188-206
            A_seq, B_seq, E_seq = C_seq @ A_seq, C_seq @ B_seq, C_seq @ E_seq
        self.A_seq, self.B_seq, self.E_seq = A_seq, B_seq, E_seq
