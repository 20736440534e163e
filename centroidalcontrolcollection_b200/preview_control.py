"""Preview control (Kajita 2003) — host classes.

Setup (once per controller): DARE by structure-preserving doubling, feedback gain K and preview gains F
(reference include/CCC/PreviewControl.h:93-172).  Online: jerk = -K x + F ref_seq (:86-89), then the ZMP of
the state one control step ahead (reference src/PreviewControlZmp.cpp:31-49).  The online part takes a
`gemv(K, F, x, ref_seq) -> jerk` callable so that the batched CUDA kernel (engine.PreviewEngine) or the
oracle can be plugged in; the default is plain numpy for the single-problem CPU configuration.
"""
import math

import numpy as np

from .linear_models import ComZmpModelJerkInput


class PreviewControl:
    def __init__(self, model, horizon_duration, horizon_dt, weight_output, weight_input):
        if horizon_duration <= 0 or horizon_dt <= 0:
            raise RuntimeError(f"[PreviewControl] Input arguments are invalid. horizon_duration: {horizon_duration}, "
                               f"horizon_dt: {horizon_dt}")  # :73-77
        self.model, self.horizon_dt = model, horizon_dt
        self.horizon_steps = int(math.ceil(horizon_duration / horizon_dt))
        self._calc_gain(np.atleast_1d(weight_output).astype(float), np.atleast_1d(weight_input).astype(float))

    def _calc_gain(self, w_out, w_in):
        m = self.model
        if m.dt != self.horizon_dt:
            m.calc_disc_matrix(self.horizon_dt)
        A, B, C = m.Ad, m.Bd, m.C
        Q, R, Rinv = np.diag(w_out), np.diag(w_in), np.diag(1.0 / w_in)
        n = A.shape[0]
        # 1. DARE by doubling (:113-144)
        A0, G0, H0 = A.copy(), B @ Rinv @ B.T, C.T @ Q @ C
        self.riccati_converged = False
        for _ in range(10000):
            W = np.linalg.inv(np.eye(n) + G0 @ H0)
            A1 = A0 @ W @ A0
            G1 = G0 + A0 @ W @ G0 @ A0.T
            H1 = H0 + A0.T @ H0 @ W @ A0
            if np.linalg.norm(H1 - H0) / np.linalg.norm(H1) < 1e-8:
                self.riccati_converged = True
                break
            A0, G0, H0 = A1, G1, H1
        self.P = H1
        self.riccati_error = np.linalg.norm(
            self.P - (A.T @ self.P @ A + C.T @ Q @ C - A.T @ self.P @ B @ np.linalg.inv(R + B.T @ self.P @ B) @ B.T @ self.P @ A))
        # 2. gains (:152-171)
        S = np.linalg.inv(R + B.T @ self.P @ B)
        self.K = S @ B.T @ self.P @ A
        p = C.shape[0]
        self.F = np.zeros((B.shape[1], self.horizon_steps * p))
        A_BK = A - B @ self.K
        f_sub = np.eye(n)
        for i in range(self.horizon_steps):
            if i < self.horizon_steps - 1:
                self.F[:, i * p:(i + 1) * p] = S @ B.T @ f_sub @ C.T @ Q
            else:
                self.F[:, i * p:(i + 1) * p] = S @ B.T @ f_sub @ self.P @ C.T
            f_sub = f_sub @ A_BK.T

    def calc_optimal_input(self, x, ref_output_seq):
        return -self.K @ x + self.F @ ref_output_seq  # :86-89


def _numpy_gemv(K, F, x, ref_seq):
    """jerk[b] = -K x[b] + F ref_seq[b] for a batch of rows."""
    return -(x @ K[0]) + ref_seq @ F[0]


class PreviewControlZmp1d(PreviewControl):
    """reference include/CCC/PreviewControlZmp.h:52-68, src/PreviewControlZmp.cpp:8-49."""

    def __init__(self, com_height, horizon_duration, horizon_dt, weight_zmp=1.0, weight_com_jerk=1e-8):
        super().__init__(ComZmpModelJerkInput(com_height), horizon_duration, horizon_dt, [weight_zmp], [weight_com_jerk])

    def proc_once(self, initial_param, ref_zmp_seq, control_dt=-1.0, gemv=_numpy_gemv):
        """initial_param [B][3], ref_zmp_seq [B][N] (sampled at t + (i+1) dt, :22-26) -> ZMP [B]."""
        ip = np.atleast_2d(np.asarray(initial_param, dtype=np.float64))
        ref = np.asarray(ref_zmp_seq, dtype=np.float64).reshape(len(ip), self.horizon_steps)
        jerk = gemv(self.K, self.F, ip, ref)
        if control_dt < 0:
            control_dt = self.horizon_dt
        com_acc = ip[:, 2] + control_dt * jerk
        com_pos = ip[:, 0] + control_dt * ip[:, 1] + 0.5 * control_dt**2 * ip[:, 2]
        return com_pos + self.model.C[0, 2] * com_acc


class PreviewControlZmp:
    """Two axes (reference src/PreviewControlZmp.cpp:51-76)."""

    def __init__(self, com_height, horizon_duration, horizon_dt, weight_zmp=1.0, weight_com_jerk=1e-8):
        self.pc_1d = PreviewControlZmp1d(com_height, horizon_duration, horizon_dt, weight_zmp, weight_com_jerk)

    def plan_batch(self, pos, vel, acc, ref_zmp_seq, control_dt=-1.0, gemv=_numpy_gemv):
        """pos/vel/acc [B][2], ref_zmp_seq [B][N][2] -> planned ZMP [B][2] (both axes as one batch of 2B rows)."""
        B = len(pos)
        ip = np.concatenate([np.stack([pos[:, a], vel[:, a], acc[:, a]], axis=1) for a in range(2)], axis=0)
        ref = np.concatenate([ref_zmp_seq[:, :, 0], ref_zmp_seq[:, :, 1]], axis=0)
        z = self.pc_1d.proc_once(ip, ref, control_dt, gemv)
        return np.stack([z[:B], z[B:]], axis=1)
