"""Host classes of the closed-form ZMP controllers: DcmTracking (reference include/CCC/DcmTracking.h,
src/DcmTracking.cpp), FootGuidedControl (include/CCC/FootGuidedControl.h, src/FootGuidedControl.cpp) and
SingularPreviewControlZmp (include/CCC/SingularPreviewControlZmp.h, src/SingularPreviewControlZmp.cpp).

`plan_once` is the reference's planOnce for one problem in numpy scalars (the check of the batched kernels);
`plan_batch` flattens P reference-data records and B initial parameters into the C-ABI structs and runs
ccc_dcm_tracking_plan / ccc_foot_guided_plan through the `run` callable (engine.dcm_tracking_plan /
engine.foot_guided_plan).
"""
import math

import numpy as np

from . import _abi
from ._abi import ptr
from .linear_models import G


class DcmTracking:
    def __init__(self, com_height, feedback_gain=2.0):
        self.feedback_gain = feedback_gain
        self.omega = math.sqrt(G / com_height)  # include/CCC/DcmTracking.h:51

    def plan_once(self, current_zmp, time_zmp_list, dcm, current_time):
        """src/DcmTracking.cpp:7-48; time_zmp_list: [(time, zmp[2]), ...] ascending."""
        current_zmp, dcm = np.asarray(current_zmp, dtype=np.float64), np.asarray(dcm, dtype=np.float64)
        if not time_zmp_list:
            return current_zmp + (1.0 + self.feedback_gain / self.omega) * (dcm - current_zmp)
        for t, _ in time_zmp_list:
            if t < current_time:
                raise RuntimeError(f"ZMP switching time must be in the future: {t} < {current_time}")
        dcm_switch = np.asarray(time_zmp_list[-1][1], dtype=np.float64)
        for i in range(len(time_zmp_list) - 2, -1, -1):
            zmp_duration = time_zmp_list[i + 1][0] - time_zmp_list[i][0]
            z = np.asarray(time_zmp_list[i][1], dtype=np.float64)
            dcm_switch = z + math.exp(-1 * self.omega * zmp_duration) * (dcm_switch - z)
        target_dcm = current_zmp + math.exp(self.omega * (current_time - time_zmp_list[0][0])) * (dcm_switch - current_zmp)
        return current_zmp + (1.0 + self.feedback_gain / self.omega) * (dcm - target_dcm)

    def plan_batch(self, run, ref_data, current_times, dcm, plan_id):
        """ref_data: P records (current_zmp[2], [(time, zmp[2]), ...]); current_times [P]; dcm [B][2]; plan_id [B]."""
        P, B = len(ref_data), len(plan_id)
        K = max([len(r[1]) for r in ref_data] + [1])
        keep = dict(plan_id=np.ascontiguousarray(plan_id, dtype=np.int32), dcm=np.ascontiguousarray(dcm, dtype=np.float64).reshape(B, 2),
                    current_time=np.ascontiguousarray(current_times, dtype=np.float64), current_zmp=np.zeros((P, 2)),
                    n_knots=np.zeros(P, dtype=np.int32), knot_time=np.zeros((P, K)), knot_zmp=np.zeros((P, K, 2)))
        for p, (cz, knots) in enumerate(ref_data):
            keep["current_zmp"][p] = cz
            keep["n_knots"][p] = len(knots)
            for i, (t, z) in enumerate(knots):
                keep["knot_time"][p, i], keep["knot_zmp"][p, i] = t, z
        bt = _abi.DcmTrackingBatch()
        bt.batch, bt.n_plans, bt.max_knots, bt.omega, bt.feedback_gain = B, P, K, self.omega, self.feedback_gain
        for k, v in keep.items():
            setattr(bt, k, ptr(v))
        return run(bt, B)


class FootGuidedControl:
    def __init__(self, com_height):
        self.omega = math.sqrt(G / com_height)  # include/CCC/FootGuidedControl.h:51

    def plan_once_1d(self, start_zmp, end_zmp, transit_start_time, transit_duration, capture_point, current_time):
        """FootGuidedControl1d::planOnce (src/FootGuidedControl.cpp:11-69)."""
        omega = self.omega
        if not transit_duration >= 0:
            raise RuntimeError(f"Transition duration must be non-negative: {transit_duration}")
        transit_end_time = transit_start_time + transit_duration
        if not transit_end_time >= current_time + 1e-6:
            raise RuntimeError("Transition end time must be in the future with some margin")
        if transit_duration == 0:
            return start_zmp + 2 * ((capture_point - start_zmp) - (end_zmp - start_zmp) * math.exp(-1 * omega * (transit_start_time - current_time))) \
                / (1.0 - math.exp(-2 * omega * (transit_start_time - current_time)))
        zmp_transit_vel = (end_zmp - start_zmp) / transit_duration
        if current_time <= transit_start_time:
            return start_zmp + (2 * (capture_point - start_zmp) + 2 * zmp_transit_vel / omega
                                * (math.exp(-1 * omega * (transit_end_time - current_time)) - math.exp(-1 * omega * (transit_start_time - current_time)))) \
                / (1.0 - math.exp(-2 * omega * (transit_end_time - current_time)))
        current_ref_zmp = start_zmp + zmp_transit_vel * (current_time - transit_start_time)
        return current_ref_zmp + (2 * (capture_point - current_ref_zmp) + 2 * zmp_transit_vel / omega
                                  * (math.exp(-1 * omega * (transit_end_time - current_time)) - 1.0)) \
            / (1.0 - math.exp(-2 * omega * (transit_end_time - current_time)))

    def plan_once(self, rd, capture_point, current_time):
        """FootGuidedControl::planOnce (:71-92); rd: dict(transit_start_zmp, transit_end_zmp, transit_start_time, transit_duration)."""
        return np.array([self.plan_once_1d(rd["transit_start_zmp"][a], rd["transit_end_zmp"][a], rd["transit_start_time"],
                                           rd["transit_duration"], capture_point[a], current_time) for a in range(2)])

    def plan_batch(self, run, ref_data, current_times, capture_point, plan_id):
        P, B = len(ref_data), len(plan_id)
        keep = dict(plan_id=np.ascontiguousarray(plan_id, dtype=np.int32),
                    capture_point=np.ascontiguousarray(capture_point, dtype=np.float64).reshape(B, 2),
                    current_time=np.ascontiguousarray(current_times, dtype=np.float64),
                    transit_start_zmp=np.array([r["transit_start_zmp"] for r in ref_data], dtype=np.float64).reshape(P, 2),
                    transit_end_zmp=np.array([r["transit_end_zmp"] for r in ref_data], dtype=np.float64).reshape(P, 2),
                    transit_start_time=np.array([r["transit_start_time"] for r in ref_data], dtype=np.float64),
                    transit_duration=np.array([r["transit_duration"] for r in ref_data], dtype=np.float64))
        bt = _abi.FootGuidedBatch()
        bt.batch, bt.n_plans, bt.omega = B, P, self.omega
        for k, v in keep.items():
            setattr(bt, k, ptr(v))
        return run(bt, B)


class SingularPreviewControlZmp:
    """Urata's singular LQ preview regulation (reference include/CCC/SingularPreviewControlZmp.h:46-51, 102-106)."""

    def __init__(self, com_height, horizon_duration, horizon_dt):
        self.horizon_dt = horizon_dt
        self.horizon_steps = int(math.ceil(horizon_duration / horizon_dt))
        self.omega = math.sqrt(G / com_height)

    def sample(self, ref_zmp_func, current_time):
        """planOnce's sampling of the reference (src/SingularPreviewControlZmp.cpp:66-71) -> [N][2]."""
        return np.array([np.asarray(ref_zmp_func(current_time + i * self.horizon_dt), dtype=np.float64) for i in range(self.horizon_steps)])

    def proc_once_1d(self, ref_zmp_seq, planned_zmp, pos, vel, control_dt):
        """src/SingularPreviewControlZmp.cpp:23-57 in Python floats, the expressions term by term (std::pow(b, 2) as b * b)."""
        w, dt = self.omega, self.horizon_dt
        k0 = (1 + w * dt) / dt + w / (1 + w * dt)
        k1 = -1 * (2 + w * dt) / dt
        k2 = -1 * (2 + w * dt) / (w * dt)
        u_fb = -1 * ((k0 * planned_zmp + k1 * pos) + k2 * vel)
        s0 = float(ref_zmp_seq[-1]) * (1 + w * dt) / (w * dt)
        b = 1 + w * dt
        for i in range(len(ref_zmp_seq) - 2, 0, -1):
            s0 = float(ref_zmp_seq[i]) + s0 / b
        u_ff = float(ref_zmp_seq[0]) / dt - w * (2 + w * dt) * s0 / (b * b)
        return planned_zmp + control_dt * (u_fb + u_ff)

    def plan_once(self, ref_zmp_func, pos, vel, planned_zmp, current_time, control_dt=-1):
        """2-D planOnce (:59-88) on the host."""
        control_dt = self.horizon_dt if control_dt < 0 else control_dt
        seq = self.sample(ref_zmp_func, current_time)
        return np.array([self.proc_once_1d(seq[:, a], float(planned_zmp[a]), float(pos[a]), float(vel[a]), control_dt) for a in range(2)])

    def plan_batch(self, run, ref_zmp, state, plan_id, control_dt=-1):
        """ref_zmp [P][N][2] (sampled sequences); state [B][2][3] per axis (planned_zmp, pos, vel); plan_id [B]."""
        keep = dict(plan_id=np.ascontiguousarray(plan_id, dtype=np.int32), state=np.ascontiguousarray(state, dtype=np.float64),
                    ref_zmp=np.ascontiguousarray(ref_zmp, dtype=np.float64))
        P, B = keep["ref_zmp"].shape[0], len(keep["plan_id"])
        assert keep["ref_zmp"].shape == (P, self.horizon_steps, 2) and keep["state"].shape == (B, 2, 3)
        bt = _abi.SingularPreviewBatch()
        bt.batch, bt.n_plans, bt.horizon_steps = B, P, self.horizon_steps
        bt.omega, bt.horizon_dt, bt.control_dt = self.omega, self.horizon_dt, (self.horizon_dt if control_dt < 0 else control_dt)
        for k, v in keep.items():
            setattr(bt, k, ptr(v))
        return run(bt, B)
