"""ctypes mirror of include/ccc_b200.h (plain-data interface structs only)."""
import ctypes as C

import numpy as np

CCC_MEM_HOST = 0
CCC_MEM_DEVICE = 1
CCC_DDP_M_MAX = 32
CCC_DDP_MAX_ALPHA = 16

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int8_p = C.POINTER(C.c_int8)
c_uint32_p = C.POINTER(C.c_uint32)


class DdpConfig(C.Structure):
    """ccc_ddp_config_t"""

    _fields_ = [
        ("with_input_constraint", C.c_int32),
        ("max_iter", C.c_int32),
        ("reg_type", C.c_int32),
        ("n_alpha", C.c_int32),
        ("initial_lambda", C.c_double),
        ("initial_dlambda", C.c_double),
        ("lambda_factor", C.c_double),
        ("lambda_min", C.c_double),
        ("lambda_max", C.c_double),
        ("k_rel_norm_thre", C.c_double),
        ("lambda_thre", C.c_double),
        ("cost_update_ratio_thre", C.c_double),
        ("cost_update_thre", C.c_double),
        ("alpha", C.c_double * CCC_DDP_MAX_ALPHA),
        ("boxqp_max_iter", C.c_int32),
        ("reserved0", C.c_int32),
        ("boxqp_grad_thre", C.c_double),
        ("boxqp_rel_improve_thre", C.c_double),
        ("boxqp_step_factor", C.c_double),
        ("boxqp_min_step", C.c_double),
        ("boxqp_armijo", C.c_double),
    ]


class DdpResult(C.Structure):
    """ccc_ddp_result_t"""

    _fields_ = [
        ("x", C.c_void_p),
        ("u", C.c_void_p),
        ("cost", C.c_void_p),
        ("iters", C.c_void_p),
        ("status", C.c_void_p),
        ("trace_len", C.c_int32),
        ("reserved0", C.c_int32),
        ("alpha_idx", C.c_void_p),
        ("lambda_trace", C.c_void_p),
        ("clamped", C.c_void_p),
    ]


class DdpCentroidalBatch(C.Structure):
    """ccc_ddp_centroidal_batch_t"""

    _fields_ = [
        ("horizon_steps", C.c_int32),
        ("batch", C.c_int32),
        ("n_sched", C.c_int32),
        ("m_max", C.c_int32),
        ("dt", C.c_double),
        ("mass", C.c_double),
        ("sched_id", C.c_void_p),
        ("m", C.c_void_p),
        ("ridge", C.c_void_p),
        ("vertex", C.c_void_p),
        ("ref_pos", C.c_void_p),
        ("w_run", C.c_double * 10),
        ("w_term", C.c_double * 9),
        ("u_lo", C.c_double),
        ("u_hi", C.c_double),
        ("x0", C.c_void_p),
        ("u_init", C.c_void_p),
    ]


class DdpSrbBatch(C.Structure):
    """ccc_ddp_srb_batch_t"""

    _fields_ = [
        ("horizon_steps", C.c_int32),
        ("batch", C.c_int32),
        ("n_sched", C.c_int32),
        ("m_max", C.c_int32),
        ("dt", C.c_double),
        ("mass", C.c_double),
        ("sched_id", C.c_void_p),
        ("m", C.c_void_p),
        ("ridge", C.c_void_p),
        ("vertex", C.c_void_p),
        ("inertia", C.c_void_p),
        ("ref", C.c_void_p),
        ("w_run", C.c_double * 13),
        ("w_term", C.c_double * 12),
        ("u_lo", C.c_double),
        ("u_hi", C.c_double),
        ("x0", C.c_void_p),
        ("u_init", C.c_void_p),
    ]


class DdpZmpBatch(C.Structure):
    """ccc_ddp_zmp_batch_t"""

    _fields_ = [("horizon_steps", C.c_int32), ("batch", C.c_int32), ("n_sched", C.c_int32), ("reserved0", C.c_int32),
                ("dt", C.c_double), ("mass", C.c_double), ("sched_id", C.c_void_p), ("ref_zmp", C.c_void_p),
                ("com_z", C.c_void_p), ("w", C.c_double * 6), ("x0", C.c_void_p), ("u_init", C.c_void_p)]


class DdpCentroidalLoop(C.Structure):
    """ccc_ddp_centroidal_loop_t"""

    _fields_ = [("horizon_steps", C.c_int32), ("batch", C.c_int32), ("n_sched", C.c_int32), ("m_max", C.c_int32),
                ("dt", C.c_double), ("mass", C.c_double), ("sim_dt", C.c_double),
                ("ticks", C.c_int32), ("stride", C.c_int32), ("grid_len", C.c_int32), ("max_iter_later", C.c_int32),
                ("sched_id", C.c_void_p), ("m", C.c_void_p), ("ridge", C.c_void_p), ("vertex", C.c_void_p), ("ref_pos", C.c_void_p),
                ("w_run", C.c_double * 10), ("w_term", C.c_double * 9), ("u_lo", C.c_double), ("u_hi", C.c_double),
                ("plant0", C.c_void_p), ("disturb_tick", C.c_int32), ("reserved0", C.c_int32), ("disturb_vel", C.c_double * 3)]


class DdpCentroidalLoopResult(C.Structure):
    """ccc_ddp_centroidal_loop_result_t"""

    _fields_ = [("plant", C.c_void_p), ("u0", C.c_void_p), ("iters", C.c_void_p)]


class QpBatch(C.Structure):
    """ccc_qp_batch_t"""

    _fields_ = [("n", C.c_int32), ("n_eq", C.c_int32), ("n_ineq", C.c_int32), ("batch", C.c_int32),
                ("Q", C.c_void_p), ("A", C.c_void_p), ("C", C.c_void_p),
                ("c", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p)]


class QpResult(C.Structure):
    """ccc_qp_result_t"""

    _fields_ = [("x", C.c_void_p), ("iters", C.c_void_p), ("status", C.c_void_p), ("n_active", C.c_void_p),
                ("active", C.c_void_p)]


class LinearMpcXyBatch(C.Structure):
    """ccc_linear_mpc_xy_batch_t"""

    _fields_ = [("horizon_steps", C.c_int32), ("batch", C.c_int32), ("n_sched", C.c_int32), ("m_max", C.c_int32),
                ("dt", C.c_double), ("mass", C.c_double), ("sched_id", C.c_void_p), ("m", C.c_void_p),
                ("ridge", C.c_void_p), ("vertex", C.c_void_p), ("com_z", C.c_void_p), ("total_force_z", C.c_void_p),
                ("ref_output", C.c_void_p), ("w_output", C.c_double * 6), ("w_force", C.c_double),
                ("force_lo", C.c_double), ("force_hi", C.c_double), ("x0", C.c_void_p)]


class LinearMpcXyResult(C.Structure):
    """ccc_linear_mpc_xy_result_t"""

    _fields_ = [("u", C.c_void_p), ("iters", C.c_void_p), ("status", C.c_void_p), ("n_active", C.c_void_p),
                ("active", C.c_void_p), ("A_seq", C.c_void_p), ("B_seq", C.c_void_p), ("obj_mat", C.c_void_p),
                ("obj_vec", C.c_void_p)]


class FootstepPlans(C.Structure):
    """ccc_footstep_plans_t"""

    _fields_ = [("n_plans", C.c_int32), ("max_steps", C.c_int32), ("horizon_steps", C.c_int32), ("eps_reps", C.c_int32),
                ("horizon_dt", C.c_double), ("manager_horizon", C.c_double), ("foot_size", C.c_double * 2),
                ("current_time", C.c_void_p), ("stance0", C.c_void_p), ("n_steps", C.c_void_p), ("foot", C.c_void_p),
                ("pos", C.c_void_p), ("times", C.c_void_p)]


class ZmpTables(C.Structure):
    """ccc_zmp_tables_t"""

    _fields_ = [("ref_zmp", C.c_void_p), ("lim_min", C.c_void_p), ("lim_max", C.c_void_p)]


class ZmpMpcBatch(C.Structure):
    """ccc_zmp_mpc_batch_t"""

    _fields_ = [("method", C.c_int32), ("horizon_steps", C.c_int32), ("batch", C.c_int32), ("n_plans", C.c_int32),
                ("control_dt", C.c_double), ("com_height_over_g", C.c_double), ("weight_zmp", C.c_double),
                ("Q", C.c_void_p), ("A", C.c_void_p), ("C", C.c_void_p), ("A_seq", C.c_void_p), ("P", C.c_void_p),
                ("plan_id", C.c_void_p), ("state", C.c_void_p), ("tables", ZmpTables)]


class ZmpMpcResult(C.Structure):
    """ccc_zmp_mpc_result_t"""

    _fields_ = [("planned_zmp", C.c_void_p), ("iters", C.c_void_p), ("status", C.c_void_p)]


class DcmTrackingBatch(C.Structure):
    """ccc_dcm_tracking_batch_t"""

    _fields_ = [("batch", C.c_int32), ("n_plans", C.c_int32), ("max_knots", C.c_int32), ("reserved0", C.c_int32),
                ("omega", C.c_double), ("feedback_gain", C.c_double), ("plan_id", C.c_void_p), ("dcm", C.c_void_p),
                ("current_time", C.c_void_p), ("current_zmp", C.c_void_p), ("n_knots", C.c_void_p), ("knot_time", C.c_void_p),
                ("knot_zmp", C.c_void_p)]


class FootGuidedBatch(C.Structure):
    """ccc_foot_guided_batch_t"""

    _fields_ = [("batch", C.c_int32), ("n_plans", C.c_int32), ("omega", C.c_double), ("plan_id", C.c_void_p),
                ("capture_point", C.c_void_p), ("current_time", C.c_void_p), ("transit_start_zmp", C.c_void_p),
                ("transit_end_zmp", C.c_void_p), ("transit_start_time", C.c_void_p), ("transit_duration", C.c_void_p)]


class SingularPreviewBatch(C.Structure):
    """ccc_singular_preview_batch_t"""

    _fields_ = [("batch", C.c_int32), ("n_plans", C.c_int32), ("horizon_steps", C.c_int32), ("reserved0", C.c_int32),
                ("omega", C.c_double), ("horizon_dt", C.c_double), ("control_dt", C.c_double), ("plan_id", C.c_void_p),
                ("state", C.c_void_p), ("ref_zmp", C.c_void_p)]


class StepMpcBatch(C.Structure):
    """ccc_step_mpc_batch_t"""

    _fields_ = [("batch", C.c_int32), ("n_plans", C.c_int32), ("max_elements", C.c_int32), ("reserved0", C.c_int32),
                ("com_height", C.c_double), ("w_free_zmp", C.c_double), ("w_fixed_zmp", C.c_double), ("w_double_support", C.c_double),
                ("w_pos", C.c_double), ("w_vel", C.c_double), ("w_capture_point_abs", C.c_double), ("w_capture_point_rel", C.c_double),
                ("plan_id", C.c_void_p), ("x_pos", C.c_void_p), ("x_vel", C.c_void_p), ("current_time", C.c_void_p),
                ("n_elements", C.c_void_p), ("single", C.c_void_p), ("zmp", C.c_void_p), ("end_time", C.c_void_p)]


class StepMpcResult(C.Structure):
    """ccc_step_mpc_result_t"""

    _fields_ = [("current_zmp", C.c_void_p), ("next_foot_zmp", C.c_void_p), ("has_next", C.c_void_p)]


def ptr(a):
    """Address of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return a.ctypes.data
