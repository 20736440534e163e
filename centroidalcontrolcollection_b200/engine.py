"""ctypes binding of libccc_b200.so — the host-side mirror of the reference's solver objects.

There is no CPU fallback: loading fails loudly if the CUDA library is missing, and every solve
returns CCC_ERR_CUDA without a device.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from .problem import DdpCentroidalProblemSet, DdpResultArrays, DdpSrbProblemSet

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libccc_b200.so")
_LIB = None

EXPORTS = [
    "ccc_abi_version", "ccc_device_count", "ccc_last_error", "ccc_ddp_config_default",
    "ccc_ddp_centroidal_create", "ccc_ddp_centroidal_destroy", "ccc_ddp_centroidal_solve",
    "ccc_ddp_centroidal_last_launches",
    "ccc_ddp_srb_create", "ccc_ddp_srb_destroy", "ccc_ddp_srb_solve", "ccc_ddp_srb_last_launches",
    "ccc_ddp_zmp_create", "ccc_ddp_zmp_destroy", "ccc_ddp_zmp_solve", "ccc_ddp_zmp_last_launches",
    "ccc_qp_create", "ccc_qp_destroy", "ccc_qp_solve", "ccc_qp_last_launches", "ccc_preview_input",
    "ccc_qp_create_grouped", "ccc_qp_solve_grouped",
    "ccc_fp64_peak_tflops",
    "ccc_dcm_tracking_plan", "ccc_foot_guided_plan", "ccc_singular_preview_plan", "ccc_step_mpc_plan",
    "ccc_footstep_compile", "ccc_zmp_mpc_create", "ccc_zmp_mpc_destroy", "ccc_zmp_mpc_plan", "ccc_zmp_mpc_last_launches",
    "ccc_linear_mpc_xy_create", "ccc_linear_mpc_xy_destroy", "ccc_linear_mpc_xy_solve", "ccc_linear_mpc_xy_last_launches",
]


class EngineError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} is missing: build it with `python -m centroidalcontrolcollection_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.ccc_abi_version.restype = C.c_int32
        L.ccc_device_count.restype = C.c_int32
        L.ccc_last_error.restype = C.c_char_p
        L.ccc_fp64_peak_tflops.restype = C.c_double
        L.ccc_fp64_peak_tflops.argtypes = [C.c_int32, C.c_void_p]
        L.ccc_ddp_config_default.argtypes = [C.c_void_p]
        L.ccc_ddp_config_default.restype = None
        L.ccc_ddp_centroidal_create.restype = C.c_void_p
        L.ccc_ddp_centroidal_create.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.ccc_ddp_centroidal_destroy.argtypes = [C.c_void_p]
        L.ccc_ddp_centroidal_destroy.restype = None
        L.ccc_ddp_centroidal_solve.restype = C.c_int32
        L.ccc_ddp_centroidal_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_ddp_centroidal_last_launches.restype = C.c_int32
        L.ccc_ddp_centroidal_last_launches.argtypes = [C.c_void_p]
        L.ccc_ddp_centroidal_set_variant.restype = C.c_int32
        L.ccc_ddp_centroidal_set_variant.argtypes = [C.c_int32]
        L.ccc_ddp_centroidal_set_chunk.restype = None
        L.ccc_ddp_centroidal_set_chunk.argtypes = [C.c_int32]
        L.ccc_ddp_set_small_batch_policy.restype = None
        L.ccc_ddp_set_small_batch_policy.argtypes = [C.c_int32, C.c_int32]
        L.ccc_ddp_set_packed_io.restype = None
        L.ccc_ddp_set_packed_io.argtypes = [C.c_int32]
        L.ccc_ddp_centroidal_last_team.restype = C.c_int32
        L.ccc_ddp_centroidal_last_team.argtypes = [C.c_void_p]
        L.ccc_ddp_centroidal_closed_loop.restype = C.c_int32
        L.ccc_ddp_centroidal_closed_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_ddp_srb_create.restype = C.c_void_p
        L.ccc_ddp_srb_create.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.ccc_ddp_srb_destroy.argtypes = [C.c_void_p]
        L.ccc_ddp_srb_destroy.restype = None
        L.ccc_ddp_srb_solve.restype = C.c_int32
        L.ccc_ddp_srb_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_ddp_srb_last_launches.restype = C.c_int32
        L.ccc_ddp_srb_last_launches.argtypes = [C.c_void_p]
        L.ccc_ddp_zmp_create.restype = C.c_void_p
        L.ccc_ddp_zmp_create.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.ccc_ddp_zmp_destroy.argtypes = [C.c_void_p]
        L.ccc_ddp_zmp_destroy.restype = None
        L.ccc_ddp_zmp_solve.restype = C.c_int32
        L.ccc_ddp_zmp_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_ddp_zmp_last_launches.restype = C.c_int32
        L.ccc_ddp_zmp_last_launches.argtypes = [C.c_void_p]
        L.ccc_qp_create.restype = C.c_void_p
        L.ccc_qp_create.argtypes = [C.c_int32] * 4
        L.ccc_qp_destroy.argtypes = [C.c_void_p]
        L.ccc_qp_destroy.restype = None
        L.ccc_qp_solve.restype = C.c_int32
        L.ccc_qp_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_qp_create_grouped.restype = C.c_void_p
        L.ccc_qp_create_grouped.argtypes = [C.c_int32] * 5
        L.ccc_qp_solve_grouped.restype = C.c_int32
        L.ccc_qp_solve_grouped.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_qp_last_launches.restype = C.c_int32
        L.ccc_qp_last_launches.argtypes = [C.c_void_p]
        L.ccc_qp_set_packed.restype = None
        L.ccc_qp_set_packed.argtypes = [C.c_int32]
        L.ccc_preview_input.restype = C.c_int32
        L.ccc_preview_input.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]
        for fn in ("ccc_dcm_tracking_plan", "ccc_foot_guided_plan", "ccc_singular_preview_plan"):
            getattr(L, fn).restype = C.c_int32
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_footstep_compile.restype = C.c_int32
        L.ccc_footstep_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_zmp_mpc_create.restype = C.c_void_p
        L.ccc_zmp_mpc_create.argtypes = [C.c_int32] * 4
        L.ccc_zmp_mpc_destroy.argtypes = [C.c_void_p]
        L.ccc_zmp_mpc_destroy.restype = None
        L.ccc_zmp_mpc_plan.restype = C.c_int32
        L.ccc_zmp_mpc_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_zmp_mpc_last_launches.restype = C.c_int32
        L.ccc_zmp_mpc_last_launches.argtypes = [C.c_void_p]
        L.ccc_linear_mpc_xy_create.restype = C.c_void_p
        L.ccc_linear_mpc_xy_create.argtypes = [C.c_int32] * 5
        L.ccc_linear_mpc_xy_destroy.argtypes = [C.c_void_p]
        L.ccc_linear_mpc_xy_destroy.restype = None
        L.ccc_linear_mpc_xy_solve.restype = C.c_int32
        L.ccc_linear_mpc_xy_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.ccc_linear_mpc_xy_last_launches.restype = C.c_int32
        L.ccc_linear_mpc_xy_last_launches.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def last_error():
    return lib().ccc_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise EngineError(f"{what} failed with code {rc}: {last_error()}")


class _DdpEngineBase:
    """Owns one device workspace of the C-ABI and solves batches with it."""

    _prefix = None

    def __init__(self, horizon_steps, max_batch, max_sched=16):
        self._fn = lambda name: getattr(lib(), f"{self._prefix}_{name}")
        self._h = self._fn("create")(int(horizon_steps), int(max_batch), int(max_sched))
        if not self._h:
            raise EngineError(f"{self._prefix}_create failed: {last_error()}")
        self.N, self.max_batch, self.max_sched = horizon_steps, max_batch, max_sched

    def close(self):
        if getattr(self, "_h", None):
            self._fn("destroy")(self._h)
            self._h = None

    __del__ = close

    def solve(self, problem_set, cfg, trace_len=0, result=None):
        """Host buffers in, host buffers out (H2D + solve + D2H, synchronous)."""
        res = result if result is not None else problem_set.new_result(trace_len)
        bs, rs = problem_set.as_struct(), res.as_struct()
        rc = self._fn("solve")(self._h, C.addressof(bs), C.addressof(cfg), C.addressof(rs), _abi.CCC_MEM_HOST, None)
        _check(rc, f"{self._prefix}_solve")
        return res

    def solve_device(self, batch_struct, cfg, result_struct, stream=0):
        """Device pointers in/out; only enqueues work on `stream` (an int cudaStream_t)."""
        rc = self._fn("solve")(self._h, C.addressof(batch_struct), C.addressof(cfg), C.addressof(result_struct),
                               _abi.CCC_MEM_DEVICE, C.c_void_p(stream))
        _check(rc, f"{self._prefix}_solve")

    @property
    def last_launches(self):
        return int(self._fn("last_launches")(self._h))


class DdpSrbEngine(_DdpEngineBase):
    """Batched counterpart of CCC::DdpSingleRigidBody's solver object (reference
    include/CCC/DdpSingleRigidBody.h:382-405)."""

    _prefix = "ccc_ddp_srb"


class DdpZmpEngine(_DdpEngineBase):
    """Batched counterpart of CCC::DdpZmp's solver object (reference include/CCC/DdpZmp.h:276-300)."""

    _prefix = "ccc_ddp_zmp"

    @staticmethod
    def set_variant(v):
        """Tuning hook, read when an engine is created: 1 = one thread per problem (default), 0 = the warp-per-problem
        engine (3 of 32 lanes live).  Returns the previous value."""
        f = lib().ccc_ddp_zmp_set_variant
        f.restype, f.argtypes = C.c_int32, [C.c_int32]
        return int(f(int(v)))


class DdpCentroidalEngine(_DdpEngineBase):
    """Batched counterpart of CCC::DdpCentroidal's solver object (reference
    include/CCC/DdpCentroidal.h:342-365): owns the device workspace, solves batches."""

    _prefix = "ccc_ddp_centroidal"

    def closed_loop(self, loop, cfg):
        """Device-resident receding-horizon loop (ccc_ddp_centroidal_closed_loop, host buffers in / out) for a
        closed_loop.CentroidalLoop; returns its LoopResultArrays."""
        res = loop.new_result()
        ls, rs = loop.as_struct(), res.as_struct()
        rc = lib().ccc_ddp_centroidal_closed_loop(self._h, C.addressof(ls), C.addressof(cfg), C.addressof(rs), _abi.CCC_MEM_HOST, None)
        _check(rc, "ccc_ddp_centroidal_closed_loop")
        return res

    @staticmethod
    def set_variant(v):
        """Tuning hook: launch shape of the solve kernel (0: 6 warps/SM, 1: 8 warps/SM as two CTAs, 2: 8 warps/SM in one CTA = default)."""
        return int(lib().ccc_ddp_centroidal_set_variant(int(v)))

    @staticmethod
    def set_chunk(iters):
        """Tuning hook: DDP iterations per visit before a solve is suspended and re-queued (0 = never)."""
        lib().ccc_ddp_centroidal_set_chunk(int(iters))

    @staticmethod
    def set_small_batch_policy(team=-1, spread=-1):
        """Tuning hook of every DDP engine (results never depend on it): team = 1 (default) runs batches of at most one
        problem per SM on the team kernel (csrc/ddp_team.cuh: one CTA per problem, concurrent line-search rollouts),
        spread = 1 (default) spreads batches smaller than the resident warps over all SMs; -1 leaves a setting as it is."""
        lib().ccc_ddp_set_small_batch_policy(int(team), int(spread))

    @staticmethod
    def set_packed_io(on):
        """Tuning hook of every DDP engine: 1 (default) = host-buffer calls whose arrays fit a 1 MB staging block move through
        it with one H2D and one D2H copy (planOnce-sized batches); 0 = one copy per array."""
        lib().ccc_ddp_set_packed_io(int(on))

    @property
    def last_team(self):
        """True if the last solve of this workspace ran on the team kernel."""
        return bool(lib().ccc_ddp_centroidal_last_team(self._h))


class QpEngine:
    """Batched counterpart of QpSolverCollection::QpSolver (reference src/LinearMpcZmp.cpp:21,69,
    src/IntrinsicallyStableMpc.cpp:28,93): one workspace per problem shape, solves batches that share Q, A, C."""

    def __init__(self, n, n_eq, n_ineq, max_batch, max_groups=1):
        self._h = lib().ccc_qp_create_grouped(int(n), int(n_eq), int(n_ineq), int(max_batch), int(max_groups))
        if not self._h:
            raise EngineError(f"ccc_qp_create failed: {last_error()}")
        self.shape = (n, n_eq, n_ineq)
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "_h", None):
            lib().ccc_qp_destroy(self._h)
            self._h = None

    __del__ = close

    def solve(self, problem_set, result=None, reuse_matrices=False):
        """Host buffers in/out (H2D + setup + solve + D2H, synchronous).  reuse_matrices: Q = A = C = NULL, i.e. the
        matrices and their factorisation of the previous call on this workspace are kept and only the per-problem vectors
        are new — what a controller does on every tick after its first (CCC/detail/QpEngine.h matrices_resident_; the
        reference fills qp_coeff_'s matrices in its constructor and only rewrites the vectors in procOnce)."""
        res = result if result is not None else problem_set.new_result()
        bs, rs = problem_set.as_struct(), res.as_struct()
        if reuse_matrices:
            bs.Q = bs.A = bs.C = None
        _check(lib().ccc_qp_solve(self._h, C.addressof(bs), C.addressof(rs), _abi.CCC_MEM_HOST, None), "ccc_qp_solve")
        return res

    def solve_grouped(self, grouped, result=None):
        """Host buffers in / out for a qp.QpGroupedProblemSet (one factorisation per matrix group)."""
        res = result if result is not None else grouped.new_result()
        bs, rs = grouped.as_struct(), res.as_struct()
        _check(lib().ccc_qp_solve_grouped(self._h, C.addressof(bs), int(grouped.G), _abi.ptr(grouped.group_id), C.addressof(rs),
                                          _abi.CCC_MEM_HOST, None), "ccc_qp_solve_grouped")
        return res

    def solve_device(self, batch_struct, result_struct, stream=0):
        _check(lib().ccc_qp_solve(self._h, C.addressof(batch_struct), C.addressof(result_struct), _abi.CCC_MEM_DEVICE,
                                  C.c_void_p(stream)), "ccc_qp_solve")

    @property
    def last_launches(self):
        return int(lib().ccc_qp_last_launches(self._h))

    @staticmethod
    def set_packed(on):
        """Tuning hook: 1 (default) = first pass with a packed R, two CTAs per SM, full-R pass for the problems whose
        active set outgrows it; 0 = the full-R kernel alone (one CTA per SM, round 1)."""
        lib().ccc_qp_set_packed(int(on))


def dcm_tracking_plan(batch_struct, batch):
    """ccc_dcm_tracking_plan with host buffers -> control ZMP [B][2] (closed_form.DcmTracking.plan_batch builds the struct)."""
    out = np.zeros((batch, 2))
    _check(lib().ccc_dcm_tracking_plan(C.addressof(batch_struct), out.ctypes.data, _abi.CCC_MEM_HOST, None), "ccc_dcm_tracking_plan")
    return out


def foot_guided_plan(batch_struct, batch):
    """ccc_foot_guided_plan with host buffers -> planned ZMP [B][2]."""
    out = np.zeros((batch, 2))
    _check(lib().ccc_foot_guided_plan(C.addressof(batch_struct), out.ctypes.data, _abi.CCC_MEM_HOST, None), "ccc_foot_guided_plan")
    return out


def singular_preview_plan(batch_struct, batch):
    """ccc_singular_preview_plan with host buffers -> planned ZMP [B][2] (closed_form.SingularPreviewControlZmp.plan_batch)."""
    out = np.zeros((batch, 2))
    _check(lib().ccc_singular_preview_plan(C.addressof(batch_struct), out.ctypes.data, _abi.CCC_MEM_HOST, None), "ccc_singular_preview_plan")
    return out


def step_mpc_plan(batch_struct, batch):
    """ccc_step_mpc_plan with host buffers -> (current_zmp [B][2], next_foot_zmp [B][2], has_next [B]) (step_mpc.StepMpc.plan_batch)."""
    cur, nxt, has = np.zeros((batch, 2)), np.zeros((batch, 2)), np.zeros(batch, dtype=np.int32)
    rs = _abi.StepMpcResult()
    rs.current_zmp, rs.next_foot_zmp, rs.has_next = cur.ctypes.data, nxt.ctypes.data, has.ctypes.data
    L = lib()
    L.ccc_step_mpc_plan.restype = C.c_int32
    L.ccc_step_mpc_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    _check(L.ccc_step_mpc_plan(C.addressof(batch_struct), C.addressof(rs), _abi.CCC_MEM_HOST, None), "ccc_step_mpc_plan")
    return cur, nxt, has


def footstep_compile(plans, tables=None):
    """Schedule compiler (ccc_footstep_compile, host buffers): schedule.FootstepPlans -> schedule.ZmpTables."""
    from .schedule import ZmpTables

    tables = tables if tables is not None else ZmpTables(plans.P, plans.N)
    ps, ts = plans.as_struct(), tables.as_struct()
    _check(lib().ccc_footstep_compile(C.addressof(ps), C.addressof(ts), _abi.CCC_MEM_HOST, None), "ccc_footstep_compile")
    return tables


class ZmpMpcEngine:
    """planOnce of CCC::LinearMpcZmp (method 0) / CCC::IntrinsicallyStableMpc (method 1) for a batch of problems
    whose reference data are rows of compiled stage tables: QP vectors assembled, QPs solved and the planned ZMP
    post-processed on the device (ccc_zmp_mpc_plan).  `mpc` is the host class of linear_mpc.py (its constructor did
    the batch-invariant setup: Q, A, C, A_seq / P)."""

    def __init__(self, mpc, max_batch, max_plans):
        from . import linear_mpc

        self.mpc = m1 = mpc.mpc_1d
        self.method = 1 if isinstance(m1, linear_mpc.IntrinsicallyStableMpc1d) else 0
        self._h = lib().ccc_zmp_mpc_create(self.method, m1.horizon_steps, int(max_batch), int(max_plans))
        if not self._h:
            raise EngineError(f"ccc_zmp_mpc_create failed: {last_error()}")
        self._mats = [np.ascontiguousarray(m1.Q), np.ascontiguousarray(m1.C)]
        if self.method == 1:
            self._mats += [np.ascontiguousarray(m1.A), np.ascontiguousarray(m1.P)]
        else:
            self._mats += [np.ascontiguousarray(m1.seq_ext.A_seq)]
        self._uploaded = False

    def close(self):
        if getattr(self, "_h", None):
            lib().ccc_zmp_mpc_destroy(self._h)
            self._h = None

    __del__ = close

    def plan(self, state, plan_id, tables, control_dt=-1.0, want_qp_info=True):
        """state: method 0 [B][2][3] (pos, vel, acc per axis), method 1 [B][2][2] (capture point, planned zmp per
        axis); plan_id [B]; tables: schedule.ZmpTables -> (planned_zmp [B][2], iters [2B], status [2B])."""
        m1 = self.mpc
        state = np.ascontiguousarray(state, dtype=np.float64)
        plan_id = np.ascontiguousarray(plan_id, dtype=np.int32)
        B = len(plan_id)
        bt = _abi.ZmpMpcBatch()
        bt.method, bt.horizon_steps, bt.batch, bt.n_plans = self.method, m1.horizon_steps, B, tables.lim_min.shape[0]
        bt.control_dt = control_dt if control_dt > 0 else m1.horizon_dt
        if self.method == 0:
            bt.com_height_over_g = -float(m1.model.C[0, 2])
        else:
            bt.weight_zmp = m1.weight_zmp
        if not self._uploaded:
            bt.Q, bt.C = _abi.ptr(self._mats[0]), _abi.ptr(self._mats[1])
            if self.method == 1:
                bt.A, bt.P = _abi.ptr(self._mats[2]), _abi.ptr(self._mats[3])
            else:
                bt.A_seq = _abi.ptr(self._mats[2])
        bt.plan_id, bt.state, bt.tables = _abi.ptr(plan_id), _abi.ptr(state), tables.as_struct()
        planned = np.zeros((B, 2))
        iters = np.zeros(2 * B, dtype=np.int32) if want_qp_info else None
        status = np.zeros(2 * B, dtype=np.int32) if want_qp_info else None
        rs = _abi.ZmpMpcResult()
        rs.planned_zmp, rs.iters, rs.status = _abi.ptr(planned), _abi.ptr(iters), _abi.ptr(status)
        _check(lib().ccc_zmp_mpc_plan(self._h, C.addressof(bt), C.addressof(rs), _abi.CCC_MEM_HOST, None), "ccc_zmp_mpc_plan")
        self._uploaded = True
        return planned, iters, status

    @property
    def last_launches(self):
        return int(lib().ccc_zmp_mpc_last_launches(self._h))


class LinearMpcXyEngine:
    """Device pipeline of CCC::LinearMpcXY::planOnce over a sweep of schedules (ccc_linear_mpc_xy_*): stage models,
    closed-form discretisation, condensing, the tensor-core B_seq' W B_seq, per-problem vectors and the QP, with one
    QP factorisation per schedule.  One workspace per QP shape (n, n_eq)."""

    def __init__(self, horizon_steps, n, n_eq, max_batch, max_sched):
        self._h = lib().ccc_linear_mpc_xy_create(int(horizon_steps), int(n), int(n_eq), int(max_batch), int(max_sched))
        if not self._h:
            raise EngineError(f"ccc_linear_mpc_xy_create failed: {last_error()}")

    def close(self):
        if getattr(self, "_h", None):
            lib().ccc_linear_mpc_xy_destroy(self._h)
            self._h = None

    __del__ = close

    def solve(self, sweep, intermediates=False, result=None):
        """Host buffers in / out for a linear_mpc_xy.XySweepProblemSet."""
        res = result if result is not None else sweep.new_result(intermediates)
        bs, rs = sweep.as_struct(), res.as_struct()
        _check(lib().ccc_linear_mpc_xy_solve(self._h, C.addressof(bs), C.addressof(rs), _abi.CCC_MEM_HOST, None),
               "ccc_linear_mpc_xy_solve")
        return res

    def solve_device(self, batch_struct, result_struct, stream=0):
        _check(lib().ccc_linear_mpc_xy_solve(self._h, C.addressof(batch_struct), C.addressof(result_struct),
                                             _abi.CCC_MEM_DEVICE, C.c_void_p(stream)), "ccc_linear_mpc_xy_solve")

    @property
    def last_launches(self):
        return int(lib().ccc_linear_mpc_xy_last_launches(self._h))


def qp_solver_for(engines=None):
    """-> qp_solve(QpProblemSet) callable backed by QpEngine workspaces cached per shape."""
    engines = {} if engines is None else engines

    def qp_solve(ps):
        key = (ps.n, ps.n_eq, ps.n_ineq)
        eng = engines.get(key)
        if eng is None or eng.max_batch < ps.batch:
            eng = engines[key] = QpEngine(ps.n, ps.n_eq, ps.n_ineq, max(ps.batch, 1))
        return eng.solve(ps)

    return qp_solve


def preview_gemv(K, F, x, ref_seq):
    """jerk[b] = -K x[b] + F ref_seq[b] on the GPU (ccc_preview_input, host buffers); plugs into
    preview_control.PreviewControlZmp1d.proc_once(gemv=...)."""
    import numpy as np

    K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1)
    F = np.ascontiguousarray(F, dtype=np.float64).reshape(-1)
    x = np.ascontiguousarray(x, dtype=np.float64)
    ref_seq = np.ascontiguousarray(ref_seq, dtype=np.float64)
    u = np.zeros(len(x))
    _check(lib().ccc_preview_input(len(x), len(F), K.ctypes.data, F.ctypes.data, x.ctypes.data, ref_seq.ctypes.data,
                                   u.ctypes.data, _abi.CCC_MEM_HOST, None), "ccc_preview_input")
    return u
