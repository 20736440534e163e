"""GPU tier: ccc_linear_mpc_xy_solve (stage models, closed-form discretisation, condensing, tensor-core
B_seq' W B_seq, per-problem vectors, one QP factorisation per schedule) vs oracle/xy.hpp, stage by stage."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import workloads
from centroidalcontrolcollection_b200.qp import QpProblemSet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine_mod():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0
    return engine


def _solve(engine_mod, sweep):
    eng = engine_mod.LinearMpcXyEngine(sweep.N, sweep.n, sweep.n_eq, sweep.batch, sweep.S)
    got = eng.solve(sweep, intermediates=True)
    launches = eng.last_launches
    eng.close()
    return got, launches


def test_xy_sweep_parity(oracle, engine_mod):
    """64 schedules x 16 initial states at the reference's size (n = 240, 15 equalities, 480 bound rows)."""
    sweep = workloads.linear_mpc_xy_sweep(n_sched=64, per_sched=16)
    ref = oracle.linear_mpc_xy_solve(sweep, n_threads=max(1, oracle.hardware_threads()))
    got, launches = _solve(engine_mod, sweep)
    assert launches >= 5  # condense, hessian, gradient, QP setup, QP solve
    # condensing: same fma chains -> same bits
    assert np.array_equal(ref.A_seq, got.A_seq)
    assert np.array_equal(ref.B_seq, got.B_seq)
    # obj_mat on the FP64 tensor cores: the oracle's chain is sequential over the rows of B_seq; the tolerance covers
    # a different accumulation order inside one k = 4 step, the print says which it is on this device
    scale = np.abs(ref.obj_mat).max()
    dH = np.abs(ref.obj_mat - got.obj_mat).max()
    print(f"obj_mat: max |dH| = {dH:.3e} (scale {scale:.3e}), bit-exact: {np.array_equal(ref.obj_mat, got.obj_mat)}")
    assert dH <= 1e-14 * scale
    assert np.array_equal(ref.obj_vec, got.obj_vec)
    assert (got.status == 0).all() and (ref.status == 0).all()
    if np.array_equal(ref.obj_mat, got.obj_mat):
        for f in ("u", "iters", "status", "n_active", "active"):
            assert np.array_equal(getattr(ref, f), getattr(got, f)), f
    else:
        # the QP stage alone, on the engine's own matrices: bit-exact against the oracle QP
        for s in (0, sweep.S // 2, sweep.S - 1):
            idx = np.where(sweep.sched_id == s)[0]
            ps0 = sweep.host_problem(s, idx)
            ps = QpProblemSet(got.obj_mat[s], ps0.C, ps0.d, ps0.A, ps0.b, got.obj_vec[idx])
            r = oracle.qp_solve(ps)
            assert np.array_equal(r.x, got.u[idx]) and np.array_equal(r.iters, got.iters[idx])
            assert np.array_equal(r.active, got.active[idx])
        assert np.abs(ref.u - got.u).max() <= 1e-6 * np.abs(ref.u).max()
        assert ref.active_sets() == got.active_sets()
    # what planOnce returns: the first stage's force scales; their vertical sum is the stage's total force
    fz = (got.u[:, :16] * sweep.ridge[sweep.sched_id, 0, :, 2]).sum(axis=1)
    assert np.abs(fz - sweep.total_force_z[sweep.sched_id, 0]).max() < 1e-6 * 1000.0
    assert (got.u >= sweep.mpc.force_range[0] - 1e-9).all() and (got.u <= sweep.mpc.force_range[1] + 1e-9).all()


def test_xy_sweep_small_horizon_and_errors(oracle, engine_mod):
    """n = 128 (8 stages): the 128-thread QP kernel; a schedule of another shape is refused."""
    sweep = workloads.linear_mpc_xy_sweep(n_sched=5, per_sched=7, horizon_steps=8)
    ref = oracle.linear_mpc_xy_solve(sweep, n_threads=4)
    got, _ = _solve(engine_mod, sweep)
    assert np.array_equal(ref.B_seq, got.B_seq) and np.array_equal(ref.obj_vec, got.obj_vec)
    assert np.abs(ref.obj_mat - got.obj_mat).max() <= 1e-14 * np.abs(ref.obj_mat).max()
    assert np.abs(ref.u - got.u).max() <= 1e-6 * np.abs(ref.u).max()
    assert ref.active_sets() == got.active_sets()
    sweep.m[3, 2] = 0  # a flight stage: n = 112, 7 equalities
    eng = engine_mod.LinearMpcXyEngine(sweep.N, 128, 8, sweep.batch, sweep.S)
    with pytest.raises(engine_mod.EngineError):
        eng.solve(sweep)
    eng.close()
