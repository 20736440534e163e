"""GPU tier: ccc_qp_solve (CUDA) vs the oracle — solutions, iteration counts and active sets bit-exact —
on configs 2 (LinearMpcZmp) and 5 (IntrinsicallyStableMpc), plus the reference's closed-loop tests."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import linear_mpc, workloads
from centroidalcontrolcollection_b200.qp import QpProblemSet

from test_linear_mpc_cpu import _closed_loop

pytestmark = pytest.mark.gpu
FIELDS = ("x", "iters", "status", "n_active", "active")


@pytest.fixture(scope="module")
def qp_solve():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0
    return engine.qp_solver_for()


def _parity(oracle, qp_solve, ps, rel_tol=1e-6):
    ref = oracle.qp_solve(ps, n_threads=max(1, oracle.hardware_threads()))
    got = qp_solve(ps)
    assert np.array_equal(ref.status, got.status) and np.array_equal(ref.iters, got.iters)
    assert ref.active_sets() == got.active_sets()
    assert np.abs(ref.x - got.x).max() <= rel_tol * max(1.0, np.abs(ref.x).max())
    for f in FIELDS:
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f
    return ref


def test_random_qps(oracle, qp_solve):
    rng = np.random.default_rng(1)
    n, mi, B = 40, 100, 300
    M = rng.standard_normal((n, n))
    ps = QpProblemSet(M @ M.T + np.eye(n), rng.standard_normal((mi, n)), rng.uniform(0.05, 0.6, (B, mi)),
                      rng.standard_normal((3, n)), 0.1 * rng.standard_normal((B, 3)), 4 * rng.standard_normal((B, n)))
    ref = _parity(oracle, qp_solve, ps)
    viol, dual = zip(*ps.subset(np.arange(20)).kkt_residuals(ref.x[:20]))
    assert max(viol) < 1e-9 and max(dual) < 1e-8


def test_config2_linear_mpc_zmp(oracle, qp_solve):
    """Config 2 at full size: 4096 ICs x 2 axes = 8192 QPs (n = 100, 200 inequalities)."""
    w = workloads.linear_mpc_zmp_config2()
    mpc = linear_mpc.LinearMpcZmp(w["com_height"], w["horizon_duration"], w["horizon_dt"])
    zmp = mpc.plan_batch(qp_solve, w["pos"], w["vel"], w["acc"], w["lim_min"], w["lim_max"], w["control_dt"])
    got = mpc.mpc_1d.last_result
    assert (got.status == 0).all()
    assert (zmp >= w["lim_min"][:, 0] - 1e-12).all() and (zmp <= w["lim_max"][:, 0] + 1e-12).all()
    zmp_o = mpc.plan_batch(lambda ps: oracle.qp_solve(ps, n_threads=max(1, oracle.hardware_threads())), w["pos"], w["vel"],
                           w["acc"], w["lim_min"], w["lim_max"], w["control_dt"])
    ref = mpc.mpc_1d.last_result
    for f in FIELDS:
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f
    assert np.array_equal(zmp, zmp_o)


def test_config5_ismpc_full_size(oracle, qp_solve):
    """Config 5 at full size: 256 plans x 512 perturbations x 2 axes = 262144 QPs with the equality constraint, every
    one of them against the oracle."""
    w = workloads.ismpc_config5()
    mpc = linear_mpc.IntrinsicallyStableMpc(w["com_height"], w["horizon_duration"], w["horizon_dt"])
    args = (w["capture_point"], w["planned_zmp"], w["ref_zmp"], w["lim_min"], w["lim_max"], w["control_dt"])
    zmp = mpc.plan_batch(qp_solve, *args)
    got = mpc.mpc_1d.last_result
    zmp_o = mpc.plan_batch(lambda ps: oracle.qp_solve(ps, n_threads=max(1, oracle.hardware_threads())), *args)
    ref = mpc.mpc_1d.last_result
    for f in FIELDS:
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f
    assert np.array_equal(zmp, zmp_o)
    assert (got.status == 0).mean() > 0.99


def test_linear_mpc_xy_batch(oracle, qp_solve):
    """LinearMpcXY at the reference's size (n = 240, 15 equalities, 480 bound rows; 256-thread CTAs with J/R in an
    L2-resident slab) and at n = 128 (128-thread CTAs with the slab): bit-exact incl. active sets."""
    from test_emu_qp import _xy_problem_set

    for horizon_steps, batch in ((15, 160), (8, 200)):
        ps = _xy_problem_set(horizon_steps, batch)
        ref = _parity(oracle, qp_solve, ps)
        assert (ref.status == 0).all()


def test_linear_mpc_xy_closed_loop(qp_solve):
    """reference tests/src/TestLinearMpcXY.cpp through the engine, with the reference's tolerances."""
    from test_linear_mpc_xy_cpu import xy_closed_loop

    sim, ok, ref_pos, iters = xy_closed_loop(qp_solve)
    assert ok
    assert np.linalg.norm(sim.pos - ref_pos) < 0.1
    assert np.linalg.norm(sim.vel) < 0.1
    assert np.linalg.norm(sim.angular_momentum) < 0.1


def test_closed_loops(qp_solve):
    """reference tests/src/TestLinearMpcZmp.cpp and TestIntrinsicallyStableMpc.cpp through the engine."""
    for kind in ("lmpc", "ismpc"):
        ok, planned, pos, zl = _closed_loop(kind, qp_solve, end_time=10.0)
        assert ok
        assert (planned - zl[0] >= 0).all() and (zl[1] - planned >= 0).all()
        assert (pos - zl[0] >= 0).all() and (zl[1] - pos >= 0).all()


def test_matrix_reuse_and_unpaired_rows(oracle):
    """Q == NULL reuses the matrices and the factorisation of the previous call on the workspace (what a
    controller's per-tick calls do); a constraint matrix that is not of the [-G; G] form takes the unpaired search."""
    import ctypes as C

    from centroidalcontrolcollection_b200 import _abi, engine

    rng = np.random.default_rng(7)
    n, mi, B = 24, 37, 96  # odd row count: cannot be paired
    M = rng.standard_normal((n, n))
    ps = QpProblemSet(M @ M.T + np.eye(n), rng.standard_normal((mi, n)), rng.uniform(0.05, 0.6, (B, mi)), None, None,
                      2 * rng.standard_normal((B, n)))
    eng = engine.QpEngine(n, 0, mi, B)
    first = eng.solve(ps.subset(np.arange(0, 48)))
    second_ps = ps.subset(np.arange(48, 96))
    res = second_ps.new_result()
    bs, rs = second_ps.as_struct(), res.as_struct()
    bs.Q = bs.A = bs.C = None
    rc = engine.lib().ccc_qp_solve(eng._h, C.addressof(bs), C.addressof(rs), _abi.CCC_MEM_HOST, None)
    assert rc == 0
    ref = oracle.qp_solve(ps, n_threads=max(1, oracle.hardware_threads()))
    for f in FIELDS:
        assert np.array_equal(getattr(ref, f)[:48], getattr(first, f)), f
        assert np.array_equal(getattr(ref, f)[48:], getattr(res, f)), f
    # a fresh workspace has nothing to reuse
    eng2 = engine.QpEngine(n, 0, mi, B)
    rc = engine.lib().ccc_qp_solve(eng2._h, C.addressof(bs), C.addressof(rs), _abi.CCC_MEM_HOST, None)
    assert rc != 0


def test_workspaces_of_several_shapes_coexist(oracle):
    """A controller whose QP size changes from tick to tick (LinearMpcZ: one variable per contact stage of the horizon)
    keeps one workspace per shape; the shared-memory attribute of the kernels must not follow the last one created."""
    from centroidalcontrolcollection_b200 import engine

    rng = np.random.default_rng(9)

    def make(n):
        M = rng.standard_normal((n, n))
        return QpProblemSet(M @ M.T + np.eye(n), np.vstack([-np.eye(n), np.eye(n)]), np.tile(np.full(2 * n, 0.3), (4, 1)), None, None,
                            2 * rng.standard_normal((4, n)))

    big, small = make(96), make(12)
    e_big = engine.QpEngine(big.n, 0, big.n_ineq, 4)
    e_small = engine.QpEngine(small.n, 0, small.n_ineq, 4)  # created later, needs far less shared memory
    for ps, eng in ((big, e_big), (small, e_small), (big, e_big)):
        got, ref = eng.solve(ps), oracle.qp_solve(ps)
        for f in FIELDS:
            assert np.array_equal(getattr(ref, f), getattr(got, f)), f


def test_matrix_groups(oracle):
    """ccc_qp_solve_grouped: every problem uses the Q / A of its group; bit-exact against the oracle group by group."""
    from centroidalcontrolcollection_b200 import engine
    from centroidalcontrolcollection_b200.qp import QpGroupedProblemSet

    rng = np.random.default_rng(12)
    n, me, mi, G, B = 20, 2, 40, 37, 500
    Q = np.zeros((G, n, n))
    for g in range(G):
        M = rng.standard_normal((n, n))
        Q[g] = M @ M.T + np.eye(n)
    gp = QpGroupedProblemSet(Q, np.vstack([-np.eye(n), np.eye(n)]), np.tile(np.full(mi, 0.4), (B, 1)), rng.integers(0, G, B),
                             rng.standard_normal((G, me, n)), 0.05 * rng.standard_normal((B, me)), 3 * rng.standard_normal((B, n)))
    eng = engine.QpEngine(n, me, mi, B, max_groups=G)
    got = eng.solve_grouped(gp)
    ref = gp.solve_by_group(lambda ps: oracle.qp_solve(ps))
    for f in FIELDS:
        assert np.array_equal(getattr(ref, f), getattr(got, f)), f
    assert (got.status == 0).all() and (got.n_active > me).any()
    with pytest.raises(engine.EngineError):
        engine.QpEngine(n, me, mi, B, max_groups=3).solve_grouped(gp)


def test_full_active_set_takes_the_dual_step(oracle, qp_solve):
    """q = n with a violated row left: the dual-only step of Goldfarb-Idnani; see tests/test_emu_qp.py for the
    two-variable case worked out."""
    Q = np.eye(2)
    C = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    ps = QpProblemSet(Q, C, np.array([[1.0, 1.0, 1.5], [1.0, 1.0, 1.2]]), c=np.array([[-3.0, -3.0], [-3.0, -2.0]]))
    ref = _parity(oracle, qp_solve, ps)
    assert (ref.status == 0).all() and np.allclose(ref.x[0], [0.75, 0.75], atol=1e-12)
    rng = np.random.default_rng(7)
    n, mi, B = 6, 40, 300
    Cn = rng.standard_normal((mi, n))
    ps2 = QpProblemSet(np.eye(n), Cn / np.linalg.norm(Cn, axis=1)[:, None], rng.uniform(0.05, 0.3, (B, mi)), c=-4 * rng.standard_normal((B, n)))
    ref = _parity(oracle, qp_solve, ps2)
    assert (ref.status == 0).all()
    viol, dual = zip(*ps2.kkt_residuals(ref.x))
    assert max(viol) < 1e-9 and max(dual) < 1e-8
