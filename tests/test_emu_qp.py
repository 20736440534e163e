"""CPU tier: the CUDA QP core (qp_cta_core.cuh) executed by the 128-fibre CTA emulator reproduces the
oracle bit for bit, incl. equality constraints, constraint drops and the LinearMpcZmp / IS-MPC structure."""
import numpy as np

from centroidalcontrolcollection_b200 import linear_mpc, workloads
from centroidalcontrolcollection_b200.qp import QpProblemSet

import emu_lib

FIELDS = ("x", "iters", "status", "n_active", "active")


def _same(oracle, ps, rcaps=(0,)):
    """rcaps: packed-R capacities of the first pass to emulate (0 = the full-R kernel alone)."""
    ref = oracle.qp_solve(ps)
    for rcap in rcaps:
        got = emu_lib.qp_solve(ps, rcap=rcap)
        for f in FIELDS:
            assert np.array_equal(getattr(ref, f), getattr(got, f)), (f, rcap)
    return ref


def ref_cap_tight(oracle, ps):
    """A packed-R capacity that some but not all problems of the batch outgrow (peak active-set size is at least the
    final one)."""
    na = oracle.qp_solve(ps).n_active
    return max(int(np.median(na)), ps.n_eq + 1)


def test_random_qps_with_equalities_and_drops(oracle):
    rng = np.random.default_rng(0)
    n, mi, B = 24, 60, 8
    M = rng.standard_normal((n, n))
    ps = QpProblemSet(M @ M.T + np.eye(n), rng.standard_normal((mi, n)), rng.uniform(0.05, 0.6, (B, mi)),
                      rng.standard_normal((2, n)), 0.1 * rng.standard_normal((B, 2)), 4 * rng.standard_normal((B, n)))
    # packed R: roomy (no overflow), and so tight that some problems overflow into the full-R pass
    ref = _same(oracle, ps, rcaps=(0, 24, int(ref_cap_tight(oracle, ps))))
    assert emu_lib.last_qp_overflows() > 0
    assert (ref.status == 0).all()
    assert (ref.iters > ref.n_active - 2).any()  # at least one problem dropped a constraint on the way
    viol, dual = zip(*ps.kkt_residuals(ref.x))
    assert max(viol) < 1e-9 and max(dual) < 1e-8


def test_linear_mpc_zmp_structure(oracle):
    w = workloads.linear_mpc_zmp_config2(batch=3)
    mpc = linear_mpc.LinearMpcZmp1d(w["com_height"], 0.3, w["horizon_dt"])  # N = 30 keeps the emulator quick
    N = mpc.horizon_steps
    ip = np.stack([w["pos"][:, 0], w["vel"][:, 0], w["acc"][:, 0]], axis=1)
    lim = np.stack([w["lim_min"][:, :N, 0], w["lim_max"][:, :N, 0]], axis=2)
    _same(oracle, mpc.build_qp(ip, lim), rcaps=(0, 20, 4))


def test_ismpc_structure(oracle):
    w = workloads.ismpc_config5(n_plans=1, n_perturb=3)
    mpc = linear_mpc.IntrinsicallyStableMpc1d(w["com_height"], 0.5, w["horizon_dt"])  # N = 25
    N = mpc.horizon_steps
    ps = mpc.build_qp(w["capture_point"][:, 1], w["planned_zmp"][:, 1], w["ref_zmp"][:, :N, 1],
                      np.stack([w["lim_min"][:, :N, 1], w["lim_max"][:, :N, 1]], axis=2))
    ref = _same(oracle, ps, rcaps=(0, 16, 3))
    assert (ref.status == 0).all() and (ref.n_active >= 1).all()


def _xy_problem_set(horizon_steps, batch):
    return workloads.linear_mpc_xy_problem_set(horizon_steps, batch)


def test_linear_mpc_xy_structure_128_threads_global_slab(oracle):
    """n = 128 (8 stages x 16 ridges): J and R no longer fit in shared memory, 128-thread CTA."""
    ps = _xy_problem_set(8, 2)
    assert ps.n == 128 and ps.n_eq == 8 and ps.n_ineq == 256
    ref = _same(oracle, ps)
    assert (ref.status == 0).all()


def test_linear_mpc_xy_structure_256_threads(oracle):
    """The reference's size: n = 240 (15 x 16), 15 equalities, 480 bound rows; 256-thread CTA, 256-leaf trees."""
    ps = _xy_problem_set(15, 1)
    assert ps.n == 240 and ps.n_eq == 15 and ps.n_ineq == 480
    ref = _same(oracle, ps)
    assert (ref.status == 0).all() and ref.iters[0] > 20
    viol, dual = zip(*ps.kkt_residuals(ref.x, tol=1e-7))
    assert max(viol) < 1e-7


def test_full_active_set_takes_the_dual_step(oracle):
    """n = 2: min |x - (3, 3)|^2 s.t. x <= 1, y <= 1, x + y <= 1.5.  The dual active-set method first makes both box rows
    active (q = n) and then meets the violated diagonal row: Goldfarb-Idnani takes a dual-only step there (z is the empty
    sum) and drops a box row; the solution is (0.75, 0.75) with only the diagonal row active.  Same path, same bits in the
    oracle and the CUDA core; a larger random family with more rows than a vertex can hold as well."""
    Q = np.eye(2)
    C = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    d = np.array([[1.0, 1.0, 1.5], [1.0, 1.0, 1.2]])
    c = np.array([[-3.0, -3.0], [-3.0, -2.0]])
    ps = QpProblemSet(Q, C, d, c=c)
    ref = _same(oracle, ps)
    assert (ref.status == 0).all()
    assert np.allclose(ref.x[0], [0.75, 0.75], atol=1e-12)
    viol, dual = zip(*ps.kkt_residuals(ref.x))
    assert max(viol) < 1e-12 and max(dual) < 1e-12
    rng = np.random.default_rng(7)
    n, mi, B = 6, 40, 12
    Cn = rng.standard_normal((mi, n))
    ps = QpProblemSet(np.eye(n), Cn / np.linalg.norm(Cn, axis=1)[:, None], rng.uniform(0.05, 0.3, (B, mi)), c=-4 * rng.standard_normal((B, n)))
    ref = _same(oracle, ps)
    assert (ref.status == 0).all()
    viol, dual = zip(*ps.kkt_residuals(ref.x))
    assert max(viol) < 1e-9 and max(dual) < 1e-8
