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
