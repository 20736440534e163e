"""The oracle's canonical arithmetic against the same algorithm in textbook arithmetic — CPU only.

The oracle (and with it the engine) departs from a straightforward Eigen/libm program in a few places for the GPU's
sake: explicit fma, tree / dot4 sums, Cholesky as L D L' without square roots, cross-multiplied Armijo
test, one-sum BoxQP objective, squared gradient-norm test, P * (1/m), fma-only sin/cos, parallel-form Givens sweep
and reciprocal diagonal in the dual active-set QP.  oracle/Makefile builds the SAME sources a second time with
-DORACLE_TEXTBOOK (oracle/num.hpp), where every one of these is written the textbook way.  These tests pin the
canonical build to the textbook one: same iteration path on (almost) every problem, trajectories equal far below the
1e-6 bar wherever the path is the same, same minimum everywhere.

What they also record: the DDP termination test (`cost_update < 1e-7`) decides the iteration count on a quantity
that rounding can move across the threshold, so on a fraction of a percent of the problems the two builds stop one
or a few iterations apart and the trajectories then differ at the solver's own tolerance (1e-3 in the inputs, with
costs equal to 1e-9) — "bit-exact iteration counts" are a property of one arithmetic, not of the algorithm.  For
DdpSingleRigidBody's cold starts (config 4) the iteration path is chaotic from the second iteration on (the first
full step from the free-fall rollout tumbles through the Euler-angle singularity), so only the
well-conditioned warm-started case is compared there.
"""
import numpy as np

from centroidalcontrolcollection_b200 import linear_mpc, problem, workloads


def _threads(oracle):
    return max(1, oracle.hardware_threads())


def test_textbook_build_is_the_textbook_build(oracle):
    assert oracle.textbook().lib().ccc_oracle_is_textbook() == 1
    assert oracle.lib().ccc_oracle_is_textbook() == 0


def test_ddp_centroidal_config3_same_iteration_path(oracle):
    """Config 3 (1024 problems of the 16384, every schedule variant): iteration counts, accepted step indices and
    BoxQP clamped sets identical on >= 99 % of the problems; where they are, states agree to 1e-8 (measured
    1.4e-10); the minimum found agrees to 1e-8 relative in cost on every problem."""
    w = workloads.ddp_centroidal_config3(batch=1024, horizon_steps=50)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config()
    a = oracle.ddp_centroidal_solve(ps, cfg, trace_len=64, n_threads=_threads(oracle))
    b = oracle.textbook().ddp_centroidal_solve(ps, cfg, trace_len=64, n_threads=_threads(oracle))
    assert (a.status == 1).all() and (b.status == 1).all()
    same = (a.iters == b.iters) & (a.alpha_idx == b.alpha_idx).all(1) & (a.clamped == b.clamped).all(1)
    assert same.mean() >= 0.99, same.mean()
    scale = np.abs(a.x).max((1, 2))
    assert (np.abs(a.x - b.x).max((1, 2)) / scale)[same].max() < 1e-8
    assert (np.abs(a.u - b.u).max((1, 2)) / np.abs(a.u).max((1, 2)))[same].max() < 1e-8
    assert (np.abs(a.cost - b.cost) / a.cost).max() < 1e-8
    # the problems that stop a few iterations apart end at the same minimum within the solver's own tolerance
    assert np.abs(a.iters - b.iters).max() <= 8
    assert (np.abs(a.x[:, :, :3] - b.x[:, :, :3]).max((1, 2))).max() < 1e-3


def test_ddp_centroidal_reference_closed_loop_in_textbook_arithmetic(oracle):
    """The reference's DdpCentroidal closed-loop test (tests/src/TestDdpCentroidal.cpp:15-156) passes in textbook
    arithmetic too, and the two closed loops stay together to 1e-6 over all 600 ticks."""
    from closed_loop import run_ddp_centroidal_closed_loop

    rec_a, rec_b = [], []
    tb = oracle.textbook()
    sim_a, _, rp, ok_a, it_a = run_ddp_centroidal_closed_loop(lambda ps, c: oracle.ddp_centroidal_solve(ps, c), record=rec_a)
    sim_b, _, _, ok_b, it_b = run_ddp_centroidal_closed_loop(lambda ps, c: tb.ddp_centroidal_solve(ps, c), record=rec_b)
    assert ok_a and ok_b and it_a == it_b
    for sim in (sim_a, sim_b):
        assert np.linalg.norm(sim.pos - rp) < 0.1 and np.linalg.norm(sim.vel) < 0.1
        assert np.linalg.norm(sim.angular_momentum) < 0.01
    pos_a, pos_b = np.array([r[1] for r in rec_a]), np.array([r[1] for r in rec_b])
    assert np.abs(pos_a - pos_b).max() < 1e-6


def test_ddp_srb_warm_started_same_iteration_path(oracle):
    """DdpSingleRigidBody where the iteration path is well conditioned: from the converged plan of the reference's
    first tick, perturbed initial states, warm-started — the regime of every tick after the first."""
    sched, _, _ = workloads.ddp_srb_test_schedule(100, 0.03, 0.0)
    w_run, w_term = workloads.srb_weights_test()
    x0 = np.zeros((1, 12))
    x0[0, 2] = 1.0
    cfg = problem.ddp_srb_config()
    first = oracle.ddp_srb_solve(problem.DdpSrbProblemSet(sched, [0], x0, 100.0, 0.03, w_run, w_term), cfg)
    assert first.status[0] == 1
    rng = np.random.Generator(np.random.PCG64(11))
    B = 64
    xs = np.tile(x0, (B, 1)) + 1e-3 * rng.standard_normal((B, 12))
    ps = problem.DdpSrbProblemSet(sched, np.zeros(B, dtype=np.int32), xs, 100.0, 0.03, w_run, w_term,
                                  u_init=np.tile(first.u, (B, 1, 1)))
    cfg3 = problem.ddp_srb_config(max_iter=3)
    a = oracle.ddp_srb_solve(ps, cfg3, trace_len=4, n_threads=_threads(oracle))
    b = oracle.textbook().ddp_srb_solve(ps, cfg3, trace_len=4, n_threads=_threads(oracle))
    same = (a.iters == b.iters) & (a.alpha_idx == b.alpha_idx).all(1)
    assert same.mean() >= 0.95, same.mean()
    assert (np.abs(a.cost - b.cost) / a.cost)[same].max() < 1e-6


def test_qp_config2_same_active_sets(oracle):
    """LinearMpcZmp (config 2, 1024 of the 8192 QPs): identical iteration counts and active sets, solutions to
    1e-10 — the parallel-form Givens sweep, the reciprocal diagonal and the tree sums change rounding only."""
    w2 = workloads.linear_mpc_zmp_config2(batch=512)
    mpc = linear_mpc.LinearMpcZmp(w2["com_height"], w2["horizon_duration"], w2["horizon_dt"])
    args = (w2["pos"], w2["vel"], w2["acc"], w2["lim_min"], w2["lim_max"], w2["control_dt"])
    tb = oracle.textbook()
    z1 = mpc.plan_batch(lambda q: oracle.qp_solve(q, n_threads=_threads(oracle)), *args)
    r1 = mpc.mpc_1d.last_result
    z2 = mpc.plan_batch(lambda q: tb.qp_solve(q, n_threads=_threads(oracle)), *args)
    r2 = mpc.mpc_1d.last_result
    assert np.array_equal(r1.iters, r2.iters) and r1.active_sets() == r2.active_sets()
    assert np.abs(r1.x - r2.x).max() < 1e-10 * max(1.0, np.abs(r1.x).max())
    assert np.abs(z1 - z2).max() < 1e-12


def test_qp_config5_same_active_sets(oracle):
    """IntrinsicallyStableMpc (config 5 sample: 16 plans x 32 perturbations, one equality + 200 inequalities)."""
    w5 = workloads.ismpc_config5(n_plans=16, n_perturb=32)
    mpc = linear_mpc.IntrinsicallyStableMpc(w5["com_height"], w5["horizon_duration"], w5["horizon_dt"])
    tb = oracle.textbook()
    args = (w5["capture_point"], w5["planned_zmp"], w5["ref_zmp"], w5["lim_min"], w5["lim_max"], w5["control_dt"])
    z1 = mpc.plan_batch(lambda q: oracle.qp_solve(q, n_threads=_threads(oracle)), *args)
    r1 = mpc.mpc_1d.last_result
    z2 = mpc.plan_batch(lambda q: tb.qp_solve(q, n_threads=_threads(oracle)), *args)
    r2 = mpc.mpc_1d.last_result
    assert np.array_equal(r1.iters, r2.iters) and r1.active_sets() == r2.active_sets()
    assert np.abs(z1 - z2).max() < 1e-12
