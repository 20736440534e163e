"""GPU tier: CCC::DdpSingleRigidBody path (12 states) — engine vs oracle, bit-exact (the model's sin/cos
are the fma-only sincos_canon on both sides)."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import problem, workloads

from closed_loop_srb import run_ddp_srb_closed_loop
from parity import assert_ddp_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0, "no CUDA device: -m gpu tests need a B200"
    return engine


def test_parity_config4_sample(eng_mod, oracle):
    """512 problems of config 4 (N = 100, A -> flight -> B), cold start to termination (with the 1536 of
    test_full_shard_properties: 2048 of the shard's 8192 checked against the oracle)."""
    w = workloads.ddp_srb_config4(batch=512)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    cfg = problem.ddp_srb_config()
    eng = eng_mod.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
    got = eng.solve(ps, cfg, trace_len=200)
    ref = oracle.ddp_srb_solve(ps, cfg, trace_len=200, n_threads=max(1, oracle.hardware_threads()))
    assert_ddp_parity(ref, got, rel_tol=1e-6, bit_exact=True)
    assert (got.status == 1).mean() > 0.5  # cold starts that need > 500 iterations stop with status 0


def test_edge_cases(eng_mod, oracle):
    cfg = problem.ddp_srb_config()
    w = workloads.ddp_srb_config4(batch=5, horizon_steps=30)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    eng = eng_mod.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
    for mi in (0, 1, 7):
        c = problem.ddp_srb_config(max_iter=mi)
        assert_ddp_parity(oracle.ddp_srb_solve(ps, c, trace_len=8), eng.solve(ps, c, trace_len=8))
    # large angles and rates, flight stages in the horizon, full inertia matrix
    sched, _, _ = workloads.ddp_srb_test_schedule(horizon_steps=30, dt=0.03, current_time=1.0)
    sched.inertia[:] = np.array([[40.0, 2.0, -1.0], [2.0, 20.0, 0.5], [-1.0, 0.5, 10.0]]).reshape(9)
    w_run, w_term = workloads.srb_weights_test()
    rng = np.random.default_rng(3)
    x0 = np.array([0.1, -0.05, 1.05, 0.9, -0.6, 0.4, 0.2, -0.1, 0.1, 1.0, -2.0, 1.5]) + 0.1 * rng.standard_normal((7, 12))
    p2 = problem.DdpSrbProblemSet(sched, np.zeros(7, np.int32), x0, 100.0, 0.03, w_run, w_term)
    e2 = eng_mod.DdpSrbEngine(30, 7, 1)
    c = problem.ddp_srb_config(max_iter=12)
    assert_ddp_parity(oracle.ddp_srb_solve(p2, c, trace_len=12), e2.solve(p2, c, trace_len=12))


def test_plan_once_closed_loop(eng_mod, oracle):
    """reference tests/src/TestDdpSingleRigidBody.cpp:15-175 through the engine (see closed_loop_srb.py for
    the one documented deviation), reference tolerances, identical to the oracle-driven loop."""
    eng = eng_mod.DdpSrbEngine(100, 1, 1)
    sim, rp, ro, tick_ok, iters = run_ddp_srb_closed_loop(lambda ps, cfg: eng.solve(ps, cfg))
    assert tick_ok
    assert np.linalg.norm(sim.x[0:3] - rp) < 0.1
    assert np.linalg.norm(sim.x[3:6] - ro) < 0.1
    assert np.linalg.norm(sim.x[6:9]) < 0.1
    assert np.linalg.norm(sim.x[9:12]) < 0.1
    sim_o, _, _, _, iters_o = run_ddp_srb_closed_loop(lambda ps, cfg: oracle.ddp_srb_solve(ps, cfg))
    assert iters == iters_o
    assert np.array_equal(sim.x, sim_o.x)


def test_full_shard_properties(eng_mod, oracle):
    """One GPU's shard of config 4 (65536 / 8 = 8192 problems): properties + oracle spot check."""
    w = workloads.ddp_srb_config4(batch=8192)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    cfg = problem.ddp_srb_config()
    eng = eng_mod.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
    got = eng.solve(ps, cfg)
    assert np.isfinite(got.x).all() and np.isfinite(got.u).all()
    assert (got.u >= ps.u_lo).all() and (got.u <= ps.u_hi).all()
    assert np.array_equal(got.x[:, 0, :], ps.x0)
    assert (got.status == 1).mean() > 0.5  # cold starts that need > 500 iterations stop with status 0
    idx = np.sort(np.random.default_rng(2).choice(ps.batch, size=1536, replace=False))
    ref = oracle.ddp_srb_solve(ps.subset(idx), cfg, n_threads=max(1, oracle.hardware_threads()))
    assert np.array_equal(ref.iters, got.iters[idx])
    assert np.array_equal(ref.x, got.x[idx]) and np.array_equal(ref.u, got.u[idx])
