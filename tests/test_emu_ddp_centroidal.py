"""CPU tier: the CUDA solver-core source, executed by the lock-step warp emulator, reproduces
the oracle bit for bit (small cases: the emulator is ~1000x slower than the oracle)."""
import numpy as np

from centroidalcontrolcollection_b200 import problem, workloads

import emu_lib
from parity import assert_ddp_parity


def _run(oracle, ps, cfg, trace_len=16, chunk=0, feats=(1, 2)):
    """feats: builds of the solver core to emulate (1 = product default, 2 = every feature bit incl. the TMA-staged
    gain lists, 0 = none); each must reproduce the oracle bit for bit."""
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=trace_len)
    for feat in feats:
        got = emu_lib.ddp_centroidal_solve(ps, cfg, trace_len=trace_len, chunk=chunk, feat=feat)
        assert_ddp_parity(ref, got)
    return ref


def test_short_horizon_to_convergence(oracle):
    w = workloads.ddp_centroidal_config3(batch=3, horizon_steps=6)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    ref = _run(oracle, ps, problem.ddp_centroidal_config())
    assert (ref.status == 1).all()


def test_suspend_and_resume_is_invisible(oracle):
    """Round-robin time slicing: suspending every 1 or 2 iterations must not change a single bit."""
    w = workloads.ddp_centroidal_config3(batch=3, horizon_steps=6)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    for chunk in (1, 2):
        ref = _run(oracle, ps, problem.ddp_centroidal_config(), chunk=chunk)
        assert ref.iters.max() > 2


def test_all_phase_kinds_few_iterations(oracle):
    """N=50 covers m = 16, 0 (flight) and 32 stages, dimension changes and BoxQP clamping."""
    w = workloads.ddp_centroidal_config3(batch=2, horizon_steps=50)
    ps = problem.DdpCentroidalProblemSet.from_workload(w).subset([0, 1])
    ref = _run(oracle, ps, problem.ddp_centroidal_config(max_iter=3), feats=(0, 1, 2))
    assert set(np.unique(ps.sched.m)) == {0, 16, 32}
    assert (ref.clamped != 0).any()


def test_warm_start_and_ragged_m_max(oracle):
    """u_init given, m_max = 16 (row stride != 32), max_iter = 1 as in the closed-loop test."""
    sched, _, _ = workloads.ddp_centroidal_test_schedule(horizon_steps=12, dt=0.03, current_time=1.3)
    sched16 = type(sched)(1, 12, m_max=16)
    sched16.m[:] = sched.m
    sched16.ridge[:] = sched.ridge[:, :, :16]
    sched16.vertex[:] = sched.vertex[:, :, :16]
    sched16.ref_pos[:] = sched.ref_pos
    w_run, w_term = workloads.centroidal_weights_test()
    rng = np.random.default_rng(5)
    x0 = np.array([[0.01, -0.02, 1.0, 3.0, -2.0, 1.0, 0.1, 0.2, -0.1]])
    u_init = rng.uniform(0, 80, size=(1, 12, 16)) * (np.arange(16)[None, None, :] < sched16.m[0][None, :, None])
    ps = problem.DdpCentroidalProblemSet(sched16, [0], x0, 100.0, 0.03, w_run, w_term, u_init=u_init)
    _run(oracle, ps, problem.ddp_centroidal_config(max_iter=1))
    _run(oracle, ps, problem.ddp_centroidal_config(max_iter=4))


def test_unconstrained_path(oracle):
    """with_input_constraint = 0: plain LLT gains (the path DdpZmp-style problems take)."""
    w = workloads.ddp_centroidal_config3(batch=1, horizon_steps=8)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config(max_iter=3)
    cfg.with_input_constraint = 0
    _run(oracle, ps, cfg)


def test_indefinite_quu_takes_the_regularisation_retry_path(oracle):
    """A negative force weight makes Quu + lambda I indefinite at the initial lambda (Fu' Vxx Fu has rank <= 6 of
    16 inputs): the factorisation must report the non-positive pivot, the backward pass is retried with a larger
    lambda until it goes through — same retries, same bits on both sides."""
    w = workloads.ddp_centroidal_config3(batch=2, horizon_steps=6)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    ps.w_run = ps.w_run.copy()
    ps.w_run[9] = -1e-4
    for constrained in (1, 0):
        cfg = problem.ddp_centroidal_config(max_iter=3)
        cfg.with_input_constraint = constrained
        ref = _run(oracle, ps, cfg)
        assert (ref.lambda_trace[:, 0] > 1e-4).all(), "lambda was not raised: the retry path did not run"
