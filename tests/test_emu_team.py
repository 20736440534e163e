"""CPU tier: the small-batch DDP kernel (csrc/ddp_team.cuh — one CTA of 8 warps per problem, the line search run as
concurrent rollouts, the smallest accepted step index taken) executed by the lock-step emulator with 256 fibres.  It must
reproduce the oracle bit for bit, including which line-search index was accepted in every iteration, in the cases where
the serial loop would have needed one, several, more than eight (two rounds) or all eleven candidates."""
import numpy as np

from centroidalcontrolcollection_b200 import problem, workloads

import emu_lib
from parity import assert_ddp_parity


def test_team_centroidal_to_convergence(oracle):
    w = workloads.ddp_centroidal_config3(batch=3, horizon_steps=6)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config()
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=16)
    got = emu_lib.ddp_centroidal_solve(ps, cfg, trace_len=16, team=1)
    assert_ddp_parity(ref, got)
    assert (ref.status == 1).all()


def test_team_all_phase_kinds_and_shortened_steps(oracle):
    """N = 50 covers m = 16, 0 (flight) and 32 stages; the cold start needs shortened steps (accepted index > 0)."""
    w = workloads.ddp_centroidal_config3(batch=2, horizon_steps=50)
    ps = problem.DdpCentroidalProblemSet.from_workload(w).subset([0])
    cfg = problem.ddp_centroidal_config(max_iter=3)
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=4)
    got = emu_lib.ddp_centroidal_solve(ps, cfg, trace_len=4, team=1)
    assert_ddp_parity(ref, got)
    assert (ref.alpha_idx[:, :3] > 0).any(), "no shortened step in this case: pick another"


def test_team_second_round_and_failed_line_search(oracle):
    """Step sizes above 2 overshoot (the cost of a near-quadratic problem rises again past alpha = 2): with nine of them in
    front of the list the accepted index lies in the team's second round (>= 8); with only such step sizes every line
    search fails, lambda is raised and the solve ends at lambda_max or max_iter — same decisions and bits as the serial loop."""
    w = workloads.ddp_centroidal_config3(batch=2, horizon_steps=6)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config(max_iter=4)
    overshoot = [8.0, 7.0, 6.0, 5.0, 4.0, 3.5, 3.0, 2.6, 2.3]
    for i, v in enumerate(overshoot + [1.0, 0.5]):
        cfg.alpha[i] = v
    cfg.n_alpha = 11
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=4)
    got = emu_lib.ddp_centroidal_solve(ps, cfg, trace_len=4, team=1)
    assert_ddp_parity(ref, got)
    assert (ref.alpha_idx[:, :4] >= 8).any(), f"no second-round acceptance: {ref.alpha_idx[:, :4]}"
    for i, v in enumerate(overshoot + [2.2, 2.1]):
        cfg.alpha[i] = v
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=4)
    got = emu_lib.ddp_centroidal_solve(ps, cfg, trace_len=4, team=1)
    assert_ddp_parity(ref, got)
    assert (ref.alpha_idx[:, :4] == -1).any(), f"no failed line search: {ref.alpha_idx[:, :4]}"


def test_team_warm_start_one_iteration(oracle):
    """The closed-loop tick: u_init given, max_iter = 1, m_max = 16 (row stride != 32)."""
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
    for max_iter in (0, 1, 3):
        cfg = problem.ddp_centroidal_config(max_iter=max_iter)
        assert_ddp_parity(oracle.ddp_centroidal_solve(ps, cfg, trace_len=4), emu_lib.ddp_centroidal_solve(ps, cfg, trace_len=4, team=1))


def test_team_unconstrained(oracle):
    w = workloads.ddp_centroidal_config3(batch=1, horizon_steps=8)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config(max_iter=3)
    cfg.with_input_constraint = 0
    assert_ddp_parity(oracle.ddp_centroidal_solve(ps, cfg, trace_len=4), emu_lib.ddp_centroidal_solve(ps, cfg, trace_len=4, team=1))


def test_team_srb(oracle):
    w = workloads.ddp_srb_config4(batch=2, horizon_steps=8)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    cfg = problem.ddp_srb_config(max_iter=5)
    assert_ddp_parity(oracle.ddp_srb_solve(ps, cfg, trace_len=8), emu_lib.ddp_srb_solve(ps, cfg, trace_len=8, team=1))
