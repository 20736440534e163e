"""GPU tier: the small-batch policies of the DDP engines.  Batches of at most one problem per SM run on the team kernel
(csrc/ddp_team.cuh: a CTA of 8 warps per problem, the line search as one round of concurrent rollouts); batches smaller
than the resident warps are spread over all SMs.  Neither may change a bit: everything is compared with the oracle
(iteration counts, accepted line-search indices, lambda trace, clamped sets, trajectories) and with the warp-per-problem
kernel."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import problem, workloads

from parity import assert_ddp_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0, "no CUDA device: -m gpu tests need a B200"
    yield engine
    engine.DdpCentroidalEngine.set_small_batch_policy(team=1, spread=1)


def _threads(oracle):
    return max(1, oracle.hardware_threads())


def test_team_kernel_centroidal_matches_oracle_and_queue_kernel(eng_mod, oracle):
    """148 cold starts of config 3 (one per SM), to convergence: 16 schedules, 19 - 300 iterations per problem."""
    w = workloads.ddp_centroidal_config3(batch=148)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config()
    eng = eng_mod.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
    eng_mod.DdpCentroidalEngine.set_small_batch_policy(team=1, spread=1)
    got = eng.solve(ps, cfg, trace_len=400)
    assert eng.last_team
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=400, n_threads=_threads(oracle))
    assert_ddp_parity(ref, got, rel_tol=1e-6, bit_exact=True)
    assert (ref.alpha_idx > 0).any(), "no shortened step in the sample"
    for team, spread in ((0, 1), (0, 0)):
        eng_mod.DdpCentroidalEngine.set_small_batch_policy(team=team, spread=spread)
        other = eng.solve(ps, cfg, trace_len=400)
        assert not eng.last_team
        assert_ddp_parity(ref, other, rel_tol=1e-6, bit_exact=True)
    eng_mod.DdpCentroidalEngine.set_small_batch_policy(team=1, spread=1)


def test_team_kernel_single_problem_tick(eng_mod, oracle):
    """The reference's caller: one problem, cold solve, then warm one-iteration ticks (u_init = the previous plan)."""
    w = workloads.ddp_centroidal_config3(batch=1, n_sched=1)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    eng = eng_mod.DdpCentroidalEngine(ps.N, 1, 1)
    cfg = problem.ddp_centroidal_config()
    got = eng.solve(ps, cfg, trace_len=64)
    assert eng.last_team
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=64)
    assert_ddp_parity(ref, got, bit_exact=True)
    ps.u_init = ref.u.copy()
    # with and without the packed staging block of small host-buffer calls (one H2D + one D2H instead of one per array)
    for packed in (1, 0):
        eng_mod.DdpCentroidalEngine.set_packed_io(packed)
        try:
            for mi in (0, 1, 2):
                c = problem.ddp_centroidal_config(max_iter=mi)
                assert_ddp_parity(oracle.ddp_centroidal_solve(ps, c, trace_len=4), eng.solve(ps, c, trace_len=4), bit_exact=True)
            # outputs the caller does not ask for stay untouched
            res = ps.new_result(0)
            keep_x = res.x.copy()
            bs, rs = ps.as_struct(), res.as_struct()
            rs.x = None
            import ctypes as C

            from centroidalcontrolcollection_b200 import _abi

            cfg1 = problem.ddp_centroidal_config(max_iter=1)
            rc = eng_mod.lib().ccc_ddp_centroidal_solve(eng._h, C.addressof(bs), C.addressof(cfg1), C.addressof(rs), _abi.CCC_MEM_HOST, None)
            assert rc == 0 and np.array_equal(res.x, keep_x)
        finally:
            eng_mod.DdpCentroidalEngine.set_packed_io(1)


def test_team_kernel_late_and_failed_line_searches(eng_mod, oracle):
    """Step-size lists that push the accepted index into the team's second round (>= 8) or make every candidate fail."""
    w = workloads.ddp_centroidal_config3(batch=32, horizon_steps=20)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    eng = eng_mod.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
    overshoot = [8.0, 7.0, 6.0, 5.0, 4.0, 3.5, 3.0, 2.6, 2.3]
    for tail, want in (([1.0, 0.5], "late"), ([2.2, 2.1], "failed")):
        cfg = problem.ddp_centroidal_config(max_iter=8)
        for i, v in enumerate(overshoot + tail):
            cfg.alpha[i] = v
        cfg.n_alpha = 11
        ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=8, n_threads=_threads(oracle))
        got = eng.solve(ps, cfg, trace_len=8)
        assert eng.last_team
        assert_ddp_parity(ref, got, bit_exact=True)
        if want == "late":
            assert (ref.alpha_idx >= 8).any()
        else:
            assert (ref.alpha_idx == -1).any()


def test_team_kernel_unconstrained_and_ragged(eng_mod, oracle):
    w = workloads.ddp_centroidal_config3(batch=20, horizon_steps=12)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    eng = eng_mod.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
    cfg = problem.ddp_centroidal_config(max_iter=5)
    cfg.with_input_constraint = 0
    assert_ddp_parity(oracle.ddp_centroidal_solve(ps, cfg, trace_len=8), eng.solve(ps, cfg, trace_len=8), bit_exact=True)
    assert eng.last_team


def test_team_kernel_srb(eng_mod, oracle):
    """DdpSingleRigidBody on the team kernel: 64 cold starts of config 4, 40 iterations."""
    w = workloads.ddp_srb_config4(batch=64)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    cfg = problem.ddp_srb_config(max_iter=40)
    eng = eng_mod.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
    got = eng.solve(ps, cfg, trace_len=40)
    ref = oracle.ddp_srb_solve(ps, cfg, trace_len=40, n_threads=_threads(oracle))
    assert_ddp_parity(ref, got, rel_tol=1e-6, bit_exact=True)


@pytest.mark.parametrize("batch", [149, 300, 1000])
def test_spread_batches_match_oracle(eng_mod, oracle, batch):
    """More problems than SMs, fewer than resident warps: the queue kernel with ceil(B / 148) warps per SM."""
    w = workloads.ddp_centroidal_config3(batch=batch, horizon_steps=20)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config(max_iter=40)
    eng = eng_mod.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
    got = eng.solve(ps, cfg, trace_len=40)
    assert not eng.last_team
    ref = oracle.ddp_centroidal_solve(ps, cfg, trace_len=40, n_threads=_threads(oracle))
    assert_ddp_parity(ref, got, rel_tol=1e-6, bit_exact=True)
