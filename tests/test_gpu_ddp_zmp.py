"""GPU tier: CCC::DdpZmp path — engine vs oracle bit-exact, and the reference's closed-loop test."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import problem

from footstep_manager import walking_plan
from parity import assert_ddp_parity
from test_ddp_zmp_cpu import run_ddp_zmp_closed_loop

pytestmark = pytest.mark.gpu
G = 9.80665


@pytest.fixture(scope="module")
def eng_mod():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0
    return engine


@pytest.mark.parametrize("variant", [1, 0])
def test_parity_batch(eng_mod, oracle, variant):
    """512 perturbed initial states on 4 reference schedules (walking plan sampled at 4 times), N = 100; variant 1 =
    one thread per problem (the default kernel), 0 = the warp-per-problem engine."""
    N, times = 100, (0.0, 1.9, 2.4, 4.95)
    ref_zmp = np.zeros((len(times), N + 1, 3))
    for s, t0 in enumerate(times):
        fm = walking_plan()
        for tick in range(int(round(t0 / 0.005)) + 1):
            fm.update(tick * 0.005)
        for k in range(N + 1):
            ref_zmp[s, k, :2] = fm.ref_zmp(t0 + k * 0.02)
    rng = np.random.default_rng(4)
    B = 512
    sched_id = (np.arange(B) % len(times)).astype(np.int32)
    x0 = np.zeros((B, 6))
    x0[:, [0, 2]] = ref_zmp[sched_id, 0, :2] + rng.uniform(-0.03, 0.03, (B, 2))
    x0[:, [1, 3]] = rng.uniform(-0.1, 0.1, (B, 2))
    x0[:, 4] = 1.0 + rng.uniform(-0.02, 0.02, B)
    u_init = np.zeros((B, N, 3))
    u_init[:, :, 0], u_init[:, :, 1], u_init[:, :, 2] = x0[:, None, 0], x0[:, None, 2], 100.0 * G
    ps = problem.DdpZmpProblemSet(ref_zmp, np.ones((len(times), N + 1)), sched_id, x0, 100.0, 0.02, u_init=u_init)
    old = eng_mod.DdpZmpEngine.set_variant(variant)
    eng = eng_mod.DdpZmpEngine(N, B, len(times))
    eng_mod.DdpZmpEngine.set_variant(old)
    for mi in (3, 40):
        cfg = problem.ddp_config(max_iter=mi)
        assert_ddp_parity(oracle.ddp_zmp_solve(ps, cfg, trace_len=40, n_threads=max(1, oracle.hardware_threads())),
                          eng.solve(ps, cfg, trace_len=40))
    with pytest.raises(eng_mod.EngineError):
        eng.solve(ps, problem.ddp_config(with_input_constraint=True))


def test_closed_loop(eng_mod):
    eng = eng_mod.DdpZmpEngine(100, 1, 1)
    ok, planned, sim, rz = run_ddp_zmp_closed_loop(lambda ps, cfg: eng.solve(ps, cfg))
    assert ok
    assert np.linalg.norm(planned[:2] - rz) < 1e-2 and abs(sim.z[0] - 1.0) < 1e-2
    assert np.linalg.norm(sim.pos[:2] - rz) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2


def test_thread_kernel_stiff_problems_and_ragged_batch(eng_mod, oracle):
    """Shortened / rejected line-search steps, lambda increases, max_iter hits; batch not a multiple of the block."""
    from test_ddp_zmp_cpu import _zmp_problem_set

    for stiff, B in ((3, 1000), (10, 77)):
        ps = _zmp_problem_set(B, 30, seed=11, stiff=stiff)
        cfg = problem.ddp_config(max_iter=60)
        ref = oracle.ddp_zmp_solve(ps, cfg, trace_len=8, n_threads=max(1, oracle.hardware_threads()))
        eng = eng_mod.DdpZmpEngine(30, B, 2)
        assert_ddp_parity(ref, eng.solve(ps, cfg, trace_len=8))
        assert (ref.alpha_idx[:, :8] > 0).any()
