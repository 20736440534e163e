"""LinearMpcZ (reference src/LinearMpcZ.cpp): the reference's closed-loop test with the oracle QP on the CPU, and the
engine vs the oracle on a batch of perturbed initial states on the GPU (solutions, iteration counts, active sets)."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import linear_mpc_z
from centroidalcontrolcollection_b200.linear_models import G, StateSpaceModel


def contact_func(t):  # tests/src/TestLinearMpcZ.cpp:26
    return not ((5.0 < t < 5.25) or (6.0 < t < 6.5))


def ref_pos_func(t):  # :27
    return 1.0 if t < 8.5 else 0.8


def _vertical_sim(mass, sim_dt):
    """tests/src/SimModels.h:44-73"""
    s = StateSpaceModel(2, 1, 0)
    s.A[0, 1] = 1
    s.B[1, 0] = 1 / mass
    s.E[1] = -1 * G
    return s.calc_disc_matrix(sim_dt)


def closed_loop(qp_solve, end_time=10.0):
    """tests/src/TestLinearMpcZ.cpp:15-84 -> (state, per-tick ok, QP iterations)."""
    horizon_dt, sim_dt, mass = 0.05, 0.04, 100.0
    mpc = linear_mpc_z.LinearMpcZ(mass, horizon_dt, int(2.0 / horizon_dt))
    sim = _vertical_sim(mass, sim_dt)
    x = np.array([ref_pos_func(0.0), 0.0])
    t, ok, iters = 0.0, True, []
    while t < end_time:
        force = float(mpc.plan_batch(qp_solve, contact_func, ref_pos_func, x[None, :], t)[0])
        if contact_func(t):
            assert mpc.last_result.status[0] == 0
            iters.append(int(mpc.last_result.iters[0]))
        ok = ok and abs(x[0] - ref_pos_func(t)) < 2.0 and abs(x[1]) < 5.0
        if not contact_func(t):
            ok = ok and abs(force) < 1e-8
        t += sim_dt
        x = sim.state_eq_disc(x, np.array([force]))
    return x, ok, iters, t


def test_closed_loop_with_oracle_qp(oracle):
    x, ok, iters, t = closed_loop(lambda ps: oracle.qp_solve(ps))
    assert ok
    assert abs(x[0] - ref_pos_func(t)) < 1e-2 and abs(x[1]) < 1e-2  # :80-81
    print(f"LinearMpcZ closed loop: {len(iters)} QPs, mean {np.mean(iters):.1f} active-set iterations")


@pytest.mark.gpu
def test_batch_parity_on_gpu(oracle):
    """256 perturbed initial states at three schedule times (a horizon with two flight phases, one inside a flight
    gap's shadow, one plain): engine vs oracle bit-exact incl. active sets."""
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    qp = engine.qp_solver_for()
    rng = np.random.default_rng(20260108)
    mpc = linear_mpc_z.LinearMpcZ(100.0, 0.05, 40)
    for t0 in (4.2, 5.3, 8.0):
        x0 = np.stack([ref_pos_func(t0) + rng.uniform(-0.05, 0.05, 256), rng.uniform(-0.5, 0.5, 256)], axis=1)
        got_f = mpc.plan_batch(qp, contact_func, ref_pos_func, x0, t0)
        got = mpc.last_result
        ref_f = mpc.plan_batch(lambda ps: oracle.qp_solve(ps, n_threads=max(1, oracle.hardware_threads())), contact_func, ref_pos_func, x0, t0)
        ref = mpc.last_result
        assert (ref.status == 0).all()
        for f in ("x", "iters", "status", "n_active", "active"):
            assert np.array_equal(getattr(ref, f), getattr(got, f)), (t0, f)
        assert np.array_equal(ref_f, got_f)
        assert mpc.last_problem.n < 40 or t0 == 8.0  # the flight stages carry no variable


@pytest.mark.gpu
def test_closed_loop_on_gpu():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    x, ok, iters, t = closed_loop(engine.qp_solver_for())
    assert ok and abs(x[0] - ref_pos_func(t)) < 1e-2 and abs(x[1]) < 1e-2
