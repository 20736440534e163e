"""CPU tier: formulation helpers and the QP-based ZMP methods with the oracle QP, pinned on the
reference's own tests (TestStateSpaceModel, TestInvariantSequentialExtension, TestLinearMpcZmp,
TestIntrinsicallyStableMpc)."""
import numpy as np

from centroidalcontrolcollection_b200 import linear_mpc
from centroidalcontrolcollection_b200.linear_models import (G, ComZmpModelJerkInput, InvariantSequentialExtension,
                                                           StateSpaceModel)

from footstep_manager import walking_plan
from sim_models import ComZmpSim2d


def test_zoh_matches_closed_form_and_euler():
    """reference tests/src/TestStateSpaceModel.cpp:64-136: ZOH vs Euler < 1e-3 at dt = 0.01; closed form App. D."""
    dt = 0.01
    m = ComZmpModelJerkInput(1.0).calc_disc_matrix(dt)
    assert np.allclose(m.Ad, [[1, dt, dt**2 / 2], [0, 1, dt], [0, 0, 1]], atol=1e-15)
    assert np.allclose(m.Bd[:, 0], [dt**3 / 6, dt**2 / 2, dt], atol=1e-15)
    x, u = np.array([1.0, 2.0, 3.0]), np.array([-0.5])
    for E in (np.zeros(3), np.array([-1.0, 2.0, -3.0])):
        s = StateSpaceModel(3, 1, 1)
        s.A[0, 1], s.A[1, 2], s.B[2, 0], s.E = 1, 1, 1, E
        s.calc_disc_matrix(dt)
        assert np.linalg.norm(s.state_eq_disc(x, u) - (x + dt * s.state_eq(x, u))) < 1e-3


def test_invariant_sequential_extension():
    """reference tests/src/TestInvariantSequentialExtension.cpp:68-92: condensed == iterated, 1e-10."""
    dt, N = 0.01, 5
    model = ComZmpModelJerkInput(1.0).calc_disc_matrix(dt)
    x0, u = np.array([1.0, 2.0, 3.0]), np.array([5.0, 2.5, 0.0, -1.0, -2.0])
    ext = InvariantSequentialExtension(model, N, False)
    xs, x = [], x0.copy()
    for k in range(N):
        x = model.state_eq_disc(x, u[k:k + 1])
        xs.append(x)
    assert np.linalg.norm(ext.A_seq @ x0 + ext.B_seq @ u + ext.E_seq - np.concatenate(xs)) < 1e-10
    ext_o = InvariantSequentialExtension(model, N, True)
    ys = np.array([model.observ_eq(xk, np.zeros(1))[0] for xk in xs])
    assert np.linalg.norm(ext_o.A_seq @ x0 + ext_o.B_seq @ u + ext_o.E_seq - ys) < 1e-10
    # Wieber's closed form
    h_g = 1.0 / G
    for i in range(N):
        for j in range(i + 1):
            p = i - j
            assert abs(ext_o.B_seq[i, j] - (dt**3 * (1 + 3 * p + 3 * p * p) / 6 - h_g * dt)) < 1e-15


def _closed_loop(kind, qp_solve, end_time=10.0):
    """reference tests/src/TestLinearMpcZmp.cpp:15-125 / TestIntrinsicallyStableMpc.cpp (same scenario)."""
    horizon_duration, horizon_dt, sim_dt, h = 2.0, 0.02, 0.005, 1.0
    mpc = linear_mpc.LinearMpcZmp(h, horizon_duration, horizon_dt) if kind == "lmpc" else linear_mpc.IntrinsicallyStableMpc(
        h, horizon_duration, horizon_dt)
    N = mpc.mpc_1d.horizon_steps
    fm = walking_plan()
    sim = ComZmpSim2d(h, sim_dt)
    planned = sim.pos.copy()
    t, ok, n_ticks = 0.0, True, 0
    while t < end_time:
        fm.update(t)
        ts = [t + i * horizon_dt for i in range(N)]
        if kind == "lmpc":
            lim = [fm.make_linear_mpc_zmp_ref_data(tt) for tt in ts]
            lo, hi = np.array([l[0] for l in lim])[None], np.array([l[1] for l in lim])[None]
            acc = G / h * (sim.pos - planned)
            planned = mpc.plan_batch(qp_solve, sim.pos[None], sim.vel[None], acc[None], lo, hi, sim_dt)[0]
        else:
            rd = [fm.make_ismpc_ref_data(tt) for tt in ts]
            ref = np.array([r[0] for r in rd])[None]
            lo, hi = np.array([r[1][0] for r in rd])[None], np.array([r[1][1] for r in rd])[None]
            cp = sim.pos + np.sqrt(h / G) * sim.vel
            planned = mpc.plan_batch(qp_solve, cp[None], planned[None], ref, lo, hi, sim_dt)[0]
        assert (mpc.mpc_1d.last_result.status == 0).all()
        zl = fm.zmp_limits(t)
        ok &= bool((planned - zl[0] >= 0).all() and (zl[1] - planned >= 0).all())
        t += sim_dt
        n_ticks += 1
        sim.update(planned)
        for dtm in (4.5, 8.5):
            if dtm <= t < dtm + sim_dt:
                sim.add_disturb(np.array([0.05, 0.05]))
    zl = fm.zmp_limits(t)
    return ok, planned, sim.pos, zl


def test_linear_mpc_zmp_closed_loop(oracle):
    ok, planned, pos, zl = _closed_loop("lmpc", lambda ps: oracle.qp_solve(ps))
    assert ok
    assert (planned - zl[0] >= 0).all() and (zl[1] - planned >= 0).all()
    assert (pos - zl[0] >= 0).all() and (zl[1] - pos >= 0).all()


def test_intrinsically_stable_mpc_closed_loop(oracle):
    ok, planned, pos, zl = _closed_loop("ismpc", lambda ps: oracle.qp_solve(ps))
    assert ok
    assert (planned - zl[0] >= 0).all() and (zl[1] - planned >= 0).all()
    assert (pos - zl[0] >= 0).all() and (zl[1] - pos >= 0).all()


def test_workload_schedule_matches_footstep_manager():
    """workloads._walking_limits (bench input generator) == the FootstepManager fixture on the same plan."""
    from centroidalcontrolcollection_b200.workloads import _walking_limits

    for t0 in (0.0, 1.8, 2.35, 4.9, 7.7):
        fm = walking_plan(0.2, 0.2)
        fm.horizon_duration = 2.0
        for tick in range(int(round(t0 / 0.005)) + 1):  # the manager must be updated every control cycle
            fm.update(tick * 0.005)
        fm.update(t0)
        ref, lo, hi = _walking_limits(0.2, 0.2, t0, 100, 0.02, eps_reps=2)
        for i in range(100):
            r, (l, h) = fm.make_ismpc_ref_data(t0 + i * 0.02)
            assert np.allclose(ref[i], r, atol=1e-12) and np.allclose(lo[i], l) and np.allclose(hi[i], h), (t0, i)
