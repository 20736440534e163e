"""StepMpc (reference include/CCC/StepMpc.h, src/StepMpc.cpp): the reference's closed-loop test with online footstep update
(tests/src/TestStepMpc.cpp:16-135) on the host restatement, and the batched kernel against it on the GPU."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import step_mpc
from centroidalcontrolcollection_b200.linear_models import G

import footstep_manager as fmx
from sim_models import ComZmpSim2d


def _closed_loop(plan):
    """plan(ctrl, elements, pos, vel, t) -> (current_zmp[2], next_foot_zmp[2] or None)."""
    sim_dt, h = 0.005, 1.0
    ctrl = step_mpc.StepMpc(h)
    fm = fmx.walking_plan()
    sim = ComZmpSim2d(h, sim_dt)
    t, ok, planned = 0.0, True, np.zeros(2)
    while t < 10.0:
        fm.update(t)
        elements = fmx.make_step_mpc_ref_data(fm, t)
        planned, next_foot = plan(ctrl, elements, sim.pos, sim.vel, t)
        ok = ok and np.linalg.norm(planned - fm.ref_zmp(t)) < 0.2
        t += sim_dt
        sim.update(planned)
        for td in (4.5, 8.5):
            if td <= t < td + sim_dt:
                sim.add_disturb(np.array([0.05, 0.05]))
        # online footstep update (:113-128)
        if fm.footstep_list:
            pre = fm.footstep_list[0].swing_end_time - 0.1
            if pre <= t < pre + sim_dt and next_foot is not None:
                fm.footstep_list[0].pos = np.array(next_foot)
    return ok, planned, sim, fm.ref_zmp(t)


def test_reference_closed_loop_on_the_host():
    ok, planned, sim, ref = _closed_loop(lambda c, el, pos, vel, t: c.plan_once(el, pos, vel, t))
    assert ok
    # reference final checks (:138-141)
    assert np.linalg.norm(planned - ref) < 1e-2 and np.linalg.norm(sim.pos - ref) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2


def test_system_is_the_normal_equation_of_the_weighted_residuals():
    """Independent check of the assembled system: its solution minimises the weighted sum of squares that the reference's
    terms describe (fixed / free ZMP, double-support smoothness, capture-point terms), evaluated directly by simulating
    the step model phase by phase."""
    ctrl = step_mpc.StepMpc(1.0, pos=0.3, vel=0.2)
    fm = fmx.walking_plan()
    for tick in range(int(2.65 / 0.005) + 1):
        fm.update(tick * 0.005)
    elements = fmx.make_step_mpc_ref_data(fm, 2.65)
    single = [e[0] for e in elements]
    end_time = [e[2] for e in elements]
    zmp = np.array([e[1][0] for e in elements])
    x0 = np.array([0.21, 0.13])
    M, v = ctrl.system_1d(single, zmp, end_time, x0, 2.65)
    n = len(zmp)
    w = ctrl.w

    def cost(u):
        x, c = x0.copy(), 0.0
        diag = np.full(n, w["free_zmp"])
        diag[0] = w["fixed_zmp"]
        if n > 1 and not single[0]:
            diag[1] = w["fixed_zmp"]
        c += 0.5 * np.sum(diag * (u - zmp) ** 2) - 0.5 * np.sum(diag * zmp ** 2)
        for i in range(1, n - 1):
            if single[i - 1] and not single[i] and single[i + 1]:
                c += 0.5 * w["double_support"] * (u[i - 1] - 2 * u[i] + u[i + 1]) ** 2
        n_future = sum(1 for i in range(1, n) if single[i])
        outs = []
        for i in range(n):
            Ad, Bd, C = ctrl._step_model(end_time[i] - (2.65 if i == 0 else end_time[i - 1]))
            x = Ad @ x + Bd * u[i]
            outs.append(C @ x)
        for i in range(n):
            c += 0.5 * w["pos"] * (outs[i][0] - zmp[min(i + 1, n - 1)]) ** 2 + 0.5 * w["vel"] * outs[i][1] ** 2
            if i >= 1 and single[i] and n_future >= 1:
                c += 0.5 * w["capture_point_abs"] * (outs[i - 1][2] - zmp[i]) ** 2
                c += 0.5 * w["capture_point_rel"] * (outs[i - 1][2] - u[i]) ** 2
        return c

    sol = np.linalg.solve(M, -v)
    rng = np.random.default_rng(3)
    c0 = cost(sol)
    for _ in range(50):
        assert cost(sol + 1e-3 * rng.standard_normal(n)) >= c0 - 1e-12
    # and the gradient of that cost at the solution vanishes (finite differences)
    g = np.array([(cost(sol + 1e-6 * np.eye(n)[i]) - cost(sol - 1e-6 * np.eye(n)[i])) / 2e-6 for i in range(n)])
    assert np.abs(g).max() < 1e-5


@pytest.mark.gpu
def test_batched_kernel_matches_host_restatement_and_closed_loop():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    rng = np.random.default_rng(20260111)
    ctrl = step_mpc.StepMpc(1.0)
    P, per = 48, 100
    recs, times = [], []
    for p in range(P):
        fm = fmx.walking_plan(step_length=rng.uniform(0.1, 0.3), step_width=rng.uniform(0.16, 0.24))
        t0 = float(rng.uniform(0.0, 9.5))
        for tick in range(int(t0 / 0.005) + 1):
            fm.update(tick * 0.005)
        fm.update(t0)
        recs.append(fmx.make_step_mpc_ref_data(fm, t0))
        times.append(t0)
    plan_id = np.repeat(np.arange(P, dtype=np.int32), per)
    pos, vel = rng.uniform(-0.2, 1.0, (P * per, 2)), rng.uniform(-0.3, 0.3, (P * per, 2))
    cur, nxt, has = ctrl.plan_batch(engine.step_mpc_plan, recs, times, pos, vel, plan_id)
    worst = 0.0
    for b in range(0, P * per, 7):
        c_ref, n_ref = ctrl.plan_once(recs[plan_id[b]], pos[b], vel[b], times[plan_id[b]])
        # the systems have condition numbers of 1e8 - 5e10 (cosh of up to 3 s of horizon, squared): two solvers (LU on the host,
        # Gaussian elimination in a different summation order on the device) agree to ~cond x 1e-16, not to rounding
        worst = max(worst, np.abs(cur[b] - c_ref).max())
        assert bool(has[b]) == (n_ref is not None)
        if n_ref is not None:
            worst = max(worst, np.abs(nxt[b] - n_ref).max())
    print(f"StepMpc kernel vs host restatement: max abs diff {worst:.3e}")
    assert worst <= 1e-5
    # also with CoM position / velocity weights switched on
    ctrl2 = step_mpc.StepMpc(1.0, pos=0.3, vel=0.2)
    cur2, nxt2, has2 = ctrl2.plan_batch(engine.step_mpc_plan, recs, times, pos, vel, plan_id)
    for b in range(0, P * per, 31):
        c_ref, n_ref = ctrl2.plan_once(recs[plan_id[b]], pos[b], vel[b], times[plan_id[b]])
        assert np.abs(cur2[b] - c_ref).max() <= 1e-5

    def plan(c, el, p_, v_, t):
        cz, nz, h = c.plan_batch(engine.step_mpc_plan, [el], [t], p_[None, :], v_[None, :], [0])
        return cz[0], (nz[0] if h[0] else None)

    ok, planned, sim, ref = _closed_loop(plan)
    assert ok
    assert np.linalg.norm(planned - ref) < 1e-2 and np.linalg.norm(sim.pos - ref) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2


def test_no_cpu_fallback():
    """Without a CUDA device the entry points of the two controllers added last refuse loudly (no CPU path in the product)."""
    from centroidalcontrolcollection_b200 import build, closed_form, engine

    build.build()
    if engine.lib().ccc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    fm = fmx.walking_plan()
    fm.update(0.0)
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        step_mpc.StepMpc(1.0).plan_batch(engine.step_mpc_plan, [fmx.make_step_mpc_ref_data(fm, 0.0)], [0.0], np.zeros((1, 2)), np.zeros((1, 2)), [0])
    spc = closed_form.SingularPreviewControlZmp(1.0, 2.0, 0.01)
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        spc.plan_batch(engine.singular_preview_plan, spc.sample(fm.ref_zmp, 0.0)[None], np.zeros((1, 2, 3)), [0])
