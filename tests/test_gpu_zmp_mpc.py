"""GPU tier: the schedule compiler (ccc_footstep_compile) against the FootstepManager fixture (tests/footstep_manager.py,
the restatement of reference tests/src/FootstepManager.h), and the device-side planOnce of LinearMpcZmp /
IntrinsicallyStableMpc on compiled tables (ccc_zmp_mpc_plan) against the host classes with the oracle QP."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import linear_mpc, schedule, workloads

import footstep_manager as fmx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine_mod():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0
    return engine


def _plan_rows(step_length, step_width, transit, swing):
    L, w = step_length, 0.5 * step_width
    return [(fmx.LEFT, (L, w), 2.0), (fmx.RIGHT, (2 * L, -w), 3.0), (fmx.LEFT, (3 * L, w), 4.0), (fmx.RIGHT, (4 * L, -w), 5.0),
            (fmx.LEFT, (3 * L, w), 6.0), (fmx.RIGHT, (3 * L, -w), 7.0)], transit, swing


def _manager_tables(rows, transit, swing, step_width, t0, N, dt, manager_horizon):
    """The stateful fixture, updated once per control cycle up to t0 (as the reference's test loops do)."""
    fm = fmx.FootstepManager({fmx.LEFT: np.array([0.0, 0.5 * step_width]), fmx.RIGHT: np.array([0.0, -0.5 * step_width])})
    fm.horizon_duration = manager_horizon
    for foot, pos, ts in rows:
        fm.append_footstep(fmx.Footstep(foot, pos, ts, transit, swing))
    for tick in range(int(round(t0 / 0.005)) + 1):
        fm.update(tick * 0.005)
    fm.update(t0)
    ref, lo, hi = np.zeros((N, 2)), np.zeros((N, 2)), np.zeros((N, 2))
    for i in range(N):
        ref[i], (lo[i], hi[i]) = fm.make_ismpc_ref_data(t0 + i * dt)
    return ref, lo, hi


def test_footstep_compile_matches_footstep_manager(engine_mod):
    N, dt = 100, 0.02
    cases = []
    for t0 in (0.0, 1.8, 2.05, 2.1, 2.35, 2.9, 2.95, 3.0, 4.9, 7.7, 8.0, 8.5, 12.0):
        for (sl, sw, tr, swg) in ((0.2, 0.2, 0.2, 0.8), (0.13, 0.23, 0.0, 1.0), (0.3, 0.16, 0.4, 0.6)):
            cases.append((t0, sl, sw, tr, swg, 10.0))
    cases.append((1.0, 0.2, 0.2, 0.2, 0.8, 2.5))  # short manager horizon: the knot list ends inside the plan
    plans = schedule.FootstepPlans(len(cases) + 1, 8, N, dt, eps_reps=2)
    want = []
    for p, (t0, sl, sw, tr, swg, mh) in enumerate(cases):
        rows, tr, swg = _plan_rows(sl, sw, tr, swg)
        plans.current_time[p] = t0
        plans.stance0[p] = [[0.0, 0.5 * sw], [0.0, -0.5 * sw]]
        for foot, pos, ts in rows:
            plans.append_footstep(p, foot, pos, ts, tr, swg)
        want.append(_manager_tables(rows, tr, swg, sw, t0, N, dt, 10.0))
    plans.current_time[-1] = 3.3  # a plan without footsteps
    want.append(_manager_tables([], 0.2, 0.8, 0.2, 3.3, N, dt, 10.0))
    cases.pop()  # the short-horizon case needs its own manager_horizon: compiled separately below
    got = engine_mod.footstep_compile(plans)
    for p in list(range(len(cases))) + [plans.P - 1]:
        ref, lo, hi = want[p if p < len(cases) else -1]
        assert np.array_equal(got.ref_zmp[p], ref), p
        assert np.array_equal(got.lim_min[p], lo) and np.array_equal(got.lim_max[p], hi), p
    # horizon (2 s) inside a 2.5 s manager horizon
    short = schedule.FootstepPlans(1, 8, N, dt, eps_reps=2, manager_horizon=2.5)
    rows, tr, swg = _plan_rows(0.2, 0.2, 0.2, 0.8)
    short.current_time[0] = 1.0
    for foot, pos, ts in rows:
        short.append_footstep(0, foot, pos, ts, tr, swg)
    g2 = engine_mod.footstep_compile(short)
    ref, lo, hi = _manager_tables(rows, tr, swg, 0.2, 1.0, N, dt, 2.5)
    assert np.array_equal(g2.ref_zmp[0], ref) and np.array_equal(g2.lim_min[0], lo) and np.array_equal(g2.lim_max[0], hi)
    # a horizon longer than the knot list is refused (the reference would read past the end of its map)
    bad = schedule.FootstepPlans(1, 8, N, dt, eps_reps=2, manager_horizon=1.0)
    with pytest.raises(engine_mod.EngineError):
        engine_mod.footstep_compile(bad)


def _walking_plans(lengths, widths, t0, N, dt):
    plans = schedule.FootstepPlans(len(lengths), 8, N, dt, eps_reps=2)
    for p, (sl, sw) in enumerate(zip(lengths, widths)):
        rows, tr, swg = _plan_rows(sl, sw, 0.2, 0.8)
        plans.current_time[p] = t0
        plans.stance0[p] = [[0.0, 0.5 * sw], [0.0, -0.5 * sw]]
        for foot, pos, ts in rows:
            plans.append_footstep(p, foot, pos, ts, tr, swg)
    return plans


def test_config2_linear_mpc_zmp_on_device_tables(oracle, engine_mod):
    """Config 2 at full size: the plan compiled on the device, 4096 x 2 QPs assembled / solved / post-processed there."""
    w = workloads.linear_mpc_zmp_config2()
    mpc = linear_mpc.LinearMpcZmp(w["com_height"], w["horizon_duration"], w["horizon_dt"])
    N = mpc.mpc_1d.horizon_steps
    tables = engine_mod.footstep_compile(_walking_plans([0.2], [0.2], 1.8, N, w["horizon_dt"]))
    assert np.array_equal(tables.lim_min[0], w["lim_min"][0]) and np.array_equal(tables.lim_max[0], w["lim_max"][0])
    B = len(w["pos"])
    state = np.stack([w["pos"], w["vel"], w["acc"]], axis=2)  # [B][2 axes][3]
    eng = engine_mod.ZmpMpcEngine(mpc, B, 1)
    planned, iters, status = eng.plan(state, np.zeros(B, dtype=np.int32), tables, w["control_dt"])
    assert eng.last_launches >= 4
    planned2, iters2, _ = eng.plan(state, np.zeros(B, dtype=np.int32), tables, w["control_dt"])  # matrices reused
    eng.close()
    assert (status == 0).all()
    threads = max(1, oracle.hardware_threads())
    ref = mpc.plan_batch(lambda ps: oracle.qp_solve(ps, n_threads=threads), w["pos"], w["vel"], w["acc"], w["lim_min"], w["lim_max"],
                         w["control_dt"])
    assert np.array_equal(planned, ref) and np.array_equal(planned2, ref)
    assert np.array_equal(iters, mpc.mpc_1d.last_result.iters) and np.array_equal(iters2, iters)


def test_config5_ismpc_on_device_tables(oracle, engine_mod):
    """Config 5's 256 footstep plans compiled on the device x 16 capture-point perturbations each."""
    w = workloads.ismpc_config5(n_plans=256, n_perturb=16)
    mpc = linear_mpc.IntrinsicallyStableMpc(w["com_height"], w["horizon_duration"], w["horizon_dt"])
    N = mpc.mpc_1d.horizon_steps
    rng = np.random.Generator(np.random.PCG64(20260104))
    lengths, widths = rng.uniform(0.1, 0.3, 16), rng.uniform(0.16, 0.24, 16)
    tables = engine_mod.footstep_compile(_walking_plans([lengths[p % 16] for p in range(256)], [widths[(p // 16) % 16] for p in range(256)],
                                                        1.8, N, w["horizon_dt"]))
    plan_id = np.repeat(np.arange(256, dtype=np.int32), 16)
    assert np.array_equal(tables.lim_min[plan_id], w["lim_min"]) and np.array_equal(tables.lim_max[plan_id], w["lim_max"])
    assert np.allclose(tables.ref_zmp[plan_id], w["ref_zmp"], atol=1e-12)
    B = len(plan_id)
    state = np.stack([w["capture_point"], w["planned_zmp"]], axis=2)  # [B][2 axes][2]
    eng = engine_mod.ZmpMpcEngine(mpc, B, 256)
    planned, iters, status = eng.plan(state, plan_id, tables, w["control_dt"])
    eng.close()
    assert (status == 0).all()
    threads = max(1, oracle.hardware_threads())
    ref = mpc.plan_batch(lambda ps: oracle.qp_solve(ps, n_threads=threads), w["capture_point"], w["planned_zmp"], tables.ref_zmp[plan_id],
                         tables.lim_min[plan_id], tables.lim_max[plan_id], w["control_dt"])
    assert np.array_equal(planned, ref)
    assert np.array_equal(iters, mpc.mpc_1d.last_result.iters)
