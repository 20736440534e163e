"""DcmTracking, FootGuidedControl and SingularPreviewControlZmp: the reference's closed-loop tests (tests/src/TestDcmTracking.cpp,
TestFootGuidedControl.cpp, TestSingularPreviewControlZmp.cpp) with the host restatement on the CPU, and the batched kernels
against it on the GPU."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import closed_form
from centroidalcontrolcollection_b200.linear_models import G

import footstep_manager as fmx
from sim_models import ComZmpSim2d


def _closed_loop(kind, plan):
    """-> (ok per tick, planned_zmp, sim, ref_zmp at the end); plan(ctrl, ref_data, initial_param, t) -> zmp[2]."""
    sim_dt, h = 0.005, 1.0
    ctrl = closed_form.DcmTracking(h) if kind == "dcm" else closed_form.FootGuidedControl(h)
    fm = fmx.walking_plan()
    sim = ComZmpSim2d(h, sim_dt)
    t, ok, planned = 0.0, True, np.zeros(2)
    while t < 10.0:
        fm.update(t)
        ip = sim.pos + np.sqrt(h / G) * sim.vel
        if kind == "dcm":
            rd = fmx.make_dcm_tracking_ref_data(fm, t)
            ref_zmp = rd[0]
        else:
            rd = fmx.make_foot_guided_control_ref_data(fm, t)
            ref_zmp = fm.ref_zmp(t)
        planned = plan(ctrl, rd, ip, t)
        ok = ok and np.linalg.norm(planned - ref_zmp) < 0.1
        t += sim_dt
        sim.update(planned)
        for td in (4.5, 8.5):
            if td <= t < td + sim_dt:
                sim.add_disturb(np.array([0.05, 0.05]))
    ref_end = fmx.make_dcm_tracking_ref_data(fm, t)[0] if kind == "dcm" else fm.ref_zmp(t)
    return ok, planned, sim, ref_end


def _host_plan(ctrl, rd, ip, t):
    if isinstance(ctrl, closed_form.DcmTracking):
        return ctrl.plan_once(rd[0], rd[1], ip, t)
    return ctrl.plan_once(rd, ip, t)


@pytest.mark.parametrize("kind", ["dcm", "fgc"])
def test_reference_closed_loop_on_the_host(kind):
    ok, planned, sim, ref = _closed_loop(kind, _host_plan)
    assert ok
    assert np.linalg.norm(planned - ref) < 1e-2 and np.linalg.norm(sim.pos - ref) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2


def _records(kind, n_plans, rng):
    """Reference data of the walking plan (varied step lengths) at random times, as the closed loops produce them."""
    recs, times = [], []
    for p in range(n_plans):
        fm = fmx.walking_plan(step_length=rng.uniform(0.1, 0.3), step_width=rng.uniform(0.16, 0.24))
        t0 = float(rng.uniform(0.0, 9.0))
        for tick in range(int(t0 / 0.005) + 1):
            fm.update(tick * 0.005)
        fm.update(t0)
        recs.append(fmx.make_dcm_tracking_ref_data(fm, t0) if kind == "dcm" else fmx.make_foot_guided_control_ref_data(fm, t0))
        times.append(t0)
    return recs, np.array(times)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dcm", "fgc"])
def test_batched_kernel_matches_host_restatement(kind):
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    rng = np.random.default_rng(20260109)
    P, per = 48, 200
    ctrl = closed_form.DcmTracking(1.0) if kind == "dcm" else closed_form.FootGuidedControl(1.0)
    recs, times = _records(kind, P, rng)
    plan_id = np.repeat(np.arange(P, dtype=np.int32), per)
    ip = rng.uniform(-0.3, 1.0, (P * per, 2))
    run = engine.dcm_tracking_plan if kind == "dcm" else engine.foot_guided_plan
    got = ctrl.plan_batch(run, recs, times, ip, plan_id)
    ref = np.array([_host_plan(ctrl, recs[p], ip[b], times[p]) for b, p in enumerate(plan_id)])
    # exp() differs by <= 1 ulp between CUDA and libm; everything else is the same sequence of operations
    assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    # the reference's exceptions come back as an error code
    bad = list(recs)
    if kind == "dcm":
        bad[3] = (recs[3][0], [(times[3] - 1.0, np.zeros(2))])
    else:
        bad[3] = dict(recs[3], transit_duration=-0.5)
    with pytest.raises(engine.EngineError):
        ctrl.plan_batch(run, bad, times, ip, plan_id)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dcm", "fgc"])
def test_reference_closed_loop_on_the_gpu(kind):
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    run = engine.dcm_tracking_plan if kind == "dcm" else engine.foot_guided_plan

    def plan(ctrl, rd, ip, t):
        return ctrl.plan_batch(run, [rd], [t], ip[None, :], [0])[0]

    ok, planned, sim, ref = _closed_loop(kind, plan)
    assert ok
    assert np.linalg.norm(planned - ref) < 1e-2 and np.linalg.norm(sim.pos - ref) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2


def _spc_closed_loop(plan):
    """reference tests/src/TestSingularPreviewControlZmp.cpp:14-113; plan(ctrl, fm, pos, vel, planned_zmp, t) -> zmp[2]."""
    sim_dt, h = 0.005, 1.0
    ctrl = closed_form.SingularPreviewControlZmp(h, 2.0, 0.01)
    fm = fmx.walking_plan()
    sim = ComZmpSim2d(h, sim_dt)
    t, ok, planned = 0.0, True, sim.pos.copy()
    while t < 10.0:
        fm.update(t)
        planned = plan(ctrl, fm, sim.pos, sim.vel, planned, t)
        ok = ok and np.linalg.norm(planned - fm.ref_zmp(t)) < 0.1
        t += sim_dt
        sim.update(planned)
        for td in (4.5, 8.5):
            if td <= t < td + sim_dt:
                sim.add_disturb(np.array([0.05, 0.05]))
    return ok, planned, sim, fm.ref_zmp(t)


def test_singular_preview_closed_loop_on_the_host():
    ok, planned, sim, ref = _spc_closed_loop(lambda c, fm, pos, vel, pz, t: c.plan_once(fm.ref_zmp, pos, vel, pz, t, 0.005))
    assert ok
    assert np.linalg.norm(planned - ref) < 1e-2 and np.linalg.norm(sim.pos - ref) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2


@pytest.mark.gpu
def test_singular_preview_kernel_matches_host_restatement_and_closed_loop():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    rng = np.random.default_rng(20260110)
    ctrl = closed_form.SingularPreviewControlZmp(1.0, 2.0, 0.01)
    P, per = 40, 300
    seqs = []
    for p in range(P):
        fm = fmx.walking_plan(step_length=rng.uniform(0.1, 0.3), step_width=rng.uniform(0.16, 0.24))
        t0 = float(rng.uniform(0.0, 9.0))
        for tick in range(int(t0 / 0.005) + 1):
            fm.update(tick * 0.005)
        fm.update(t0)
        seqs.append(ctrl.sample(fm.ref_zmp, t0))
    seqs = np.array(seqs)
    plan_id = np.repeat(np.arange(P, dtype=np.int32), per)
    state = rng.uniform(-0.3, 1.0, (P * per, 2, 3))
    got = ctrl.plan_batch(engine.singular_preview_plan, seqs, state, plan_id, 0.005)
    ref = np.array([[ctrl.proc_once_1d(seqs[p][:, a], *state[b, a], 0.005) for a in range(2)] for b, p in enumerate(plan_id)])
    assert np.array_equal(got, ref), f"max diff {np.abs(got - ref).max()}"  # no transcendental on the path: same bits
    with pytest.raises(engine.EngineError):
        ctrl.plan_batch(engine.singular_preview_plan, seqs, state, plan_id + P, 0.005)

    def plan(c, fm, pos, vel, pz, t):
        st = np.stack([pz, pos, vel], axis=1)[None]
        return c.plan_batch(engine.singular_preview_plan, c.sample(fm.ref_zmp, t)[None], st, [0], 0.005)[0]

    ok, planned, sim, ref_end = _spc_closed_loop(plan)
    assert ok
    assert np.linalg.norm(planned - ref_end) < 1e-2 and np.linalg.norm(sim.pos - ref_end) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2
