"""PreviewControlCentroidal: the reference's closed-loop test (tests/src/TestPreviewControlCentroidal.cpp:15-150) with
numpy preview rows and the oracle QP on the CPU; on the GPU the preview rows (ccc_preview_input) and the wrench
distribution (ccc_qp_solve_grouped, one matrix group per problem) against the oracle."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import contact
from centroidalcontrolcollection_b200.linear_models import G
from centroidalcontrolcollection_b200.preview_control_centroidal import PreviewControlCentroidal

from sim_models import CentroidalSim

EPS_T = 1e-6


def contacts_at(t):  # :31-53
    t += EPS_T
    rect = ((-0.1, -0.1), (0.1, 0.1)) if t < 1.4 else ((0.15, 0.15), (0.35, 0.35)) if t < 1.6 else ((0.4, -0.1), (0.6, 0.1))
    return contact.contact_from_rect(*rect)


def ref_data(t):  # :54-74 -> (pos (angular, linear), wrench (moment, force))
    t += EPS_T
    lin = (0.0, 0.0, 1.0) if t < 1.4 else (0.25, 0.0, 1.2) if t < 1.6 else (0.5, 0.0, 1.0)
    return np.array([0.0, 0.0, 0.0, *lin]), np.zeros(6)


def closed_loop(qp_solve_grouped, gemv=None, end_time=3.0):
    """-> (sim, per-tick ok, final reference position)."""
    sim_dt, mass, inertia = 0.005, 100.0, (40.0, 20.0, 10.0)
    pc = PreviewControlCentroidal(mass, inertia, 2.0, 0.01)
    sim = CentroidalSim(mass, inertia, sim_dt)
    sim.x[0:3] = ref_data(0.0)[0][3:6]
    wrench = np.array([0, 0, 0, 0, 0, mass * G])
    kw = {} if gemv is None else {"gemv": gemv}
    t, ok = 0.0, True
    while t < end_time:
        pos = np.concatenate([sim.x[3:6], sim.x[0:3]])  # (angular, linear)
        vel = np.concatenate([sim.x[9:12], sim.x[6:9]])
        acc = np.concatenate([wrench[0:3] / np.array(inertia), wrench[3:6] / mass - np.array([0, 0, G])])  # :104-105
        vtx, rdg = contacts_at(t)
        wrench = pc.plan_batch(qp_solve_grouped, vtx, rdg, ref_data, pos[None], vel[None], acc[None], t, sim_dt, **kw)[0]
        assert pc.wrench_dist.last_result.status[0] == 0
        rp = ref_data(t)[0]
        ok = ok and np.linalg.norm(pos - rp) < 2.0 and np.linalg.norm(vel) < 2.0
        t += sim_dt
        sim.update(wrench[3:6], wrench[0:3])
        if 1.0 <= t < 1.0 + sim_dt:
            sim.add_disturb(np.array([0.05, 0.05, 0.0]), np.zeros(3))
    return sim, ok, ref_data(t)[0]


def test_reference_closed_loop_with_oracle_qp(oracle):
    sim, ok, rp = closed_loop(lambda gp: gp.solve_by_group(lambda ps: oracle.qp_solve(ps)))
    assert ok
    pos = np.concatenate([sim.x[3:6], sim.x[0:3]])
    vel = np.concatenate([sim.x[9:12], sim.x[6:9]])
    assert np.linalg.norm(pos - rp) < 0.1 and np.linalg.norm(vel) < 0.1  # :147-148


@pytest.mark.gpu
def test_batch_on_gpu_matches_oracle(oracle):
    """2048 perturbed states at three times of the scenario: preview rows and wrench-distribution QPs on the GPU."""
    from centroidalcontrolcollection_b200 import build, engine

    from test_preview_control_cpu import _oracle_gemv

    build.build()
    rng = np.random.default_rng(20260110)
    pc = PreviewControlCentroidal(100.0, (40.0, 20.0, 10.0), 2.0, 0.01)
    B = 2048
    eng = engine.QpEngine(16, 0, 32, B, max_groups=B)
    for t0 in (0.3, 1.45, 2.2):
        rp = ref_data(t0)[0]
        pos = rp[None, :] + rng.uniform(-0.05, 0.05, (B, 6))
        vel = rng.uniform(-0.2, 0.2, (B, 6))
        acc = rng.uniform(-0.5, 0.5, (B, 6))
        vtx, rdg = contacts_at(t0)
        got = pc.plan_batch(eng.solve_grouped, vtx, rdg, ref_data, pos, vel, acc, t0, 0.005, gemv=engine.preview_gemv)
        got_res = pc.wrench_dist.last_result
        ref = pc.plan_batch(lambda gp: gp.solve_by_group(lambda ps: oracle.qp_solve(ps)), vtx, rdg, ref_data, pos, vel, acc, t0, 0.005,
                            gemv=_oracle_gemv(oracle))
        ref_res = pc.wrench_dist.last_result
        assert (ref_res.status == 0).all()
        for f in ("x", "iters", "status", "n_active", "active"):
            assert np.array_equal(getattr(ref_res, f), getattr(got_res, f)), (t0, f)
        assert np.array_equal(ref, got)
        assert (got_res.n_active > 0).any()  # some ridge forces sit on their lower bound


@pytest.mark.gpu
def test_reference_closed_loop_on_gpu():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    eng = engine.QpEngine(16, 0, 32, 1, max_groups=1)
    sim, ok, rp = closed_loop(eng.solve_grouped, gemv=engine.preview_gemv)
    pos = np.concatenate([sim.x[3:6], sim.x[0:3]])
    vel = np.concatenate([sim.x[9:12], sim.x[6:9]])
    assert ok and np.linalg.norm(pos - rp) < 0.1 and np.linalg.norm(vel) < 0.1
