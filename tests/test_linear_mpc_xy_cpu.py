"""CPU tier: VariantSequentialExtension and LinearMpcXY with the oracle QP, pinned on the reference's own tests
(TestVariantSequentialExtension, TestLinearMpcXY)."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import contact, linear_mpc_xy
from centroidalcontrolcollection_b200.linear_models import G, StateSpaceModel, VariantSequentialExtension

from sim_models import CentroidalSim


def _model1(E=None):
    s = StateSpaceModel(3, 1, 1)
    s.A[0, 1], s.A[1, 2], s.B[2, 0] = 1, 1, 1
    s.C[0, 0], s.C[0, 2] = 1, 2
    if E is not None:
        s.E = np.asarray(E, dtype=np.float64)
    return s


def _model3(idx):
    s = StateSpaceModel(3, 0, 1)
    s.A[0, 1], s.A[1, 2] = 0.1 * idx, 0.2 * idx * idx
    s.C[0, 0], s.C[0, 1] = 1, -1
    s.E = np.array([-0.1 * idx, 0.2 * idx * idx, -0.3])
    return s


def _iterate(models, u_seq, x0, for_output):
    out, x, acc = [], x0.copy(), 0
    for m in models:
        x = m.state_eq_disc(x, u_seq[acc:acc + m.input_dim])
        out.append(m.C @ x if for_output else x)  # extension for outputs: C x (D = 0, F dropped)
        acc += m.input_dim
    return np.concatenate(out)


@pytest.mark.parametrize("for_output", [False, True])
def test_variant_sequential_extension(for_output):
    """reference tests/src/TestVariantSequentialExtension.cpp:128-224: condensed == iterated, 1e-10."""
    dt, x0 = 1e-2, np.array([1.0, 2.0, 3.0])
    for E in (None, (-1.0, 2.0, -3.0)):
        models = [_model1(E).calc_disc_matrix(dt)] * 5
        u = np.array([5.0, 2.5, 0.0, -1.0, -2.0])
        ext = VariantSequentialExtension(models, for_output)
        assert np.linalg.norm(ext.A_seq @ x0 + ext.B_seq @ u + ext.E_seq - _iterate(models, u, x0, for_output)) < 1e-10
    models = [(_model3(i) if i in (4, 5, 6, 8) else _model1()).calc_disc_matrix(dt) for i in range(10)]
    u = np.array([0.5, 1.0, 0.5, 1.0, -1.0, -2.0])
    ext = VariantSequentialExtension(models, for_output)
    assert ext.total_input_dim == 6 and ext.B_seq.shape == ((10 if for_output else 30), 6)
    assert np.linalg.norm(ext.A_seq @ x0 + ext.B_seq @ u + ext.E_seq - _iterate(models, u, x0, for_output)) < 1e-10


def xy_motion_param(mass, t):
    """reference tests/src/TestLinearMpcXY.cpp:29-58"""
    if t < 3.0:
        rect = ((0.9, -0.15), (1.1, 0.15))
    elif t < 4.0:
        rect = ((0.9, 0.05), (1.1, 0.15))
    elif t < 5.0:
        rect = ((1.15, -0.15), (1.35, -0.05))
    elif t < 6.0:
        rect = ((1.4, 0.05), (1.6, 0.15))
    else:
        rect = ((1.4, -0.15), (1.6, 0.15))
    vertex, ridge = contact.contact_from_rect(*rect)
    return linear_mpc_xy.MotionParam(1.0, mass * G, vertex, ridge)


def xy_ref_data(t):
    """reference tests/src/TestLinearMpcXY.cpp:59-83"""
    pos = (1.0, 0.0) if t < 3.0 else (1.0, 0.1) if t < 4.0 else (1.25, -0.1) if t < 5.0 else (1.5, 0.1) if t < 6.0 else (1.5, 0.0)
    return np.array(pos), np.zeros(2), np.zeros(2)


def xy_closed_loop(qp_solve, end_time=8.0):
    """reference tests/src/TestLinearMpcXY.cpp:15-146 -> (sim, per-tick ok, QP iteration counts)."""
    horizon_dt, sim_dt, mass = 0.1, 0.05, 100.0
    horizon_steps = int(1.5 / horizon_dt)
    mpc = linear_mpc_xy.LinearMpcXY(mass, horizon_dt, horizon_steps)
    sim = CentroidalSim(mass, (40.0, 20.0, 10.0), sim_dt)
    sim.x[0:3] = [*xy_ref_data(0.0)[0], 1.0]
    t, ok, iters = 0.0, True, []
    mp = lambda tt: xy_motion_param(mass, tt)
    while t < end_time:
        x0 = linear_mpc_xy.to_state(mass, sim.pos[:2], sim.vel[:2], sim.angular_momentum[:2])
        scales = mpc.plan_batch(qp_solve, mp, xy_ref_data, x0[None, :], t)[0]
        assert mpc.last_result.status[0] == 0
        iters.append(int(mpc.last_result.iters[0]))
        cur = mp(t)
        ref_pos = np.array([*xy_ref_data(t)[0], cur.com_z])
        force, moment = contact.total_wrench(cur.vertex, cur.ridge, scales, sim.pos)
        ok = ok and np.linalg.norm(sim.pos - ref_pos) < 2.0 and np.linalg.norm(sim.vel) < 2.0
        ok = ok and np.linalg.norm(sim.angular_momentum) < 5.0
        t += sim_dt
        sim.update(force, moment)
    ref_pos = np.array([*xy_ref_data(t)[0], 1.0])
    return sim, ok, ref_pos, iters


def test_linear_mpc_xy_closed_loop(oracle):
    """The reference scenario with the reference's tolerances (TestLinearMpcXY.cpp:126-142), oracle QP (n = 240)."""
    import time

    t0 = time.time()
    sim, ok, ref_pos, iters = xy_closed_loop(lambda ps: oracle.qp_solve(ps))
    print(f"XY closed loop: {time.time() - t0:.1f} s, QP iterations mean {np.mean(iters):.1f} max {max(iters)}")
    assert ok
    assert np.linalg.norm(sim.pos - ref_pos) < 0.1
    assert np.linalg.norm(sim.vel) < 0.1
    assert np.linalg.norm(sim.angular_momentum) < 0.1


def test_xy_sweep_oracle_vs_host_path(oracle):
    """oracle/xy.hpp (closed-form zero-order hold of the nilpotent model, canonical condensing, obj_mat / obj_vec)
    against the independent host path (scipy expm of the augmented matrix as StateSpaceModel::calcDiscMatrix does,
    numpy condensing): the matrices agree to rounding, and the QP solutions / active sets of the two agree."""
    from centroidalcontrolcollection_b200 import workloads

    sweep = workloads.linear_mpc_xy_sweep(n_sched=6, per_sched=3)
    res = oracle.linear_mpc_xy_solve(sweep, n_threads=4)
    assert (res.status == 0).all()
    n = sweep.n
    assert n == 240 and sweep.n_eq == 15
    for s in range(sweep.S):
        idx = np.where(sweep.sched_id == s)[0]
        ps = sweep.host_problem(s, idx)
        scale = np.abs(ps.Q).max()
        assert np.abs(res.obj_mat[s] - ps.Q).max() < 1e-12 * scale
        assert np.abs(res.obj_vec[idx] - ps.c).max() < 1e-11 * np.abs(ps.c).max()
        ref = oracle.qp_solve(ps)
        assert (ref.status == 0).all()
        assert np.abs(ref.x - res.u[idx]).max() < 1e-6 * np.abs(ref.x).max()
        assert ref.active_sets() == [res.active_sets()[i] for i in idx]
    # the condensed prediction equals the iterated discrete model (TestVariantSequentialExtension's identity)
    s, b = 2, int(np.where(sweep.sched_id == 2)[0][0])
    ps = sweep.host_problem(s, [b])
    x_pred = res.A_seq[s] @ sweep.x0[b] + res.B_seq[s] @ res.u[b]
    models = [linear_mpc_xy.Model(sweep.mpc.mass, linear_mpc_xy.MotionParam(
        sweep.com_z[s, i], sweep.total_force_z[s, i], sweep.vertex[s, i, :16], sweep.ridge[s, i, :16])).calc_disc_matrix(0.1)
        for i in range(sweep.N)]
    assert np.linalg.norm(x_pred - _iterate(models, res.u[b], sweep.x0[b], False)) < 1e-9 * np.linalg.norm(x_pred)
