"""Closed-loop drivers restating the reference's PlanOnce tests; the solver is a parameter so
the same loop pins the oracle (CPU) and checks the engine (GPU)."""
import numpy as np

from centroidalcontrolcollection_b200 import problem, workloads
from centroidalcontrolcollection_b200.contact import total_wrench

from sim_models import CentroidalSim


def run_ddp_centroidal_closed_loop(solve, end_time=3.0, horizon_steps=100, horizon_dt=0.03, record=None):
    """reference tests/src/TestDdpCentroidal.cpp:15-156.

    solve(problem_set, cfg) -> DdpResultArrays.  Returns (sim, t, ref_pos_at_end, per-tick checks ok).
    """
    sim_dt, mass = 0.005, 100.0
    sim = CentroidalSim(mass, (40.0, 20.0, 10.0), sim_dt)
    _, motion, ref = workloads.ddp_centroidal_test_schedule(horizon_steps, horizon_dt)
    w_run, w_term = workloads.centroidal_weights_test()
    sim.x[0:3] = ref(0.0)
    cfg = problem.ddp_centroidal_config()
    t, u_prev, m_prev, tick_ok, iters = 0.0, None, None, True, []
    while t < end_time:
        sched, _, _ = workloads.ddp_centroidal_test_schedule(horizon_steps, horizon_dt, t)
        x0 = np.concatenate([sim.pos, mass * sim.vel, sim.angular_momentum])[None, :]
        u_init = None
        if u_prev is not None:
            # warm start: previous u_list unshifted, re-zeroed where the input dimension changed (:102-114)
            u_init = u_prev.copy()
            u_init[0, sched.m[0] != m_prev] = 0.0
        ps = problem.DdpCentroidalProblemSet(sched, [0], x0, mass, horizon_dt, w_run, w_term, u_init=u_init)
        res = solve(ps, cfg)
        cfg.max_iter = 1  # :116
        u_prev, m_prev = res.u.copy(), sched.m[0].copy()
        iters.append(int(res.iters[0]))
        m0 = int(sched.m[0, 0])
        f, n = total_wrench(sched.vertex[0, 0, :m0], sched.ridge[0, 0, :m0], res.u[0, 0, :m0], sim.pos)
        if record is not None:
            record.append((t, sim.pos.copy(), sim.vel.copy(), sim.angular_momentum.copy(), f, n, iters[-1]))
        rp = np.array(ref(t))
        tick_ok &= np.linalg.norm(sim.pos - rp) < 2.0 and np.linalg.norm(sim.vel) < 2.0
        tick_ok &= np.linalg.norm(sim.angular_momentum) < 1.0
        t += sim_dt
        sim.update(f, n)
        if 1.0 <= t < 1.0 + sim_dt:
            # sva::ForceVecd(moment = 0, force = (0.05, 0.05, 0)) per unit mass (:23-24)
            sim.add_disturb(np.array([0.05, 0.05, 0.0]), np.zeros(3))
    return sim, t, np.array(ref(t)), bool(tick_ok), iters
