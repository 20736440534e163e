"""CPU tier: the SRB instantiation of the CUDA solver core, run by the warp emulator, reproduces the
oracle bit for bit (sin/cos included: both sides use the same fma-only sincos)."""
import numpy as np

from centroidalcontrolcollection_b200 import problem, workloads

import emu_lib
from parity import assert_ddp_parity


def test_srb_short_horizon(oracle):
    w = workloads.ddp_srb_config4(batch=2, horizon_steps=8)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    cfg = problem.ddp_srb_config(max_iter=6)
    assert_ddp_parity(oracle.ddp_srb_solve(ps, cfg, trace_len=8), emu_lib.ddp_srb_solve(ps, cfg, trace_len=8, chunk=2))


def test_srb_flight_and_large_angles(oracle):
    """Stages with no input (flight) and Euler angles / rates far from zero."""
    sched, _, _ = workloads.ddp_srb_test_schedule(horizon_steps=14, dt=0.03, current_time=1.25)
    assert 0 in sched.m and 16 in sched.m
    w_run, w_term = workloads.srb_weights_test()
    x0 = np.array([[0.1, -0.05, 1.05, 0.9, -0.6, 0.4, 0.2, -0.1, 0.1, 1.0, -2.0, 1.5]])
    ps = problem.DdpSrbProblemSet(sched, [0], x0, 100.0, 0.03, w_run, w_term)
    cfg = problem.ddp_srb_config(max_iter=3)
    ref = oracle.ddp_srb_solve(ps, cfg, trace_len=4)
    for feat in (1, 2):  # product default; every feature bit (TMA-staged gains: 14-double rows, 2-slot ring)
        assert_ddp_parity(ref, emu_lib.ddp_srb_solve(ps, cfg, trace_len=4, feat=feat))


def test_srb_two_contacts_32_inputs(oracle):
    """Two simultaneous rectangle contacts: 32 ridge inputs per stage, then one contact (16), then flight (0).  The SRB
    instantiation runs the compact-code forms of the Quu assembly / Quu K loops (rolled over the columns, through the tile):
    this is their only case beyond 16 columns, with a change of the input dimension between stages on top."""
    from centroidalcontrolcollection_b200.contact import contact_from_rect
    from centroidalcontrolcollection_b200.schedule import SrbSchedule

    A = contact_from_rect((-0.1, -0.5), (0.1, -0.1))
    Bc = contact_from_rect((-0.1, 0.1), (0.1, 0.5))
    I = np.array([[40.0, 1.0, -0.5], [1.0, 20.0, 0.3], [-0.5, 0.3, 10.0]])

    def motion(t):
        if t < 0.1:
            return [A, Bc], I
        if t < 0.2:
            return [A], I
        return [], I

    def ref(t):
        return (0.25, 0.4, 1.0), (0.0, 0.0, 0.05 if t > 0.1 else 0.0)  # far to the side: ridges on the near side unload

    sched = SrbSchedule(1, 9)
    sched.sample(0, motion, ref, 0.0, 0.03)
    assert set(np.unique(sched.m)) == {0, 16, 32}
    w_run, w_term = workloads.srb_weights_test()
    x0 = np.array([[0.02, -0.01, 1.01, 0.05, -0.03, 0.02, 0.1, -0.6, 0.02, 0.8, -0.1, 0.1],
                   [-0.01, 0.02, 0.98, -0.04, 0.06, -0.02, -0.4, 0.5, 0.0, -0.1, 0.9, -0.05]])
    ps = problem.DdpSrbProblemSet(sched, np.zeros(2, np.int32), x0, 100.0, 0.03, w_run, w_term)
    cfg = problem.ddp_srb_config(max_iter=5)
    ref_res = oracle.ddp_srb_solve(ps, cfg, trace_len=8)
    assert (ref_res.clamped != 0).any()
    assert_ddp_parity(ref_res, emu_lib.ddp_srb_solve(ps, cfg, trace_len=8))
    assert_ddp_parity(ref_res, emu_lib.ddp_srb_solve(ps, cfg, trace_len=8, team=1))
