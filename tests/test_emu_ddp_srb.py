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
