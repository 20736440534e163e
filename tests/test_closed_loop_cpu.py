"""CPU tier: the oracle's batched closed loop (ccc_oracle_ddp_centroidal_closed_loop) on the reference's scenario:
reference tolerances, and agreement with the tick-by-tick Python driver of tests/closed_loop.py (which samples the
callbacks per cycle and steps a numpy plant) — two independent formulations of the same loop."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import problem

from closed_loop import run_ddp_centroidal_closed_loop
from closed_loop_spec import check_reference_tolerances, reference_scenario


def test_oracle_closed_loop_reference_scenario(oracle):
    lp, ref = reference_scenario(ticks=600, batch=1)
    res = oracle.ddp_centroidal_closed_loop(lp, problem.ddp_centroidal_config(), n_threads=1)
    check_reference_tolerances(res, lp, ref)
    assert res.iters[0, 0] > 1 and (res.iters[0, 1:] == 1).all()
    # the tick-by-tick driver: same solver, callbacks re-sampled every cycle, numpy ZOH plant
    record = []
    run_ddp_centroidal_closed_loop(lambda ps, cfg: oracle.ddp_centroidal_solve(ps, cfg), end_time=600 * 0.005 - 1e-9, record=record)
    assert len(record) == 600
    pos = np.array([r[1] for r in record])
    assert np.abs(pos - res.plant[0, :600, 0:3]).max() < 1e-6
    assert [r[6] for r in record] == res.iters[0].tolist()


def test_oracle_closed_loop_batch_is_independent(oracle):
    """Problems of a batch do not interact: a batch of 3 equals three runs of 1."""
    lp, _ = reference_scenario(ticks=40, batch=3, horizon_steps=30, perturb=0.01)
    cfg = problem.ddp_centroidal_config()
    res = oracle.ddp_centroidal_closed_loop(lp, cfg, n_threads=3)
    for b in range(3):
        one, _ = reference_scenario(ticks=40, batch=3, horizon_steps=30, perturb=0.01)
        one.set_plants([0], one.plant0[b:b + 1, 0:3], one.plant0[b:b + 1, 3:6], one.plant0[b:b + 1, 6:9])
        r1 = oracle.ddp_centroidal_closed_loop(one, cfg, n_threads=1)
        assert np.array_equal(r1.plant[0], res.plant[b]) and np.array_equal(r1.iters[0], res.iters[b])


def test_loop_description_is_validated():
    """horizon_dt must be an integer multiple of sim_dt; the time grid covers the last cycle's terminal stage."""
    import pytest

    from centroidalcontrolcollection_b200 import closed_loop, workloads

    w_run, w_term = workloads.centroidal_weights_test()
    with pytest.raises(ValueError):
        closed_loop.CentroidalLoop(10, 0.03, 0.007, 5, 100.0, w_run, w_term)
    lp = closed_loop.CentroidalLoop(10, 0.03, 0.005, 5, 100.0, w_run, w_term)
    assert lp.stride == 6 and lp.grid_len == 5 - 1 + 10 * 6 + 1
    assert lp.sched.m.shape == (1, lp.grid_len)


def test_oracle_rejects_a_short_time_grid(oracle):
    lp, _ = reference_scenario(ticks=10, batch=1, horizon_steps=20)
    lp.grid_len -= 1  # one entry short of the last horizon's terminal stage
    with pytest.raises(RuntimeError):
        oracle.ddp_centroidal_closed_loop(lp, problem.ddp_centroidal_config(), n_threads=1)
