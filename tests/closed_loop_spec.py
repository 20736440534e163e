"""The reference's DdpCentroidal closed-loop scenario (tests/src/TestDdpCentroidal.cpp:15-156) as a CentroidalLoop
(centroidalcontrolcollection_b200/closed_loop.py): shared by the CPU tier (oracle) and the GPU tier (engine)."""
import numpy as np

from centroidalcontrolcollection_b200 import closed_loop, workloads


def reference_scenario(ticks=600, batch=1, horizon_steps=100, perturb=0.0, seed=3):
    horizon_dt, sim_dt, mass = 0.03, 0.005, 100.0
    _, motion, ref = workloads.ddp_centroidal_test_schedule(horizon_steps, horizon_dt)
    w_run, w_term = workloads.centroidal_weights_test()
    lp = closed_loop.CentroidalLoop(horizon_steps, horizon_dt, sim_dt, ticks, mass, w_run, w_term)
    lp.sample(0, motion, ref, 0.0)
    rng = np.random.default_rng(seed)
    pos = np.tile(np.array(ref(0.0)), (batch, 1)) + perturb * rng.standard_normal((batch, 3))
    vel = 2.5 * perturb * rng.standard_normal((batch, 3))
    lp.set_plants(np.zeros(batch, dtype=np.int32), pos, vel, np.zeros((batch, 3)))
    lp.set_disturbance(199, [0.05, 0.05, 0.0])  # the cycle after which t lands in [1.0, 1.0 + sim_dt) (:145-149)
    return lp, ref


def check_reference_tolerances(res, lp, ref):
    """Per-cycle and final checks of the reference test (:133-135, :154-156)."""
    for b in range(lp.batch):
        for tick in range(lp.ticks):
            st = res.plant[b, tick]
            assert np.linalg.norm(st[0:3] - np.array(ref(tick * lp.sim_dt))) < 2.0
            assert np.linalg.norm(st[3:6]) < 2.0 and np.linalg.norm(st[6:9]) < 1.0
        end = res.plant[b, -1]
        assert np.linalg.norm(end[0:3] - np.array(ref(lp.ticks * lp.sim_dt))) < 0.1
        assert np.linalg.norm(end[3:6]) < 0.1 and np.linalg.norm(end[6:9]) < 0.01
