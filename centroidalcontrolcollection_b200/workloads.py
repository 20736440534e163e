"""Synthetic workloads of BASELINE.json's configs, shared by tests and bench.py.

Everything here is host-side input generation with fixed seeds; no solver arithmetic.
"""
import numpy as np

from .contact import contact_from_rect
from .schedule import CentroidalSchedule, SrbSchedule

EPS_T = 1e-6  # "small value to avoid numerical instability at bounds", reference tests/src/TestDdpCentroidal.cpp:38


def centroidal_weights_test():
    """WeightParam of reference tests/src/TestDdpCentroidal.cpp:28-31 over the defaults of
    include/CCC/DdpCentroidal.h:69-75."""
    w_run = np.array([1.0, 1.0, 10.0, 0, 0, 0, 1.0, 1.0, 1.0, 1e-6])
    w_term = np.array([1.0, 1.0, 10.0, 0, 0, 0, 1.0, 1.0, 1.0])
    return w_run, w_term


def ddp_centroidal_config3(batch=16384, n_sched=16, horizon_steps=50, dt=0.03, seed=20260102):
    """Config 3 (north star): DdpCentroidal, N=50, 4-phase contact schedule, cold start.

    Phases over the 1.5 s horizon: [0,0.45) rect A (m=16); [0.45,0.6) flight (m=0);
    [0.6,1.05) rect B (m=16); [1.05,1.5) A'+B (m=32).  n_sched variants jitter the phase
    boundaries by k*dt, k in {-2..1}, and shift rect B in x by {0.4,0.45,0.5,0.55}.
    ICs: c0 = (0,0,1)+N(0,0.02^2), v0 = N(0,0.05^2), L0 = N(0,0.5^2); schedules are assigned
    round-robin so that every schedule gets batch/n_sched problems.
    """
    mass = 100.0
    sched = CentroidalSchedule(n_sched, horizon_steps)
    for s in range(n_sched):
        jit = ((s % 4) - 2) * dt
        bx = 0.4 + 0.05 * ((s // 4) % 4)
        t1, t2, t3 = 0.45 + jit, 0.6 + jit, 1.05 + jit
        A = contact_from_rect((-0.1, -0.1), (0.1, 0.1))
        Bc = contact_from_rect((bx, -0.1), (bx + 0.2, 0.1))
        A2 = contact_from_rect((bx - 0.25, -0.1), (bx - 0.05, 0.1))

        def motion(t, t1=t1, t2=t2, t3=t3, A=A, Bc=Bc, A2=A2):
            t += EPS_T
            if t < t1:
                return [A]
            if t < t2:
                return []
            if t < t3:
                return [Bc]
            return [A2, Bc]

        def ref(t, t1=t1, t2=t2, bx=bx):
            t += EPS_T
            if t < t1:
                return (0.0, 0.0, 1.0)
            if t < t2:
                return (0.5 * (bx + 0.1), 0.0, 1.2)
            return (bx + 0.1, 0.0, 1.0)

        sched.sample(s, motion, ref, 0.0, dt)
    rng = np.random.Generator(np.random.PCG64(seed))
    x0 = np.zeros((batch, 9))
    x0[:, 0:3] = np.array([0.0, 0.0, 1.0]) + 0.02 * rng.standard_normal((batch, 3))
    x0[:, 3:6] = mass * 0.05 * rng.standard_normal((batch, 3))
    x0[:, 6:9] = 0.5 * rng.standard_normal((batch, 3))
    sched_id = (np.arange(batch) % n_sched).astype(np.int32)
    w_run, w_term = centroidal_weights_test()
    return dict(
        name=f"DdpCentroidal N={horizon_steps} dt={dt} 4-phase schedule x{n_sched} batch={batch}",
        mass=mass, dt=dt, N=horizon_steps, sched=sched, sched_id=sched_id, x0=x0,
        w_run=w_run, w_term=w_term, u_lo=0.0, u_hi=1e6,
    )


def ddp_centroidal_test_schedule(horizon_steps=100, dt=0.03, current_time=0.0):
    """Single schedule of reference tests/src/TestDdpCentroidal.cpp:35-76 sampled at current_time."""
    A = contact_from_rect((-0.1, -0.1), (0.1, 0.1))
    Bc = contact_from_rect((0.4, -0.1), (0.6, 0.1))

    def motion(t):
        t += EPS_T
        if t < 1.4:
            return [A]
        if t < 1.6:
            return []
        return [Bc]

    def ref(t):
        t += EPS_T
        if t < 1.4:
            return (0.0, 0.0, 1.0)
        if t < 1.6:
            return (0.25, 0.0, 1.2)
        return (0.5, 0.0, 1.0)

    sched = CentroidalSchedule(1, horizon_steps)
    sched.sample(0, motion, ref, current_time, dt)
    return sched, motion, ref


def srb_weights_test():
    """WeightParam of reference tests/src/TestDdpSingleRigidBody.cpp:28-33 over the defaults of
    include/CCC/DdpSingleRigidBody.h:89-97."""
    w_run = np.array([1.0, 1.0, 10.0, 0.5, 0.5, 0.5, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 1e-6])
    w_term = np.array([1.0, 1.0, 10.0, 0.5, 0.5, 0.5, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01])
    return w_run, w_term


def ddp_srb_test_schedule(horizon_steps=100, dt=0.03, current_time=0.0, inertia=(40.0, 20.0, 10.0)):
    """Schedule of reference tests/src/TestDdpSingleRigidBody.cpp:37-87 sampled at current_time."""
    A = contact_from_rect((-0.1, -0.5), (0.1, 0.5))
    Bc = contact_from_rect((0.4, -0.5), (0.6, 0.5))
    I = np.diag(inertia)

    def motion(t):
        t += EPS_T
        if t < 1.4:
            return [A], I
        if t < 1.6:
            return [], I
        return [Bc], I

    def ref(t):
        t += EPS_T
        if t < 1.4:
            pos = (0.0, 0.0, 1.0)
        elif t < 1.6:
            pos = (0.25, 0.0, 1.2)
        else:
            pos = (0.5, 0.0, 1.0)
        ori = (0.0, 0.0, 0.3) if 2.2 < t < 2.4 else (0.0, 0.0, 0.0)
        return pos, ori

    sched = SrbSchedule(1, horizon_steps)
    sched.sample(0, motion, ref, current_time, dt)
    return sched, motion, ref


def ddp_srb_config4(batch=65536, horizon_steps=100, dt=0.03, seed=20260103):
    """Config 4: DdpSingleRigidBody, 12 states, N=100, the reference test's A -> flight -> B schedule,
    perturbed initial states: c0 = (0,0,1)+N(0,0.02^2), euler0 = N(0,0.05^2), v0 = N(0,0.05^2),
    omega0 = N(0,0.1^2)."""
    sched, _, _ = ddp_srb_test_schedule(horizon_steps, dt, 0.0)
    rng = np.random.Generator(np.random.PCG64(seed))
    x0 = np.zeros((batch, 12))
    x0[:, 0:3] = np.array([0.0, 0.0, 1.0]) + 0.02 * rng.standard_normal((batch, 3))
    x0[:, 3:6] = 0.05 * rng.standard_normal((batch, 3))
    x0[:, 6:9] = 0.05 * rng.standard_normal((batch, 3))
    x0[:, 9:12] = 0.1 * rng.standard_normal((batch, 3))
    w_run, w_term = srb_weights_test()
    return dict(name=f"DdpSingleRigidBody N={horizon_steps} dt={dt} A-flight-B schedule batch={batch}", mass=100.0, dt=dt,
                N=horizon_steps, sched=sched, sched_id=np.zeros(batch, dtype=np.int32), x0=x0, w_run=w_run,
                w_term=w_term, u_lo=0.0, u_hi=1e6)
