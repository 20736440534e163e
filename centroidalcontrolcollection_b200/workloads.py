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


# ---- QP-based ZMP methods (configs 2 and 5) --------------------------------------------------------
def _walking_limits(step_length, step_width, t0, horizon_steps, horizon_dt, eps_reps):
    """ZMP limits and reference ZMP of the six-step walking plan (reference
    tests/src/TestLinearMpcZmp.cpp:30-41) over a horizon starting at t0.  Self-contained restatement of the
    schedule the reference's FootstepManager produces (tests/footstep_manager.py is the fixture copy used to
    cross-check it): returns (ref_zmp[N][2], lim_min[N][2], lim_max[N][2])."""
    foot_size = np.array([0.1, 0.05])
    L, w = step_length, 0.5 * step_width
    stance = {0: np.array([0.0, w]), 1: np.array([0.0, -w])}
    steps = [(0, (L, w), 2.0), (1, (2 * L, -w), 3.0), (0, (3 * L, w), 4.0), (1, (4 * L, -w), 5.0), (0, (3 * L, w), 6.0),
             (1, (3 * L, -w), 7.0)]
    transit, swing = 0.2, 0.8
    mid = lambda st: next(iter(st.values())) if len(st) == 1 else 0.5 * (st[0] + st[1])
    zmp_knots, stance_knots = [(-1e9, mid(stance))], [(-1e9, dict(stance))]
    tmp = dict(stance)
    for foot, pos, ts in steps:
        pos = np.array(pos)
        t_sw0, t_sw1, t_end = ts + 0.5 * transit, ts + 0.5 * transit + swing, ts + transit + swing
        zmp_knots.append((ts, mid(tmp)))
        stance_knots.append((ts, dict(tmp)))
        tmp.pop(foot)
        zmp_knots.append((t_sw0, tmp[1 - foot].copy()))
        stance_knots.append((t_sw0, dict(tmp)))
        tmp[foot] = pos
        zmp_knots.append((t_sw1, tmp[1 - foot].copy()))
        stance_knots.append((t_sw1, dict(tmp)))
        zmp_knots.append((t_end, mid(tmp)))
    zmp_knots.append((1e9, mid(tmp)))
    ref, lo, hi = np.zeros((horizon_steps, 2)), np.zeros((horizon_steps, 2)), np.zeros((horizon_steps, 2))
    zt = [k[0] for k in zmp_knots]
    stt = [k[0] for k in stance_knots]
    for i in range(horizon_steps):
        t = t0 + i * horizon_dt + eps_reps * EPS_T
        j = int(np.searchsorted(zt, t, side="right"))
        (ta, za), (tb, zb) = zmp_knots[j - 1], zmp_knots[j]
        ratio = 0.0 if tb - ta > 1e8 else (t - ta) / (tb - ta)
        ref[i] = (1 - ratio) * za + ratio * zb
        st = stance_knots[int(np.searchsorted(stt, t, side="right")) - 1][1]
        pts = np.array(list(st.values()))
        lo[i], hi[i] = pts.min(axis=0) - 0.5 * foot_size, pts.max(axis=0) + 0.5 * foot_size
    return ref, lo, hi


def linear_mpc_zmp_config2(batch=4096, t0=1.8, seed=20260101):
    """Config 2: LinearMpcZmp, h = 1.0, N = 100, dt = 0.01, control_dt = 0.005, one shared footstep plan sampled
    at t0 (the horizon contains a support change), perturbed ICs around mid-stance."""
    h, N, dt = 1.0, 100, 0.01
    _, lo, hi = _walking_limits(0.2, 0.2, t0, N, dt, eps_reps=2)
    rng = np.random.Generator(np.random.PCG64(seed))
    mid = np.array([0.0, 0.0])
    pos = mid + rng.uniform(-0.03, 0.03, (batch, 2))
    vel = rng.uniform(-0.15, 0.15, (batch, 2))
    acc = 9.80665 / h * (pos - mid)
    return dict(name=f"LinearMpcZmp N={N} dt={dt} batch={batch} (x2 axes)", com_height=h, horizon_duration=N * dt,
                horizon_dt=dt, control_dt=0.005, pos=pos, vel=vel, acc=acc, lim_min=np.tile(lo[None], (batch, 1, 1)),
                lim_max=np.tile(hi[None], (batch, 1, 1)))


def ismpc_config5(n_plans=256, n_perturb=512, t0=1.8, seed=20260104):
    """Config 5: IntrinsicallyStableMpc, h = 1.0, N = 100 (2.0 / 0.02), control_dt = 0.005; n_plans footstep plans
    (16 step lengths x 16 step widths for the default 256) x n_perturb capture-point perturbations."""
    h, N, dt = 1.0, 100, 0.02
    rng = np.random.Generator(np.random.PCG64(seed))
    side = int(round(np.sqrt(n_plans)))
    lengths, widths = rng.uniform(0.1, 0.3, side), rng.uniform(0.16, 0.24, max(n_plans // side, 1))
    refs, los, his = [], [], []
    for p in range(n_plans):
        r, lo, hi = _walking_limits(lengths[p % side], widths[(p // side) % len(widths)], t0, N, dt, eps_reps=2)
        refs.append(r), los.append(lo), his.append(hi)
    plan = np.repeat(np.arange(n_plans), n_perturb)
    B = n_plans * n_perturb
    cp = rng.uniform(-0.04, 0.04, (B, 2))
    return dict(name=f"IntrinsicallyStableMpc N={N} dt={dt} {n_plans} plans x {n_perturb} perturbations (x2 axes)",
                com_height=h, horizon_duration=N * dt, horizon_dt=dt, control_dt=0.005, capture_point=cp,
                planned_zmp=np.zeros((B, 2)), ref_zmp=np.array(refs)[plan], lim_min=np.array(los)[plan],
                lim_max=np.array(his)[plan])


def linear_mpc_xy_batch(batch=256, horizon_steps=15, t0=2.8, seed=20260105):
    """LinearMpcXY problems on the reference test's schedule (tests/src/TestLinearMpcXY.cpp:17-83: mass 100,
    horizon_dt 0.1, single rectangle contact that changes at t = 3, 4, 5, 6 s) sampled at t0, with `batch`
    perturbed initial states: pos = ref + U(-0.03, 0.03)^2, vel = U(-0.15, 0.15)^2, L_xy = U(-0.5, 0.5)^2.
    n = 16 x horizon_steps decision variables (240 at the reference's horizon)."""
    from . import contact, linear_mpc_xy
    from .linear_models import G

    mass, horizon_dt = 100.0, 0.1

    def motion_param(t):
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

    def ref_data(t):
        pos = (1.0, 0.0) if t < 3.0 else (1.0, 0.1) if t < 4.0 else (1.25, -0.1) if t < 5.0 else (1.5, 0.1) if t < 6.0 else (1.5, 0.0)
        return np.array(pos), np.zeros(2), np.zeros(2)

    rng = np.random.default_rng(seed)
    pos = ref_data(t0)[0][None, :] + rng.uniform(-0.03, 0.03, (batch, 2))
    vel = rng.uniform(-0.15, 0.15, (batch, 2))
    am = rng.uniform(-0.5, 0.5, (batch, 2))
    return {"name": "LinearMpcXY test schedule", "mass": mass, "horizon_dt": horizon_dt, "horizon_steps": horizon_steps,
            "t0": t0, "motion_param_func": motion_param, "ref_data_func": ref_data,
            "x0": linear_mpc_xy.to_state(mass, pos, vel, am)}


def linear_mpc_xy_problem_set(horizon_steps=15, batch=1184):
    """The assembled QP batch (qp.QpProblemSet) of `batch` LinearMpcXY problems on the reference test's schedule:
    n = 16 x horizon_steps variables, one equality per stage, 2n bound rows (reference src/LinearMpcXY.cpp:116-182)."""
    from . import linear_mpc_xy

    w = linear_mpc_xy_batch(batch=batch, horizon_steps=horizon_steps)
    mpc = linear_mpc_xy.LinearMpcXY(w["mass"], w["horizon_dt"], horizon_steps)
    ts = [w["t0"] + i * w["horizon_dt"] for i in range(horizon_steps)]
    ref = np.concatenate([linear_mpc_xy.to_state(w["mass"], *w["ref_data_func"](t)) for t in ts])
    return mpc.build_qp([w["motion_param_func"](t) for t in ts], ref, w["x0"])


def linear_mpc_xy_sweep(n_sched=64, per_sched=16, horizon_steps=15, seed=20260106):
    """A sweep of contact / reference schedules for LinearMpcXY (linear_mpc_xy.XySweepProblemSet): the reference
    test's scenario (tests/src/TestLinearMpcXY.cpp:17-83) sampled at n_sched start times t0 = 2.0 + 4.2 s / n_sched
    (so the horizons see every contact change), each with its own foot-width scale U(0.8, 1.2), CoM height
    U(0.9, 1.1) and vertical force (1 + U(-0.1, 0.1)) m g per stage, times per_sched perturbed initial states as
    in linear_mpc_xy_batch.  Every stage has one rectangle: n = 16 x horizon_steps, one equality per stage."""
    from . import contact, linear_mpc_xy
    from .linear_models import G

    mass, horizon_dt = 100.0, 0.1
    rng = np.random.default_rng(seed)
    mpc = linear_mpc_xy.LinearMpcXY(mass, horizon_dt, horizon_steps)
    sweep = linear_mpc_xy.XySweepProblemSet(mpc, horizon_steps, n_sched, m_max=16)
    x0, sid = [], []
    for s in range(n_sched):
        t0 = 2.0 + 4.2 * s / n_sched
        wy, cz = rng.uniform(0.8, 1.2), rng.uniform(0.9, 1.1)
        fz_scale = 1.0 + rng.uniform(-0.1, 0.1, horizon_steps)

        def motion_param(t, wy=wy, cz=cz, fz_scale=fz_scale, t0=t0):
            if t < 3.0:
                rect = ((0.9, -0.15 * wy), (1.1, 0.15 * wy))
            elif t < 4.0:
                rect = ((0.9, 0.05 * wy), (1.1, 0.15 * wy))
            elif t < 5.0:
                rect = ((1.15, -0.15 * wy), (1.35, -0.05 * wy))
            elif t < 6.0:
                rect = ((1.4, 0.05 * wy), (1.6, 0.15 * wy))
            else:
                rect = ((1.4, -0.15 * wy), (1.6, 0.15 * wy))
            vertex, ridge = contact.contact_from_rect(*rect)
            k = min(max(int(round((t - t0) / horizon_dt)), 0), horizon_steps - 1)
            return linear_mpc_xy.MotionParam(cz, fz_scale[k] * mass * G, vertex, ridge)

        def ref_data(t, wy=wy):
            pos = (1.0, 0.0) if t < 3.0 else (1.0, 0.1 * wy) if t < 4.0 else (1.25, -0.1 * wy) if t < 5.0 else (1.5, 0.1 * wy) if t < 6.0 else (1.5, 0.0)
            return np.array(pos), np.zeros(2), np.zeros(2)

        sweep.sample(s, motion_param, ref_data, t0)
        pos = ref_data(t0)[0][None, :] + rng.uniform(-0.03, 0.03, (per_sched, 2))
        vel = rng.uniform(-0.15, 0.15, (per_sched, 2))
        am = rng.uniform(-0.5, 0.5, (per_sched, 2))
        x0.append(linear_mpc_xy.to_state(mass, pos, vel, am))
        sid.append(np.full(per_sched, s, dtype=np.int32))
    sweep.set_initial_states(np.concatenate(x0), np.concatenate(sid))
    return sweep


def ddp_zmp_batch(batch=65536, horizon_steps=100, dt=0.02, seed=20260107):
    """DdpZmp problems as the reference's test loop poses them (tests/src/TestDdpZmp.cpp:15-135: mass 100, horizon
    2 s / 0.02 s, the six-step walking plan): 4 schedules = the plan's reference ZMP sampled at t0 = 0, 1.9, 2.4 and
    4.95 s, `batch` perturbed states (pos +- 0.03 around the reference ZMP, vel +- 0.1, height 1 +- 0.02) and the
    warm start of a controller that held the CoM over the ZMP (u = (c_x, c_y, m g))."""
    from .schedule import FootstepPlans

    G = 9.80665
    times = (0.0, 1.9, 2.4, 4.95)
    # reference ZMP on the horizon grid through the schedule compiler's definition (_walking_limits: one epsilon, refZmp)
    ref_zmp = np.zeros((len(times), horizon_steps + 1, 3))
    for s, t0 in enumerate(times):
        r, _, _ = _walking_limits(0.2, 0.2, t0, horizon_steps + 1, dt, eps_reps=1)
        ref_zmp[s, :, :2] = r
    rng = np.random.default_rng(seed)
    sched_id = (np.arange(batch) % len(times)).astype(np.int32)
    x0 = np.zeros((batch, 6))
    x0[:, [0, 2]] = ref_zmp[sched_id, 0, :2] + rng.uniform(-0.03, 0.03, (batch, 2))
    x0[:, [1, 3]] = rng.uniform(-0.1, 0.1, (batch, 2))
    x0[:, 4] = 1.0 + rng.uniform(-0.02, 0.02, batch)
    u_init = np.zeros((batch, horizon_steps, 3))
    u_init[:, :, 0], u_init[:, :, 1], u_init[:, :, 2] = x0[:, None, 0], x0[:, None, 2], 100.0 * G
    return dict(name=f"DdpZmp N={horizon_steps} dt={dt} 4 walking-plan schedules batch={batch} warm start max_iter=3", mass=100.0, dt=dt,
                ref_zmp=ref_zmp, com_z=np.ones((len(times), horizon_steps + 1)), sched_id=sched_id, x0=x0, u_init=u_init)
