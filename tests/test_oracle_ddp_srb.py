"""Pin the CPU oracle's DdpSingleRigidBody restatement against the reference's own known-answer
material (reference tests/src/TestDdpSingleRigidBody.cpp) — CPU only."""
import ctypes as C
import math

import numpy as np

from centroidalcontrolcollection_b200 import problem, workloads
from centroidalcontrolcollection_b200.contact import contact_from_rect
from centroidalcontrolcollection_b200.schedule import SrbSchedule

from closed_loop_srb import run_ddp_srb_closed_loop


def _kat_problem():
    # tests/src/TestDdpSingleRigidBody.cpp:201-231: dt 0.03, mass 100, default weights, inertia diag(15,10,5),
    # ref pos (0.1,-0.2,1.0), ori (-0.1,0.2,-0.3), x = (1,-2,...,-12), u = (1..16)
    sched = SrbSchedule(1, 1)
    A = contact_from_rect((-0.1, -0.1), (0.1, 0.1))
    sched.sample(0, lambda t: ([A], np.diag([15.0, 10.0, 5.0])), lambda t: ((0.1, -0.2, 1.0), (-0.1, 0.2, -0.3)), 0.0, 0.03)
    w_run = np.array([1.0] * 6 + [0.01] * 6 + [1e-6])
    w_term = np.array([1.0] * 6 + [0.01] * 6)
    x = np.array([1.0, -2.0, 3.0, -4.0, 5.0, -6.0, 7.0, -8.0, 9.0, -10.0, 11.0, -12.0])
    u = np.arange(1.0, 17.0)
    ps = problem.DdpSrbProblemSet(sched, [0], x[None, :], 100.0, 0.03, w_run, w_term)
    return ps, x, u


def _eval(oracle, ps, x, u):
    bs = ps.as_struct()
    m = 16
    out = dict(xn=np.zeros(12), rc=C.c_double(), tc=C.c_double(), Fx=np.zeros((12, 12)), Fu=np.zeros((12, m)),
               Lx=np.zeros(12), Lu=np.zeros(m), Vx=np.zeros(12))
    x, u = np.ascontiguousarray(x), np.ascontiguousarray(u)
    rc = oracle.lib().ccc_oracle_srb_eval(
        C.addressof(bs), 0, x.ctypes.data, u.ctypes.data, out["xn"].ctypes.data, C.addressof(out["rc"]),
        C.addressof(out["tc"]), out["Fx"].ctypes.data, out["Fu"].ctypes.data, out["Lx"].ctypes.data,
        out["Lu"].ctypes.data, out["Vx"].ctypes.data)
    assert rc == 0
    out["rc"], out["tc"] = out["rc"].value, out["tc"].value
    return out


def test_sincos_canon_accuracy(oracle):
    """The oracle's sin/cos (fma-only, reproducible on the GPU) agree with libm to ~1 ulp."""
    L = oracle.lib()
    L.ccc_oracle_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    s, c = C.c_double(), C.c_double()
    xs = np.concatenate([np.linspace(-12, 12, 4001), np.random.default_rng(0).uniform(-500, 500, 4000), [0.0, -0.0]])
    for x in xs:
        L.ccc_oracle_sincos(float(x), C.byref(s), C.byref(c))
        assert abs(s.value - math.sin(x)) <= 2.3e-16 and abs(c.value - math.cos(x)) <= 2.3e-16


def test_check_derivatives(oracle):
    """tests/src/TestDdpSingleRigidBody.cpp:197-308: analytic vs central differences, eps 1e-6, tol 1e-6."""
    ps, x, u = _kat_problem()
    eps = 1e-6
    a = _eval(oracle, ps, x, u)
    Fx_num, Fu_num = np.zeros((12, 12)), np.zeros((12, 16))
    Lx_num, Lu_num, Vx_num = np.zeros(12), np.zeros(16), np.zeros(12)
    for i in range(12):
        e = np.zeros(12)
        e[i] = eps
        p, q = _eval(oracle, ps, x + e, u), _eval(oracle, ps, x - e, u)
        Fx_num[:, i] = (p["xn"] - q["xn"]) / (2 * eps)
        Lx_num[i] = (p["rc"] - q["rc"]) / (2 * eps)
        Vx_num[i] = (p["tc"] - q["tc"]) / (2 * eps)
    for i in range(16):
        e = np.zeros(16)
        e[i] = eps
        p, q = _eval(oracle, ps, x, u + e), _eval(oracle, ps, x, u - e)
        Fu_num[:, i] = (p["xn"] - q["xn"]) / (2 * eps)
        Lu_num[i] = (p["rc"] - q["rc"]) / (2 * eps)
    assert np.linalg.norm(a["Fx"] - Fx_num) < 1e-6
    assert np.linalg.norm(a["Fu"] - Fu_num) < 1e-6
    assert np.linalg.norm(a["Lx"] - Lx_num) < 1e-6
    assert np.linalg.norm(a["Lu"] - Lu_num) < 1e-6
    assert np.linalg.norm(a["Vx"] - Vx_num) < 1e-6


def test_state_eq_matches_plain_numpy(oracle):
    """stateEq (reference src/DdpSingleRigidBody.cpp:52-91) evaluated independently in numpy with libm."""
    ps, x, u = _kat_problem()
    a = _eval(oracle, ps, x, u)
    r, v = ps.sched.ridge[0, 0, :16], ps.sched.vertex[0, 0, :16]
    I = np.diag([15.0, 10.0, 5.0])
    ca, sa, cb, sb = math.cos(x[3]), math.sin(x[3]), math.cos(x[4]), math.sin(x[4])
    E = np.array([[ca * sb / cb, sb * sa / cb, 1.0], [-sa, ca, 0.0], [ca / cb, sa / cb, 0.0]])
    w = x[9:12]
    f = (u[:, None] * r).sum(0)
    n = (u[:, None] * np.cross(v - x[None, 0:3], r)).sum(0)
    wdot = np.linalg.solve(I, -np.cross(w, I @ w) + n)
    xdot = np.concatenate([x[6:9], E @ w, f / 100.0 - np.array([0, 0, 9.80665]), wdot])
    assert np.allclose(a["xn"], x + 0.03 * xdot, rtol=1e-12, atol=1e-12)


def test_solution_is_a_kkt_point_of_the_reference_problem(oracle):
    """Solver-independent pin: dynamics and costs of reference src/DdpSingleRigidBody.cpp:52-112 in plain numpy
    (libm sin/cos, numpy solve for the inertia), over a horizon that spans contact A, the flight phase and contact
    B.  The returned inputs must reproduce the returned cost, and the central-difference gradient of the total
    cost w.r.t. every input, projected on the box [u_lo, u_hi], must vanish."""
    N, dt, mass = 30, 0.03, 100.0
    sched, _, _ = workloads.ddp_srb_test_schedule(N, dt, 0.9)
    assert set(sched.m[0]) == {0, 16}
    w_run, w_term = workloads.srb_weights_test()
    rng = np.random.Generator(np.random.PCG64(7))
    x0 = np.zeros((2, 12))
    x0[:, 2] = 1.0
    x0[1] += np.concatenate([0.02 * rng.standard_normal(3), 0.05 * rng.standard_normal(6), 0.1 * rng.standard_normal(3)])
    ps = problem.DdpSrbProblemSet(sched, [0, 0], x0, mass, dt, w_run, w_term)
    res = oracle.ddp_srb_solve(ps, problem.ddp_srb_config())
    assert (res.status == 1).all()

    def total_cost(x_start, U):  # U[P][N][m_max]: P input sequences rolled out side by side
        x, c = np.tile(x_start, (U.shape[0], 1)), np.zeros(U.shape[0])
        for k in range(N):
            m = sched.m[0, k]
            r, v, inertia = sched.ridge[0, k, :m], sched.vertex[0, k, :m], sched.inertia[0, k].reshape(3, 3)
            u = U[:, k, :m]
            e = x.copy()
            e[:, 0:6] -= sched.ref[0, k]
            c += 0.5 * (w_run[:12] * e * e).sum(1) + 0.5 * w_run[12] * (u * u).sum(1)
            ca, sa, cb, sb = np.cos(x[:, 3]), np.sin(x[:, 3]), np.cos(x[:, 4]), np.sin(x[:, 4])
            om = x[:, 9:12]
            euler_dot = np.stack([ca * sb / cb * om[:, 0] + sb * sa / cb * om[:, 1] + om[:, 2],
                                  -sa * om[:, 0] + ca * om[:, 1], ca / cb * om[:, 0] + sa / cb * om[:, 1]], 1)
            moment = (u[:, :, None] * np.cross(v[None] - x[:, None, 0:3], r[None])).sum(1)
            om_dot = np.linalg.solve(inertia, (-np.cross(om, om @ inertia.T) + moment).T).T
            x = x + dt * np.concatenate([x[:, 6:9], euler_dot, (u @ r) / mass - np.array([0, 0, 9.80665]), om_dot], 1)
        e = x.copy()
        e[:, 0:6] -= sched.ref[0, N]
        return c + 0.5 * (w_term * e * e).sum(1)

    idx = [(k, j) for k in range(N) for j in range(sched.m[0, k])]
    for b in range(2):
        u = res.u[b]
        cost = total_cost(x0[b], u[None])[0]
        assert abs(cost - res.cost[b]) < 1e-12 * cost
        up, um, h = np.tile(u, (len(idx), 1, 1)), np.tile(u, (len(idx), 1, 1)), np.zeros(len(idx))
        for i, (k, j) in enumerate(idx):
            h[i] = 1e-4 * max(1.0, abs(u[k, j]))
            up[i, k, j] += h[i]
            um[i, k, j] -= h[i]
        grad = (total_cost(x0[b], up) - total_cost(x0[b], um)) / (2 * h)
        at = np.array([u[k, j] for k, j in idx])
        proj = grad.copy()
        proj[(at <= ps.u_lo) & (grad > 0)] = 0
        proj[(at >= ps.u_hi) & (grad < 0)] = 0
        assert (at <= ps.u_lo).any() and np.abs(grad).max() > 1e-5
        assert np.abs(proj).max() < 1e-6 and np.abs(proj).max() < 1e-2 * np.abs(grad).max()


def test_plan_once_closed_loop(oracle):
    """tests/src/TestDdpSingleRigidBody.cpp:15-175 with the reference's tolerances."""
    sim, rp, ro, tick_ok, iters = run_ddp_srb_closed_loop(lambda ps, cfg: oracle.ddp_srb_solve(ps, cfg))
    assert tick_ok
    assert np.linalg.norm(sim.x[0:3] - rp) < 0.1
    assert np.linalg.norm(sim.x[3:6] - ro) < 0.1
    assert np.linalg.norm(sim.x[6:9]) < 0.1
    assert np.linalg.norm(sim.x[9:12]) < 0.1
    assert iters[0] > 2 and all(i == 1 for i in iters[1:])  # max_iter = 1 after the first tick (:133)
