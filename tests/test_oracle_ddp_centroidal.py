"""Pin the CPU oracle against the known-answer material of the reference's own tests
(reference tests/src/TestDdpCentroidal.cpp) — CPU only."""
import ctypes as C

import numpy as np

from centroidalcontrolcollection_b200 import problem, workloads
from centroidalcontrolcollection_b200.contact import contact_from_rect
from centroidalcontrolcollection_b200.schedule import CentroidalSchedule

from closed_loop import run_ddp_centroidal_closed_loop


def _kat_problem():
    # tests/src/TestDdpCentroidal.cpp:180-208: dt 0.03, mass 100, default weights, one rect, ref (0.1,-0.2,1.0)
    sched = CentroidalSchedule(1, 1)
    A = contact_from_rect((-0.1, -0.1), (0.1, 0.1))
    sched.sample(0, lambda t: [A], lambda t: (0.1, -0.2, 1.0), 0.0, 0.03)
    w_run = np.array([1.0, 1, 1, 0, 0, 0, 1, 1, 1, 1e-6])
    w_term = np.array([1.0, 1, 1, 0, 0, 0, 1, 1, 1])
    x = np.array([1.0, -2.0, 3.0, -4.0, 5.0, -6.0, 7.0, -8.0, 9.0])
    u = np.arange(1.0, 17.0)
    ps = problem.DdpCentroidalProblemSet(sched, [0], x[None, :], 100.0, 0.03, w_run, w_term)
    return ps, x, u


def _eval(oracle, ps, x, u):
    bs = ps.as_struct()
    m = 16
    out = dict(xn=np.zeros(9), rc=C.c_double(), tc=C.c_double(), Fx=np.zeros((9, 9)), Fu=np.zeros((9, m)),
               Lx=np.zeros(9), Lu=np.zeros(m), Vx=np.zeros(9))
    x = np.ascontiguousarray(x)
    u = np.ascontiguousarray(u)
    rc = oracle.lib().ccc_oracle_centroidal_eval(
        C.addressof(bs), 0, x.ctypes.data, u.ctypes.data, out["xn"].ctypes.data, C.addressof(out["rc"]),
        C.addressof(out["tc"]), out["Fx"].ctypes.data, out["Fu"].ctypes.data, out["Lx"].ctypes.data,
        out["Lu"].ctypes.data, out["Vx"].ctypes.data)
    assert rc == 0
    out["rc"], out["tc"] = out["rc"].value, out["tc"].value
    return out


def test_input_dim_is_16_for_one_rect():
    # tests/src/TestDdpCentroidal.cpp:208 — u has 16 entries for one rectangle
    ps, _, _ = _kat_problem()
    assert ps.sched.m[0, 0] == 16
    # ridges are unit vectors inside the mu = 0.5 cone, 4 per vertex
    r = ps.sched.ridge[0, 0, :16]
    assert np.allclose(np.linalg.norm(r, axis=1), 1.0)
    assert np.allclose(np.hypot(r[:, 0], r[:, 1]) / r[:, 2], 0.5)


def test_check_derivatives(oracle):
    """tests/src/TestDdpCentroidal.cpp:176-284: analytic vs central differences, eps 1e-6, tol 1e-6."""
    ps, x, u = _kat_problem()
    eps = 1e-6
    a = _eval(oracle, ps, x, u)
    Fx_num, Fu_num = np.zeros((9, 9)), np.zeros((9, 16))
    Lx_num, Lu_num, Vx_num = np.zeros(9), np.zeros(16), np.zeros(9)
    for i in range(9):
        e = np.zeros(9)
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
    """stateEq (reference src/DdpCentroidal.cpp:32-64) evaluated independently in numpy."""
    ps, x, u = _kat_problem()
    a = _eval(oracle, ps, x, u)
    r, v = ps.sched.ridge[0, 0, :16], ps.sched.vertex[0, 0, :16]
    f = (u[:, None] * r).sum(0)
    n = (u[:, None] * np.cross(v - x[None, 0:3], r)).sum(0)
    xdot = np.concatenate([x[3:6] / 100.0, f - np.array([0, 0, 100.0 * 9.80665]), n])
    assert np.allclose(a["xn"], x + 0.03 * xdot, rtol=1e-13, atol=1e-13)
    cost = 0.5 * ((x[0:3] - [0.1, -0.2, 1.0]) ** 2).sum() + 0.5 * (x[6:9] ** 2).sum() + 0.5e-6 * (u**2).sum()
    assert np.isclose(a["rc"], cost, rtol=1e-13)


def test_plan_once_closed_loop(oracle):
    """tests/src/TestDdpCentroidal.cpp:15-156 with the reference's tolerances."""
    solve = lambda ps, cfg: oracle.ddp_centroidal_solve(ps, cfg, trace_len=0, n_threads=1)
    sim, t, ref, tick_ok, iters = run_ddp_centroidal_closed_loop(solve)
    assert tick_ok
    assert np.linalg.norm(sim.pos - ref) < 0.1
    assert np.linalg.norm(sim.vel) < 0.1
    assert np.linalg.norm(sim.angular_momentum) < 0.01
    assert iters[0] > 1 and all(i == 1 for i in iters[1:])


def test_cold_start_converges_config3_sample(oracle):
    """Every problem of a config-3 sample terminates with status 1 and u within the limits."""
    w = workloads.ddp_centroidal_config3(batch=32)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    res = oracle.ddp_centroidal_solve(ps, problem.ddp_centroidal_config(), trace_len=8, n_threads=4)
    assert (res.status == 1).all()
    assert (res.u >= 0.0).all() and (res.u <= 1e6).all()
    # inputs beyond the stage dimension stay zero; flight stages have no force
    m = ps.sched.m[ps.sched_id]
    for j in range(32):
        assert (res.u[:, :, j][m <= j] == 0).all()
    # the rollout of the returned u reproduces the returned x (stateEq consistency)
    assert np.isfinite(res.x).all() and np.isfinite(res.cost).all()


def test_solution_is_a_kkt_point_of_the_reference_problem(oracle):
    """Solver-independent pin: the oracle's answer is a stationary point of the reference's box-constrained problem.

    Dynamics and costs of reference src/DdpCentroidal.cpp:32-64 / :100-135 are written out in plain numpy; the
    returned u is rolled out (must reproduce the returned x and cost), the gradient of the total cost w.r.t. every
    input comes from a numpy adjoint sweep, and its projection on the box [u_lo, u_hi] must vanish.  Nothing here
    shares code with oracle/."""
    B = 8
    ps = problem.DdpCentroidalProblemSet.from_workload(workloads.ddp_centroidal_config3(batch=B))
    res = oracle.ddp_centroidal_solve(ps, problem.ddp_centroidal_config(), trace_len=0, n_threads=4)
    N, dt, mass, sc = ps.N, ps.dt, ps.mass, ps.sched
    grav = np.array([0.0, 0.0, mass * 9.80665])
    for b in range(B):
        s, u = ps.sched_id[b], res.u[b]
        x = np.zeros((N + 1, 9))
        x[0] = ps.x0[b]
        cost = 0.0
        for k in range(N):
            m = sc.m[s, k]
            r, arm = sc.ridge[s, k, :m], sc.vertex[s, k, :m] - x[k, None, 0:3]
            force = (u[k, :m, None] * r).sum(0)
            moment = (u[k, :m, None] * np.cross(arm, r)).sum(0)
            e = x[k].copy()
            e[0:3] -= sc.ref_pos[s, k]
            cost += 0.5 * (ps.w_run[:9] * e * e).sum() + 0.5 * ps.w_run[9] * (u[k, :m] ** 2).sum()
            x[k + 1] = x[k] + dt * np.concatenate([x[k, 3:6] / mass, force - grav, moment])
        e = x[N].copy()
        e[0:3] -= sc.ref_pos[s, N]
        cost += 0.5 * (ps.w_term * e * e).sum()
        assert np.abs(x - res.x[b]).max() < 1e-11
        assert abs(cost - res.cost[b]) < 1e-12 * abs(cost)

        costate = ps.w_term * e
        grad = np.zeros_like(u)
        for k in range(N - 1, -1, -1):
            m = sc.m[s, k]
            r, arm = sc.ridge[s, k, :m], sc.vertex[s, k, :m] - x[k, None, 0:3]
            Fu = np.zeros((9, m))
            Fu[3:6], Fu[6:9] = dt * r.T, dt * np.cross(arm, r).T
            grad[k, :m] = ps.w_run[9] * u[k, :m] + Fu.T @ costate
            f = (u[k, :m, None] * r).sum(0)
            Fx = np.eye(9)
            Fx[0:3, 3:6] = dt / mass * np.eye(3)
            Fx[6:9, 0:3] = dt * np.array([[0, -f[2], f[1]], [f[2], 0, -f[0]], [-f[1], f[0], 0]])  # d(-p x f)/dp
            e = x[k].copy()
            e[0:3] -= sc.ref_pos[s, k]
            costate = ps.w_run[:9] * e + Fx.T @ costate
        proj = grad.copy()
        proj[(u <= ps.u_lo) & (grad > 0)] = 0
        proj[(u >= ps.u_hi) & (grad < 0)] = 0
        assert np.abs(grad).max() > 1e-4  # the bounds are what holds the solution: the raw gradient is not small
        assert np.abs(proj).max() < 1e-5 and np.abs(proj).max() < 1e-2 * np.abs(grad).max()


def test_oracle_matches_golden_fixture(oracle):
    """tests/golden/ddp_centroidal_config3_b16.npz (tests/golden/make_golden.py): drift pin."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddp_centroidal_config3_b16.npz"))
    w = workloads.ddp_centroidal_config3(batch=16)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    assert np.array_equal(ps.x0, g["x0"])
    res = oracle.ddp_centroidal_solve(ps, problem.ddp_centroidal_config(), trace_len=g["alpha_idx"].shape[1], n_threads=4)
    for k in ("iters", "status", "alpha_idx", "clamped", "x", "u", "cost", "lambda_trace"):
        assert np.array_equal(getattr(res, k), g[k]), k
