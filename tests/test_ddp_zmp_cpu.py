"""CPU tier for CCC::DdpZmp: oracle pinned on the reference's known-answer material
(tests/src/TestDdpZmp.cpp), and the ZMP instantiation of the CUDA core under the warp emulator."""
import ctypes as C

import numpy as np
import pytest

from centroidalcontrolcollection_b200 import problem

import emu_lib
from footstep_manager import walking_plan
from parity import assert_ddp_parity
from sim_models import ComZmpSim3d

G = 9.80665


def _eval(oracle, ps, x, u):
    L = oracle.lib()
    L.ccc_oracle_zmp_eval.restype = C.c_int32
    L.ccc_oracle_zmp_eval.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
    bs = ps.as_struct()
    out = dict(xn=np.zeros(6), rc=C.c_double(), tc=C.c_double(), Fx=np.zeros((6, 6)), Fu=np.zeros((6, 3)), Lx=np.zeros(6),
               Lu=np.zeros(3), Vx=np.zeros(6))
    x, u = np.ascontiguousarray(x), np.ascontiguousarray(u)
    assert L.ccc_oracle_zmp_eval(C.addressof(bs), 0, x.ctypes.data, u.ctypes.data, out["xn"].ctypes.data,
                                 C.addressof(out["rc"]), C.addressof(out["tc"]), out["Fx"].ctypes.data, out["Fu"].ctypes.data,
                                 out["Lx"].ctypes.data, out["Lu"].ctypes.data, out["Vx"].ctypes.data) == 0
    out["rc"], out["tc"] = out["rc"].value, out["tc"].value
    return out


def test_check_derivatives(oracle):
    """tests/src/TestDdpZmp.cpp:153-248: dt 0.005, ref zmp (0.1,-0.2,0.3), com_z 1.0, x, u literals; tol 1e-6."""
    ref_zmp = np.tile(np.array([0.1, -0.2, 0.3]), (1, 2, 1))
    com_z = np.ones((1, 2))
    x = np.array([1.0, 0.1, -2.1, -0.5, 1.1, 0.5])
    u = np.array([1.0, -2.0, 1000.0])
    ps = problem.DdpZmpProblemSet(ref_zmp, com_z, [0], x[None], 100.0, 0.005)
    eps = 1e-6
    a = _eval(oracle, ps, x, u)
    Fx, Fu, Lx, Lu, Vx = np.zeros((6, 6)), np.zeros((6, 3)), np.zeros(6), np.zeros(3), np.zeros(6)
    for i in range(6):
        e = np.zeros(6)
        e[i] = eps
        p, q = _eval(oracle, ps, x + e, u), _eval(oracle, ps, x - e, u)
        Fx[:, i], Lx[i], Vx[i] = (p["xn"] - q["xn"]) / (2 * eps), (p["rc"] - q["rc"]) / (2 * eps), (p["tc"] - q["tc"]) / (2 * eps)
    for i in range(3):
        e = np.zeros(3)
        e[i] = eps
        p, q = _eval(oracle, ps, x, u + e), _eval(oracle, ps, x, u - e)
        Fu[:, i], Lu[i] = (p["xn"] - q["xn"]) / (2 * eps), (p["rc"] - q["rc"]) / (2 * eps)
    for k, num in (("Fx", Fx), ("Fu", Fu), ("Lx", Lx), ("Lu", Lu), ("Vx", Vx)):
        assert np.linalg.norm(a[k] - num) < 1e-6, k
    # running cost against the reference formula (src/DdpZmp.cpp:20-27)
    rc = 1e2 * 0.5 * (x[4] - 1.0) ** 2 + 1e-1 * 0.5 * ((u[:2] - [0.1, -0.2]) ** 2).sum() + 1e-4 * 0.5 * (u[2] - 100 * G) ** 2
    assert np.isclose(a["rc"], rc, rtol=1e-13)


def run_ddp_zmp_closed_loop(solve, end_time=10.0):
    """tests/src/TestDdpZmp.cpp:15-137.  solve(problem_set, cfg) -> DdpResultArrays."""
    horizon_dt, N, sim_dt, mass, h = 0.02, 100, 0.005, 100.0, 1.0
    fm = walking_plan()
    sim = ComZmpSim3d(mass, sim_dt)
    sim.z[0] = h
    cfg = problem.ddp_config(max_iter=3)  # :28 (nmpc_ddp defaults otherwise)
    t, u_prev, ok, planned = 0.0, None, True, None
    while t < end_time:
        fm.update(t)
        ref_zmp = np.zeros((1, N + 1, 3))
        for k in range(N + 1):
            ref_zmp[0, k, :2] = fm.ref_zmp(t + k * horizon_dt)
        com_z = np.full((1, N + 1), h)
        x0 = np.array([[sim.x[0], sim.x[1], sim.y[0], sim.y[1], sim.z[0], sim.z[1]]])
        u_init = np.tile(np.array([sim.x[0], sim.y[0], mass * G]), (1, N, 1)) if u_prev is None else u_prev  # :84-93
        ps = problem.DdpZmpProblemSet(ref_zmp, com_z, [0], x0, mass, horizon_dt, u_init=u_init)
        res = solve(ps, cfg)
        u_prev = res.u.copy()
        planned = res.u[0, 0]
        rz = fm.ref_zmp(t)
        ok &= bool(np.linalg.norm(planned[:2] - rz) < 0.1 and abs(sim.z[0] - h) < 0.1)  # :108-109
        t += sim_dt
        sim.update(planned[:2], planned[2])
        for dtm in (4.5, 8.5):
            if dtm <= t < dtm + sim_dt:
                sim.add_disturb(np.array([0.05, 0.05]))
    return ok, planned, sim, fm.ref_zmp(t)


def test_closed_loop(oracle):
    ok, planned, sim, rz = run_ddp_zmp_closed_loop(lambda ps, cfg: oracle.ddp_zmp_solve(ps, cfg))
    assert ok
    assert np.linalg.norm(planned[:2] - rz) < 1e-2          # :131
    assert abs(sim.z[0] - 1.0) < 1e-2                        # :132
    assert np.linalg.norm(sim.pos[:2] - rz) < 1e-2           # :133
    assert np.linalg.norm(sim.vel) < 1e-2                    # :134


def test_solution_is_a_stationary_point_of_the_reference_problem(oracle):
    """Solver-independent pin: total cost of reference src/DdpZmp.cpp:8-45 written out in plain numpy, rolled out
    from the returned inputs (must give the returned cost); its central-difference gradient w.r.t. every input is
    two orders of magnitude or more below the gradient at the initial guess (DdpZmp has no input limits)."""
    fm = walking_plan()
    fm.update(1.9)
    N, mass, dt = 20, 100.0, 0.02
    ref_zmp = np.zeros((1, N + 1, 3))
    for k in range(N + 1):
        ref_zmp[0, k, :2] = fm.ref_zmp(1.9 + k * dt)
    x0 = np.array([[0.01, 0.1, -0.02, -0.05, 1.02, 0.03], [0.0, 0.0, 0.0, 0.0, 1.0, 0.0]])
    u_init = np.tile(np.array([0.0, 0.0, mass * G]), (2, N, 1))
    ps = problem.DdpZmpProblemSet(ref_zmp, np.ones((1, N + 1)), [0, 0], x0, mass, dt, u_init=u_init)
    res = oracle.ddp_zmp_solve(ps, problem.ddp_config())
    assert (res.status == 1).all()
    w = ps.weights

    def total_cost(b, u):
        x, c = x0[b].copy(), 0.0
        for k in range(N):
            rz = ref_zmp[0, k]
            c += 0.5 * (w[0] * (x[4] - 1.0) ** 2 + w[1] * ((u[k, :2] - rz[:2]) ** 2).sum() + w[2] * (u[k, 2] - mass * G) ** 2)
            h = mass * (x[4] - rz[2])
            x = x + dt * np.array([x[1], (x[0] - u[k, 0]) * u[k, 2] / h, x[3], (x[2] - u[k, 1]) * u[k, 2] / h, x[5],
                                   u[k, 2] / mass - G])
        rz = ref_zmp[0, N]
        return c + 0.5 * (w[3] * ((x[[0, 2]] - rz[:2]) ** 2).sum() + w[4] * (x[4] - 1.0) ** 2 + w[5] * (x[[1, 3, 5]] ** 2).sum())

    def gradient(b, u):
        g = np.zeros_like(u)
        for k in range(N):
            for j in range(3):
                e = np.zeros_like(u)
                e[k, j] = 1e-5 * max(1.0, abs(u[k, j]))
                g[k, j] = (total_cost(b, u + e) - total_cost(b, u - e)) / (2 * e[k, j])
        return g

    for b in range(2):
        assert abs(total_cost(b, res.u[b]) - res.cost[b]) < 1e-12 * res.cost[b]
        g_start, g_end = np.abs(gradient(b, u_init[b])).max(), np.abs(gradient(b, res.u[b])).max()
        assert g_start > 5e-3 and g_end < 1e-5 and g_end < 1e-2 * g_start


def test_emulated_kernel_matches_oracle(oracle):
    fm = walking_plan()
    fm.update(1.9)
    N = 20
    ref_zmp = np.zeros((1, N + 1, 3))
    for k in range(N + 1):
        ref_zmp[0, k, :2] = fm.ref_zmp(1.9 + k * 0.02)
    x0 = np.array([[0.01, 0.1, -0.02, -0.05, 1.02, 0.03], [0.0, 0.0, 0.0, 0.0, 1.0, 0.0]])
    u_init = np.tile(np.array([0.0, 0.0, 100 * G]), (2, N, 1))
    ps = problem.DdpZmpProblemSet(ref_zmp, np.ones((1, N + 1)), [0, 0], x0, 100.0, 0.02, u_init=u_init)
    cfg = problem.ddp_config(max_iter=4)
    ref = oracle.ddp_zmp_solve(ps, cfg, trace_len=4)
    for feat in (1, 2):
        assert_ddp_parity(ref, emu_lib.ddp_zmp_solve(ps, cfg, trace_len=4, chunk=2, feat=feat))


def _zmp_problem_set(B, N, seed, cold_frac=0.3, stiff=0):
    """B DdpZmp problems over two walking-plan schedules: perturbed states, warm starts near m g and cold starts."""
    rng = np.random.default_rng(seed)
    ref_zmp, com_z = np.zeros((2, N + 1, 3)), np.ones((2, N + 1))
    for s, t0 in enumerate((1.9, 2.6)):
        fm = walking_plan()
        fm.update(t0)
        for k in range(N + 1):
            ref_zmp[s, k, :2] = fm.ref_zmp(t0 + k * 0.02)
        com_z[s] = 1.0 + 0.02 * s
    x0 = np.zeros((B, 6))
    x0[:, 0::2] = np.array([0.0, 0.0, 1.0]) + rng.uniform(-0.03, 0.03, (B, 3))
    x0[:, 1::2] = rng.uniform(-0.1, 0.1, (B, 3))
    u_init = np.tile(np.array([0.0, 0.0, 100 * G]), (B, N, 1)) * (1 + 0.05 * rng.standard_normal((B, N, 3)))
    cold = rng.uniform(size=B) < cold_frac
    u_init[cold, :, :2] = 0.3  # far-off ZMP and force guesses: more iterations, shortened steps
    u_init[cold, :, 2] *= rng.uniform(0.3, 2.5, (int(cold.sum()), 1))
    if stiff:  # low, fast-moving CoM: shortened and rejected line-search steps, lambda increases
        x0[:, 4] = 1.0 - 0.07 * stiff * rng.uniform(0, 1, B)
        x0[:, 5] = -0.5 * stiff
        x0[:, 1] = 0.3 * stiff
    return problem.DdpZmpProblemSet(ref_zmp, com_z, rng.integers(0, 2, B), x0, 100.0, 0.02, u_init=u_init)


@pytest.mark.parametrize("max_iter", [1, 3, 60])
def test_thread_per_problem_solver_matches_oracle(oracle, max_iter):
    """csrc/ddp_thread_zmp.cuh, the source of the one-thread-per-problem kernel, compiled for the host."""
    for stiff in (0, 3, 10):
        ps = _zmp_problem_set(24, 30, seed=7, stiff=stiff)
        cfg = problem.ddp_config(max_iter=max_iter)
        ref = oracle.ddp_zmp_solve(ps, cfg, trace_len=8)
        assert_ddp_parity(ref, emu_lib.ddp_zmp_thread_solve(ps, cfg, trace_len=8))
        if max_iter == 60 and stiff:
            assert ref.iters.max() > 6 and (ref.alpha_idx[:, :8] > 0).any()
