"""Config 1 (CPU plumbing): TestPreviewControlZmp of the reference (tests/src/TestPreviewControlZmp.cpp:15-105)
through the host classes — single problem, no GPU."""
import numpy as np

from centroidalcontrolcollection_b200.linear_models import G
from centroidalcontrolcollection_b200.preview_control import PreviewControlZmp

from footstep_manager import walking_plan
from sim_models import ComZmpSim2d


def run_preview_control_closed_loop(gemv=None, end_time=10.0):
    horizon_duration, horizon_dt, sim_dt, h = 2.0, 0.01, 0.005, 1.0
    pc = PreviewControlZmp(h, horizon_duration, horizon_dt)
    N = pc.pc_1d.horizon_steps
    fm = walking_plan()
    sim = ComZmpSim2d(h, sim_dt)
    planned = sim.pos.copy()
    t, ok = 0.0, True
    fm.update(0.0)
    kw = {} if gemv is None else {"gemv": gemv}
    while t < end_time:
        fm.update(t)
        ref_seq = np.array([fm.ref_zmp(t + (i + 1) * horizon_dt) for i in range(N)])[None]  # note (i + 1): :57-62
        acc = G / h * (sim.pos - planned)
        planned = pc.plan_batch(sim.pos[None], sim.vel[None], acc[None], ref_seq, sim_dt, **kw)[0]
        ok &= bool(np.linalg.norm(planned - fm.ref_zmp(t)) < 0.1)  # :84
        t += sim_dt
        sim.update(planned)
        for dtm in (4.5, 8.5):
            if dtm <= t < dtm + sim_dt:
                sim.add_disturb(np.array([0.05, 0.05]))
    return pc, ok, planned, sim, fm.ref_zmp(t)


def test_gains():
    pc, *_ = run_preview_control_closed_loop(end_time=0.0)
    p = pc.pc_1d
    assert p.riccati_converged and p.horizon_steps == 200
    assert p.riccati_error < 1e-6 * np.linalg.norm(p.P)
    assert p.K.shape == (1, 3) and p.F.shape == (1, 200)
    assert (np.abs(p.F[0, :50]) > np.abs(p.F[0, 150:]).max()).all()  # preview gains decay along the horizon


def test_preview_control_zmp_closed_loop():
    pc, ok, planned, sim, ref = run_preview_control_closed_loop()
    assert ok
    assert np.linalg.norm(planned - ref) < 1e-2      # :103
    assert np.linalg.norm(sim.pos - ref) < 1e-2      # :104
    assert np.linalg.norm(sim.vel) < 1e-2            # :105


def _oracle_gemv(oracle):
    import ctypes as C

    L = oracle.lib()
    L.ccc_oracle_preview_input.restype = C.c_int32
    L.ccc_oracle_preview_input.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5

    def gemv(K, F, x, ref):
        K, F = np.ascontiguousarray(K).reshape(-1), np.ascontiguousarray(F).reshape(-1)
        x, ref = np.ascontiguousarray(x), np.ascontiguousarray(ref)
        u = np.zeros(len(x))
        assert L.ccc_oracle_preview_input(len(x), len(F), K.ctypes.data, F.ctypes.data, x.ctypes.data, ref.ctypes.data,
                                          u.ctypes.data) == 0
        return u

    return gemv


def test_oracle_and_emulated_kernel_agree(oracle):
    """C++ oracle of the online dot product == numpy to rounding, == the emulated CUDA row kernel bit for bit."""
    import ctypes as C

    import emu_lib

    rng = np.random.default_rng(0)
    B, N = 5, 200
    K, F = rng.standard_normal(3), rng.standard_normal(N) * np.exp(-np.arange(N) / 40.0)
    x, ref = rng.standard_normal((B, 3)), rng.standard_normal((B, N))
    u_o = _oracle_gemv(oracle)(K, F, x, ref)
    assert np.allclose(u_o, -(x @ K) + ref @ F, rtol=1e-13, atol=1e-13)
    L = emu_lib.lib()
    L.ccc_emu_preview_input.restype = C.c_int32
    L.ccc_emu_preview_input.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5
    u_e = np.zeros(B)
    assert L.ccc_emu_preview_input(B, N, K.ctypes.data, F.ctypes.data, x.ctypes.data, ref.ctypes.data, u_e.ctypes.data) == 0
    assert np.array_equal(u_o, u_e)
