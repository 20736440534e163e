"""Parity assertions shared by the emulator (CPU) and engine (GPU) tests."""
import numpy as np

FIELDS = ["iters", "status", "alpha_idx", "clamped", "x", "u", "cost", "lambda_trace"]


def assert_ddp_parity(ref, got, rel_tol=1e-6, bit_exact=True):
    """north_star bar: iteration counts, accepted step indices and clamped sets bit-exact;
    CoM/state, input and cost trajectories within rel_tol (1e-6 relative, L-infinity).
    With bit_exact=True additionally require identical bits everywhere (canonical arithmetic)."""
    for name in ["iters", "status", "alpha_idx", "clamped"]:
        a, b = getattr(ref, name), getattr(got, name)
        assert np.array_equal(a, b), f"{name} differs in {np.count_nonzero(a != b)} entries"
    for name in ["x", "u", "cost", "lambda_trace"]:
        a, b = getattr(ref, name), getattr(got, name)
        scale = max(1.0, float(np.abs(a).max()))
        err = float(np.abs(a - b).max()) / scale
        assert err <= rel_tol, f"{name}: relative L-inf error {err:.3e} > {rel_tol}"
        if bit_exact:
            assert np.array_equal(a, b), f"{name} not bit-exact (max |d| = {np.abs(a - b).max():.3e})"


def assert_ddp_parity_tol(ref, got, rel_tol=1e-6):
    """Tolerance-only variant (north_star bar) for paths where bit-exactness is not claimed."""
    assert_ddp_parity(ref, got, rel_tol=rel_tol, bit_exact=False)


def com_linf(ref, got):
    """CoM-trajectory L-infinity error (BASELINE.json's second headline figure)."""
    return float(np.abs(ref.x[:, :, 0:3] - got.x[:, :, 0:3]).max())
