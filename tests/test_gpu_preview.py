"""GPU tier: batched preview-control input (ccc_preview_input) vs the oracle, and the config-1 closed loop
driven through the GPU kernel."""
import numpy as np
import pytest

from test_preview_control_cpu import _oracle_gemv, run_preview_control_closed_loop

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gemv():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0
    return engine.preview_gemv


def test_parity(oracle, gemv):
    rng = np.random.default_rng(3)
    for B, N in ((1, 200), (37, 200), (4096, 100), (513, 33)):
        K, F = rng.standard_normal(3), rng.standard_normal(N)
        x, ref = rng.standard_normal((B, 3)), rng.standard_normal((B, N))
        assert np.array_equal(gemv(K, F, x, ref), _oracle_gemv(oracle)(K, F, x, ref))


def test_closed_loop_through_gpu(gemv):
    pc, ok, planned, sim, ref = run_preview_control_closed_loop(gemv=gemv, end_time=10.0)
    assert ok
    assert np.linalg.norm(planned - ref) < 1e-2 and np.linalg.norm(sim.pos - ref) < 1e-2 and np.linalg.norm(sim.vel) < 1e-2
