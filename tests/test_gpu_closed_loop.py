"""GPU tier: ccc_ddp_centroidal_closed_loop (device-resident receding-horizon loop) against the oracle's closed loop:
plant trajectories, applied force scales and iteration counts bit-exact; reference tolerances on the full scenario."""
import numpy as np
import pytest

from centroidalcontrolcollection_b200 import problem

from closed_loop_spec import check_reference_tolerances, reference_scenario

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from centroidalcontrolcollection_b200 import build, engine

    build.build()
    assert engine.lib().ccc_device_count() > 0
    return engine


def _parity(ref, got):
    assert np.array_equal(ref.iters, got.iters), "DDP iteration counts differ"
    assert np.array_equal(ref.u0, got.u0), f"applied force scales differ by {np.abs(ref.u0 - got.u0).max():.3e}"
    assert np.array_equal(ref.plant, got.plant), f"plant trajectories differ by {np.abs(ref.plant - got.plant).max():.3e}"


def test_closed_loop_parity_short(eng_mod, oracle):
    """8 perturbed plants, 80 cycles across the first contact switch of a 30-stage horizon."""
    lp, _ = reference_scenario(ticks=80, batch=8, horizon_steps=30, perturb=0.01)
    lp.set_disturbance(20, [0.05, 0.05, 0.0])
    cfg = problem.ddp_centroidal_config()
    eng = eng_mod.DdpCentroidalEngine(lp.N, lp.batch, 1)
    got = eng.closed_loop(lp, cfg)
    ref = oracle.ddp_centroidal_closed_loop(lp, cfg, n_threads=max(1, oracle.hardware_threads()))
    _parity(ref, got)
    assert eng.last_launches >= 3 * 80


def test_closed_loop_reference_scenario(eng_mod, oracle):
    """The reference test itself (600 cycles, horizon 100, disturbance at 1 s) for 4 plants on the device."""
    lp, ref_fn = reference_scenario(ticks=600, batch=4, perturb=0.005)
    cfg = problem.ddp_centroidal_config()
    eng = eng_mod.DdpCentroidalEngine(lp.N, lp.batch, 1)
    got = eng.closed_loop(lp, cfg)
    check_reference_tolerances(got, lp, ref_fn)
    ref = oracle.ddp_centroidal_closed_loop(lp, cfg, n_threads=max(1, oracle.hardware_threads()))
    _parity(ref, got)
