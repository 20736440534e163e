"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

The reference holds no golden vectors for this path and cannot be built or imported
here, so the fixtures are oracle outputs on seeded inputs: they pin the oracle against silent
drift (CPU tier) and give the GPU tier a check that does not depend on running the oracle.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from centroidalcontrolcollection_b200 import problem, workloads  # noqa: E402
from oracle import binding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ddp_centroidal():
    w = workloads.ddp_centroidal_config3(batch=16)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    res = binding.ddp_centroidal_solve(ps, problem.ddp_centroidal_config(), trace_len=96, n_threads=8)
    assert res.iters.max() <= 96
    np.savez_compressed(os.path.join(HERE, "ddp_centroidal_config3_b16.npz"), x0=ps.x0, x=res.x, u=res.u,
                        cost=res.cost, iters=res.iters, status=res.status, alpha_idx=res.alpha_idx,
                        lambda_trace=res.lambda_trace, clamped=res.clamped)


if __name__ == "__main__":
    binding.build()
    ddp_centroidal()
    print("golden fixtures written to", HERE)
