"""Full-size parity of the two DDP BASELINE configurations against the CPU oracle (run on the GPU box; the oracle side takes
about a minute for config 3 and two to three minutes for the config-4 shard on 16 host threads):
    python tools/full_parity.py [3] [4]     -> one JSON line per configuration
Every output of every problem is compared: iteration counts, status, accepted line-search index and lambda of every
iteration, BoxQP clamped sets of the last backward pass, state / input trajectories and costs, bit for bit."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from centroidalcontrolcollection_b200 import build, engine, problem, workloads  # noqa: E402
from oracle import binding  # noqa: E402

build.build()
threads = max(1, binding.hardware_threads())
which = sys.argv[1:] or ["3", "4"]


def compare(ref, got):
    out = {}
    for f in ("iters", "status", "alpha_idx", "clamped", "x", "u", "cost", "lambda_trace"):
        a, b = getattr(ref, f), getattr(got, f)
        out[f] = bool(np.array_equal(a, b))
    out["com_linf"] = float(np.abs(ref.x[:, :, :3] - got.x[:, :, :3]).max())
    return out


if "3" in which:
    w = workloads.ddp_centroidal_config3(batch=16384)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config()
    eng = engine.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
    got = eng.solve(ps, cfg, trace_len=400)
    t0 = time.perf_counter()
    ref = binding.ddp_centroidal_solve(ps, cfg, trace_len=400, n_threads=threads)
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": "config 3: " + w["name"], "problems": int(ps.batch), "bit_exact": compare(ref, got),
                      "mean_ddp_iters": float(got.iters.mean()), "max_ddp_iters": int(got.iters.max()),
                      "oracle_seconds": dt, "oracle_threads": threads}), flush=True)
    eng.close()
if "4" in which:
    w = workloads.ddp_srb_config4(batch=8192)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    cfg = problem.ddp_srb_config()
    eng = engine.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
    got = eng.solve(ps, cfg, trace_len=500)
    t0 = time.perf_counter()
    ref = binding.ddp_srb_solve(ps, cfg, trace_len=500, n_threads=threads)
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": "config 4 (one of 8 shards): " + w["name"], "problems": int(ps.batch), "bit_exact": compare(ref, got),
                      "mean_ddp_iters": float(got.iters.mean()), "max_ddp_iters": int(got.iters.max()),
                      "converged_frac": float((got.status == 1).mean()), "oracle_seconds": dt, "oracle_threads": threads}), flush=True)
    eng.close()
