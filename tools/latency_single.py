"""Latency of planOnce-sized DdpCentroidal calls (small batches) through the host-buffer C-ABI, under the three
small-batch policies of the engine (ccc_ddp_set_small_batch_policy; results are bit-identical under all of them):
    python tools/latency_single.py
  team    — batches of at most one problem per SM on the team kernel (csrc/ddp_team.cuh): the default
  spread  — the warp-per-problem kernel with small batches spread over all SMs
  packed  — the warp-per-problem kernel filling SMs with eight problems each (the round-1 behaviour)
cold = cold start to convergence; warm = one DDP iteration from the converged plan (what every control cycle after the
first does in the reference's test loop, tests/src/TestDdpCentroidal.cpp:116)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
cfg = problem.ddp_centroidal_config()
cfg1 = problem.ddp_centroidal_config(max_iter=1)
cfg0 = problem.ddp_centroidal_config(max_iter=0)
POLICIES = (("team", 1, 1), ("spread", 0, 1), ("packed", 0, 0))
if "--io" in sys.argv:
    # A/B of the packed staging block of small host-buffer calls (default on) under the default policy
    for io in (1, 0, 1, 0):
        engine.DdpCentroidalEngine.set_packed_io(io)
        for B in (1, 8):
            wb = workloads.ddp_centroidal_config3(batch=B, n_sched=1)
            pb = problem.DdpCentroidalProblemSet.from_workload(wb)
            eb = engine.DdpCentroidalEngine(pb.N, B, 1)
            rb = eb.solve(pb, cfg)
            pb.u_init = rb.u.copy()
            ts = []
            for _ in range(5):
                eb.solve(pb, cfg1)
            for _ in range(200):
                t0 = time.perf_counter()
                eb.solve(pb, cfg1)
                ts.append(time.perf_counter() - t0)
            print(f"packed staging {io}: batch {B}, warm tick median {np.median(ts) * 1e3:.3f} ms, best {min(ts) * 1e3:.3f} ms", flush=True)
            eb.close()
    engine.DdpCentroidalEngine.set_packed_io(1)
    sys.exit(0)


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return r, np.median(ts) * 1e3, min(ts) * 1e3


ref = {}
for name, team, spread in POLICIES:
    engine.DdpCentroidalEngine.set_small_batch_policy(team=team, spread=spread)
    for B in (1, 8, 64, 148, 296, 592):
        wb = workloads.ddp_centroidal_config3(batch=B, n_sched=1)
        pb = problem.DdpCentroidalProblemSet.from_workload(wb)
        eb = engine.DdpCentroidalEngine(pb.N, B, 1)
        rb, med, best = timed(lambda: eb.solve(pb, cfg), 5 if B <= 8 else 3)
        it = rb.iters
        key = (B,)
        if key in ref:
            same = all(np.array_equal(getattr(rb, f), getattr(ref[key], f)) for f in ("x", "u", "iters", "status", "cost"))
        else:
            ref[key] = rb
            same = True
        line = (f"{name:6s} batch {B:4d}: cold solve {med:8.2f} ms (best {best:.2f}; iterations mean {it.mean():.1f} max {it.max()}"
                f", {med / it.max():.3f} ms per iteration of the longest solve)")
        pb.u_init = rb.u.copy()
        _, m1, b1 = timed(lambda: eb.solve(pb, cfg1), 30)
        _, m0, b0 = timed(lambda: eb.solve(pb, cfg0), 30)
        print(f"{line}; warm tick (max_iter 1) {m1:.3f} ms (best {b1:.3f}); max_iter 0 (copies, packing, initial rollout) {m0:.3f} ms;"
              f" team kernel {eb.last_team}; identical to the first policy's result {same}", flush=True)
        eb.close()
engine.DdpCentroidalEngine.set_small_batch_policy(team=1, spread=1)
