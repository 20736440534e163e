"""Latency of one DdpCentroidal planOnce-sized call (batch of one) through the host-buffer C-ABI, and where it goes:
    python tools/latency_single.py
cold = cold start to convergence; warm = one DDP iteration from the converged plan (what every control cycle after the
first does in the reference's test loop, tests/src/TestDdpCentroidal.cpp:116)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
w = workloads.ddp_centroidal_config3(batch=1, n_sched=1)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
eng = engine.DdpCentroidalEngine(ps.N, 1, 1)
cfg = problem.ddp_centroidal_config()


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return r, np.median(ts) * 1e3, min(ts) * 1e3


res, med, best = timed(lambda: eng.solve(ps, cfg), 10)
print(f"cold solve, batch 1: {int(res.iters[0])} DDP iterations, median {med:.2f} ms (best {best:.2f}), {med / max(int(res.iters[0]), 1):.3f} ms per iteration")
ps.u_init = res.u.copy()
cfg1 = problem.ddp_centroidal_config(max_iter=1)
res1, med1, best1 = timed(lambda: eng.solve(ps, cfg1), 50)
print(f"warm tick (max_iter 1), batch 1: median {med1:.3f} ms (best {best1:.3f})")
cfg0 = problem.ddp_centroidal_config(max_iter=0)
_, med0, best0 = timed(lambda: eng.solve(ps, cfg0), 50)
print(f"max_iter 0 (copies, table packing, initial rollout only): median {med0:.3f} ms (best {best0:.3f})")
for B in (8, 64, 148, 592):
    wb = workloads.ddp_centroidal_config3(batch=B, n_sched=1)
    pb = problem.DdpCentroidalProblemSet.from_workload(wb)
    eb = engine.DdpCentroidalEngine(pb.N, B, 1)
    rb = eb.solve(pb, cfg)
    pb.u_init = rb.u.copy()
    _, mb, bb = timed(lambda: eb.solve(pb, cfg1), 20)
    print(f"warm tick, batch {B}: median {mb:.3f} ms")
    eb.close()
