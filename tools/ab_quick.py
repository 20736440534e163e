"""Quick A/B timing of the two warp-per-problem DDP kernels on fixed workloads (host-buffer API, best of 3), with a hash of
the results so that builds can be compared bit for bit:
    python tools/ab_quick.py [label] [centroidal|srb|""] [variant]
  centroidal: config 3, 16384 cold starts to convergence;  srb: config 4, 4096 cold starts, 40 iterations."""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
label = sys.argv[1] if len(sys.argv) > 1 else ""


def digest(res):
    h = hashlib.sha256()
    for f in ("x", "u", "cost", "iters", "status"):
        h.update(np.ascontiguousarray(getattr(res, f)).tobytes())
    return h.hexdigest()[:16]


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return r, min(ts)


only = sys.argv[2] if len(sys.argv) > 2 else ""
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # ddp_host.cuh Variants (0 = product default)
engine.DdpCentroidalEngine.set_variant(variant)
label = f"{label} variant {variant}"
if only in ("", "centroidal"):
    w = workloads.ddp_centroidal_config3(batch=16384)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    eng = engine.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
    res, t = best(lambda: eng.solve(ps, problem.ddp_centroidal_config()))
    print(f"{label:28s} centroidal 16384: {t * 1e3:8.1f} ms  {16384 / t:8.0f} solves/s  digest {digest(res)}", flush=True)
    eng.close()
if only == "centroidal":
    sys.exit(0)
w = workloads.ddp_srb_config4(batch=4096)
ps = problem.DdpSrbProblemSet.from_workload(w)
eng = engine.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
cfg = problem.ddp_srb_config(max_iter=40)
res, t = best(lambda: eng.solve(ps, cfg))
print(f"{label:28s} srb 4096 x 40 iterations: {t * 1e3:8.1f} ms  {float(res.iters.sum()) / t / 1e3:8.2f} k iterations/s  digest {digest(res)}", flush=True)
