"""Probe: kernel durations of a single-problem warm tick under the small-batch policies (run under
ncu --metrics gpu__time_duration.sum to list the launches)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = workloads.ddp_centroidal_config3(batch=B, n_sched=1)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
eng = engine.DdpCentroidalEngine(ps.N, B, 1)
res = eng.solve(ps, problem.ddp_centroidal_config())
ps.u_init = res.u.copy()
for name, team, spread in (("team", 1, 1), ("packed", 0, 0)):
    engine.DdpCentroidalEngine.set_small_batch_policy(team=team, spread=spread)
    for mi in (0, 1, 2):
        r = eng.solve(ps, problem.ddp_centroidal_config(max_iter=mi), trace_len=4)
        print(name, "max_iter", mi, "iters", r.iters[:4], "status", r.status[:4], "alpha", r.alpha_idx[:2].tolist(), flush=True)
