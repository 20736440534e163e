"""Small DDP calls for compute-sanitizer runs (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Team kernel (batch 3), queue kernel with spreading (batch 150, short horizon), SRB team kernel (batch 2), packed staging path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
w = workloads.ddp_centroidal_config3(batch=3, horizon_steps=50)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
eng = engine.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
r = eng.solve(ps, problem.ddp_centroidal_config(max_iter=4), trace_len=4)
assert eng.last_team
print("team centroidal", r.iters, r.alpha_idx.tolist())
eng.close()
w = workloads.ddp_centroidal_config3(batch=150, horizon_steps=12)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
eng = engine.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
r = eng.solve(ps, problem.ddp_centroidal_config(max_iter=3))
assert not eng.last_team
print("queue centroidal (spread)", int(r.iters.sum()))
eng.close()
w = workloads.ddp_srb_config4(batch=2, horizon_steps=20)
ps = problem.DdpSrbProblemSet.from_workload(w)
eng = engine.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
r = eng.solve(ps, problem.ddp_srb_config(max_iter=3))
print("team srb", r.iters)
eng.close()
