"""One cold solve and warm ticks of a small DdpCentroidal batch on the team kernel, for ncu captures:
    python tools/profile_team.py [batch] [cold max_iter]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 12
w = workloads.ddp_centroidal_config3(batch=B)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
eng = engine.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
res = eng.solve(ps, problem.ddp_centroidal_config(max_iter=max_iter))
assert eng.last_team
print("iters mean", res.iters.mean(), "max", res.iters.max())
