"""One solve of a small config-3 batch, for ncu captures:  python tools/profile_solve.py [batch] [variant] [max_iter]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # index into ddp_host.cuh Variants<M>::table (0 = product default)
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 500
if len(sys.argv) > 4:
    engine.DdpCentroidalEngine.set_chunk(int(sys.argv[4]))
engine.DdpCentroidalEngine.set_variant(variant)
w = workloads.ddp_centroidal_config3(batch=B)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
eng = engine.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
res = eng.solve(ps, problem.ddp_centroidal_config(max_iter=max_iter))
print("iters mean", res.iters.mean(), "max", res.iters.max())
print("MEAN_ITERS", float(res.iters.mean()))
