"""One batch of DdpZmp problems through ccc_ddp_zmp_solve (thread-per-problem kernel), for ncu captures:
    python tools/profile_zmp.py [batch] [max_iter] [variant]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = workloads.ddp_zmp_batch(batch=B)
ps = problem.DdpZmpProblemSet(w["ref_zmp"], w["com_z"], w["sched_id"], w["x0"], w["mass"], w["dt"], u_init=w["u_init"])
engine.DdpZmpEngine.set_variant(variant)
eng = engine.DdpZmpEngine(ps.N, ps.batch, len(w["ref_zmp"]))
cfg = problem.ddp_config(max_iter=max_iter)
res = eng.solve(ps, cfg)
if os.environ.get("NV_COMPUTE_PROFILER_PERFWORKS_DIR") or "--once" in sys.argv:
    sys.exit(0)
t0 = time.time()
res = eng.solve(ps, cfg)
print(f"DdpZmp: {B} problems, mean iterations {res.iters.mean():.2f}, {B / (time.time() - t0):.0f} solves/s through the host-buffer API")
