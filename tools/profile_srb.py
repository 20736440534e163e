"""One shard of config 4 (DdpSingleRigidBody, N = 100, cold start) through the host-buffer API, for timing and ncu:
    python tools/profile_srb.py [batch] [max_iter]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 500
w = workloads.ddp_srb_config4(batch=B)
ps = problem.DdpSrbProblemSet.from_workload(w)
eng = engine.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
cfg = problem.ddp_srb_config(max_iter=max_iter)
res = eng.solve(ps, cfg)
print("MEAN_ITERS", float(res.iters.mean()))
if os.environ.get("NV_COMPUTE_PROFILER_PERFWORKS_DIR") or "--once" in sys.argv:
    sys.exit(0)  # under ncu: one launch is enough
t0 = time.time()
res = eng.solve(ps, cfg)
dt = time.time() - t0
print(f"config 4 shard: {B} problems, N = {ps.N}, mean DDP iterations {res.iters.mean():.1f} (max {res.iters.max()}), "
      f"converged {float((res.status == 1).mean()):.3f}, {B / dt:.0f} solves/s through ccc_ddp_srb_solve (host buffers)")
