"""One LinearMpcXY sweep through ccc_linear_mpc_xy_solve, for ncu captures and quick timing:
    python tools/profile_xy.py [n_sched] [per_sched]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, workloads

build.build()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 148
per = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sweep = workloads.linear_mpc_xy_sweep(n_sched=S, per_sched=per)
eng = engine.LinearMpcXyEngine(sweep.N, sweep.n, sweep.n_eq, sweep.batch, sweep.S)
res = eng.solve(sweep)
if os.environ.get("NV_COMPUTE_PROFILER_PERFWORKS_DIR") or "--once" in sys.argv:
    sys.exit(0)
t0 = time.time()
res = eng.solve(sweep)
dt = time.time() - t0
print(f"LinearMpcXY sweep: {S} schedules x {per} states (n = {sweep.n}), solved {float((res.status == 0).mean()):.3f}, "
      f"mean active-set iterations {res.iters.mean():.1f}, {sweep.batch / dt:.0f} solves/s, {eng.last_launches} launches")
