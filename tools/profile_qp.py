"""One batch of config-2 (LinearMpcZmp) or config-5 (IntrinsicallyStableMpc) QPs, for ncu captures and quick timing:
    python tools/profile_qp.py [config 2|5] [batch] [repeats]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, linear_mpc, workloads

build.build()
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 296
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
_qp = engine.qp_solver_for()
_t = {"engine": 0.0}


def qp(ps):
    t = time.time()
    r = _qp(ps)
    _t["engine"] += time.time() - t
    return r


if cfg == 2:
    w = workloads.linear_mpc_zmp_config2(batch=B)
    mpc = linear_mpc.LinearMpcZmp(w["com_height"], w["horizon_duration"], w["horizon_dt"])
    run = lambda: mpc.plan_batch(qp, w["pos"], w["vel"], w["acc"], w["lim_min"], w["lim_max"], w["control_dt"])
    last = lambda: mpc.mpc_1d.last_result
else:
    side = max(int(np.sqrt(B // 64)), 1)
    w = workloads.ismpc_config5(n_plans=side * side, n_perturb=B // (side * side))
    mpc = linear_mpc.IntrinsicallyStableMpc(w["com_height"], w["horizon_duration"], w["horizon_dt"])
    run = lambda: mpc.plan_batch(qp, w["capture_point"], w["planned_zmp"], w["ref_zmp"], w["lim_min"], w["lim_max"], w["control_dt"])
    last = lambda: mpc.mpc_1d.last_result
z = run()
_t["engine"] = 0.0
t0 = time.time()
for _ in range(reps):
    z = run()
dt = (time.time() - t0) / max(reps, 1)
r = last()
print(f"config {cfg}: {len(z)} two-axis problems ({2 * len(z)} QPs), status ok {bool((r.status == 0).all())}, "
      f"mean active-set iterations {r.iters.mean():.1f}, {len(z) / dt:.0f} two-axis solves/s through the Python host class, "
      f"{len(z) * max(reps, 1) / max(_t['engine'], 1e-9):.0f} two-axis solves/s inside ccc_qp_solve (host buffers: H2D + setup + solve + D2H)")
