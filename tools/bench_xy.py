"""LinearMpcXY batch through the host-buffer C ABI:  python tools/bench_xy.py [batch]  -> one JSON line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from centroidalcontrolcollection_b200 import engine  # noqa: E402
from centroidalcontrolcollection_b200 import workloads  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
ps = workloads.linear_mpc_xy_problem_set(15, B)
qp = engine.qp_solver_for()
qp(ps)
t0 = time.time()
res = qp(ps)
dt = time.time() - t0
print(json.dumps({"workload": "LinearMpcXY n=240, 15 eq, 480 bound rows", "batch": B, "seconds": dt, "solves_per_s": B / dt,
                  "mean_iters": float(res.iters.mean()), "max_iters": int(res.iters.max()),
                  "solved_frac": float((res.status == 0).mean()), "mean_active": float(res.n_active.mean())}))
