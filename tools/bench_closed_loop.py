"""Closed-loop trajectories/s of ccc_ddp_centroidal_closed_loop on the reference's test scenario
(tests/src/TestDdpCentroidal.cpp: horizon 100 x 0.03 s, 600 control cycles of 5 ms, disturbance at 1 s) for a batch of
perturbed plants:  python tools/bench_closed_loop.py [batch] [ticks]  -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from centroidalcontrolcollection_b200 import build, engine, problem  # noqa: E402
from closed_loop_spec import reference_scenario  # noqa: E402

build.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 600
lp, ref = reference_scenario(ticks=ticks, batch=B, perturb=0.005)
cfg = problem.ddp_centroidal_config()
eng = engine.DdpCentroidalEngine(lp.N, lp.batch, 1)
t0 = time.perf_counter()
res = eng.closed_loop(lp, cfg)
dt = time.perf_counter() - t0
end = res.plant[:, -1]
err = np.linalg.norm(end[:, 0:3] - np.array(ref(ticks * lp.sim_dt))[None, :], axis=1)
print(json.dumps({"workload": f"DdpCentroidal closed loop, horizon {lp.N} x {lp.dt} s, {ticks} cycles of {lp.sim_dt} s, batch {B}",
                  "seconds": dt, "trajectories_per_s": B / dt, "control_cycles_per_s": B * ticks / dt,
                  "ms_per_cycle_of_the_batch": 1e3 * dt / ticks, "kernel_launches": eng.last_launches,
                  "first_cycle_iters_mean": float(res.iters[:, 0].mean()), "final_pos_err_max": float(err.max()),
                  "api": "ccc_ddp_centroidal_closed_loop(CCC_MEM_HOST)"}))
