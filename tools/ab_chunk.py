"""A/B of the time-slice length (DDP iterations per visit before a solve is suspended and re-queued) on the two full-size DDP
workloads, host-buffer API, best of 2:  python tools/ab_chunk.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import build, engine, problem, workloads

build.build()


def best(fn, reps=2):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


w = workloads.ddp_srb_config4(batch=8192)
ps = problem.DdpSrbProblemSet.from_workload(w)
eng = engine.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
cfg = problem.ddp_srb_config()
for chunk in ((32, 64, 128, 0, 32) if '--centroidal' not in sys.argv and '--srb' not in sys.argv else (32, 16, 24, 48, 96, 128, 192, 256, 32) if '--srb' in sys.argv else ()):
    engine.DdpCentroidalEngine.set_chunk(chunk)
    t = best(lambda: eng.solve(ps, cfg), 1)
    print(f"srb 8192 (config-4 shard), chunk {chunk:3d}: {t * 1e3:8.1f} ms  {8192 / t:7.0f} solves/s", flush=True)
eng.close()
if '--srb' in sys.argv:
    sys.exit(0)
w = workloads.ddp_centroidal_config3(batch=16384)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
eng = engine.DdpCentroidalEngine(ps.N, ps.batch, ps.sched.S)
cfg = problem.ddp_centroidal_config()
for chunk in ((32, 64, 16, 0, 32) if '--centroidal' not in sys.argv else (32, 4, 8, 12, 16, 20, 24, 32, 8, 16, 24, 32)):
    engine.DdpCentroidalEngine.set_chunk(chunk)
    t = best(lambda: eng.solve(ps, cfg))
    print(f"centroidal 16384 (config 3), chunk {chunk:3d}: {t * 1e3:8.1f} ms  {16384 / t:7.0f} solves/s", flush=True)
engine.DdpCentroidalEngine.set_chunk(64)
