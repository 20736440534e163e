"""Throughput of the other BASELINE.json configurations (they are parity-test cases, not the bench line; this tool
records where they stand):  python tools/bench_configs.py [--reps R] [--no-cpu]  -> one JSON line per configuration.

  config 2   LinearMpcZmp            N = 100, dt = 0.01, batch 4096 (x 2 axes)            ccc_qp_solve
  config 4   DdpSingleRigidBody      N = 100, one 8192-problem shard of the 65536 batch   ccc_ddp_srb_solve
  config 5   IntrinsicallyStableMpc  256 plans x 512 perturbations (x 2 axes)             ccc_qp_solve
  xy         LinearMpcXY             reference test schedule, n = 240, batch 1184         ccc_qp_solve

`value` is measured inside the C-ABI call with host buffers (H2D + setup + solve + D2H, time.perf_counter around the
call, best of R); `cpu_baseline` is the CPU oracle (oracle/, all host threads) on a bounded sample of the same inputs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from centroidalcontrolcollection_b200 import build, engine, linear_mpc, problem, workloads  # noqa: E402


def timed(fn, reps):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out, best


def qp_case(name, mpc, run, units, reps, cpu):
    """run(qp_solve) drives the host class once; the engine call inside it is what is timed."""
    qp = engine.qp_solver_for()
    captured = {}

    def grab(ps):
        captured["ps"] = ps
        return qp(ps)

    run(grab)  # warm-up, and captures the assembled QP batch
    ps = captured["ps"]
    res, dt = timed(lambda: qp(ps), reps)
    line = {"workload": name, "unit": "2-axis solves/s" if units == 2 else "solves/s", "value": ps.batch / units / dt,
            "qps": ps.batch, "n": ps.n, "n_eq": ps.n_eq, "n_ineq": ps.n_ineq, "mean_active_set_iterations": float(res.iters.mean()),
            "solved_frac": float((res.status == 0).mean()), "seconds": dt, "api": "ccc_qp_solve(CCC_MEM_HOST)"}
    if cpu:
        from oracle import binding

        threads = binding.hardware_threads()
        k = min(ps.batch, 64 * threads)
        sub = ps.subset(np.arange(k))
        _, cdt = timed(lambda: binding.qp_solve(sub, n_threads=threads), 1)
        line["cpu_baseline"] = {"value": k / units / cdt, "unit": line["unit"], "cores": threads, "kind": "port",
                                "sample": f"first {k} QPs of the batch"}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--qp-packed", type=int, default=1, help="0: full-R QP kernel alone (round 1), 1: packed-R first pass (default)")
    args = ap.parse_args()
    build.build()
    engine.QpEngine.set_packed(args.qp_packed)
    want = lambda k: not args.only or k in args.only.split(",")

    if want("2"):
        w = workloads.linear_mpc_zmp_config2()
        mpc = linear_mpc.LinearMpcZmp(w["com_height"], w["horizon_duration"], w["horizon_dt"])
        run = lambda q: mpc.plan_batch(q, w["pos"], w["vel"], w["acc"], w["lim_min"], w["lim_max"], w["control_dt"])
        print(json.dumps(qp_case("config 2: " + w["name"], mpc, run, 2, args.reps, not args.no_cpu)), flush=True)
    if want("5"):
        w = workloads.ismpc_config5()
        mpc = linear_mpc.IntrinsicallyStableMpc(w["com_height"], w["horizon_duration"], w["horizon_dt"])
        run = lambda q: mpc.plan_batch(q, w["capture_point"], w["planned_zmp"], w["ref_zmp"], w["lim_min"], w["lim_max"], w["control_dt"])
        print(json.dumps(qp_case("config 5: " + w["name"], mpc, run, 2, args.reps, not args.no_cpu)), flush=True)
    if want("xy"):
        ps = workloads.linear_mpc_xy_problem_set(15, 1184)
        print(json.dumps(qp_case("LinearMpcXY: reference test schedule, n = 240, 15 equalities, 480 bound rows, batch 1184", None,
                                 lambda q: q(ps), 1, args.reps, not args.no_cpu)), flush=True)
    if want("4"):
        w = workloads.ddp_srb_config4(batch=8192)
        ps = problem.DdpSrbProblemSet.from_workload(w)
        eng = engine.DdpSrbEngine(ps.N, ps.batch, ps.sched.S)
        cfg = problem.ddp_srb_config()
        eng.solve(ps, cfg)
        res, dt = timed(lambda: eng.solve(ps, cfg), max(1, args.reps - 1))
        line = {"workload": "config 4 (one of 8 shards): " + w["name"], "unit": "solves/s", "value": ps.batch / dt, "seconds": dt,
                "mean_ddp_iters": float(res.iters.mean()), "max_ddp_iters": int(res.iters.max()),
                "converged_frac": float((res.status == 1).mean()), "api": "ccc_ddp_srb_solve(CCC_MEM_HOST)"}
        if not args.no_cpu:
            from oracle import binding

            threads = binding.hardware_threads()
            sub = ps.subset(np.arange(8 * threads))
            _, cdt = timed(lambda: binding.ddp_srb_solve(sub, cfg, n_threads=threads), 1)
            line["cpu_baseline"] = {"value": sub.batch / cdt, "unit": "solves/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sub.batch} problems of the shard"}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
