"""Why do ~15 % of config 4's cold starts stop at max_iter = 500?  ("what's weak" #1b.)

python tools/srb_cold_start_study.py [n_problems] > profiles/r02_srb_cold_start.txt

Runs the CPU oracle on the first n problems of config 4 (DdpSingleRigidBody, N = 100, u_init = 0, perturbed initial
orientation / angular velocity) and tabulates, per outcome class, the final cost, the largest |pitch| (the ZYX Euler
angle whose cosine divides the rate matrix, reference src/DdpSingleRigidBody.cpp:26-38) and the largest angular
velocity along the returned trajectory, and the same for a warm start from the converged plan of the unperturbed
reference scenario (the situation of every control cycle but the first).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from centroidalcontrolcollection_b200 import problem, workloads  # noqa: E402
from oracle import binding  # noqa: E402


def table(tag, res):
    x = res.x
    beta = np.abs(x[:, :, 4]).max(axis=1)
    om = np.abs(x[:, :, 9:12]).max(axis=(1, 2))
    classes = [("converged, cost < 10", (res.status == 1) & (res.cost < 10)), ("converged, cost >= 10", (res.status == 1) & (res.cost >= 10)),
               ("max_iter reached", res.status == 0), ("lambda_max reached", res.status < 0)]
    print(f"## {tag}: {len(res.iters)} problems, mean iterations {res.iters.mean():.1f}")
    print(f"{'class':26s} {'count':>6s} {'iters (median)':>15s} {'cost (median)':>14s} {'max|pitch| median / max [rad]':>30s} {'max|omega| median [rad/s]':>26s}")
    for name, sel in classes:
        if sel.sum() == 0:
            print(f"{name:26s} {0:6d}")
            continue
        print(f"{name:26s} {int(sel.sum()):6d} {np.median(res.iters[sel]):15.0f} {np.median(res.cost[sel]):14.3f} "
              f"{np.median(beta[sel]):14.3f} / {beta[sel].max():6.3f} {np.median(om[sel]):26.2f}")
    tumbled = beta > np.pi / 2
    print(f"trajectories whose pitch passes +-pi/2: {int(tumbled.sum())}; of those, cost < 10: {int((tumbled & (res.cost < 10)).sum())}; "
          f"problems with cost >= 10 or max_iter that did NOT pass pi/2: {int((~tumbled & ((res.cost >= 10) | (res.status != 1))).sum())}\n")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    threads = binding.hardware_threads()
    w = workloads.ddp_srb_config4(batch=n)
    ps = problem.DdpSrbProblemSet.from_workload(w)
    cfg = problem.ddp_srb_config()
    print("# DdpSingleRigidBody config 4 (N = 100, dt = 0.03, A - flight - B schedule), CPU oracle, max_iter = 500\n")
    cold = binding.ddp_srb_solve(ps, cfg, n_threads=threads)
    table("cold start (u_init = 0: the rollout of the first iteration is 3 s of free fall)", cold)
    # warm start: the converged plan of the unperturbed scenario
    w1 = workloads.ddp_srb_config4(batch=1)
    w1["x0"][0] = [0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    ps1 = problem.DdpSrbProblemSet.from_workload(w1)
    nominal = binding.ddp_srb_solve(ps1, cfg, n_threads=1)
    print(f"nominal plan (unperturbed initial state, cold start): {int(nominal.iters[0])} iterations, cost {nominal.cost[0]:.3f}\n")
    ps.u_init = np.ascontiguousarray(np.tile(nominal.u, (n, 1, 1)))
    warm = binding.ddp_srb_solve(ps, cfg, n_threads=threads)
    table("warm start from the nominal plan (what every control cycle after the first one does)", warm)


if __name__ == "__main__":
    main()
