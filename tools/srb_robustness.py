"""Robustness study of the reference's DdpSingleRigidBody closed-loop test (tests/src/TestDdpSingleRigidBody.cpp:15-175,
one DDP iteration per tick after the first) under 1e-9 perturbations of the first initial state, for every unpinned
choice of the restated solver (oracle/num.hpp `Choice`) and for the textbook-arithmetic build.

    python tools/srb_robustness.py [--runs 16] [--later-max-iter 1] > profiles/r02_srb_robustness.txt

TEST TOOL: runs the CPU oracle only.
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANTS = [("canonical (oracle definition)", 0, False), ("textbook arithmetic", 0, True),
            ("BoxQP cold start", 2, False), ("BoxQP warm start from the same stage", 16, False),
            ("Vxx / dV with Quu + lambda I", 4, False), ("descent test s'g > 1e-10", 8, False)]


def one(args):
    name, bits, textbook, later, k, scale = args
    from oracle import binding
    import closed_loop_srb

    orc = binding.textbook() if textbook else binding
    orc.set_choices(bits)
    first = []
    costs = []

    def solve(ps, cfg):
        if not first:
            ps.x0[0, 0] += scale * k
            first.append(1)
        r = orc.ddp_srb_solve(ps, cfg, trace_len=4)
        costs.append((float(r.cost[0]), int(r.alpha_idx[0, 0])))
        return r

    sim, rp, ro, ok, iters = closed_loop_srb.run_ddp_srb_closed_loop(solve, later_max_iter=later)
    fin = [np.linalg.norm(sim.x[0:3] - rp), np.linalg.norm(sim.x[3:6] - ro), np.linalg.norm(sim.x[6:9]),
           np.linalg.norm(sim.x[9:12])]
    passed = bool(ok and all(np.isfinite(fin)) and max(fin) < 0.1)
    rejected = sum(1 for c, a in costs[1:] if a == -1)
    finite = [c for c, _ in costs if np.isfinite(c)]
    return name, k, passed, fin, rejected, max(finite) if finite else float("nan"), iters[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=16)
    ap.add_argument("--later-max-iter", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1e-9)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    tasks = [(n, b, t, a.later_max_iter, k, a.scale) for n, b, t in VARIANTS for k in range(a.runs)]
    with mp.get_context("spawn").Pool(a.procs) as pool:
        out = pool.map(one, tasks, chunksize=1)
    print(f"# DdpSingleRigidBody closed loop, max_iter = {a.later_max_iter} after the first tick, {a.runs} runs per variant, "
          f"first x0[0] perturbed by k * {a.scale:g}")
    for n, _, _ in VARIANTS:
        rows = [r for r in out if r[0] == n]
        npass = sum(r[2] for r in rows)
        pos = [r[3][0] for r in rows if r[2]]
        print(f"{n:40s} pass {npass:2d}/{len(rows)}  final pos err of passing runs {min(pos):.4f}..{max(pos):.4f}  "
              f"rejected ticks {min(r[4] for r in rows)}..{max(r[4] for r in rows)}  first-tick iters "
              f"{sorted(set(r[6] for r in rows))}  failing k = {[r[1] for r in rows if not r[2]]}")


if __name__ == "__main__":
    main()
