"""DdpZmp throughput: the thread-per-problem kernel against the warp-per-problem engine (3 of 32 lanes live).
python tools/bench_zmp.py [batch] -> one JSON line per variant (host-buffer C-ABI call, best of 3)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from centroidalcontrolcollection_b200 import build, engine, problem  # noqa: E402
from footstep_manager import walking_plan  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    N, G = 100, 9.80665
    build.build()
    times = (0.0, 1.9, 2.4, 4.95)
    ref_zmp = np.zeros((len(times), N + 1, 3))
    for s, t0 in enumerate(times):
        fm = walking_plan()
        for tick in range(int(round(t0 / 0.005)) + 1):
            fm.update(tick * 0.005)
        for k in range(N + 1):
            ref_zmp[s, k, :2] = fm.ref_zmp(t0 + k * 0.02)
    rng = np.random.default_rng(4)
    sched_id = (np.arange(B) % len(times)).astype(np.int32)
    x0 = np.zeros((B, 6))
    x0[:, [0, 2]] = ref_zmp[sched_id, 0, :2] + rng.uniform(-0.03, 0.03, (B, 2))
    x0[:, [1, 3]] = rng.uniform(-0.1, 0.1, (B, 2))
    x0[:, 4] = 1.0 + rng.uniform(-0.02, 0.02, B)
    u_init = np.zeros((B, N, 3))
    u_init[:, :, 0], u_init[:, :, 1], u_init[:, :, 2] = x0[:, None, 0], x0[:, None, 2], 100.0 * G
    ps = problem.DdpZmpProblemSet(ref_zmp, np.ones((len(times), N + 1)), sched_id, x0, 100.0, 0.02, u_init=u_init)
    out = {}
    for variant, label in ((1, "thread per problem"), (0, "warp per problem (3 of 32 lanes)")):
        Bv = B if variant == 1 else min(B, 8192)
        psv = ps if Bv == B else problem.DdpZmpProblemSet(ref_zmp, np.ones((len(times), N + 1)), sched_id[:Bv], x0[:Bv], 100.0, 0.02,
                                                          u_init=u_init[:Bv])
        old = engine.DdpZmpEngine.set_variant(variant)
        eng = engine.DdpZmpEngine(N, Bv, len(times))
        engine.DdpZmpEngine.set_variant(old)
        for mi in (3, 40):
            cfg = problem.ddp_config(max_iter=mi)
            res = eng.solve(psv, cfg)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                res = eng.solve(psv, cfg)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out[(variant, mi)] = res
            # device-resident: inputs and outputs stay in HBM, CUDA events around the launch
            import torch

            from centroidalcontrolcollection_b200 import _abi

            dev = torch.device("cuda", 0)
            keep = {k: torch.from_numpy(np.ascontiguousarray(getattr(psv, k))).to(dev) for k in ("ref_zmp", "com_z", "sched_id", "x0", "u_init")}
            bs = psv.as_struct()
            for k, t in keep.items():
                setattr(bs, k, t.data_ptr())
            d_out = dict(x=torch.empty((Bv, N + 1, 6), dtype=torch.float64, device=dev), u=torch.empty((Bv, N, 3), dtype=torch.float64, device=dev),
                         cost=torch.empty(Bv, dtype=torch.float64, device=dev), iters=torch.empty(Bv, dtype=torch.int32, device=dev),
                         status=torch.empty(Bv, dtype=torch.int32, device=dev))
            rs = _abi.DdpResult()
            for k, t in d_out.items():
                setattr(rs, k, t.data_ptr())
            stream = torch.cuda.current_stream(dev)
            ms = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.solve_device(bs, cfg, rs, stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                ms.append(e0.elapsed_time(e1))
            assert np.array_equal(d_out["u"].cpu().numpy(), res.u)
            dev_rate = Bv / (min(ms[1:]) / 1e3)
            print(json.dumps({"kernel": label, "device_resident_solves_per_s": dev_rate, "device_ms": min(ms[1:]), "workload": f"DdpZmp N=100 dt=0.02, 4 schedules, batch {Bv}, warm start, max_iter {mi}",
                              "solves_per_s": Bv / best, "seconds": best, "mean_ddp_iters": float(res.iters.mean()),
                              "api": "ccc_ddp_zmp_solve(CCC_MEM_HOST)"}), flush=True)
        eng.close()
    k = min(B, 8192)
    print(json.dumps({"variants_bit_identical_on_first": k, "u": bool(np.array_equal(out[(1, 40)].u[:k], out[(0, 40)].u[:k])),
                      "iters": bool(np.array_equal(out[(1, 40)].iters[:k], out[(0, 40)].iters[:k]))}))


if __name__ == "__main__":
    main()
