"""A/B of the DDP solver core's feature bits on BASELINE config 3 (device-resident, CUDA events, L2 flush between
steps).  Needs a library built with CCC_AB_VARIANTS=1 (centroidalcontrolcollection_b200/build.py).

    CCC_AB_VARIANTS=1 python -m centroidalcontrolcollection_b200.build --force
    python tools/ab_variants.py [batch] [steps] > gpurun_out/ab.jsonl

Every variant must return the same bits as variant 0 (iteration counts, states, inputs): the features change how the
work is scheduled and staged, not what is computed.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centroidalcontrolcollection_b200 import _abi, engine, problem, workloads  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
names = sys.argv[3].split(",") if len(sys.argv) > 3 else None
dev = torch.device("cuda", 0)
w = workloads.ddp_centroidal_config3(batch=B)
ps = problem.DdpCentroidalProblemSet.from_workload(w)
cfg = problem.ddp_centroidal_config()
N, S, mm = ps.N, ps.sched.S, ps.m_max
eng = engine.DdpCentroidalEngine(N, B, S)
dev_t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_in = dict(sched_id=dev_t(ps.sched_id), m=dev_t(ps.sched.m), ridge=dev_t(ps.sched.ridge), vertex=dev_t(ps.sched.vertex),
            ref_pos=dev_t(ps.sched.ref_pos), x0=dev_t(ps.x0))
d_out = dict(x=torch.empty((B, N + 1, 9), dtype=torch.float64, device=dev), u=torch.empty((B, N, mm), dtype=torch.float64, device=dev),
             cost=torch.empty(B, dtype=torch.float64, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev),
             status=torch.empty(B, dtype=torch.int32, device=dev))
bs = ps.as_struct()
for k, t in d_in.items():
    setattr(bs, k, t.data_ptr())
bs.u_init = None
rs = _abi.DdpResult()
for k, t in d_out.items():
    setattr(rs, k, t.data_ptr())
rs.trace_len = 0
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream(dev)
nvar = engine.DdpCentroidalEngine.set_variant(0)
labels = ["default (abort + tma)", "feat 0 (round-1 core)", "abort only", "tma only"] + [f"variant {i}" for i in range(4, nvar)]
ref = None
for v in range(nvar):
    engine.DdpCentroidalEngine.set_variant(v)
    for _ in range(1):
        flush.zero_()
        eng.solve_device(bs, cfg, rs, stream.cuda_stream)
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.solve_device(bs, cfg, rs, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out = {k: t.cpu().numpy().copy() for k, t in d_out.items()}
    if ref is None:
        ref = out
    same = all(np.array_equal(out[k], ref[k]) for k in out)
    print(json.dumps({"variant": v, "label": labels[v], "batch": B, "ms_per_step": ms, "solves_per_s": B / (min(ms) / 1e3),
                      "mean_iters": float(out["iters"].mean()), "bit_identical_to_variant_0": bool(same)}), flush=True)
