"""GPU tier, needs >= 2 devices (gpurun --gpus 2; skipped on a single-GPU box): the sharded data plane on NCCL.
Rank 0 owns the batch; distributed.ShardedDdp scatters it, every rank solves its shard on its own GPU through the
C-ABI with device buffers, rank 0 gathers — the result must equal the single-GPU solve of the whole batch bit for bit
(problems are independent and every GPU runs the same arithmetic)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, kind, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from centroidalcontrolcollection_b200 import distributed as D
    from centroidalcontrolcollection_b200 import engine, problem, workloads

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if kind == "centroidal":
        w = workloads.ddp_centroidal_config3(batch=batch, horizon_steps=20)
        cls, ecls, cfg = problem.DdpCentroidalProblemSet, engine.DdpCentroidalEngine, problem.ddp_centroidal_config(max_iter=8)
    else:
        w = workloads.ddp_srb_config4(batch=batch, horizon_steps=20)
        cls, ecls, cfg = problem.DdpSrbProblemSet, engine.DdpSrbEngine, problem.ddp_srb_config(max_iter=6)
    ps_full = cls.from_workload(w) if rank == 0 else None
    meta = [(w["sched"].N, w["sched"].S, w["sched"].m_max)]
    sh = D.ShardedDdp(ecls, cls, meta[0][0], batch, meta[0][1], meta[0][2], dev)
    sh.setup(ps_full)
    got = sh.solve(ps_full, cfg)
    sc, ga = sh.bytes_moved()
    assert sc > 0 and ga > 0
    if rank == 0:
        ref = ecls(ps_full.N, batch, ps_full.sched.S).solve(ps_full, cfg)
        np.savez(out_path, **{"got_" + k: v for k, v in got.items()},
                 **{"ref_" + k: getattr(ref, k) for k in ("x", "u", "cost", "iters", "status")})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,batch", [("centroidal", 301), ("srb", 96)])
def test_two_gpus_scatter_solve_gather(tmp_path, kind, batch):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), batch, kind, out), nprocs=2, join=True)
    z = np.load(out)
    for f in ("x", "u", "cost", "iters", "status"):
        assert np.array_equal(z["got_" + f], z["ref_" + f]), f
