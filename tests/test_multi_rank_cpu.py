"""CPU tier: the N > 1 sharding path with world_size 2 over gloo.  Each rank solves its shard (with the
oracle standing in for the GPU engine, which is not available here) and rank 0 gathers; the result
must equal the single-process solve."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from centroidalcontrolcollection_b200 import distributed as D
    from centroidalcontrolcollection_b200 import problem, workloads
    from oracle import binding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ps_full = None
    if rank == 0:
        ps_full = problem.DdpCentroidalProblemSet.from_workload(workloads.ddp_centroidal_config3(batch=batch, horizon_steps=12))
    ps = D.scatter_problem_set(ps_full, problem.DdpCentroidalProblemSet, src=0)
    lo, hi = D.shard_range(batch, rank, world)
    assert ps.batch == hi - lo
    res = binding.ddp_centroidal_solve(ps, problem.ddp_centroidal_config(max_iter=5))
    got = D.gather_result(res, dst=0)
    if rank == 0:
        np.savez(out_path, **got)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from centroidalcontrolcollection_b200.distributed import shard_range

    for B in (1, 7, 16, 16384, 65536):
        for W in (1, 2, 3, 4, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_ranks_scatter_solve_gather(tmp_path, oracle):
    from centroidalcontrolcollection_b200 import problem, workloads

    batch = 11  # odd: shards of 6 and 5
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), batch, out), nprocs=2, join=True)
    got = np.load(out)
    ps = problem.DdpCentroidalProblemSet.from_workload(workloads.ddp_centroidal_config3(batch=batch, horizon_steps=12))
    ref = oracle.ddp_centroidal_solve(ps, problem.ddp_centroidal_config(max_iter=5))
    for f in ("x", "u", "cost", "iters", "status"):
        assert np.array_equal(got[f], getattr(ref, f)), f
