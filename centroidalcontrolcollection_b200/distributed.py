"""Batch sharding across ranks (one process per GPU, torch.distributed; NCCL on GPUs, gloo on CPU).

Problems are independent, so the data path has no collective: rank r solves the contiguous slice
shard_range(B, r, W).  The only communication is plumbing at the edges — scatter of the per-problem
inputs (initial states, schedule ids) from rank 0, broadcast of the shared schedule tables, gather of
the result arrays back to rank 0.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous, balanced slice [lo, hi) of `batch` problems for `rank` of `world`."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dev(device):
    return torch.device(device) if device is not None else torch.device("cpu")


def broadcast_array(a, src=0, device=None):
    """Broadcast a numpy array from `src` (shape and dtype travel first)."""
    rank = dist.get_rank()
    meta = [(a.shape, str(a.dtype))] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    shape, dtype = meta[0]
    t = torch.from_numpy(np.ascontiguousarray(a)).to(_dev(device)) if rank == src else torch.empty(
        shape, dtype=getattr(torch, dtype.replace("float64", "float64")), device=_dev(device))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def scatter_rows(a, src=0, device=None):
    """Scatter the rows of `a` (only read on `src`) by shard_range; returns this rank's rows."""
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [(a.shape, str(a.dtype))] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    shape, dtype = meta[0]
    lo, hi = shard_range(shape[0], rank, world)
    out = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=getattr(torch, dtype), device=_dev(device))
    if rank == src:
        full = torch.from_numpy(np.ascontiguousarray(a)).to(_dev(device))
        parts = [full[slice(*shard_range(shape[0], r, world))].contiguous() for r in range(world)]
        # point-to-point sends: shards may differ in length by one row
        reqs = [dist.isend(parts[r], dst=r) for r in range(world) if r != src]
        out.copy_(parts[src])
        for q in reqs:
            q.wait()
    else:
        dist.recv(out, src=src)
    return out.cpu().numpy()


def gather_rows(a, dst=0, device=None):
    """Inverse of scatter_rows: concatenates every rank's rows on `dst` (None elsewhere)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    counts = [None] * world
    dist.all_gather_object(counts, int(a.shape[0]))
    t = torch.from_numpy(np.ascontiguousarray(a)).to(_dev(device))
    if rank == dst:
        parts = []
        for r in range(world):
            if r == dst:
                parts.append(t)
            else:
                buf = torch.empty((counts[r],) + tuple(a.shape[1:]), dtype=t.dtype, device=t.device)
                dist.recv(buf, src=r)
                parts.append(buf)
        return torch.cat(parts, dim=0).cpu().numpy()
    dist.send(t, dst=dst)
    return None


def scatter_problem_set(ps_full, cls, src=0, device=None):
    """Rank `src` holds the full problem set; every rank gets its shard (schedules are broadcast)."""
    rank = dist.get_rank()
    is_src = rank == src
    sched_meta = [(type(ps_full.sched), ps_full.sched.S, ps_full.sched.N, ps_full.sched.m_max)] if is_src else [None]
    dist.broadcast_object_list(sched_meta, src=src)
    sched_cls, S, N, m_max = sched_meta[0]
    sched = ps_full.sched if is_src else sched_cls(S, N, m_max)
    for name in ("m", "ridge", "vertex", "ref_pos", "inertia", "ref"):
        if hasattr(sched, name):
            setattr(sched, name, broadcast_array(getattr(sched, name), src, device))
    scal = [(ps_full.mass, ps_full.dt, ps_full.w_run, ps_full.w_term, ps_full.u_lo, ps_full.u_hi)] if is_src else [None]
    dist.broadcast_object_list(scal, src=src)
    mass, dt, w_run, w_term, u_lo, u_hi = scal[0]
    x0 = scatter_rows(ps_full.x0 if is_src else None, src, device)
    sched_id = scatter_rows(ps_full.sched_id if is_src else None, src, device)
    return cls(sched, sched_id, x0, mass, dt, w_run, w_term, u_lo, u_hi)


def gather_result(res, dst=0, device=None, fields=("x", "u", "cost", "iters", "status")):
    """Gather the per-problem result arrays on `dst`; returns a dict there, None elsewhere."""
    out = {f: gather_rows(getattr(res, f), dst, device) for f in fields}
    return out if dist.get_rank() == dst else None
