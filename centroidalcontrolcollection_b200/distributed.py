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


# ---- device path: NCCL collectives on device tensors, no host staging on the non-root ranks ----------------------
_TABLES = {"DdpCentroidalProblemSet": ("m", "ridge", "vertex", "ref_pos"), "DdpSrbProblemSet": ("m", "ridge", "vertex", "inertia", "ref")}


class ShardedDdp:
    """One batch owned by rank `src`, solved on every rank's GPU (reference shape: one caller-side
    batch in, all trajectories out, src/DdpSingleRigidBody.cpp:283-307 per problem).

    solve() on every rank:  src copies the batch host -> device; the shared stage tables are broadcast, the
    per-problem inputs (x0, sched_id) scattered in equal shards (the batch is padded with copies of its last problem
    up to a multiple of the world size); every rank runs its engine on its shard with device buffers in and out;
    x / u / cost / iters / status are gathered on src and copied device -> host.  All collectives are NCCL calls on
    device tensors on the current stream; there is no collective inside the solve itself (problems are independent).
    """

    def __init__(self, engine_cls, problem_cls, horizon_steps, batch_total, n_sched, m_max, device, src=0):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.src, self.dev = src, torch.device(device)
        self.problem_cls, self.N, self.S, self.mm = problem_cls, horizon_steps, n_sched, m_max
        self.nx = problem_cls.nx
        self.B = batch_total
        self.per = -(-batch_total // self.world)  # shard size (last shards padded)
        self.engine = engine_cls(horizon_steps, self.per, n_sched)
        f64, i32 = torch.float64, torch.int32
        d = self.dev
        per, N, nx, mm = self.per, self.N, self.nx, m_max
        self.x0 = torch.empty((per, nx), dtype=f64, device=d)
        self.sid = torch.empty(per, dtype=i32, device=d)
        self.out = dict(x=torch.empty((per, N + 1, nx), dtype=f64, device=d), u=torch.empty((per, N, mm), dtype=f64, device=d),
                        cost=torch.empty(per, dtype=f64, device=d), iters=torch.empty(per, dtype=i32, device=d),
                        status=torch.empty(per, dtype=i32, device=d))
        self.tables = None
        if self.rank == src:
            W = self.world
            self.g_x0 = torch.empty((W * per, nx), dtype=f64, device=d)
            self.g_sid = torch.empty(W * per, dtype=i32, device=d)
            self.g_out = {k: torch.empty((W * per,) + tuple(v.shape[1:]), dtype=v.dtype, device=d) for k, v in self.out.items()}
            self.h_out = {k: torch.empty((W * per,) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory() for k, v in self.out.items()}
        self.last_timing = {}

    def _alloc_tables(self, shapes):
        self.tables = {k: torch.empty(shape, dtype=getattr(torch, dt), device=self.dev) for k, (shape, dt) in shapes.items()}

    def setup(self, ps_full):
        """Once per problem family (not per solve): scalars and table shapes travel as Python objects."""
        names = _TABLES[self.problem_cls.__name__]
        meta = [None]
        if self.rank == self.src:
            s = ps_full.sched
            meta = [dict(shapes={k: (tuple(getattr(s, k).shape), str(getattr(s, k).dtype)) for k in names}, mass=ps_full.mass,
                         dt=ps_full.dt, w_run=ps_full.w_run, w_term=ps_full.w_term, u_lo=ps_full.u_lo, u_hi=ps_full.u_hi)]
        dist.broadcast_object_list(meta, src=self.src, device=self.dev)
        self.meta = meta[0]
        self._alloc_tables(self.meta["shapes"])

    def _batch_struct(self):
        from . import _abi
        from ._abi import ptr

        b = _abi.DdpCentroidalBatch() if self.nx == 9 else _abi.DdpSrbBatch()
        b.horizon_steps, b.batch, b.n_sched, b.m_max = self.N, self.per, self.S, self.mm
        b.dt, b.mass = self.meta["dt"], self.meta["mass"]
        for k, t in self.tables.items():
            setattr(b, k, t.data_ptr())
        for i, v in enumerate(self.meta["w_run"]):
            b.w_run[i] = v
        for i, v in enumerate(self.meta["w_term"]):
            b.w_term[i] = v
        b.u_lo, b.u_hi = self.meta["u_lo"], self.meta["u_hi"]
        b.sched_id, b.x0, b.u_init = self.sid.data_ptr(), self.x0.data_ptr(), None
        r = _abi.DdpResult()
        for k, t in self.out.items():
            setattr(r, k, t.data_ptr())
        r.trace_len = 0
        return b, r

    def solve(self, ps_full, cfg, timing=True):
        """ps_full is read on rank src only (host arrays, ideally pinned).  Returns {x, u, cost, iters, status} as numpy
        views of pinned host buffers on src (first B rows), None elsewhere.  self.last_timing holds the device-timed
        phases of this rank in ms (scatter incl. H2D and table broadcast, solve, gather incl. D2H)."""
        is_src = self.rank == self.src
        stream = torch.cuda.current_stream(self.dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        if is_src:
            s = ps_full.sched
            for k, t in self.tables.items():
                t.copy_(torch.from_numpy(getattr(s, k)), non_blocking=True)
            B = ps_full.batch
            self.g_x0[:B].copy_(torch.from_numpy(ps_full.x0), non_blocking=True)
            self.g_sid[:B].copy_(torch.from_numpy(ps_full.sched_id), non_blocking=True)
            if B < self.g_x0.shape[0]:
                self.g_x0[B:] = self.g_x0[B - 1]
                self.g_sid[B:] = self.g_sid[B - 1]
        for t in self.tables.values():
            dist.broadcast(t, src=self.src)
        dist.scatter(self.x0, list(self.g_x0.chunk(self.world)) if is_src else None, src=self.src)
        dist.scatter(self.sid, list(self.g_sid.chunk(self.world)) if is_src else None, src=self.src)
        ev[1].record(stream)
        b, r = self._batch_struct()
        self.engine.solve_device(b, cfg, r, stream.cuda_stream)
        ev[2].record(stream)
        for k, t in self.out.items():
            dist.gather(t, list(self.g_out[k].chunk(self.world)) if is_src else None, dst=self.src)
        res = None
        if is_src:
            for k in self.out:
                self.h_out[k].copy_(self.g_out[k], non_blocking=True)
        ev[3].record(stream)
        stream.synchronize()
        if is_src:
            res = {k: v.numpy()[: self.B] for k, v in self.h_out.items()}
        if timing:
            self.last_timing = {"scatter_ms": ev[0].elapsed_time(ev[1]), "solve_ms": ev[1].elapsed_time(ev[2]),
                                "gather_ms": ev[2].elapsed_time(ev[3])}
        return res

    def bytes_moved(self):
        """(scatter bytes leaving src, gather bytes arriving at src) per solve over NVLink, excluding src's own shard."""
        per, W = self.per, self.world
        tables = sum(t.numel() * t.element_size() for t in self.tables.values()) * (W - 1)
        sc = (self.x0.numel() * 8 + self.sid.numel() * 4) * (W - 1) + tables
        ga = sum(t.numel() * t.element_size() for t in self.out.values()) * (W - 1)
        return int(sc), int(ga)
