#!/usr/bin/env python
"""bench.py — MPC solves/sec of the batched DdpCentroidal solve (BASELINE.json north star).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" is one pass of the hot path over one batch: `batch` (default 16384) independent
cold-start DdpCentroidal problems, horizon 50, 4-phase contact schedule x 16 schedule variants
(SURVEY.md §8d config 3), each solved to DDP termination.  Under torchrun (N > 1) every rank
owns one GPU and its own batch (weak scaling, no data-path collective: problems are independent);
the timed region is bracketed by a barrier + device synchronize and the slowest rank counts.

Prints ONE JSON line (see the driver's contract): `value` = solves/s with inputs resident in HBM
(CUDA events on the launching stream), `e2e` = the same through the host-buffer C-ABI call
(pinned host memory, H2D + D2H inside the timed region), `roofline` for the solve kernel against
the measured HBM peak, `cpu_baseline` = the CPU oracle timed on a bounded sample of the same
workload on this box's host cores.

`--impl reference` times the reference's CPU path — here the oracle port, because the reference
itself does not compile in this image — with all host threads, same JSON line.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from centroidalcontrolcollection_b200 import _abi, problem, workloads  # noqa: E402

_REAL_STDOUT = 1
METRIC = "MPC solves/sec (DdpCentroidal horizon=50)"
UNIT = "solves/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes_per_solve(ps):
    """Compulsory HBM traffic of one solve if every intermediate stayed on chip:
    read x0 (72 B) + sched_id (4 B); write the x trajectory, the u trajectory (sum of stage
    dimensions) and cost/iters/status (16 B).  Shared schedule tables amortise to ~0."""
    N = ps.N
    m_sum = ps.sched.m[ps.sched_id].sum(axis=1).mean()
    return 72 + 4 + (N + 1) * 9 * 8 + float(m_sum) * 8 + 16


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_rate(ps, cfg, n_problems, threads):
    """Solves/s of the CPU oracle (oracle/) on the first n_problems of the workload."""
    from oracle import binding

    sub = ps.subset(np.arange(min(n_problems, ps.batch)))
    t0 = time.perf_counter()
    binding.ddp_centroidal_solve(sub, cfg, trace_len=0, n_threads=threads)
    dt = time.perf_counter() - t0
    return sub.batch / dt, sub.batch, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on all host threads, rank 0 only."""
    if rank != 0:
        return
    from oracle import binding

    binding.build()
    threads = binding.hardware_threads()
    w = workloads.ddp_centroidal_config3(batch=args.batch)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config()
    per_step = min(args.batch, max(threads * args.ref_problems_per_thread, 1))
    for _ in range(args.warmup):
        cpu_oracle_rate(ps, cfg, max(threads, 1), threads)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        _, k, _ = cpu_oracle_rate(ps, cfg, per_step, threads)
        n += k
    dt = time.perf_counter() - t0
    value = n / dt
    sample = f"{per_step} of {args.batch} problems per step (first problems of the seeded batch), {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "note": "reference does not compile here (Eigen/nmpc_ddp absent): "
                   "timed the CPU oracle port of its algorithm; each step is a bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def _emit(line):
    """The JSON line goes to the real stdout; everything else a library may print there (NCCL's version banner
    under torchrun) was redirected to stderr at start-up, so that stdout carries exactly one line."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def run_ours(args, rank, world, local_rank):
    import torch

    from centroidalcontrolcollection_b200 import build, engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    build.build()
    engine.lib()
    if args.variant is not None:
        engine.DdpCentroidalEngine.set_variant(args.variant)
    if args.chunk is not None:
        engine.DdpCentroidalEngine.set_chunk(args.chunk)

    B = args.batch
    w = workloads.ddp_centroidal_config3(batch=B, seed=20260102 + rank)
    ps = problem.DdpCentroidalProblemSet.from_workload(w)
    cfg = problem.ddp_centroidal_config()
    if args.max_iter is not None:
        cfg.max_iter = args.max_iter
    N, S, mm = ps.N, ps.sched.S, ps.m_max
    eng = engine.DdpCentroidalEngine(N, B, S)

    # ---- device-resident inputs / outputs (torch = device memory + stream plumbing only) ----
    def dev_t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    d_in = dict(sched_id=dev_t(ps.sched_id), m=dev_t(ps.sched.m), ridge=dev_t(ps.sched.ridge),
                vertex=dev_t(ps.sched.vertex), ref_pos=dev_t(ps.sched.ref_pos), x0=dev_t(ps.x0))
    d_out = dict(x=torch.empty((B, N + 1, 9), dtype=torch.float64, device=dev),
                 u=torch.empty((B, N, mm), dtype=torch.float64, device=dev),
                 cost=torch.empty(B, dtype=torch.float64, device=dev),
                 iters=torch.empty(B, dtype=torch.int32, device=dev),
                 status=torch.empty(B, dtype=torch.int32, device=dev))
    bs = ps.as_struct()
    for k, t in d_in.items():
        setattr(bs, k, t.data_ptr())
    bs.u_init = None
    rs = _abi.DdpResult()
    for k, t in d_out.items():
        setattr(rs, k, t.data_ptr())
    rs.trace_len = 0
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def step_device():
        eng.solve_device(bs, cfg, rs, stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for s0, s1 in ev:
        flush.zero_()  # evict the previous step's lines from L2 (outside the event pair)
        s0.record(stream)
        step_device()
        s1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(ms)
    launches_per_step = eng.last_launches
    iters = d_out["iters"].cpu().numpy()
    status = d_out["status"].cpu().numpy()

    # ---- end to end through the host-buffer C-ABI call, pinned host memory ----
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    hs = type(ps.sched)(S, N, mm)
    for name in ("m", "ridge", "vertex", "ref_pos"):
        t, arr = pin(getattr(ps.sched, name))
        keep.append(t)
        setattr(hs, name, arr)
    t_sid, a_sid = pin(ps.sched_id)
    t_x0, a_x0 = pin(ps.x0)
    keep += [t_sid, t_x0]
    ps_h = problem.DdpCentroidalProblemSet(hs, a_sid, a_x0, ps.mass, ps.dt, ps.w_run, ps.w_term, ps.u_lo, ps.u_hi)
    res_h = ps_h.new_result(0)
    for name in ("x", "u", "cost", "iters", "status", "clamped"):
        t, arr = pin(getattr(res_h, name))
        keep.append(t)
        setattr(res_h, name, arr)
    h2d = ps_h.sched_id.nbytes + hs.m.nbytes + hs.ridge.nbytes + hs.vertex.nbytes + hs.ref_pos.nbytes + ps_h.x0.nbytes
    d2h = res_h.x.nbytes + res_h.u.nbytes + res_h.cost.nbytes + res_h.iters.nbytes + res_h.status.nbytes + res_h.clamped.nbytes
    e2e_steps = max(1, min(args.steps, 3))
    eng.solve(ps_h, cfg, result=res_h)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.solve(ps_h, cfg, result=res_h)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks: slowest rank counts ----
    t_dev = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t_dev[0]), float(t_dev[1])
    value = world * B * args.steps / (total_ms_max / 1e3)
    e2e_value = world * B * e2e_steps / e2e_s_max

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        abytes = algorithmic_bytes_per_solve(ps)
        kernel_ms = total_ms / args.steps
        achieved = abytes * B / (kernel_ms / 1e3) / 1e9
        traffic, util = None, None
        tp = os.path.join(ROOT, "profiles", "ddp_centroidal_traffic.json")
        if os.path.exists(tp):
            prof = json.load(open(tp))
            traffic = prof.get("dram_bytes_per_launch")
            # what actually binds this kernel: issue slots and the FP64 pipe, from the ncu pass on this launch
            util = {"issue_slots_pct": prof.get("issue_active_pct"), "fp64_pipe_pct": prof.get("fp64_pipe_active_pct"),
                    "l2_hit_rate_pct": prof.get("l2_hit_rate_pct"), "source": "profiles/ddp_centroidal_traffic.json (ncu, same launch)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "per_gpu_batch": B, "cold_start": True, "max_iter": int(cfg.max_iter),
                       "l2": "256 MiB flush write between timed steps; per-step working set (gain lists, 2.6 GB) > L2",
                       "mean_ddp_iters": float(iters.mean()), "max_ddp_iters": int(iters.max()),
                       "converged_frac": float((status == 1).mean()), "wall_s_timed_region": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "ccc_ddp_centroidal_solve(CCC_MEM_HOST), pinned host buffers"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "ccc_host::ddp_solve_kernel<ccc::CentroidalModel,8,1,true>",
                         "algorithmic_bytes_per_solve": abytes, "kernel_ms_per_launch": kernel_ms, "measured_utilisation": util,
                         "note": "latency/FP64-issue bound serial recursion: HBM fraction is small by nature, see DESIGN.md §5"},
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import binding

            binding.build()
            threads = binding.hardware_threads()
            n_sample = min(B, max(threads * args.cpu_baseline_problems_per_thread, 8))
            v, k, dt = cpu_oracle_rate(ps, cfg, n_sample, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"first {k} of {B} problems of the same seeded batch, {dt:.1f} s wall"}
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="problems per GPU per step")
    ap.add_argument("--variant", type=int, default=None, help="launch-shape variant of the solve kernel (tuning)")
    ap.add_argument("--chunk", type=int, default=None, help="DDP iterations per visit before a solve is re-queued (tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-iter", type=int, default=None, help="experiments only: cap DDP iterations (default 500 = nmpc_ddp's)")
    ap.add_argument("--ref-problems-per-thread", type=int, default=24,
                    help="--impl reference: problems per host thread per step (~0.1-0.2 s each)")
    ap.add_argument("--cpu-baseline-problems-per-thread", type=int, default=96,
                    help="cpu_baseline leg: bounded sample, problems per host thread (~10-20 s in total)")
    args = ap.parse_args()
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray library output on stdout -> stderr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
