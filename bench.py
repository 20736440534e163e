#!/usr/bin/env python
"""bench.py — MPC solves/sec of the batched DdpCentroidal solve (BASELINE.json north star).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" is one pass of the hot path over one batch: `batch` (default 16384) independent
cold-start DdpCentroidal problems, horizon 50, 4-phase contact schedule x 16 schedule variants
(SURVEY.md §8d config 3), each solved to DDP termination.  Under torchrun (N > 1) every rank
owns one GPU and `batch` problems (weak scaling; the solve itself has no collective: problems are
independent); the timed region is bracketed by a barrier + device synchronize and the slowest rank counts.

Prints ONE JSON line (see the driver's contract):
  `value`   solves/s with inputs resident in HBM (CUDA events on the launching stream; per-rank times in `per_rank`);
  `e2e`     N = 1: the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region);
            N > 1: the sharded data plane — rank 0 owns all N x batch problems in host memory, distributed.ShardedDdp
            scatters them over NCCL, every rank solves its shard, rank 0 gathers the trajectories and copies them to
            the host (scatter / solve / gather times and NVLink bytes in `e2e.sharded`);
  `roofline` the solve kernel against the measured HBM peak, plus `roofline.fp64`: the kernel's DFMA count (ncu
            counters of the same launch, profiles/r02_ddp_centroidal_counters.json, stamped with the code version)
            against the FP64 fma peak measured in this run by ccc_fp64_peak_tflops;
  `cpu_baseline` the CPU oracle (a scalar port, not Eigen) on a bounded sample of the same workload on this box's cores;
  `config.other` (N = 1) the other BASELINE configurations in the same run: 2 (LinearMpcZmp), 4 (DdpSingleRigidBody,
            one 8192-problem shard), 5 (IntrinsicallyStableMpc), LinearMpcXY — value through the host-buffer API,
            algorithmic bytes, and a parity flag against the oracle.

`--workload config4` benches DdpSingleRigidBody instead (BASELINE config 4: N = 100, batch 65536 sharded across the
ranks: `--batch` is per GPU, 8192 by default there) with the same line layout.

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


def algorithmic_bytes_per_solve(ps, nx=9):
    """Compulsory HBM traffic of one solve if every intermediate stayed on chip:
    read x0 (8 nx B) + sched_id (4 B); write the x trajectory, the u trajectory (sum of stage
    dimensions) and cost/iters/status (16 B).  Shared schedule tables amortise to ~0."""
    N = ps.N
    m_sum = ps.sched.m[ps.sched_id].sum(axis=1).mean()
    return 8 * nx + 4 + (N + 1) * nx * 8 + float(m_sum) * 8 + 16


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


def cpu_oracle_rate(ps, cfg, n_problems, threads, kind="centroidal"):
    """Solves/s of the CPU oracle (oracle/) on the first n_problems of the workload."""
    from oracle import binding

    sub = ps.subset(np.arange(min(n_problems, ps.batch)))
    t0 = time.perf_counter()
    (binding.ddp_centroidal_solve if kind == "centroidal" else binding.ddp_srb_solve)(sub, cfg, trace_len=0, n_threads=threads)
    dt = time.perf_counter() - t0
    return sub.batch / dt, sub.batch, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on all host threads, rank 0 only."""
    if rank != 0:
        return
    from oracle import binding

    binding.build()
    threads = binding.hardware_threads()
    wl_fn, ps_name, _, cfg_name, default_batch = WORKLOADS[args.workload]
    kind = "centroidal" if args.workload == "config3" else "srb"
    batch = args.batch if args.batch is not None else default_batch
    w = getattr(workloads, wl_fn)(batch=batch)
    ps = getattr(problem, ps_name).from_workload(w)
    cfg = getattr(problem, cfg_name)()
    per_step = min(batch, max(threads * args.ref_problems_per_thread // (1 if kind == "centroidal" else 12), 1))
    for _ in range(args.warmup):
        cpu_oracle_rate(ps, cfg, max(threads, 1), threads, kind)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        _, k, _ = cpu_oracle_rate(ps, cfg, per_step, threads, kind)
        n += k
    dt = time.perf_counter() - t0
    value = n / dt
    sample = (f"{per_step} of {batch} problems per step (first problems of the seeded batch), {threads} threads; scalar C++ port of the "
              "reference's algorithm (-O3 -mavx2 -mfma, no SIMD linear algebra), not the reference's Eigen build")
    line = {
        "impl": "reference", "metric": METRIC if kind == "centroidal" else "MPC solves/sec (DdpSingleRigidBody horizon=100)", "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
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


WORKLOADS = {
    # name -> (workload factory, problem-set class name, engine class name, config factory name, default per-GPU batch)
    "config3": ("ddp_centroidal_config3", "DdpCentroidalProblemSet", "DdpCentroidalEngine", "ddp_centroidal_config", 16384),
    "config4": ("ddp_srb_config4", "DdpSrbProblemSet", "DdpSrbEngine", "ddp_srb_config", 8192),
}


def _timed(fn, reps):
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out, best


def other_configs(threads):
    """The other BASELINE configurations, each through the host-buffer C-ABI call (H2D + solve + D2H inside the timed
    call, best of 2) with a parity check of the engine's output against the CPU oracle on the same inputs."""
    from centroidalcontrolcollection_b200 import engine, linear_mpc
    from oracle import binding

    out = []

    def qp_case(name, run, axes, abytes, parity_n):
        qp = engine.qp_solver_for()
        cap = {}

        def grab(ps):
            cap["ps"] = ps
            return qp(ps)

        run(grab)  # warm-up; captures the assembled QP batch
        ps = cap["ps"]
        # per-problem vectors and results in pinned host memory, as in the headline e2e leg
        import torch

        keep = []

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            keep.append(t)
            return t.numpy()

        for f in ("c", "b", "d"):
            if getattr(ps, f) is not None:
                setattr(ps, f, pin(getattr(ps, f)))
        res = ps.new_result()
        for f in ("x", "iters", "status", "n_active", "active"):
            setattr(res, f, pin(getattr(res, f)))
        eng = engine.QpEngine(ps.n, ps.n_eq, ps.n_ineq, ps.batch)
        eng.solve(ps, result=res)
        _, dt_setup = _timed(lambda: eng.solve(ps, result=res), 3)
        x_with_setup = res.x.copy()
        # every tick after a controller's first: matrices and factorisation resident, per-problem vectors new (what the
        # drop-in classes do, CCC/detail/QpEngine.h; the reference rewrites only qp_coeff_'s vectors in procOnce)
        _, dt = _timed(lambda: eng.solve(ps, result=res, reuse_matrices=True), 3)
        assert np.array_equal(x_with_setup, res.x)
        # device-resident: per-problem vectors and results stay in HBM, matrices factorised by the call above (Q = NULL)
        dev = torch.device("cuda", torch.cuda.current_device())
        dkeep = {f: torch.from_numpy(np.ascontiguousarray(getattr(ps, f))).to(dev) for f in ("c", "b", "d") if getattr(ps, f) is not None}
        dout = dict(x=torch.empty((ps.batch, ps.n), dtype=torch.float64, device=dev), iters=torch.empty(ps.batch, dtype=torch.int32, device=dev),
                    status=torch.empty(ps.batch, dtype=torch.int32, device=dev))
        bs, rs = ps.as_struct(), _abi.QpResult()
        bs.Q = bs.A = bs.C = None
        for f in ("c", "b", "d"):
            setattr(bs, f, dkeep[f].data_ptr() if f in dkeep else None)
        for f, t in dout.items():
            setattr(rs, f, t.data_ptr())
        stq = torch.cuda.current_stream(dev)
        msq = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stq)
            eng.solve_device(bs, rs, stq.cuda_stream)
            e1.record(stq)
            torch.cuda.synchronize(dev)
            msq.append(e0.elapsed_time(e1))
        dev_ok = bool(np.array_equal(dout["x"].cpu().numpy(), res.x))
        eng.close()
        k = min(ps.batch, parity_n)
        sub = ps.subset(np.arange(k))
        ref = binding.qp_solve(sub, n_threads=threads)
        parity = bool(np.array_equal(ref.x, res.x[:k]) and np.array_equal(ref.iters, res.iters[:k])
                      and ref.active_sets() == [tuple(sorted(int(v) for v in row[:n])) for row, n in zip(res.active[:k], res.n_active[:k])])
        unit = "2-axis solves/s" if axes == 2 else "solves/s"
        out.append({"workload": name, "unit": unit, "value": ps.batch / axes / (min(msq[1:]) / 1e3), "kernel_ms": min(msq[1:]),
                    "device_path_equals_host_path": dev_ok, "e2e": ps.batch / axes / dt, "e2e_with_setup": ps.batch / axes / dt_setup, "qps": int(ps.batch), "n": ps.n, "n_eq": ps.n_eq,
                    "n_ineq": ps.n_ineq, "mean_active_set_iterations": float(res.iters.mean()), "solved_frac": float((res.status == 0).mean()),
                    "algorithmic_bytes_per_solve": abytes, "parity": {"bit_exact_vs_oracle": parity, "checked": int(k), "of": int(ps.batch)},
                    "api": "ccc_qp_solve(CCC_MEM_HOST), pinned host buffers; e2e = a controller's tick (matrices and factorisation "
                           "resident, Q = NULL), e2e_with_setup = matrices uploaded and factorised in the call"})

    w = workloads.linear_mpc_zmp_config2()
    mpc2 = linear_mpc.LinearMpcZmp(w["com_height"], w["horizon_duration"], w["horizon_dt"])
    qp_case("config 2: " + w["name"], lambda q, w=w: mpc2.plan_batch(q, w["pos"], w["vel"], w["acc"], w["lim_min"], w["lim_max"], w["control_dt"]),
            2, 2 * (24 + 1600) + 2 * (800 + 8 + 32), 8192)
    w5 = workloads.ismpc_config5()
    mpc5 = linear_mpc.IntrinsicallyStableMpc(w5["com_height"], w5["horizon_duration"], w5["horizon_dt"])
    qp_case("config 5: " + w5["name"], lambda q: mpc5.plan_batch(q, w5["capture_point"], w5["planned_zmp"], w5["ref_zmp"], w5["lim_min"],
                                                               w5["lim_max"], w5["control_dt"]), 2, 2 * 16 + 9 + 2 * (800 + 8 + 32), 16384)
    psxy = workloads.linear_mpc_xy_problem_set(15, 1184)
    qp_case("LinearMpcXY: reference test schedule, n = 240, 15 equalities, 480 bound rows, batch 1184", lambda q: q(psxy), 1,
            6 * 8 + 240 * 8 + 240 * 8 + 8 + 32, 128)

    # config 5 again, full size, with the footstep plans compiled and the QP vectors assembled on the device: the host
    # hands over 256 plans and 131072 (capture point, planned ZMP) pairs and reads 131072 planned ZMPs back
    from centroidalcontrolcollection_b200 import schedule

    rng5 = np.random.Generator(np.random.PCG64(20260104))
    lengths, widths = rng5.uniform(0.1, 0.3, 16), rng5.uniform(0.16, 0.24, 16)
    plans = schedule.FootstepPlans(256, 8, mpc5.mpc_1d.horizon_steps, w5["horizon_dt"], eps_reps=2)
    for p_ in range(256):
        sl, sw = lengths[p_ % 16], widths[(p_ // 16) % 16]
        plans.current_time[p_] = 1.8
        plans.stance0[p_] = [[0.0, 0.5 * sw], [0.0, -0.5 * sw]]
        for foot, x, ts in [(0, sl, 2.0), (1, 2 * sl, 3.0), (0, 3 * sl, 4.0), (1, 4 * sl, 5.0), (0, 3 * sl, 6.0), (1, 3 * sl, 7.0)]:
            plans.append_footstep(p_, foot, (x, 0.5 * sw if foot == 0 else -0.5 * sw), ts, 0.2, 0.8)
    plan_id = np.repeat(np.arange(256, dtype=np.int32), 512)
    state5 = np.ascontiguousarray(np.stack([w5["capture_point"], w5["planned_zmp"]], axis=2))
    eng5 = engine.ZmpMpcEngine(mpc5, len(plan_id), 256)

    def run5():
        tables = engine.footstep_compile(plans)
        return eng5.plan(state5, plan_id, tables, w5["control_dt"], want_qp_info=True)

    run5()
    (planned5, iters5, status5), dt5 = _timed(run5, 3)
    launches5 = eng5.last_launches + 1
    eng5.close()
    k = 4096
    ref5 = mpc5.plan_batch(lambda ps: binding.qp_solve(ps, n_threads=threads), w5["capture_point"][:k], w5["planned_zmp"][:k],
                           w5["ref_zmp"][:k], w5["lim_min"][:k], w5["lim_max"][:k], w5["control_dt"])
    out.append({"workload": "config 5 on the device: 256 footstep plans compiled + " + w5["name"], "unit": "2-axis solves/s",
                "e2e": len(plan_id) / dt5, "qps": int(2 * len(plan_id)), "n": 100, "n_eq": 1, "n_ineq": 200,
                "mean_active_set_iterations": float(iters5.mean()), "solved_frac": float((status5 == 0).mean()), "gpu_launches": launches5,
                "h2d_bytes": int(state5.nbytes + plan_id.nbytes + plans.times.nbytes + plans.pos.nbytes), "d2h_bytes": int(planned5.nbytes + iters5.nbytes + status5.nbytes),
                "algorithmic_bytes_per_solve": 2 * 16 + 4 + 16,
                "parity": {"planned_zmp_max_abs_diff_vs_host_class_with_oracle_qp": float(np.abs(ref5 - planned5[:k]).max()), "checked": k,
                           "of": int(len(plan_id))},
                "api": "ccc_footstep_compile + ccc_zmp_mpc_plan (CCC_MEM_HOST)"})

    # LinearMpcXY over a sweep of schedules, everything after the callback sampling on the device
    sweep = workloads.linear_mpc_xy_sweep(n_sched=256, per_sched=16)
    engxy = engine.LinearMpcXyEngine(sweep.N, sweep.n, sweep.n_eq, sweep.batch, sweep.S)
    resxy = sweep.new_result(intermediates=False)
    engxy.solve(sweep, result=resxy)
    _, dtxy = _timed(lambda: engxy.solve(sweep, result=resxy), 2)
    launches_xy = engxy.last_launches
    engxy.close()
    sub = sweep.first_schedules(8)
    refxy = binding.linear_mpc_xy_solve(sub, n_threads=threads, intermediates=False)
    k = sub.batch
    parity_xy = bool(np.array_equal(refxy.u, resxy.u[sub.index]) and np.array_equal(refxy.iters, resxy.iters[sub.index])
                     and np.array_equal(refxy.active, resxy.active[sub.index]))
    t0 = time.perf_counter()
    for s_ in range(4):
        sweep.host_problem(s_, np.where(sweep.sched_id == s_)[0])
    host_setup_s = (time.perf_counter() - t0) / 4
    out.append({"workload": "LinearMpcXY sweep: 256 schedules x 16 initial states, n = 240, 15 equalities, 480 bound rows",
                "unit": "solves/s", "e2e": sweep.batch / dtxy, "qps": int(sweep.batch), "schedules": int(sweep.S), "n": sweep.n,
                "n_eq": sweep.n_eq, "n_ineq": 2 * sweep.n, "mean_active_set_iterations": float(resxy.iters.mean()),
                "solved_frac": float((resxy.status == 0).mean()), "gpu_launches": launches_xy,
                "algorithmic_bytes_per_solve": 6 * 8 + 4 + 240 * 8 + 240 * 4 + 12,
                "host_setup_s_per_schedule_numpy": host_setup_s,
                "parity": {"bit_exact_vs_oracle": parity_xy, "checked": int(k), "of": int(sweep.batch)},
                "api": "ccc_linear_mpc_xy_solve(CCC_MEM_HOST): stage models, ZOH, condensing, DMMA B'WB, QP setup per schedule, QP"})

    # DdpZmp (the seventh north-star method): 65536 warm-started problems on 4 sampled walking-plan schedules, max_iter 3
    # as the reference's test loop uses (tests/src/TestDdpZmp.cpp:88); one thread per problem
    import torch

    wz = workloads.ddp_zmp_batch(batch=65536)
    psz = problem.DdpZmpProblemSet(wz["ref_zmp"], wz["com_z"], wz["sched_id"], wz["x0"], wz["mass"], wz["dt"], u_init=wz["u_init"])
    cfgz = problem.ddp_config(max_iter=3)
    engz = engine.DdpZmpEngine(psz.N, psz.batch, len(wz["ref_zmp"]))
    resz = engz.solve(psz, cfgz)
    _, dtz = _timed(lambda: engz.solve(psz, cfgz), 2)
    dev = torch.device("cuda", torch.cuda.current_device())
    keepz = {k_: torch.from_numpy(np.ascontiguousarray(getattr(psz, k_))).to(dev) for k_ in ("ref_zmp", "com_z", "sched_id", "x0", "u_init")}
    bsz = psz.as_struct()
    for k_, t_ in keepz.items():
        setattr(bsz, k_, t_.data_ptr())
    d_outz = dict(x=torch.empty((psz.batch, psz.N + 1, 6), dtype=torch.float64, device=dev),
                  u=torch.empty((psz.batch, psz.N, 3), dtype=torch.float64, device=dev),
                  iters=torch.empty(psz.batch, dtype=torch.int32, device=dev), status=torch.empty(psz.batch, dtype=torch.int32, device=dev))
    rsz = _abi.DdpResult()
    for k_, t_ in d_outz.items():
        setattr(rsz, k_, t_.data_ptr())
    stz = torch.cuda.current_stream(dev)
    msz = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stz)
        engz.solve_device(bsz, cfgz, rsz, stz.cuda_stream)
        e1.record(stz)
        torch.cuda.synchronize(dev)
        msz.append(e0.elapsed_time(e1))
    engz.close()
    kz = 1024
    psz_sub = problem.DdpZmpProblemSet(wz["ref_zmp"], wz["com_z"], wz["sched_id"][:kz], wz["x0"][:kz], wz["mass"], wz["dt"], u_init=wz["u_init"][:kz])
    refz = binding.ddp_zmp_solve(psz_sub, cfgz, n_threads=threads)
    parity_z = bool(np.array_equal(refz.u, resz.u[:kz]) and np.array_equal(refz.x, resz.x[:kz]) and np.array_equal(refz.iters, resz.iters[:kz]))
    out.append({"workload": wz["name"], "unit": "solves/s", "value": psz.batch / (min(msz[1:]) / 1e3), "e2e": psz.batch / dtz,
                "kernel_ms": min(msz[1:]), "mean_ddp_iters": float(resz.iters.mean()), "gpu_launches": 1,
                "algorithmic_bytes_per_solve": 48 + 4 + 2400 + 4848 + 2400 + 16,
                "parity": {"bit_exact_vs_oracle": parity_z, "checked": kz, "of": int(psz.batch)},
                "api": "ccc_ddp_zmp_solve: value = CCC_MEM_DEVICE (CUDA events), e2e = CCC_MEM_HOST (pageable host buffers, PCIe bound)"})

    w4 = workloads.ddp_srb_config4(batch=8192)
    ps4 = problem.DdpSrbProblemSet.from_workload(w4)
    eng4 = engine.DdpSrbEngine(ps4.N, ps4.batch, ps4.sched.S)
    cfg4 = problem.ddp_srb_config()
    eng4.solve(ps4.subset(np.arange(256)), cfg4)  # warm-up on a slice
    res4, dt4 = _timed(lambda: eng4.solve(ps4, cfg4, trace_len=0), 1)
    k = 4 * threads
    ref4 = binding.ddp_srb_solve(ps4.subset(np.arange(k)), cfg4, n_threads=threads)
    parity4 = bool(np.array_equal(ref4.iters, res4.iters[:k]) and np.array_equal(ref4.x, res4.x[:k]) and np.array_equal(ref4.u, res4.u[:k]))
    m_sum = float(ps4.sched.m[ps4.sched_id].sum(axis=1).mean())
    out.append({"workload": "config 4 (one of 8 shards): " + w4["name"], "unit": "solves/s", "e2e": ps4.batch / dt4,
                "mean_ddp_iters": float(res4.iters.mean()), "max_ddp_iters": int(res4.iters.max()),
                "converged_frac": float((res4.status == 1).mean()),
                "algorithmic_bytes_per_solve": 96 + 4 + 101 * 12 * 8 + m_sum * 8 + 16,
                "parity": {"bit_exact_vs_oracle": parity4, "checked": int(k), "of": int(ps4.batch)}, "api": "ccc_ddp_srb_solve(CCC_MEM_HOST)"})
    eng4.close()

    # planOnce-sized calls (the reference's own caller: one problem per control tick): small batches run on the team
    # kernel (one CTA of 8 warps per problem, the line search as one round of concurrent rollouts)
    def small_batch(B):
        wb = workloads.ddp_centroidal_config3(batch=B, n_sched=1)
        pb = problem.DdpCentroidalProblemSet.from_workload(wb)
        eb = engine.DdpCentroidalEngine(pb.N, B, 1)
        cfg = problem.ddp_centroidal_config()
        cold, t_cold = _timed(lambda: eb.solve(pb, cfg), 3)
        refb = binding.ddp_centroidal_solve(pb, cfg, n_threads=threads)
        par = bool(np.array_equal(refb.x, cold.x) and np.array_equal(refb.u, cold.u) and np.array_equal(refb.iters, cold.iters))
        pb.u_init = cold.u.copy()
        cfg1 = problem.ddp_centroidal_config(max_iter=1)
        _, t_warm = _timed(lambda: eb.solve(pb, cfg1), 20)
        team = bool(eb.last_team)
        eb.close()
        return {"batch": B, "cold_solve_ms": t_cold * 1e3, "cold_ddp_iters_max": int(cold.iters.max()), "warm_tick_ms": t_warm * 1e3,
                "team_kernel": team, "bit_exact_vs_oracle": par}

    out.append({"workload": "DdpCentroidal N=50, planOnce-sized batches (latency, not throughput)", "unit": "ms per call",
                "cases": [small_batch(1), small_batch(148)],
                "api": "ccc_ddp_centroidal_solve(CCC_MEM_HOST), pageable buffers, time.perf_counter around the call, best of 3 / 20; "
                       "warm tick = max_iter 1 from the converged plan (tests/src/TestDdpCentroidal.cpp:116)"})
    return out


def load_counters(kind):
    """ncu counters of the bench launch on a stated code version (tools/capture_counters.sh): DRAM bytes, DFMA count."""
    p = os.path.join(ROOT, "profiles", f"r02_ddp_{kind}_counters.json")
    return json.load(open(p)) if os.path.exists(p) else None


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
    L = engine.lib()
    if args.variant is not None:
        engine.DdpCentroidalEngine.set_variant(args.variant)
    if args.chunk is not None:
        engine.DdpCentroidalEngine.set_chunk(args.chunk)

    wl_fn, ps_name, eng_name, cfg_name, default_batch = WORKLOADS[args.workload]
    ps_cls, eng_cls = getattr(problem, ps_name), getattr(engine, eng_name)
    kind = "centroidal" if args.workload == "config3" else "srb"
    B = args.batch if args.batch is not None else default_batch
    seed0 = 20260102 if args.workload == "config3" else 20260103
    w = getattr(workloads, wl_fn)(batch=B, seed=seed0 + rank)
    ps = ps_cls.from_workload(w)
    cfg = getattr(problem, cfg_name)()
    if args.max_iter is not None:
        cfg.max_iter = args.max_iter
    N, S, mm, nx = ps.N, ps.sched.S, ps.m_max, ps_cls.nx
    eng = eng_cls(N, B, S)

    # ---- device-resident inputs / outputs (torch = device memory + stream plumbing only) ----
    def dev_t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    table_names = ("m", "ridge", "vertex", "ref_pos") if kind == "centroidal" else ("m", "ridge", "vertex", "inertia", "ref")
    d_in = {k: dev_t(getattr(ps.sched, k)) for k in table_names}
    d_in.update(sched_id=dev_t(ps.sched_id), x0=dev_t(ps.x0))
    d_out = dict(x=torch.empty((B, N + 1, nx), dtype=torch.float64, device=dev),
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
    fp64_peak = float(L.ccc_fp64_peak_tflops(3, C.c_void_p(stream.cuda_stream))) if rank == 0 else 0.0

    # ---- end to end ----
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    e2e_steps = max(1, min(args.steps, 3))
    sharded = None
    if world == 1:
        # the host-buffer C-ABI call, pinned host memory
        hs = type(ps.sched)(S, N, mm)
        for name in table_names:
            t, arr = pin(getattr(ps.sched, name))
            keep.append(t)
            setattr(hs, name, arr)
        t_sid, a_sid = pin(ps.sched_id)
        t_x0, a_x0 = pin(ps.x0)
        keep += [t_sid, t_x0]
        ps_h = ps_cls(hs, a_sid, a_x0, ps.mass, ps.dt, ps.w_run, ps.w_term, ps.u_lo, ps.u_hi)
        res_h = ps_h.new_result(0)
        for name in ("x", "u", "cost", "iters", "status", "clamped"):
            t, arr = pin(getattr(res_h, name))
            keep.append(t)
            setattr(res_h, name, arr)
        h2d = ps_h.sched_id.nbytes + ps_h.x0.nbytes + sum(getattr(hs, n).nbytes for n in table_names)
        d2h = res_h.x.nbytes + res_h.u.nbytes + res_h.cost.nbytes + res_h.iters.nbytes + res_h.status.nbytes + res_h.clamped.nbytes
        eng.solve(ps_h, cfg, result=res_h)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.solve(ps_h, cfg, result=res_h)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_api = f"ccc_ddp_{kind}_solve(CCC_MEM_HOST), pinned host buffers"
    else:
        # the sharded data plane: rank 0 owns all world x B problems (the per-rank batches of the timed region above,
        # concatenated), NCCL scatter -> solve on every GPU -> NCCL gather -> rank 0's host memory
        from centroidalcontrolcollection_b200 import distributed as D

        eng.close()
        Bt = world * B
        ps_full = None
        if rank == 0:
            parts = [ps_cls.from_workload(getattr(workloads, wl_fn)(batch=B, seed=seed0 + r)) for r in range(world)]
            t_x0, a_x0 = pin(np.concatenate([q.x0 for q in parts]))
            t_sid, a_sid = pin(np.concatenate([q.sched_id for q in parts]))
            keep += [t_x0, t_sid]
            ps_full = ps_cls(ps.sched, a_sid, a_x0, ps.mass, ps.dt, ps.w_run, ps.w_term, ps.u_lo, ps.u_hi)
        sh = D.ShardedDdp(eng_cls, ps_cls, N, Bt, S, mm, dev)
        sh.setup(ps_full)
        got = sh.solve(ps_full, cfg)  # warm-up; also the cross-check below
        if rank == 0:
            # the shard this rank solved device-resident above must come back identical through the sharded path
            assert np.array_equal(got["iters"][:B], iters) and np.array_equal(got["x"][:B], d_out["x"].cpu().numpy()), \
                "sharded solve differs from the single-rank solve of the same problems"
        barrier()
        t0 = time.perf_counter()
        phases = []
        for _ in range(e2e_steps):
            sh.solve(ps_full, cfg)
            phases.append(dict(sh.last_timing))
        barrier()
        e2e_s = time.perf_counter() - t0
        sc_b, ga_b = sh.bytes_moved()
        h2d = (Bt * nx * 8 + Bt * 4 + sum(getattr(ps.sched, n).nbytes for n in table_names)) if rank == 0 else 0
        d2h = sum(v.numel() * v.element_size() for v in sh.h_out.values()) if rank == 0 else 0
        mean_phase = {k: float(np.mean([p[k] for p in phases])) for k in phases[0]}
        ph_t = torch.tensor([mean_phase["scatter_ms"], mean_phase["solve_ms"], mean_phase["gather_ms"]], dtype=torch.float64, device=dev)
        ph_all = [torch.zeros_like(ph_t) for _ in range(world)]
        dist.all_gather(ph_all, ph_t)
        sharded = {"total_problems": Bt, "scatter_bytes_from_rank0": sc_b, "gather_bytes_into_rank0": ga_b,
                   "per_rank_ms": [{"rank": r, "scatter": float(t[0]), "solve": float(t[1]), "gather": float(t[2])} for r, t in enumerate(ph_all)],
                   "note": "scatter = H2D on rank 0 + table broadcast + ncclScatter of x0 / sched_id; gather = ncclGather of "
                           "x, u, cost, iters, status into rank 0 + D2H (rank 0 waits for the slowest shard inside 'gather')"}
        e2e_api = "distributed.ShardedDdp.solve (rank 0 host buffers -> NCCL scatter -> ccc_ddp_*_solve(CCC_MEM_DEVICE) per rank -> NCCL gather -> rank 0 host)"

    # ---- reduce over ranks: slowest rank counts ----
    t_dev = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    per_rank = None
    if dist is not None:
        allr = [torch.zeros_like(t_dev) for _ in range(world)]
        dist.all_gather(allr, t_dev)
        it_t = torch.tensor([float(iters.mean()), float(iters.max())], dtype=torch.float64, device=dev)
        alli = [torch.zeros_like(it_t) for _ in range(world)]
        dist.all_gather(alli, it_t)
        per_rank = [{"rank": r, "kernel_ms_per_step": float(t[0]) / args.steps, "mean_ddp_iters": float(i[0]), "max_ddp_iters": int(i[1])}
                    for r, (t, i) in enumerate(zip(allr, alli))]
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_s_max = float(t_dev[0]), float(t_dev[1])
    value = world * B * args.steps / (total_ms_max / 1e3)
    e2e_value = world * B * e2e_steps / e2e_s_max

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        abytes = algorithmic_bytes_per_solve(ps, nx)
        kernel_ms = total_ms / args.steps
        achieved = abytes * B / (kernel_ms / 1e3) / 1e9
        cnt = load_counters(kind)
        traffic, util, fp64 = None, None, None
        if cnt:
            # counters were captured on `cnt["batch"]` problems with `cnt["mean_ddp_iters"]` iterations: scale to this launch
            scale = (B * float(iters.mean())) / (cnt["batch"] * cnt["mean_ddp_iters"])
            traffic = cnt["dram_bytes_per_launch"] * scale
            dfma = cnt["dfma_thread_inst_per_launch"] * scale
            ach_tf = 2.0 * dfma / (kernel_ms / 1e3) / 1e12
            fp64 = {"achieved_tflops": ach_tf, "peak_tflops": fp64_peak, "frac": ach_tf / fp64_peak if fp64_peak > 0 else None,
                    "dfma_thread_inst_per_solve": dfma / B, "peak_source": "ccc_fp64_peak_tflops: pure-DFMA kernel timed in this run",
                    "counter_source": cnt["source"]}
            util = {"issue_slots_pct": cnt.get("issue_active_pct"), "fp64_pipe_pct": cnt.get("fp64_pipe_active_pct"),
                    "source": cnt["source"]}
        line = {
            "metric": METRIC if kind == "centroidal" else "MPC solves/sec (DdpSingleRigidBody horizon=100)", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "per_gpu_batch": B, "cold_start": True, "max_iter": int(cfg.max_iter),
                       "l2": "256 MiB flush write between timed steps; per-step working set (gain lists, > 2 GB) > L2",
                       "mean_ddp_iters": float(iters.mean()), "max_ddp_iters": int(iters.max()),
                       "converged_frac": float((status == 1).mean()), "wall_s_timed_region": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": e2e_api},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": f"ccc_host::ddp_solve_kernel<ccc::{'CentroidalModel' if kind == 'centroidal' else 'SrbModel'},8,1,true>",
                         "algorithmic_bytes_per_solve": abytes, "kernel_ms_per_launch": kernel_ms, "fp64": fp64, "measured_utilisation": util,
                         "note": "latency/FP64-issue bound serial recursion: HBM fraction is small by nature, see DESIGN.md §5"},
        }
        if per_rank:
            line["per_rank"] = per_rank
        if sharded:
            line["e2e"]["sharded"] = sharded
        if not args.no_cpu_baseline and world == 1:
            from oracle import binding

            binding.build()
            threads = binding.hardware_threads()
            n_sample = min(B, max(threads * args.cpu_baseline_problems_per_thread // (1 if kind == "centroidal" else 12), 8))
            v, k, dt = cpu_oracle_rate(ps, cfg, n_sample, threads, kind)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"first {k} of {B} problems of the same seeded batch, {dt:.1f} s wall; scalar C++ port "
                                              "of the reference's algorithm (-O3 -mavx2 -mfma, no SIMD linear algebra)"}
            if not args.no_other and kind == "centroidal":
                eng.close()
                del d_out, flush
                torch.cuda.empty_cache()
                line["config"]["other"] = other_configs(threads)
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
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU per step (default 16384; 8192 for --workload config4)")
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS), help="config3 = DdpCentroidal (north star), config4 = DdpSingleRigidBody")
    ap.add_argument("--no-other", action="store_true", help="skip config.other (the other BASELINE configurations)")
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
