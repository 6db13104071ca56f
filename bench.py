#!/usr/bin/env python
"""Benchmark of the tornadox filter step on B200 (contract: see the task statement / DESIGN.md section Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mine|reference] [--solver dek1|ek0] [--grid 2048]

metric: filter state-dims * steps / second ("step" = one attempt_step) on BASELINE.json config 5:
ivp.fhn_2d on a 2048 x 2048 grid (d = 8 388 608), nu = 4, fp64.  The headline ``value`` is
DiagonalEK1; the KroneckerEK0 numbers of the same run are reported under ``also``.
One JSON line on stdout (rank 0).  For N > 1 launch with torchrun; d is sharded by grid rows (strong scaling).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NU = 4
ALG_BYTES = {"dek1": lambda n, w: 2 * (n + n * n) * w, "ek0": lambda n, w: 2 * n * w}   # per state dimension


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mine", choices=["mine", "reference"])
    ap.add_argument("--solver", default="dek1", choices=["dek1", "ek0"])
    ap.add_argument("--grid", type=int, default=2048, help="fhn_2d grid side (BASELINE config 5: 2048)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU reference arm: the NumPy oracle (a port of the reference; JAX is not installable here) on all host cores
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    solver, side, attempts, warmup, seed = job
    from oracle.ref_np import filters, problems

    n = NU + 1
    rng = np.random.default_rng(seed)
    d = 2 * side * side
    prob = problems.fhn_2d(dx=1.0 / side, y0=rng.uniform(size=d))
    mean = rng.standard_normal((n, d))
    mean[0] = prob.y0
    dt = 1e-7
    if solver == "dek1":
        state = filters.FilterState(0.0, mean, 1e-3 * rng.standard_normal((d, n, n)), None, None)
        step = lambda s: filters.diagonal_ek1_attempt(s, dt, prob.f, prob.df_diagonal, NU)[0]
    else:
        state = filters.FilterState(0.0, mean, 1e-3 * rng.standard_normal((n, n)), None, None)
        step = lambda s: filters.kronecker_ek0_attempt(s, dt, prob.f, NU)[0]
    for _ in range(max(1, warmup)):   # warm caches / page in
        state = step(state)
    t0 = time.perf_counter()
    for _ in range(attempts):
        state = step(state)
    return time.perf_counter() - t0, d * attempts


def cpu_reference(solver, steps, warmup):
    """All host cores, one independent fhn_2d sub-grid per worker process, ``warmup`` untimed and ``steps`` timed attempts
    each (a DiagonalEK1 attempt on the 384x384 sub-grid takes ~0.6 s per core); returns throughput + description."""
    import multiprocessing as mp

    cores = min(os.cpu_count() or 1, 16)   # bounded: every worker holds ~1 GB of NumPy temporaries
    side = {"dek1": 384, "ek0": 1024}[solver]
    attempts = max(1, steps)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    jobs = [(solver, side, attempts, warmup, 1000 + i) for i in range(cores)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    units = sum(r[1] for r in res)
    value = units / slowest
    sample = (f"{cores} worker processes x fhn_2d {side}x{side} (d={2 * side * side} each) x {attempts} "
              f"{'DiagonalEK1' if solver == 'dek1' else 'KroneckerEK0'} attempts of the NumPy oracle "
              f"(LAPACK dgeqrf via numpy.linalg.qr); wall {wall:.1f}s incl. process start")
    return value, cores, sample, slowest / attempts * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 200))          # a step = one attempt on every worker's sub-grid (bounded sample, ~0.6 s)
    warmup = max(1, min(args.warmup, 20))
    value, cores, sample, ms = cpu_reference(args.solver, steps, warmup)
    n, w = NU + 1, 8
    line = {
        "impl": "reference",
        "metric": "filter state-dims*steps/sec (fhn_2d, nu=4, fp64)",
        "value": value, "unit": "state-dims*steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"fhn_2d {args.grid}x{args.grid} (d={2 * args.grid ** 2}), "
                               f"{'DiagonalEK1' if args.solver == 'dek1' else 'KroneckerEK0'} nu=4 attempt_step; "
                               "CPU arm timed on a bounded sample of it", "solver": args.solver},
        "cpu_baseline": {"value": value, "unit": "state-dims*steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "state-dims*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is pure Python on JAX, which cannot be installed here (no network, no wheel); this arm is "
                "the op-for-op NumPy restatement under oracle/ (LAPACK geqrf like JAX-CPU) on all host cores",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_mine(args):
    import torch

    from tornadox_b200 import ek0, ek1, ivp, step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; tornadox_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    group = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    dtype = {"f64": torch.float64, "f32": torch.float32}[args.dtype]
    w = 8 if dtype == torch.float64 else 4
    n = NU + 1
    side = args.grid
    d = 2 * side * side
    y0 = np.random.default_rng(2).uniform(size=d)
    problem = ivp.fhn_2d(dx=1.0 / side, y0=y0)

    def make(solver_name):
        cls = ek1.DiagonalEK1 if solver_name == "dek1" else ek0.KroneckerEK0
        sol = cls(num_derivatives=NU, steprule=step.AdaptiveSteps(), dtype=dtype, process_group=group)
        state = sol.initialize(*problem)
        return sol, state

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_steps(sol, state, dt, k, warm):
        """W untimed + K timed attempt_step calls through the public API, state resident in HBM."""
        for _ in range(warm):
            state, _ = sol.attempt_step(state, dt, *problem)
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = sol.engine.launch_count()
        start.record()
        h0 = time.perf_counter()
        for _ in range(k):
            state, _ = sol.attempt_step(state, dt, *problem)
        host_ms = (time.perf_counter() - h0) * 1e3 / k          # host time to enqueue one attempt (Python + C ABI)
        stop.record()
        barrier()
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, state, sol.engine.launch_count() - launches0, host_ms

    def kernel_time(sol, state, dt, k):
        """Average duration of the dominant kernel: CUDA events on the launch stream around k back-to-back launches into
        preallocated outputs (a single launch after a synchronize would also time the host's enqueue latency)."""
        eng = sol.engine
        is_dek1 = isinstance(sol, ek1.DiagonalEK1)
        mo = torch.empty_like(state.y.mean)
        so = torch.empty_like(state.y.cov_sqrtm) if is_dek1 else None

        def launch():
            if is_dek1:
                eng.dek1_attempt(state.t, dt, state.y.mean, state.y.cov_sqrtm, 1e-4, 1e-2, mean_out=mo, sc_out=so,
                                 want_err_ref=sol.materialize_error_vectors)
            else:
                eng.ek0_attempt(state.t, dt, state.y.mean, state.y.cov_sqrtm_2, 1e-4, 1e-2, mean_out=mo,
                                want_ref=sol.materialize_reference_state)

        for _ in range(2):
            launch()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            launch()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / k

    results = {}
    sampler = ClockSampler(local_rank)
    peak, peak_src = measured_peak()
    order = [args.solver] + [s for s in ("dek1", "ek0") if s != args.solver]
    for name in order:
        sol, state = make(name)
        # the step the adaptive rule itself would start with (step.py:111-115); a fixed larger dt makes the
        # explicit (off-diagonal) part of the stencil unstable within a few dozen unchecked attempts
        dt = 0.5 * sol.steprule.first_dt(*problem)
        if name == args.solver and rank == 0:
            sampler.start()
        ms, state, launches, host_ms = timed_steps(sol, state, dt, args.steps, args.warmup)
        clocks = sampler.stop() if (name == args.solver and rank == 0) else None
        kms = kernel_time(sol, state, dt, min(args.steps, 10))
        alg = ALG_BYTES[name](n, w) * (d / world)        # per launch, per GPU
        results[name] = {
            "value": d * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps, "launches": launches, "clocks": clocks,
            "host_enqueue_ms_per_step": host_ms,
            "roofline": {"bound": "hbm", "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (kms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "kernel_ms": kms, "algorithmic_bytes_per_launch": alg},
        }
        if name == args.solver and not args.no_e2e:
            results[name]["e2e_resident"] = e2e_resident_solve(torch, sol, state, dt, problem, world, dev, d, args.steps)
            results[name]["e2e_device_loop"] = e2e_device_loop(torch, sol, state, dt, problem, d, args.steps)
            torch.cuda.empty_cache()
            results[name]["e2e"] = e2e_host_buffers(torch, sol, state, dt, problem, world, dev, d, max(2, min(args.steps, 5)))
        del sol, state
        torch.cuda.empty_cache()

    if world == 1 and side == 2048 and args.dtype == "f64":
        for name in results:
            results[name]["roofline"]["traffic"] = _recorded_traffic(name)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference(args.solver, 16, 1)      # ~10 s of work on every host core
        cpu = {"value": v, "unit": "state-dims*steps/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        main = results[args.solver]
        other = [s for s in results if s != args.solver][0]
        line = {
            "metric": "filter state-dims*steps/sec (fhn_2d, nu=4, fp64)" if args.dtype == "f64"
            else "filter state-dims*steps/sec (fhn_2d, nu=4, fp32)",
            "value": main["value"], "unit": "state-dims*steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"ivp.fhn_2d {side}x{side} grid (d={d}), "
                            f"{'ek1.DiagonalEK1' if args.solver == 'dek1' else 'ek0.KroneckerEK0'}(num_derivatives=4) "
                            "attempt_step, y0 ~ U[0,1) seed 2, TaylorMode init, AdaptiveSteps tolerances fused",
                "solver": args.solver,
                "sharding": f"grid rows over {world} GPU(s); halo rows and scalar all-reduce through NVLink peer memory "
                            "inside the step kernel (torch.distributed/NCCL only for rendezvous)",
                "l2": "state (mean+factors) is 2.0 GB per attempt for DiagonalEK1 / 0.34 GB for KroneckerEK0, larger than the 126 MB L2; no flush needed",
            },
            "clocks": main["clocks"],
            "e2e": main.get("e2e"),
            "e2e_resident": main.get("e2e_resident"),
            "e2e_device_loop": main.get("e2e_device_loop"),
            "gpu_launches": main["launches"],
            "roofline": main["roofline"],
            "cpu_baseline": cpu,
            "host_enqueue_ms_per_step": main["host_enqueue_ms_per_step"],
            "also": {other: {k: results[other][k] for k in ("value", "ms_per_step", "roofline", "launches", "host_enqueue_ms_per_step")}},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def _recorded_traffic(solver):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(solver)
    return None


def e2e_host_buffers(torch, sol, state, dt, problem, world, dev, d, k):
    """Same metric through the public attempt_step with the state in HOST (pinned) memory: every step moves the whole
    input state host->device and the whole output state (+ the error norm) device->host.

    DiagonalEK1: the library's host-buffer entry point (tdx_dek1_attempt_step_host[_sharded]) streams the (shard of the)
    state through the device in chunks with the copies in both directions overlapping the kernels; a sharded run first
    exchanges the halos of the shards' edge rows.  KroneckerEK0: plain copy-in / step / copy-out around the device-resident
    call."""
    from tornadox_b200 import ek1, odefilter, rv

    is_dek1 = isinstance(sol, ek1.DiagonalEK1)
    y = state.y
    fac = y.cov_sqrtm if is_dek1 else y.cov_sqrtm_2
    h_mean_in = torch.empty(y.mean.shape, dtype=y.mean.dtype, pin_memory=True).copy_(y.mean)
    h_fac_in = torch.empty(fac.shape, dtype=fac.dtype, pin_memory=True).copy_(fac)
    h2d = h_mean_in.numel() * h_mean_in.element_size() + h_fac_in.numel() * h_fac_in.element_size()
    pipelined = is_dek1

    if pipelined:
        host_state = odefilter.ODEFilterState(state.t, rv.BatchedMultivariateNormal(h_mean_in, h_fac_in), None, None)
        vec_bytes = 2 * y.mean.shape[1] * y.mean.element_size() if sol.materialize_error_vectors else 0
        if world > 1:   # + the shard's edge rows for the halo exchange that precedes the pipeline
            h2d += 4 * problem.spec.grid_cols * y.mean.shape[0] * y.mean.element_size()

        def one():
            new, _ = sol.attempt_step(host_state, dt, *problem)      # synchronous: outputs are complete in host memory
            return new

        what = ("DiagonalEK1.attempt_step on a state in pinned host memory -> tdx_dek1_attempt_step_host: chunked "
                "H2D | kernel | D2H pipeline (whole state in, whole state + error vectors + norm out, every step)")
        d2h = h2d + vec_bytes + 8
    else:
        h_mean_out = torch.empty_like(h_mean_in, pin_memory=True)
        h_fac_out = torch.empty_like(h_fac_in, pin_memory=True)
        h_norm = torch.empty(1, dtype=torch.float64, pin_memory=True)
        d_mean, d_fac = torch.empty_like(y.mean), torch.empty_like(fac)

        def one():
            d_mean.copy_(h_mean_in, non_blocking=True)
            d_fac.copy_(h_fac_in, non_blocking=True)
            if is_dek1:
                st = odefilter.ODEFilterState(state.t, rv.BatchedMultivariateNormal(d_mean, d_fac), None, None)
            else:
                st = odefilter.ODEFilterState(state.t, rv.LeftIsotropicMatrixNormal(d_mean, y.d, d_fac), None, None)
            new, _ = sol.attempt_step(st, dt, *problem)
            h_mean_out.copy_(new.y.mean, non_blocking=True)
            h_fac_out.copy_(new.y.cov_sqrtm if is_dek1 else new.y.cov_sqrtm_2, non_blocking=True)
            h_norm.copy_(new.scaled_error_norm_sq_sum.sum().reshape(1), non_blocking=True)

        what = "attempt_step with the whole state in pinned host memory: H2D of mean+factors, kernel, D2H of mean+factors+norm, every step"
        d2h = h2d + 8

    one()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        one()
    b.record()
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(b), (time.perf_counter() - t0) * 1e3)   # the pipelined call blocks the host: wall clock counts
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": d * k / (ms * 1e-3), "unit": "state-dims*steps/s", "h2d_bytes_per_step": h2d * world,
            "d2h_bytes_per_step": d2h * world, "ms_per_step": ms / k, "what": what}


def e2e_resident_solve(torch, sol, state, dt, problem, world, dev, d, k):
    """The solver's own loop (perform_full_step): state stays in HBM, the host reads the error norm (8 B D2H)
    and decides accept/reject + next dt every attempt."""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    attempts = 0
    cur, cur_dt = state, dt
    for _ in range(k):
        cur, cur_dt, info = sol.perform_full_step(cur, cur_dt, problem.f, problem.t0, 1e9, problem.y0, None, problem.df_diagonal)
        attempts += info["num_attempted_steps"]
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    return {"value": d * attempts / sec, "unit": "state-dims*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
            "attempts": attempts, "accepted": k, "ms_per_attempt": sec / attempts * 1e3,
            "what": "public perform_full_step loop (adaptive accept/reject on the host), wall clock incl. the per-attempt sync"}


def e2e_device_loop(torch, sol, state, dt, problem, d, k):
    """k attempts of the adaptive solve under the device-resident step controller: dt, accept / reject and the buffer flip
    are decided by the kernels; the host enqueues the batch and reads the controller block once at the end."""
    from tornadox_b200 import ek1

    eng = sol.engine
    rule = eng.native_steprule(sol.steprule)
    is_dek1 = isinstance(sol, ek1.DiagonalEK1)
    pair_a = (state.y.mean, state.y.cov_sqrtm if is_dek1 else state.y.cov_sqrtm_2)
    pair_b = (torch.empty_like(pair_a[0]), torch.empty_like(pair_a[1]))
    err = torch.empty((), dtype=pair_a[0].dtype, device=pair_a[0].device)

    def batch():
        eng.ctl_begin(rule, state.t, dt, 1e300)          # no end time: every launch of the batch is a real attempt
        if is_dek1:
            eng.dek1_enqueue_attempts(k, pair_a, pair_b, None, None)
        else:
            eng.ek0_enqueue_attempts(k, pair_a, pair_b, err, None)
        return eng.ctl_read()

    batch()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctl = batch()
    sec = time.perf_counter() - t0
    attempts = max(1, int(ctl.attempts))
    return {"value": d * attempts / sec, "unit": "state-dims*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
            "attempts": attempts, "accepted": int(ctl.accepted), "ms_per_attempt": sec / attempts * 1e3,
            "what": "adaptive attempts under the device-resident step controller (tdx_*_enqueue_attempts): wall clock per "
                    "attempt incl. enqueueing the batch and one controller read; error vectors not materialised"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mine(args)


if __name__ == "__main__":
    main()
