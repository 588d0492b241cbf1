#!/usr/bin/env python
"""Benchmark of the hot path: Solver::factor() + Solver::solve() of the supernodal sparse Cholesky.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload bal|...]

One "step" = factor() + solve(nRHS=1) of ONE matrix. At N=1 the workload is BASELINE.json configs[1]: the BAL-shaped
synthetic problem (871 cameras x 527480 points, block sizes 6/3, points eliminated by the sparse-elimination path,
fp64). At N>1 every rank owns one such matrix (a batch of N identically structured problems sharded one per GPU,
no data-path collective) -> weak scaling; value = algorithmic GF of all ranks / max-over-ranks device time.

Timing: CUDA events on the solver's stream around factor()+solve() of every step; the in-place factor is restored
from a pristine device copy between steps (untimed; the 0.57 GB copy also evicts L2). Inputs are larger than L2.
`--impl reference` times the CPU restatement of the reference's BLAS backend (oracle/, kind "port": the reference
itself cannot be compiled in this image) on the host cores, same config/metric.
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
# stdout carries exactly one JSON line: NCCL's version banner / debug lines go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.path.insert(0, ROOT)

METRIC = "factor()+solve() GF/s (algorithmic fp64 flops of the skeleton / time)"

WORKLOADS = {
    # name: (generator kind, params, (bsize_min, bsize_max), given elimination range or None, find_auto_ranges)
    "bal": dict(kind=3, params=[527480, 871, 2, 3.28, 40, 0.1], bsize=(3, 6), n_elim=527480, auto=True,
                desc="BAL-shaped synthetic 871 cams x 527480 points (6/3 blocks), sparse-elim {0,numPts}, fp64"),
    "bal_small": dict(kind=3, params=[60000, 200, 2, 3.28, 40, 0.1], bsize=(3, 6), n_elim=60000, auto=True,
                      desc="BAL-shaped synthetic 200 cams x 60000 points (6/3 blocks), sparse-elim, fp64"),
    "grid": dict(kind=1, params=[120, 120, 1.0, 2], bsize=(6, 6), n_elim=0, auto=False,
                 desc="GRID 120x120 block=6 conn=2 pure supernodal (no sparse elimination), fp64"),
    "flat": dict(kind=0, params=[1000, 0.05], bsize=(3, 3), n_elim=0, auto=True,
                 desc="FLAT size=1000 block=3 fill=0.05, fp64"),
    "flat_batch": dict(kind=0, params=[2000, 0.03], bsize=(3, 3), n_elim=0, auto=True, batch=64,
                       desc="batch=64 identical-structure FLAT size=2000 block=3 fill=0.03, batch sharded across the GPUs, fp64"),
    "stress": dict(kind=3, params=[1000000, 200, 2, 3.0, 200, 1.0], bsize=(3, 6), n_elim=1000000, auto=True,
                   desc="sparse-elim stress: 1M independent 3x3 points + 200 cameras, fp64"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout must carry exactly ONE line, the JSON record. Libraries write there behind Python's back (NCCL prints its version
# banner on fd 1 when the environment sets NCCL_DEBUG=VERSION), so fd 1 is pointed at stderr for the whole run and the
# record goes to the saved descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line). NVML in a thread
    (a query every 2 ms: the timed region of the headline workload is only tens of ms long, too short for a freshly
    spawned `nvidia-smi -lms`), falling back to an `nvidia-smi` subprocess when pynvml is unusable. The thread starts
    before the warm-up; `mark()` opens the timed window and `stop()` closes it - only samples inside count."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.thread, self.run, self.t0, self.source = gpu_index, [], None, None, False, None, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except Exception:
                pass
        return self.idx

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", None) or nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", None) or nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", None) or nv.nvmlClocksThrottleReasonSwPowerCap}
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while self.run:
                    try:
                        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        mask = int(reasons_fn(h))
                        self.rows.append((time.perf_counter(), sm, mx, [n for n, b in bits.items() if mask & b]))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.run, self.source = True, "nvml"
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.run = False
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self._physical_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                  [n for n, v in zip(self.NAMES, r[5:9]) if v.lower().startswith("active")]))
            except Exception:
                pass

    def mark(self):
        self.t0 = time.perf_counter()

    def stop(self):
        t1 = time.perf_counter()
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        if self.source == "nvidia-smi":
            time.sleep(0.05)
            self.proc.terminate()
        self.run = False
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [r for r in self.rows if t0 <= r[0] <= t1 + 0.03]
        window = "timed region"
        if not inside:  # region shorter than one sampling period: the closest samples around it
            inside, window = self.rows[-3:], "nearest samples (timed region shorter than the sampling period)"
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": max(r[2] for r in inside) if inside else None, "reasons": reasons, "samples": len(inside),
                "source": self.source, "window": window}


def gen_problem(api, wl):
    w = WORKLOADS[wl]
    sizes, ptrs, inds = api.gen_pattern_arrays(w["kind"], w["params"], w["bsize"][0], w["bsize"][1], 37)
    ranges = [0, w["n_elim"]] if w["n_elim"] else []
    return sizes, ptrs, inds, ranges, w


def algorithmic_work(solver_cls, api, backend_symbolic, wl, model):
    """Algorithmic flops of the workload (from the skeleton alone, the excess an implementation chooses to
    execute does not count): the smaller of the flop counts of (a) the skeleton the B200 arm factors (`model`) and (b) the
    skeleton the reference's own default builds (BackendFast -> model_OpenBlas_i7, Solver.cpp:679-683: the least
    supernode merging of the presets). On the BAL-shaped and FLAT problems the two agree to < 1 %; on GRID 120^2 the
    B200 preset merges far more (198.7 GF executed) than necessary (65.6 GF) - GF/s is quoted on the necessary work.
    `solver_cls` / `api` / `backend_symbolic`: the library that counts (product for the b200 arm, oracle for the CPU arm)."""
    sizes, ptrs, inds, ranges, w = gen_problem(api, wl)
    out = {}
    for name, m in (("executed", model), ("reference_default", 0)):
        s = solver_cls.create(sizes, ptrs, inds, ranges, backend=backend_symbolic, computation_model=m,
                              find_sparse_elim_ranges=w["auto"])
        out[name] = s.work_estimate()
        out[name]["lumps"] = s.num_lumps
        out[name]["order"] = s.order
    best = min(("executed", "reference_default"), key=lambda k: out[k]["factor_flops"] + out[k]["solve_flops_per_rhs"])
    out["algorithmic"] = out[best]
    return out


def make_config(w, work, batch):
    """the `config` object, identical (keys and values) in every arm of the same workload"""
    a = work["algorithmic"]
    return {"workload": w["desc"], "batch": batch, "n_rhs": 1, "order": a["order"],
            "algorithmic_gflop": round((a["factor_flops"] + a["solve_flops_per_rhs"]) / 1e9, 3)}


def dist_setup(n_gpus):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def pin_to_gpu_numa_node(local):
    """One process per GPU: run (and first-touch / pin host buffers) on the CPUs of the NUMA node the GPU hangs off, so
    that the end-to-end leg's pinned staging buffers and the H2D DMA stay on the GPU's side of the socket interconnect
    (8 ranks uploading 566 MB each per step otherwise share one socket's memory controllers). Best effort: silently
    does nothing when sysfs / NVML do not tell, or when the node's CPUs are not in this process's affinity mask."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis else local
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None


def run_reference(args):
    """the reference's CPU path (restated BLAS backend, all host threads) on the same workload; loads oracle/ only"""
    rank, world, _ = dist_setup(args.gpus)
    if rank != 0:
        return
    from oracle import cpu as ocpu
    from baspacho_b200 import _capi  # constants + ctypes binding class only: the product library is NOT loaded
    api = ocpu.api()
    cores = os.cpu_count()
    work = algorithmic_work(ocpu.OracleSolver, api, _capi.BACKEND_SYMBOLIC_ONLY, args.workload, args.model)
    sizes, ptrs, inds, ranges, w = gen_problem(api, args.workload)
    t0 = time.time()
    s = ocpu.OracleSolver.create(sizes, ptrs, inds, ranges, backend=_capi.BACKEND_FAST, num_threads=cores,
                                 find_sparse_elim_ranges=w["auto"])
    analysis_s = time.time() - t0
    data0 = api.random_data_array(s.data_size, -1, 1, 37)
    s.damp(data0, 0.0, s.order * 1.2)
    rhs0 = api.random_data_array(s.order, -1, 1, 38).reshape(1, s.order)
    a = work["algorithmic"]
    flops = a["factor_flops"] + a["solve_flops_per_rhs"]
    times = []
    for it in range(args.warmup + args.steps):
        data, x = data0.copy(), rhs0.copy()
        t0 = time.perf_counter()
        s.factor(data)
        s.solve(data, x)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = args.steps * flops / tot / 1e9
    batch = w.get("batch", 0) or world
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GF/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if w.get("batch") else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(w, work, batch),
        "detail": {"analysis_s": round(analysis_s, 3), "lumps": s.num_lumps, "one_matrix_per_step": True},
        "cpu_baseline": {"value": value, "unit": "GF/s", "cores": cores, "kind": "port",
                         "sample": "one matrix of the workload per step (1 factor+solve), restated reference BackendFast: OpenBLAS "
                                   f"{os.path.basename(ocpu.blas_path() or 'none')} + {cores} threads"},
        "e2e": {"value": value, "unit": "GF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def measure_dgemm_peak(torch, dev):
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    best = 0.0
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        if i >= 1:
            best = max(best, 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b, c
    torch.cuda.empty_cache()
    return best


def measure_workload(args, wl, steps, warmup, rank, world, dev, sampler=None, with_profile=False, with_unfused=False):
    """times `steps` steps of one workload on this rank's GPU (CUDA events on the solver's stream, max over ranks) and its
    end-to-end form through the host-buffer C-ABI calls; returns a dict (identical on every rank up to rank-local keys)"""
    import torch
    import baspacho_b200 as bsp
    dist = torch.distributed if world > 1 else None
    api = bsp.api()
    sizes, ptrs, inds, ranges, w = gen_problem(api, wl)
    t0 = time.time()
    s = bsp.Solver.create(sizes, ptrs, inds, ranges, computation_model=args.model, find_sparse_elim_ranges=w["auto"])
    analysis_s = time.time() - t0
    stream = torch.cuda.Stream(device=dev)
    s.set_stream(stream)
    work = algorithmic_work(bsp.Solver, api, bsp.BACKEND_SYMBOLIC_ONLY, wl, args.model) if rank == 0 else None
    if world > 1:
        box = [work]
        dist.broadcast_object_list(box, src=0)
        work = box[0]
    a = work["algorithmic"]
    flops = a["factor_flops"] + a["solve_flops_per_rhs"]
    batch_total = w.get("batch", 0)
    if batch_total:
        # config 4: a fixed batch of identically structured matrices, contiguous shard per rank (strong scaling)
        from baspacho_b200.sharding import shard_range
        lo, hi = shard_range(batch_total, rank, world)
        n_items = hi - lo
        data_h = np.stack([api.random_data_array(s.data_size, -1, 1, 37 + q) for q in range(lo, hi)]) if n_items else np.zeros((0, s.data_size))
        for q in range(n_items):
            s.damp(data_h[q], 0.0, s.order * 1.3)
        rhs_h = np.stack([api.random_data_array(s.order, -1, 1, 1038 + q).reshape(1, s.order) for q in range(lo, hi)]) if n_items else np.zeros((0, 1, s.order))
    else:
        # every rank owns one matrix of the batch: same structure, different values (weak scaling)
        n_items = 1
        data_h = api.random_data_array(s.data_size, -1, 1, 37 + rank)
        s.damp(data_h, 0.0, s.order * 1.2)
        rhs_h = api.random_data_array(s.order, -1, 1, 38 + rank).reshape(1, s.order)
    pin_data = torch.from_numpy(data_h).pin_memory()
    pin_rhs = torch.from_numpy(rhs_h.copy()).pin_memory()
    pristine = pin_data.to(dev)
    work_d = torch.empty_like(pristine)
    rhs_d = pin_rhs.to(dev)
    x_d = torch.empty_like(rhs_d)

    def do_factor():
        if batch_total:
            if n_items:
                s.factor_batched(work_d)
        else:
            s.factor(work_d)

    def do_solve():
        if batch_total:
            if n_items:
                s.solve_batched(work_d, x_d)
        else:
            s.solve(work_d, x_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.cuda.stream(stream):
            work_d.copy_(pristine, non_blocking=True)
            x_d.copy_(rhs_d, non_blocking=True)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            do_factor()
            e1.record(stream)
            do_solve()
            e2.record(stream)
        return e0, e1, e2

    for _ in range(warmup):
        step()
    barrier()
    if sampler is not None:
        sampler.mark()
    n0 = s.launch_count()
    t_wall = time.perf_counter()
    evs = [step() for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = s.launch_count() - n0
    clocks = sampler.stop() if sampler is not None else None
    fac_ms = [a_.elapsed_time(b_) for a_, b_, _ in evs]
    sol_ms = [b_.elapsed_time(c_) for _, b_, c_ in evs]
    tot_s = (sum(fac_ms) + sum(sol_ms)) * 1e-3
    t = torch.tensor([tot_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_max = float(t.item())

    # residual check of the last step (result correctness inside the bench, cheap): ||A x - b|| through addMvFrom
    resid = 0.0
    if n_items:
        p0, x0, r0 = (pristine[0], x_d[0], rhs_d[0]) if batch_total else (pristine, x_d, rhs_d)
        y = torch.zeros_like(r0)
        with torch.cuda.stream(stream):
            s.add_mv_from(p0, 0, x0, y)
        torch.cuda.synchronize()
        resid = float((y - r0).norm() / r0.norm())
    x_head = (x_d[0, 0, :4] if batch_total else x_d[0, :4]).cpu().numpy().tolist() if n_items else None
    x_gpu = x_d.cpu().numpy() if (rank == 0 and not batch_total) else None

    # ---- the fine-grained op sequence (what INTEGRATION.md's one-line getBackend change delivers without the fused hooks)
    unfused_ms = None
    if with_unfused:
        s.set_fused(False)
        ts = []
        for it in range(4):
            e0, _, e2 = step()
            torch.cuda.synchronize()
            if it >= 1:
                ts.append(e0.elapsed_time(e2))
        s.set_fused(True)
        unfused_ms = float(np.mean(ts))

    # ---- e2e: the C-ABI host-buffer call (pinned host memory in, solution out), H2D + D2H inside the timed region
    x_host = torch.empty_like(pin_rhs).pin_memory()
    e2e_times = []
    h2d = d2h = 0
    for it in range(2 + max(3, steps // 2)):
        x_host.copy_(pin_rhs)
        barrier()
        t0 = time.perf_counter()
        if batch_total:
            if n_items:
                s.factor_solve_host_batched(pin_data, x_host)
        else:
            s.factor_solve_host(pin_data, x_host, None)
        dt = time.perf_counter() - t0
        if it >= 2:
            e2e_times.append(dt)
    if n_items:
        h2d, d2h = s.host_copy_bytes()
    e2e_resid = None
    if n_items and not batch_total:  # the end-to-end call's own answer: residual of its solution
        xe = x_host.to(dev)
        y = torch.zeros_like(rhs_d)
        with torch.cuda.stream(stream):
            s.add_mv_from(pristine, 0, xe, y)
        torch.cuda.synchronize()
        e2e_resid = float((y - rhs_d).norm() / rhs_d.norm())
    te = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=dev)
    tb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    e2e_s = float(te.item())

    prof = None
    if with_profile and rank == 0:
        # per-kernel-class profile of one more step (events around every launch of our kernels)
        api.profile(True)
        step()
        torch.cuda.synchronize()
        prof = api.profile_json()
        api.profile(False)

    total_flops = flops * batch_total if batch_total else world * flops
    ex = work["executed"]
    res = {
        "w": w, "work": work, "solver": s, "problem": (sizes, ptrs, inds, ranges), "rhs_h": rhs_h, "x_gpu": x_gpu,
        "batch": batch_total or world, "n_items": n_items, "steps": steps, "warmup": warmup,
        "value": steps * total_flops / tot_max / 1e9, "ms_per_step": tot_max / steps * 1e3,
        "factor_ms": float(np.mean(fac_ms)), "solve_ms": float(np.mean(sol_ms)),
        "factor_gfs": a["factor_flops"] * (n_items if batch_total else 1) / (np.mean(fac_ms) * 1e-3) / 1e9,
        "residual": resid, "wall_s_timed_region": t_wall, "x_head": x_head, "launches": launches, "clocks": clocks,
        "unfused_ms_per_step": unfused_ms, "kernel_classes": prof, "analysis_s": round(analysis_s, 3),
        "e2e": {"value": total_flops / e2e_s / 1e9, "unit": "GF/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": int(tb[1].item()),
                "bytes_note": "summed over ranks; counted by the library from the copies it issues (upper triangles of wide "
                              "diagonal blocks are not uploaded)",
                "residual": e2e_resid,
                "api": "bspb200_factor_solve_host_batched" if batch_total else "bspb200_factor_solve_host (C ABI, pinned host "
                       "buffers; point columns uploaded in chunks, the elimination of a chunk overlaps the next upload)"},
        "detail": {"order": s.order, "data_size": s.data_size, "lumps": s.num_lumps, "items_on_rank0": n_items,
                   "executed_gflop": round((ex["factor_flops"] + ex["solve_flops_per_rhs"]) / 1e9, 3),
                   "algorithmic_factor_gflop": round(a["factor_flops"] / 1e9, 3), "nnz_l": ex["nnz_l"],
                   "dense_lump_sizes": np.diff(s.lumpStart[w["n_elim"]:]).tolist()[-8:] if w["n_elim"] else None,
                   "l2_policy": "inputs larger than L2; the in-place factor is restored from a pristine device copy between steps",
                   "timing": "cuda events per step on the solver stream; sum over steps; max over ranks",
                   "analysis_s": round(analysis_s, 3)},
    }
    return res


def ref_cuda_leg(args, wl, steps=3, warmup=2):
    """the reference's CUDA algorithm restated on cuSOLVER / cuBLAS (oracle/RefCudaOps.cu) on the same GPU, same inputs: on
    the B200 arm's skeleton (same supernodes) and on its own preset's (model_Cuda117_2080Ti, what the reference picks)"""
    import torch
    from oracle import refcuda
    api = refcuda.api()
    sizes, ptrs, inds, ranges, w = gen_problem(api, wl)
    out = {}
    for name, model in (("same_skeleton", args.model), ("own_preset_2080ti", 1)):
        s = refcuda.RefCudaSolver.create(sizes, ptrs, inds, ranges, computation_model=model, find_sparse_elim_ranges=w["auto"])
        stream = torch.cuda.Stream()
        s.set_stream(stream)
        data_h = api.random_data_array(s.data_size, -1, 1, 37)
        s.damp(data_h, 0.0, s.order * 1.2)
        pristine = torch.from_numpy(data_h).cuda()
        work_d = torch.empty_like(pristine)
        rhs_d = torch.from_numpy(api.random_data_array(s.order, -1, 1, 38).reshape(1, s.order)).cuda()
        x_d = torch.empty_like(rhs_d)
        fac, sol = [], []
        for it in range(warmup + steps):
            with torch.cuda.stream(stream):
                work_d.copy_(pristine, non_blocking=True)
                x_d.copy_(rhs_d, non_blocking=True)
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record(stream)
                s.factor(work_d)
                e1.record(stream)
                s.solve(work_d, x_d)
                e2.record(stream)
            torch.cuda.synchronize()
            if it >= warmup:
                fac.append(e0.elapsed_time(e1)), sol.append(e1.elapsed_time(e2))
        out[name] = {"ms_per_step": float(np.mean(fac) + np.mean(sol)), "factor_ms": float(np.mean(fac)),
                     "solve_ms": float(np.mean(sol)), "lumps": s.num_lumps, "x_head": x_d[0, :4].cpu().numpy().tolist()}
        del s, pristine, work_d
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    rank, world, local = dist_setup(args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl b200 (no CPU fallback)"
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    log(f"rank {rank}: numa pinning {numa}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()  # before the warm-up, so that it is already sampling when the timed region opens
    m = measure_workload(args, args.workload, args.steps, args.warmup, rank, world, dev, sampler=sampler,
                         with_profile=True, with_unfused=(rank == 0 and world == 1))
    # BASELINE config 4 (batch = 64 identical-structure FLAT-2000, sharded over the GPUs, strong scaling) rides along
    # with the headline workload so that every N of the driver's scaling run records it
    c4 = None
    if args.workload == "bal" and not args.no_config4:
        c4 = measure_workload(args, "flat_batch", min(args.steps, 5), 3, rank, world, dev)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    w, work, s = m["w"], m["work"], m["solver"]
    a = work["algorithmic"]
    prof = m["kernel_classes"]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    dgemm_peak = measure_dgemm_peak(torch, dev)
    # dominant kernel: the tensor-bound class with the largest share of the step
    dom = max(("lump_chol", "gemm"), key=lambda k: prof[k]["ms"])
    g = prof[dom]
    dom_tf = g["flops"] / max(g["ms"], 1e-9) / 1e9
    names = {"gemm": "gemm_nt_f64_kernel (DMMA m8n8k4 SYRK/GEMM tiles of the blocked supernode Cholesky)",
             "lump_chol": "lump_chol_kernel (persistent left-looking tile Cholesky of a wide supernode: TMA -> smem -> DMMA "
                          "m8n8k4 tiles, device-side dependency flags; algorithmic flops n^3/3 + rows n^2 of the lump column)"}
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if args.workload == "bal" and os.path.exists(tpath):
        tj = json.load(open(tpath)).get({"gemm": "gemm_nt_f64_kernel", "lump_chol": "lump_chol_kernel"}[dom])
        if tj:
            traffic, traffic_note = tj["dram_bytes_per_launch"], {k: tj[k] for k in ("source", "launches", "largest_launch") if k in tj}
    # floor of the whole step: dense flops at the DGEMM peak + elimination and solve bytes at the HBM peak
    ex = work["executed"]
    elim_bytes = a["elim_bytes"]
    solve_bytes = 8.0 * (2.0 * s.data_size + 4.0 * s.order)
    dense_flops = a["factor_flops"] - a["elim_flops"]
    floor_ms = (dense_flops / (dgemm_peak * 1e12) + (elim_bytes + solve_bytes) / (hbm_peak * 1e9)) * 1e3 if dgemm_peak else None
    roofline = {"kernel": names[dom], "bound": "tensor", "achieved": dom_tf, "peak": dgemm_peak, "unit": "TFLOP/s",
                "frac": dom_tf / dgemm_peak if dgemm_peak else None, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json holds no fp64 figure; "
                               "tcgen05 has no f64 kind, DMMA is the fp64 tensor path)",
                "launches_per_step": g["launches"], "share_of_step_ms": g["ms"],
                "frac_step": floor_ms / m["ms_per_step"] if floor_ms else None,
                "floor_ms_step": floor_ms,
                "floor_note": "dense flops / DGEMM peak + (elimination bytes + solve bytes of SURVEY 8d) / HBM peak"}
    eg, ef = prof["elim_gather"], prof["elim_factor"]
    roofline_hbm = None
    if ef["launches"]:
        ms = ef["ms"] + eg["ms"]
        roofline_hbm = {"kernels": "elim_factor_lumps + elim_gather (sparse elimination)", "bound": "hbm",
                        "achieved": elim_bytes / max(ms, 1e-9) / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": elim_bytes / max(ms, 1e-9) / 1e6 / hbm_peak,
                        "bytes": elim_bytes, "bytes_note": "SURVEY 8d: read+write every eliminated column once, read-modify-write each target entry once",
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback",
                        "ms": ms}

    # ---- CPU baseline (oracle port of the reference's BLAS backend) on a bounded sample: the full workload once
    cpu_baseline = None
    if not args.no_cpu_baseline:
        try:
            from baspacho_b200 import _capi
            from oracle import cpu as ocpu
            sizes, ptrs, inds, ranges = m["problem"]
            cores = os.cpu_count()
            o = ocpu.OracleSolver.create(sizes, ptrs, inds, ranges, backend=_capi.BACKEND_FAST, num_threads=cores,
                                         find_sparse_elim_ranges=w["auto"])
            d = ocpu.api().random_data_array(o.data_size, -1, 1, 37)
            o.damp(d, 0.0, o.order * 1.2)
            xr = (m["rhs_h"][0] if w.get("batch") else m["rhs_h"]).copy()
            t0 = time.perf_counter()
            o.factor(d)
            o.solve(d, xr)
            dt = time.perf_counter() - t0
            flops = a["factor_flops"] + a["solve_flops_per_rhs"]
            cpu_baseline = {"value": flops / dt / 1e9, "unit": "GF/s", "cores": cores, "kind": "port",
                            "sample": f"full workload, 1 factor+solve ({dt:.2f} s), restated reference BackendFast (OpenBLAS + threads)",
                            # comparable only when both arms built the same skeleton (the CPU arm picks its own
                            # supernode-merge model, as the reference does: Solver.cpp:679-683)
                            "solution_max_abs_diff_vs_gpu": float(np.abs(xr - m["x_gpu"]).max())
                            if m["x_gpu"] is not None and np.array_equal(o.lumpStart, s.lumpStart)
                            and np.array_equal(o.array("permutation"), s.array("permutation")) else None}
        except Exception as e:  # the baseline must never take the GPU number down
            cpu_baseline = {"value": None, "unit": "GF/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    ref_cuda = None
    if not args.no_ref_cuda and m["n_items"] == 1 and not w.get("batch"):
        try:
            ref_cuda = ref_cuda_leg(args, args.workload)
        except Exception as e:  # noqa: BLE001
            ref_cuda = {"failed": repr(e)}

    world_desc = f"; batch of {world} such matrices sharded 1 per GPU" if world > 1 and not w.get("batch") else ""
    line = {
        "metric": METRIC, "value": m["value"], "unit": "GF/s", "n_gpus": world, "steps": m["steps"], "warmup": m["warmup"],
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong" if w.get("batch") else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(w, work, m["batch"]), "detail": dict(m["detail"], note=world_desc.strip("; ")),
        "factor_ms": m["factor_ms"], "solve_ms": m["solve_ms"], "factor_gfs": m["factor_gfs"],
        "residual": m["residual"], "wall_s_timed_region": m["wall_s_timed_region"], "x_head": m["x_head"],
        "gpu_launches": m["launches"], "e2e": m["e2e"], "unfused_ms_per_step": m["unfused_ms_per_step"],
        "roofline": roofline, "roofline_hbm": roofline_hbm, "kernel_classes": prof,
        "cpu_baseline": cpu_baseline, "ref_cuda": ref_cuda, "clocks": m["clocks"],
    }
    if c4 is not None:
        line["config4"] = {"config": make_config(c4["w"], c4["work"], c4["batch"]), "scaling": "strong",
                           "value": c4["value"], "unit": "GF/s", "ms_per_step": c4["ms_per_step"],
                           "factor_ms": c4["factor_ms"], "solve_ms": c4["solve_ms"], "items_on_rank0": c4["n_items"],
                           "steps": c4["steps"], "warmup": c4["warmup"], "residual": c4["residual"],
                           "gpu_launches": c4["launches"], "e2e": c4["e2e"], "detail": c4["detail"]}
    emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_ref_cuda(args):
    """second GPU baseline as its own arm: see ref_cuda_leg"""
    import torch
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return
    assert torch.cuda.is_available(), "--impl ref_cuda needs a GPU"
    torch.cuda.set_device(local)
    from oracle import refcuda
    from baspacho_b200 import _capi
    work = algorithmic_work(refcuda.RefCudaSolver, refcuda.api(), _capi.BACKEND_SYMBOLIC_ONLY, args.workload, args.model)
    w = WORKLOADS[args.workload]
    legs = ref_cuda_leg(args, args.workload, steps=args.steps, warmup=args.warmup)
    a = work["algorithmic"]
    flops = a["factor_flops"] + a["solve_flops_per_rhs"]
    best = min(legs.values(), key=lambda r: r["ms_per_step"])
    line = {"impl": "ref_cuda", "metric": METRIC, "value": flops / (best["ms_per_step"] * 1e-3) / 1e9, "unit": "GF/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(w, work, 1),
            "detail": {"backend": "restated reference MatOpsCuda.cu: cusolverDnDpotrf + cublasDtrsm/Dgemm per lump, "
                                  "thread-per-pair elimination with fp64 atomics, per-lump synchronous span-table copy",
                       "legs": legs}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "ref_cuda"])
    ap.add_argument("--workload", default="bal", choices=sorted(WORKLOADS))
    ap.add_argument("--model", type=int, default=2, help="supernode-merge cost model preset (2 = B200)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "ref_cuda":
        run_ref_cuda(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
