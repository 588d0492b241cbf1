"""fp32 line for BASELINE config 2 and the mixed-precision variant: factor in fp32 (3xTF32 GEMMs,
SIMT sparse elimination), then iterative refinement of the solution with fp64 residuals (r = b - A x through addMvFrom on
the fp64 data, correction solved in fp32) until the fp64 residual stops improving. Prints one JSON line:
  python tools/mixed_precision.py [workload]      (B200 box; workloads of bench.py)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import baspacho_b200 as bsp  # noqa: E402
from bench import gen_problem  # noqa: E402


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "bal"
    api = bsp.api()
    sizes, ptrs, inds, ranges, w = gen_problem(api, wl)
    s = bsp.Solver.create(sizes, ptrs, inds, ranges, computation_model=2, find_sparse_elim_ranges=w["auto"])
    data0 = api.random_data_array(s.data_size, -1, 1, 37)
    s.damp(data0, 0.0, s.order * 1.2)
    rhs0 = api.random_data_array(s.order, -1, 1, 38)
    A64 = torch.from_numpy(data0).cuda()
    b64 = torch.from_numpy(rhs0).cuda()
    A32p = A64.float()
    out = {"workload": w["desc"], "order": int(s.order)}

    # fp64 reference: factor + solve
    f64, x64 = A64.clone(), b64.clone()

    def step64():
        f64.copy_(A64)
        x64.copy_(b64)
        s.factor(f64)
        s.solve(f64, x64)
    out["fp64_ms"] = timed(step64)
    y = torch.zeros_like(b64)
    s.add_mv_from(A64, 0, x64, y)
    out["fp64_residual"] = float((y - b64).norm() / b64.norm())

    # fp32: factor + solve (the fp32 line of config 2)
    f32, x32 = A32p.clone(), b64.float()

    def step32():
        f32.copy_(A32p)
        x32.copy_(b64.float())
        s.factor(f32)
        s.solve(f32, x32)
    out["fp32_ms"] = timed(step32)
    y.zero_()
    s.add_mv_from(A64, 0, x32.double(), y)
    out["fp32_residual_in_fp64"] = float((y - b64).norm() / b64.norm())
    t_fac32 = timed(lambda: (f32.copy_(A32p), s.factor(f32)))
    out["fp32_factor_ms"] = t_fac32

    # mixed precision: fp32 factor (kept), fp64 residual, fp32 correction
    f32.copy_(A32p)
    s.factor(f32)
    x = torch.zeros_like(b64)
    r = b64.clone()
    hist = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for it in range(8):
        d = r.float()
        s.solve(f32, d)
        x += d.double()
        y.zero_()
        s.add_mv_from(A64, 0, x, y)
        r = b64 - y
        hist.append(float(r.norm() / b64.norm()))
        if len(hist) > 1 and hist[-1] > 0.5 * hist[-2]:
            break
    e1.record()
    torch.cuda.synchronize()
    out["refine_residuals"] = hist
    out["refine_iterations"] = len(hist)
    out["refine_ms_after_factor"] = e0.elapsed_time(e1)
    out["mixed_total_ms"] = t_fac32 + out["refine_ms_after_factor"]
    out["solution_rel_diff_vs_fp64"] = float((x - x64).norm() / x64.norm())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
