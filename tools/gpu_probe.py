"""GPU probe (run under gpurun): fp64 peak denominators (cuBLAS DGEMM via torch, HBM copy), our dense kernels at
the shapes the factorization uses, and a mid-size sparse run. Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp  # noqa: E402
from baspacho_b200 import _capi  # noqa: E402

out = {}
dev = torch.device("cuda:0")
api = bsp.api()


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), float(np.median(ts))


# ---- peaks
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev)
b = torch.randn(n, n, dtype=torch.float64, device=dev)
c = torch.empty_like(a)
tmin, tmed = timeit(lambda: torch.matmul(a, b, out=c), reps=5)
out["cublas_dgemm_8192_tflops"] = {"best": 2 * n**3 / tmin / 1e12, "median": 2 * n**3 / tmed / 1e12}
t0 = time.time()
reps = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    torch.matmul(a, b, out=c)
    reps += 1
    if reps % 4 == 0:
        torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
out["cublas_dgemm_8192_tflops"]["sustained"] = reps * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
x = torch.empty(1 << 28, dtype=torch.float64, device=dev)
y = torch.empty_like(x)
tmin, _ = timeit(lambda: y.copy_(x), reps=5)
out["hbm_copy_gbs"] = 2 * x.numel() * 8 / tmin / 1e9
del x, y


def our_gemm(m, n_, k, lower=False, alpha=1.0, beta=0.0):
    A = torch.randn(m, k, dtype=torch.float64, device=dev)
    B = torch.randn(n_, k, dtype=torch.float64, device=dev)
    C = torch.zeros(m, n_, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: api.check(api.dev_gemm_nt(0, m, n_, k, alpha, A.data_ptr(), k, B.data_ptr(), k, beta, C.data_ptr(), n_,
                                           int(lower), st))
    tmin, tmed = timeit(fn)
    ref = A @ B.T
    if lower:
        err = (torch.tril(C) - torch.tril(ref)).abs().max().item()
    else:
        err = (C - ref).abs().max().item()
    flops = 2.0 * m * n_ * k * (0.5 if lower else 1.0)
    tref, _ = timeit(lambda: torch.matmul(A, B.T, out=ref))
    return {"m": m, "n": n_, "k": k, "lower": lower, "tflops": flops / tmin / 1e12, "ms": tmin * 1e3, "max_err": err,
            "cublas_ms": tref * 1e3}


out["gemm"] = [our_gemm(8192, 8192, 8192), our_gemm(4096, 4096, 4096), our_gemm(5226, 5226, 96, lower=True),
               our_gemm(5226, 5226, 192, lower=True), our_gemm(5226, 5226, 512, lower=True),
               our_gemm(4000, 96, 2000), our_gemm(300, 60, 48), our_gemm(5001, 4999, 131)]
print(json.dumps(out, indent=1), flush=True)


def our_potrf(nn):
    M = torch.randn(nn, nn, dtype=torch.float64, device=dev)
    A0 = M @ M.T + nn * torch.eye(nn, dtype=torch.float64, device=dev)
    A = A0.clone()
    st = torch.cuda.current_stream().cuda_stream

    def fn():
        A.copy_(A0)
        api.check(api.dev_potrf(0, nn, 0, A.data_ptr(), nn, st))
    tcopy, _ = timeit(lambda: A.copy_(A0), reps=3, warm=1)
    tmin, tmed = timeit(fn, reps=3, warm=1)
    Lref = torch.linalg.cholesky(A0)
    tref, _ = timeit(lambda: torch.linalg.cholesky(A0), reps=3, warm=1)
    err = (torch.tril(A) - Lref).abs().max().item() / Lref.abs().max().item()
    return {"n": nn, "ms": (tmin - tcopy) * 1e3, "tflops": nn**3 / 3 / (tmin - tcopy) / 1e12, "rel_err": err,
            "torch_cholesky_ms": tref * 1e3}


out["potrf"] = [our_potrf(96), our_potrf(1024), our_potrf(5226)]
print(json.dumps(out["potrf"], indent=1), flush=True)

# ---- mid-size BA-shaped problem end to end (scaled-down config 2)
from tests import helpers as H  # noqa: E402

n_pts, n_cams = 60000, 200
sizes, ptrs, inds = H.ba_problem(n_pts, n_cams)
t0 = time.time()
g = bsp.Solver.create(sizes, ptrs, inds, [0, n_pts])
out["ba_mid"] = {"analysis_s": time.time() - t0, "order": g.order, "data_size": g.data_size, "lumps": g.num_lumps,
                 "work": g.work_estimate()}
data = H.make_data(g, 37, np.float64, 1.2)
d0 = torch.from_numpy(data).cuda()
d = d0.clone()


def fac():
    d.copy_(d0)
    g.factor(d)


tcopy, _ = timeit(lambda: d.copy_(d0), reps=3, warm=1)
tmin, _ = timeit(fac, reps=3, warm=1)
out["ba_mid"]["factor_ms"] = (tmin - tcopy) * 1e3
rhs = torch.randn(1, g.order, dtype=torch.float64, device=dev)
xx = rhs.clone()


def sol():
    xx.copy_(rhs)
    g.solve(d, xx)


tmin, _ = timeit(sol, reps=3, warm=1)
out["ba_mid"]["solve_ms"] = tmin * 1e3
ref = data.copy()
o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, [0, n_pts], backend=_capi.BACKEND_FAST, num_threads=os.cpu_count())
t0 = time.time()
o.factor(ref)
out["ba_mid"]["cpu_fast_factor_ms"] = (time.time() - t0) * 1e3
out["ba_mid"]["max_abs_diff_vs_cpu"] = float(np.abs(d.cpu().numpy() - ref)[np.abs(ref) > 0].max())
out["host_cores"] = os.cpu_count()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out["ba_mid"], indent=1))
