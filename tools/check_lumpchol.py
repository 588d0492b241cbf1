"""GPU check + timing of the tile-DAG lump Cholesky (LumpCholKernel.cu) through bspb200_dev_potrf: correctness against
torch (cuSOLVER) and against the recursive schedule of the same library (BSPB200_LUMPCHOL=0), bitwise determinism, and
CUDA-event timings of both paths and of torch.linalg.cholesky on the same matrices."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp

api = bsp.api()
torch.manual_seed(0)
out = {}


def make(n, rb, ld=None):
    M = torch.randn(n, n, dtype=torch.float64, device="cuda")
    A11 = M @ M.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    A21 = torch.randn(rb, n, dtype=torch.float64, device="cuda")
    return A11, A21, torch.cat([A11, A21]).contiguous()


def run(W, n, rb, mode):
    os.environ["BSPB200_LUMPCHOL"] = "1" if mode else "0"
    api.check(api.dev_potrf(0, n, rb, W.data_ptr(), n, torch.cuda.current_stream().cuda_stream))


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


shapes = [(384, 0), (480, 0), (426, 300), (1000, 700), (600, 5000), (1200, 38), (2000, 2), (5226, 0)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for n, rb in shapes:
    A11, A21, A = make(n, rb)
    L = torch.linalg.cholesky(A11)
    X = torch.linalg.solve_triangular(L, A21.T, upper=False).T if rb else A21
    res = {}
    for mode in (1, 0):
        W = A.clone()
        run(W, n, rb, mode)
        torch.cuda.synchronize()
        e1 = (torch.tril(W[:n]) - L).abs().max().item() / L.abs().max().item()
        e2 = (W[n:] - X).abs().max().item() / max(1e-300, X.abs().max().item()) if rb else 0.0
        res["err_new" if mode else "err_old"] = max(e1, e2)
        if mode:
            W2 = A.clone()
            run(W2, n, rb, 1)
            torch.cuda.synchronize()
            res["deterministic"] = bool(torch.equal(torch.tril(W[:n]), torch.tril(W2[:n])) and torch.equal(W[n:], W2[n:]))
        W0 = A.clone()
        tmin, tmed = timeit(lambda: (W0.copy_(A), run(W0, n, rb, mode)))
        tcopy, _ = timeit(lambda: W0.copy_(A))
        res["ms_new" if mode else "ms_old"] = tmin - tcopy
    if rb == 0:
        tmin, _ = timeit(lambda: torch.linalg.cholesky(A11))
        res["ms_cusolver"] = tmin
    flops = n**3 / 3 + rb * n * n
    res["tflops_new"] = flops / (res["ms_new"] * 1e-3) / 1e12
    out[f"{n}x{rb}"] = res
    print(n, rb, json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/check_lumpchol.json", "w"), indent=1)
