"""GPU probe (run under gpurun): phase stamps + event time of ONE panel launch (96 columns, 5130 rows below) and the
whole dense factorization of a 5226 x 5226 block, for the panel-kernel variant selected by the environment
(BSPB200_PANEL, BSPB200_PANEL_UF). Prints one JSON line."""
import json
import os
import sys

os.environ.setdefault("BSPB200_PANEL_CLK", "1")
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp  # noqa: E402

api = bsp.api()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5, warm=2, setup=None):
    ts = []
    for i in range(warm + reps):
        if setup:
            setup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return min(ts)


def spd(n, rows_below=0):
    M = torch.randn(n, n, dtype=torch.float64, device=dev)
    A = M @ M.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    if rows_below:
        A = torch.cat([A, torch.randn(rows_below, n, dtype=torch.float64, device=dev)], 0)
    return A.contiguous()


out = {"env": {k: v for k, v in os.environ.items() if k.startswith("BSPB200_")}}
n, rb = 96, 5130
A0 = spd(n, rb)
A = A0.clone()
for _ in range(3):
    A.copy_(A0)
    api.check(api.dev_potrf(0, n, rb, A.data_ptr(), n, st))
torch.cuda.synchronize()
buf = np.zeros(64, dtype=np.int64)
api.debug_read(0, buf.ctypes.data, buf.nbytes)
k = int(np.count_nonzero(buf))
out["panel_event_us"] = timeit(lambda: api.check(api.dev_potrf(0, n, rb, A.data_ptr(), n, st)), setup=lambda: A.copy_(A0)) * 1e3
out["panel_stamps"] = np.diff(buf[:k]).tolist()
Lref = torch.linalg.cholesky(A0[:n])
out["panel_err"] = (torch.tril(A[:n]) - Lref).abs().max().item()
nn = 5226
A0 = spd(nn)
A = A0.clone()
out["potrf_5226_ms"] = timeit(lambda: api.check(api.dev_potrf(0, nn, 0, A.data_ptr(), nn, st)), reps=4, warm=2, setup=lambda: A.copy_(A0))
Lref = torch.linalg.cholesky(A0)
out["potrf_5226_rel_err"] = (torch.tril(A) - Lref).abs().max().item() / Lref.abs().max().item()
print(json.dumps(out), flush=True)
