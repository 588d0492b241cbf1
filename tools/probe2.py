"""GPU probe 2 (run under gpurun): phase breakdown of the panel kernel (clock64 stamps, BSPB200_PANEL_CLK=1), the GEMM
shapes of the blocked factorization of a dense n x n lump timed one by one against cuBLAS, and the whole dense
factorization against cuSOLVER (torch.linalg.cholesky). Writes gpurun_out/probe2.json."""
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
out = {}
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
    return min(ts), float(np.median(ts))


def spd(n, rows_below=0):
    M = torch.randn(n, n, dtype=torch.float64, device=dev)
    A = M @ M.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    if rows_below:
        A = torch.cat([A, torch.randn(rows_below, n, dtype=torch.float64, device=dev)], 0)
    return A.contiguous()


# ---- 1. panel kernel phases
for (n, rb) in [(96, 0), (96, 64), (96, 5130), (48, 5130)]:
    A0 = spd(n, rb)
    A = A0.clone()
    for _ in range(3):
        A.copy_(A0)
        api.check(api.dev_potrf(0, n, rb, A.data_ptr(), n, st))
    torch.cuda.synchronize()
    buf = np.zeros(64, dtype=np.int64)
    api.debug_read(0, buf.ctypes.data, buf.nbytes)
    k = int(np.count_nonzero(buf))
    tmin, _ = timeit(lambda: api.check(api.dev_potrf(0, n, rb, A.data_ptr(), n, st)), setup=lambda: A.copy_(A0))
    Lref = torch.linalg.cholesky(A0[:n])
    err = (torch.tril(A[:n]) - Lref).abs().max().item()
    out[f"panel_n{n}_rows{rb}"] = {"event_us": tmin * 1e3, "stamps_cycles": np.diff(buf[:k]).tolist() if k > 1 else [],
                                   "total_cycles": int(buf[k - 1] - buf[0]) if k > 1 else 0, "chol_abs_err": err}
print(json.dumps(out, indent=1), flush=True)

# ---- 2. GEMMs of the blocked factorization of an n x n lump (same recursion as potrfRec)
NB = 96


def rec(total, c0, w, acc):
    if w <= NB:
        acc.append(("panel", total - (c0 + w), w, 0))
        return
    blocks = (w + NB - 1) // NB
    w1 = (blocks // 2) * NB

    def tiles(left):
        m, n2 = total - (c0 + left), w - left
        tn, tm = (n2 + 127) // 128, (m + 127) // 128
        return tn * (tn + 1) // 2 + (tm - tn) * tn
    best = -1
    for b in range(max(1, blocks * 3 // 10), min(blocks - 1, blocks * 7 // 10) + 1):
        t = tiles(b * NB)
        if t < 148:
            continue
        eff = t / (148.0 * ((t + 147) // 148)) - 0.002 * abs(2 * b - blocks)
        if eff > best:
            best, w1 = eff, b * NB
    rec(total, c0, w1, acc)
    r0 = c0 + w1
    acc.append(("gemm", total - r0, w - w1, w1))
    rec(total, r0, w - w1, acc)


n = 5226
ops = []
rec(n, 0, n, ops)
gemms = [(m, nn, k) for (kind, m, nn, k) in ops if kind == "gemm"]
res = []
tot_ours = tot_cub = 0.0
for (m, nn, k) in sorted(set(gemms), key=lambda s: -s[0] * s[1] * s[2]):
    cnt = gemms.count((m, nn, k))
    P = torch.randn(m, k, dtype=torch.float64, device=dev)
    Cm = torch.randn(m, nn, dtype=torch.float64, device=dev)
    fn = lambda: api.check(api.dev_gemm_nt(0, m, nn, k, -1.0, P.data_ptr(), k, P.data_ptr(), k, 1.0, Cm.data_ptr(), nn, 1, st))
    tmin, _ = timeit(fn, reps=4, warm=2)
    Pt = P[:nn].T.contiguous()
    tcub, _ = timeit(lambda: torch.addmm(Cm, P, P[:nn].T, alpha=-1.0, out=Cm), reps=4, warm=2)
    low = nn * (nn + 1) / 2 + max(0, m - nn) * nn
    fl = 2.0 * low * k
    res.append({"m": m, "n": nn, "k": k, "count": cnt, "ms": tmin, "tflops_lower": fl / tmin / 1e9, "cublas_full_ms": tcub,
                "cublas_tflops_full": 2.0 * m * nn * k / tcub / 1e9})
    tot_ours += cnt * tmin
    tot_cub += cnt * tcub
out["gemm_shapes"] = res
out["gemm_total_ms"] = {"ours": tot_ours, "cublas_full": tot_cub, "launches": len(gemms)}
out["panels"] = len([o for o in ops if o[0] == "panel"])
print(json.dumps(out["gemm_total_ms"]), flush=True)

# ---- 3. whole dense factorization vs cuSOLVER
for nn in (1200, 5226):
    A0 = spd(nn)
    A = A0.clone()
    tcopy, _ = timeit(lambda: A.copy_(A0), reps=3, warm=1)
    tmin, _ = timeit(lambda: api.check(api.dev_potrf(0, nn, 0, A.data_ptr(), nn, st)), reps=4, warm=2, setup=lambda: A.copy_(A0))
    tref, _ = timeit(lambda: torch.linalg.cholesky(A0), reps=4, warm=2)
    Lref = torch.linalg.cholesky(A0)
    err = (torch.tril(A) - Lref).abs().max().item() / Lref.abs().max().item()
    out[f"potrf_{nn}"] = {"ms": tmin, "tflops": nn**3 / 3 / tmin / 1e9, "cusolver_ms": tref, "rel_err": err}
print(json.dumps({k: v for k, v in out.items() if k.startswith("potrf")}), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe2.json", "w"), indent=1)
