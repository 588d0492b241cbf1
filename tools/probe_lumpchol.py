"""Phase breakdown of the tile-DAG lump Cholesky (BSPB200_LUMPCHOL_DBG=1 stamps): per diagonal job the cycles spent in
each part of the chain, per CTA the totals of main loop vs epilogue; and timings of n = 5226."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp

api = bsp.api()
torch.manual_seed(0)
n, rb = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5226, 0)
M = torch.randn(n, n, dtype=torch.float64, device="cuda")
A11 = M @ M.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
A = torch.cat([A11, torch.randn(rb, n, dtype=torch.float64, device="cuda")]).contiguous()
st = torch.cuda.current_stream().cuda_stream
ufs = os.environ.get("PROBE_UFS", "0").split(",")
for uf in ufs:
    os.environ["BSPB200_LUMPCHOL_UF"] = uf
    print("==== UF", uf)
    os.environ["BSPB200_LUMPCHOL"] = "1"
    W = A.clone()
    for rep in range(3):
        W.copy_(A)
        api.check(api.dev_potrf(0, n, rb, W.data_ptr(), n, st))
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        W.copy_(A)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        api.check(api.dev_potrf(0, n, rb, W.data_ptr(), n, st))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("ms (min, median):", min(ts), sorted(ts)[2])
    os.environ["BSPB200_LUMPCHOL_DBG"] = "1"
    W.copy_(A)
    api.check(api.dev_potrf(0, n, rb, W.data_ptr(), n, st))
    torch.cuda.synchronize()
    buf = np.zeros(64 * 16 + 1024 * 4, dtype=np.int64)
    got = api.debug_read(1, buf.ctypes.data_as(C.c_void_p), buf.nbytes)
    os.environ["BSPB200_LUMPCHOL_DBG"] = "0"
    st_ = buf[:64 * 16].reshape(64, 16)
    segs = [("mainloop", 0, 1), ("stage_M_P", 1, 2), ("wait_W", 2, 3), ("trsm", 3, 4), ("store_L1+zeroS", 4, 5), ("syrk+D", 5, 7),
            ("potrf", 7, 8), ("invert", 8, 9), ("W_store", 9, 10), ("L_store", 10, 11)]
    nb = (n + 95) // 96
    rows = []
    for d in range(1, min(nb, 64) - 1):
        s = st_[d]
        if s[0] == 0:
            continue
        rows.append([int(s[b_] - s[a_]) if s[b_] and s[a_] else 0 for _, a_, b_ in segs])
    rows = np.array(rows)
    print("diag jobs:", len(rows))
    print("phase (mean cycles over diag jobs):")
    for i, (nm, _, _) in enumerate(segs):
        print(f"  {nm:16s} {rows[:, i].mean():10.0f}   (min {rows[:, i].min()}, max {rows[:, i].max()})")
    chain = rows[:, 3:9].sum(axis=1)
    print("chain (trsm .. W_store): mean cycles", chain.mean(), "=", chain.mean() / 1.965e3, "us")
    # W_d flag time differences between consecutive diag jobs = the realized chain step
    w10 = st_[1:nb, 12]  # %globaltimer (ns) when W_d was published: the per-SM clock64 stamps do not compare across CTAs
    w10 = w10[w10 > 0]
    print("W flag to W flag: mean", np.diff(w10).mean() / 1e3, "us per block column (median", np.median(np.diff(w10)) / 1e3, ")")
    sub = st_[2:nb - 1]
    print("store_L1+zeroS split [x->E0, proxy fence, zero S, barrier+issue]:", [int((sub[:, b_] - sub[:, a_]).mean()) for a_, b_ in ((4, 13), (13, 14), (14, 15), (15, 5))])
    ps = st_[63]
    print("potrf panel stamps (d = 20), panel 0:", [int(ps[i + 1] - ps[i]) for i in range(5)], "panel 1:", [int(ps[8 + i + 1] - ps[8 + i]) for i in range(5)],
          " [solve rows, barrier, warp-0 tile update, factor 8x8, barrier]")
    cta = buf[64 * 16:].reshape(1024, 4)
    cta = cta[cta[:, 3] > 0]
    print("CTAs:", len(cta), "jobs/CTA mean", cta[:, 3].mean(), "main loop cycles mean", cta[:, 0].mean(), "epilogue cycles mean", cta[:, 1].mean(), "flag-wait cycles mean", cta[:, 2].mean(),
          "total", (cta[:, 0] + cta[:, 1]).mean(), "=", (cta[:, 0] + cta[:, 1]).mean() / 1.965e6, "ms busy per CTA")
