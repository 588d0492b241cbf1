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
    # chain CTA (one SM: its clock64 stamps compare across steps): [0] step start, [3] M1 landed, [4] triangular product,
    # [5] L1 staged + diagonal block zeroed + row copies issued, [7] D built + L1 published, [8] potrf, [9] L(d,d) stored +
    # inverse, [10] W published + next operands requested. %globaltimer (ns): [1] operands of block d published by the
    # accumulate job, [2] chain reaches block d, [12] W_d published. Accumulate jobs: [13] start, [14] main loop, [15] staged.
    segs = [("wait_M1", 0, 3), ("trsm", 3, 4), ("stage_L1+zeroS", 4, 5), ("syrk+D+flag", 5, 7), ("potrf", 7, 8),
            ("L_store+invert", 8, 9), ("W_publish+request", 9, 10)]
    nb = (n + 95) // 96
    rows = []
    for d in range(2, min(nb, 63) - 1):
        s = st_[d]
        rows.append([int(s[b_] - s[a_]) if s[b_] and s[a_] else 0 for _, a_, b_ in segs])
    rows = np.array(rows)
    print("chain steps:", len(rows))
    print("phase (mean cycles over the chain steps):")
    for i, (nm, _, _) in enumerate(segs):
        print(f"  {nm:18s} {rows[:, i].mean():10.0f}   (min {rows[:, i].min()}, max {rows[:, i].max()})")
    print("wait for the next operands per step (k cycles):", " ".join(str(int(x // 1000)) for x in rows[:, 6]))
    s0 = st_[1:min(nb, 63), 0]
    print("chain step (start to start): mean", np.diff(s0).mean(), "cycles =", np.diff(s0).mean() / 1.965e3, "us; median", np.median(np.diff(s0)))
    slack = (st_[2:min(nb, 63), 2] - np.maximum(st_[2:min(nb, 63), 1], st_[2:min(nb, 63), 6])) / 1e3
    print("M1 published after P (us), mean:", ((st_[2:min(nb, 63), 1] - st_[2:min(nb, 63), 6]) / 1e3).mean())
    print("operands ready before the chain arrives (us): mean %.1f min %.1f, first blocks %s" % (slack.mean(), slack.min(), np.round(slack[:6], 1)))
    acc = st_[2:min(nb, 63)]
    print("accumulate jobs: main loop mean", int((acc[:, 14] - acc[:, 13]).mean()), "staging mean", int((acc[:, 15] - acc[:, 14]).mean()))
    ps = st_[63]
    print("potrf panel stamps (d = 20), panel 0:", [int(ps[i + 1] - ps[i]) for i in range(5)], "panel 1:", [int(ps[8 + i + 1] - ps[8 + i]) for i in range(5)],
          " [solve rows, barrier, warp-0 tile update, factor 8x8, barrier]")
    cta = buf[64 * 16:].reshape(1024, 4)
    cta = cta[cta[:, 3] > 0]
    print("CTAs:", len(cta), "jobs/CTA mean", cta[:, 3].mean(), "main loop cycles mean", cta[:, 0].mean(), "epilogue cycles mean", cta[:, 1].mean(), "flag-wait cycles mean", cta[:, 2].mean(),
          "total", (cta[:, 0] + cta[:, 1]).mean(), "=", (cta[:, 0] + cta[:, 1]).mean() / 1.965e6, "ms busy per CTA")
