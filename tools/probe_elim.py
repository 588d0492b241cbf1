"""Sparse-elimination kernels on the BAL-shaped and the stress workload: time and cross-check the gather variants
(BSPB200_GATHER = 1: per-lane direct + staged heavy, 3: warp-cooperative coalesced loads + shuffle exchange)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp
from baspacho_b200 import _capi
from bench import WORKLOADS

api = bsp.api()
out = {}
for wl in (sys.argv[1:] or ["bal", "stress"]):
    w = WORKLOADS[wl]
    sizes, ptrs, inds = api.gen_pattern_arrays(w["kind"], w["params"], w["bsize"][0], w["bsize"][1], 37)
    s = bsp.Solver.create(sizes, ptrs, inds, [0, w["n_elim"]], computation_model=_capi.MODEL_B200, find_sparse_elim_ranges=w["auto"])
    data_h = api.random_data_array(s.data_size, -1, 1, 37)
    s.damp(data_h, 0.0, s.order * 1.2)
    pristine = torch.from_numpy(data_h).cuda()
    work = torch.empty_like(pristine)
    res = {}
    ref = None
    for mode in ("1", "6"):
        os.environ["BSPB200_GATHER"] = mode
        ts = []
        for it in range(6):
            work.copy_(pristine)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            s.do_elimination(work, 0)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        res[f"ms_mode{mode}"] = min(ts)
        cur = work.clone()
        if ref is None:
            ref = cur
        else:
            fin = torch.isfinite(ref) & torch.isfinite(cur)
            res["max_rel_diff_between_modes"] = float(((cur - ref).abs()[fin]).max() / ref.abs()[fin].max())
            work2 = pristine.clone()
            s.do_elimination(work2, 0)
            res["mode6_deterministic"] = bool(torch.equal(work2[fin], cur[fin]))
    api.profile(True)
    work.copy_(pristine)
    s.do_elimination(work, 0)
    torch.cuda.synchronize()
    prof = api.profile_json()
    api.profile(False)
    res["classes_mode6"] = {k: v for k, v in prof.items() if v["launches"]}
    out[wl] = res
    print(wl, json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_elim.json", "w"), indent=1)
