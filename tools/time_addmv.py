"""time of Solver::addMvFrom from span 0 on a bench workload (block-sparse symmetric MV incl. the elimination ranges)"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp
from bench import gen_problem
wl = sys.argv[1] if len(sys.argv) > 1 else "bal"
api = bsp.api()
sizes, ptrs, inds, ranges, w = gen_problem(api, wl)
s = bsp.Solver.create(sizes, ptrs, inds, ranges, computation_model=2, find_sparse_elim_ranges=w["auto"])
A = torch.empty(s.data_size, dtype=torch.float64, device="cuda").uniform_(-1, 1)
x = torch.empty(s.order, dtype=torch.float64, device="cuda").uniform_(-1, 1)
y = torch.zeros_like(x)
ts = []
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); s.add_mv_from(A, 0, x, y); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(json.dumps({"workload": w["desc"], "add_mv_from_span0_ms": min(ts), "all_ms": ts, "bytes": int(s.data_size) * 8,
                  "GBps": s.data_size * 8 / (min(ts) * 1e-3) / 1e9}))
