import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import baspacho_b200 as bsp
from baspacho_b200 import _capi
from tests import helpers as H
api = bsp.api()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 60
sizes, ptrs, inds = api.gen_pattern_arrays(1, [G, G, 1.0, 2], 6, 6, 37)
s = bsp.Solver.create(sizes, ptrs, inds, computation_model=int(os.environ.get("MODEL", 2)), find_sparse_elim_ranges=False)
o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, backend=_capi.BACKEND_FAST, num_threads=16, computation_model=int(os.environ.get("MODEL", 2)), find_sparse_elim_ranges=False)
print("lumps", s.num_lumps, "widths", np.diff(s.lumpStart)[-12:])
data = H.make_data(s, 37, np.float64, 1.2)
ref = data.copy(); o.factor(ref)
stream = torch.cuda.Stream(); s.set_stream(stream)
for rep in range(3):
    d = torch.from_numpy(data).cuda()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        s.factor(d)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    mask = np.tril(s.densify(np.ones_like(data))) > 0
    # per lump error
    cd = s.chainData; ccp = s.chainColPtr
    worst = []
    for l in range(s.num_lumps):
        a, b = cd[ccp[l]], cd[ccp[l + 1]]
        w = s.lumpStart[l + 1] - s.lumpStart[l]
        blk_g = got[a:b].reshape(-1, w).copy(); blk_r = ref[a:b].reshape(-1, w).copy()
        iu = np.triu_indices(w, 1); blk_g[iu] = 0; blk_r[iu] = 0
        e = np.abs(blk_g - blk_r).max()
        if e > 1e-9:
            rows, cols = np.where(np.abs(blk_g - blk_r) > 1e-9)
            worst.append((l, int(w), blk_g.shape[0], float(e), int(rows.min()), int(rows.max()), int(cols.min()), int(cols.max())))
    print("rep", rep, "bad lumps:", worst[:6], "count", len(worst))
