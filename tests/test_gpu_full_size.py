"""Parity at BASELINE.json's FULL sizes (pytest -m gpu, B200 box): every configuration of BASELINE.json that runs on a GPU
is factored and solved through the C ABI with device pointers and compared ELEMENTWISE - every stored lower-triangle
entry of the factor and every entry of the solution - with the CPU oracle (the restated reference BLAS backend,
oracle/CpuOps.cpp) on the same seeded inputs and the same skeleton (index arrays asserted identical first).
Stated tolerance (tests/helpers.py): FACTOR_ULPS / SOLVE_ULPS * eps(fp64) * max|reference|.
The observed errors are written to gpurun_out/parity_observed.json (committed copy: profiles/r02_parity_observed.json)."""
import json
import os

import numpy as np
import pytest

import baspacho_b200 as bsp
from baspacho_b200 import _capi
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **vals):
    path = os.path.join(ROOT, "gpurun_out", "parity_observed.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[name] = vals
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def make_pair(wl, **kw):
    from bench import WORKLOADS
    w = WORKLOADS[wl]
    api = bsp.api()
    sizes, ptrs, inds = api.gen_pattern_arrays(w["kind"], w["params"], w["bsize"][0], w["bsize"][1], 37)
    ranges = [0, w["n_elim"]] if w["n_elim"] else []
    kw = dict(computation_model=_capi.MODEL_B200, find_sparse_elim_ranges=w["auto"], **kw)
    g = bsp.Solver.create(sizes, ptrs, inds, ranges, **kw)
    o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, ranges, backend=_capi.BACKEND_FAST,
                                         num_threads=os.cpu_count(), **kw)
    for name in _capi.ARRAY_IDS:  # same skeleton, permutation and elimination ranges on both sides, bit for bit
        assert np.array_equal(g.array(name), o.array(name)), name
    return g, o, w


def factor_solve_compare(name, g, o, data, rhs, fused=True, factor_ulps=H.FACTOR_ULPS):
    import torch
    g.set_fused(fused)
    d = torch.from_numpy(data).cuda()
    x = torch.from_numpy(rhs).cuda()
    g.factor(d)
    g.solve(d, x)
    torch.cuda.synchronize()
    got_f, got_x = d.cpu().numpy(), x.cpu().numpy()
    del d, x
    ref_f, ref_x = data.copy(), rhs.copy()
    o.factor(ref_f)
    o.solve(ref_f, ref_x)
    mask = H.flat_lower_mask(g)
    assert np.isfinite(got_f[mask]).all() and np.isfinite(got_x).all()
    ef, ex = H.ulp_err(got_f, ref_f, mask), H.ulp_err(got_x, ref_x)
    record(name, factor_ulps=ef, solve_ulps=ex, order=int(g.order), data_size=int(g.data_size),
           max_abs_factor=float(np.abs(ref_f[mask]).max()), max_abs_x=float(np.abs(ref_x).max()), fused=bool(fused))
    assert ef <= factor_ulps, (name, "factor", ef)
    assert ex <= H.SOLVE_ULPS, (name, "solve", ex)
    return got_f, got_x


@pytest.mark.parametrize("fused", [True, False])
def test_config2_bal_full_size_vs_oracle(fused):
    """BASELINE config 2: BAL-shaped 871 cameras x 527 480 points, sparse elimination of the points + the 5226-wide dense
    camera lump; fused (default) path and the fine-grained op sequence INTEGRATION.md's one-line drop-in delivers"""
    g, o, w = make_pair("bal")
    assert g.order == 527480 * 3 + 871 * 6
    data = bsp.api().random_data_array(g.data_size, -1, 1, 37)
    g.damp(data, 0.0, g.order * 1.2)
    rhs = bsp.api().random_data_array(g.order, -1, 1, 38).reshape(1, g.order)
    factor_solve_compare(f"config2_bal_{'fused' if fused else 'unfused'}", g, o, data, rhs, fused)


def test_config3_grid_full_size_vs_oracle():
    """BASELINE config 3: GRID 120 x 120, block 6, connectivity 2, pure supernodal (no sparse elimination)"""
    g, o, w = make_pair("grid")
    assert g.order == 120 * 120 * 6 and g.num_elim_ranges == 0
    data = bsp.api().random_data_array(g.data_size, -1, 1, 37)
    g.damp(data, 0.0, g.order * 1.2)
    rhs = bsp.api().random_data_array(g.order * 2, -1, 1, 38).reshape(2, g.order)
    factor_solve_compare("config3_grid", g, o, data, rhs)


def test_config4_flat_batch64_full_size_vs_oracle():
    """BASELINE config 4: batch = 64 identically structured FLAT size 2000 block 3 fill 0.03 problems, all 64 items of ONE
    factor_batched / solve_batched call, every item against the oracle"""
    import torch
    g, o, w = make_pair("flat_batch")
    batch = w["batch"]
    api = bsp.api()
    datas = np.stack([api.random_data_array(g.data_size, -1, 1, 37 + q) for q in range(batch)])
    for q in range(batch):
        g.damp(datas[q], 0.0, g.order * 1.3)
    rhs = np.stack([api.random_data_array(g.order, -1, 1, 1038 + q).reshape(1, g.order) for q in range(batch)])
    dev = torch.from_numpy(datas).cuda()
    xs = torch.from_numpy(rhs).cuda()
    g.factor_batched(dev)
    g.solve_batched(dev, xs)
    torch.cuda.synchronize()
    got_f, got_x = dev.cpu().numpy(), xs.cpu().numpy()
    del dev, xs
    mask = H.flat_lower_mask(g)
    worst_f = worst_x = 0.0
    for q in range(batch):
        ref_f, ref_x = datas[q].copy(), rhs[q].copy()
        o.factor(ref_f)
        o.solve(ref_f, ref_x)
        worst_f = max(worst_f, H.ulp_err(got_f[q], ref_f, mask))
        worst_x = max(worst_x, H.ulp_err(got_x[q], ref_x))
    record("config4_flat_batch64", factor_ulps=worst_f, solve_ulps=worst_x, order=int(g.order), batch=batch)
    assert worst_f <= H.FACTOR_ULPS and worst_x <= H.SOLVE_ULPS, (worst_f, worst_x)


def test_config5_stress_full_size_vs_oracle():
    """BASELINE config 5: 1 M independent 3x3 point blocks + 200 cameras in one dense supernode (scatter bound)"""
    g, o, w = make_pair("stress")
    assert g.order == 1000000 * 3 + 200 * 6
    data = bsp.api().random_data_array(g.data_size, -1, 1, 37)
    g.damp(data, 0.0, g.order * 1.2)
    rhs = bsp.api().random_data_array(g.order, -1, 1, 38).reshape(1, g.order)
    factor_solve_compare("config5_stress", g, o, data, rhs, factor_ulps=H.FACTOR_ULPS_LONG_SUMS)


def test_config1_flat_vs_oracle():
    """BASELINE config 1 (the reference's CPU-runnable plumbing case) on the device path: FLAT 1000, block 3, fill 0.05"""
    g, o, w = make_pair("flat")
    data = bsp.api().random_data_array(g.data_size, -1, 1, 37)
    g.damp(data, 0.0, g.order * 1.2)
    rhs = bsp.api().random_data_array(g.order * 10, -1, 1, 38).reshape(10, g.order)
    factor_solve_compare("config1_flat", g, o, data, rhs)
