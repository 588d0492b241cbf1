"""GPU parity tests (run on the B200 box: pytest -m gpu). They mirror the reference's CUDA test files
(tests/CudaFactorTest.cpp, BatchedCudaFactorTest.cpp, CudaSolveTest.cpp, BatchedCudaSolveTest.cpp,
CudaPartialTest.cpp, PartialFactorSolveTest.cpp): same problem families, same checks against a dense
Cholesky / dense triangular solves with the reference's tolerances, plus elementwise agreement with the CPU oracle.
Everything goes through the C ABI of libbaspacho_b200.so with device pointers."""
import numpy as np
import pytest

import baspacho_b200 as bsp
from baspacho_b200 import _capi
from tests import helpers as H

pytestmark = pytest.mark.gpu
DTYPES = [np.float64, np.float32]


def torch_of(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_pair(sizes, ptrs, inds, ranges=(), **kw):
    # createSolver picks the supernode-merge cost model from the backend (reference Solver.cpp:679-683): pin one model
    # so that the device solver and the CPU checker build the same skeleton
    kw.setdefault("computation_model", _capi.MODEL_CUDA_2080TI)
    g = bsp.Solver.create(sizes, ptrs, inds, ranges, **kw)
    o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, ranges, backend=_capi.BACKEND_REF, **kw)
    for name in _capi.ARRAY_IDS:  # the index structure both sides work on is identical, bit for bit
        np.testing.assert_array_equal(g.array(name), o.array(name), err_msg=name)
    return g, o


def check_factor(g, o, data, dtype, eps, fused):
    g.set_fused(fused)
    dense = H.sym_from_lower(g.densify(data))
    L = np.linalg.cholesky(dense.astype(np.float64))
    d = torch_of(data)
    g.factor(d)
    got = d.cpu().numpy()
    assert H.lower_fro_err(g, got, L) < eps
    ref = data.copy()
    o.factor(ref)
    scale = np.abs(ref).max()
    lower_mask = np.tril(g.densify(np.ones_like(data))) > 0  # compare stored lower-triangle entries only
    dg, dr = np.tril(g.densify(got)), np.tril(g.densify(ref))
    assert np.abs(dg - dr)[lower_mask].max() <= H.ORACLE_RTOL[dtype] * scale * 20
    return got


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fused", [False, True])
def test_coalesced_factor_fixture(dtype, fused):
    """CudaFactor.CoalescedFactor (CudaFactorTest.cpp:44-73)"""
    g = bsp.Solver.from_skel(**H.fixture_skel())
    o = H.oracle_cpu.OracleSolver.from_skel(**H.fixture_skel())
    data = np.arange(13, 13 + g.data_size, dtype=dtype)
    g.damp(data, 5.0, 50.0)
    check_factor(g, o, data, dtype, H.EPS[dtype][0] * 50, fused)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fused", [False, True])
def test_coalesced_factor_many(dtype, fused):
    """CudaFactor.CoalescedFactor_Many (CudaFactorTest.cpp:75-123): 20 random 115-node problems, no sparse elimination"""
    for i in range(20):
        sizes, ptrs, inds = H.random_problem(i)
        g, o = make_pair(sizes, ptrs, inds, find_sparse_elim_ranges=False)
        data = H.make_data(g, 9 + i, dtype)
        check_factor(g, o, data, dtype, H.eps2(dtype, g.order), fused)


@pytest.mark.parametrize("dtype", DTYPES)
def test_sparse_elim_many(dtype):
    """CudaFactor.SparseElim_Many (CudaFactorTest.cpp:129-183): doElimination alone on a given independent set; only the
    eliminated columns are compared (plus the Schur complement left in the rest, vs the oracle)"""
    for i in range(10):
        n_pts, n_cams = 150 + 10 * i, 12 + i
        sizes, ptrs, inds = H.ba_problem(n_pts, n_cams, seed=57 + i, window=4)
        g, o = make_pair(sizes, ptrs, inds, [0, n_pts], add_fill_policy=_capi.FILL_FOR_GIVEN_ELIMS)
        data = H.make_data(g, 9 + i, dtype)
        L = np.linalg.cholesky(H.sym_from_lower(g.densify(data)).astype(np.float64))
        d = torch_of(data)
        g.do_elimination(d, 0)
        got = d.cpu().numpy()
        ncols = int(g.spanStart[n_pts])
        dg = np.tril(g.densify(got).astype(np.float64))
        assert np.linalg.norm(dg[:, :ncols] - L[:, :ncols]) < H.eps2(dtype, g.order)
        ref = data.copy()
        o.do_elimination(ref, 0)
        dr = np.tril(o.densify(ref).astype(np.float64))
        assert np.abs(dg - dr).max() <= H.ORACLE_RTOL[dtype] * np.abs(dr).max() * 20


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fused", [False, True])
def test_sparse_elim_and_factor_many(dtype, fused):
    """CudaFactor.SparseElimAndFactor_Many (CudaFactorTest.cpp:185-233): given + automatically found elimination ranges"""
    for i in range(10):
        n_pts, n_cams = 200 + 10 * i, 15 + i
        sizes, ptrs, inds = H.ba_problem(n_pts, n_cams, seed=57 + i, window=5)
        g, o = make_pair(sizes, ptrs, inds, [0, n_pts])
        data = H.make_data(g, 9 + i, dtype)
        check_factor(g, o, data, dtype, H.eps2(dtype, g.order), fused)
    for i in range(5):  # automatic ranges: FLAT + schur set (Bench.cpp:303-321 family) and grids
        sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_FLAT_SCHUR, [40, 0.2, 600, 0.02], 2, 4, 37 + i)
        g, o = make_pair(sizes, ptrs, inds)
        assert g.num_elim_ranges >= 1
        data = H.make_data(g, 9 + i, dtype)
        check_factor(g, o, data, dtype, H.eps2(dtype, g.order), fused)


@pytest.mark.parametrize("dtype", DTYPES)
def test_large_dense_supernode(dtype):
    """one wide supernode (blocked potrf/trsm on DMMA tiles): flat pattern with fill 1 -> a single dense lump"""
    sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_FLAT, [130, 1.0], 3, 3, 37)
    g, o = make_pair(sizes, ptrs, inds, find_sparse_elim_ranges=False)
    assert g.num_lumps <= 3
    data = H.make_data(g, 11, dtype, 1.2)
    got = check_factor(g, o, data, dtype, H.eps2(dtype, g.order) * 10, True)
    # and the solve on it
    rhs = H.oapi().random_data_array(g.order * 3, -1, 1, 38, dtype=dtype).reshape(3, g.order)
    x = torch_of(rhs)
    g.solve(torch_of(got), x)
    A = H.sym_from_lower(g.densify(data)).astype(np.float64)
    res = np.linalg.norm(A @ x.cpu().numpy().astype(np.float64).T - rhs.T) / np.linalg.norm(rhs)
    assert res < (1e-12 if dtype == np.float64 else 1e-4)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", [bsp.SOLVE_L, bsp.SOLVE_LT, bsp.SOLVE_LLT])
@pytest.mark.parametrize("fused", [False, True])
def test_solve_many(dtype, mode, fused):
    """CudaSolve.{SolveL,SolveLt}_SparseElimAndFactor_Many (CudaSolveTest.cpp:44-236): nRHS=5 on UNFACTORED damped
    data treated as L, against dense triangular solves"""
    nrhs = 5
    cases = []
    for i in range(6):
        sizes, ptrs, inds = H.random_problem(i, fill=0.03)
        cases.append((sizes, ptrs, inds, ()))
    for i in range(6):
        n_pts = 120 + 10 * i
        sizes, ptrs, inds = H.ba_problem(n_pts, 10 + i, seed=57 + i, window=4)
        cases.append((sizes, ptrs, inds, (0, n_pts)))
    for ci, (sizes, ptrs, inds, ranges) in enumerate(cases):
        g, o = make_pair(sizes, ptrs, inds, list(ranges))
        g.set_fused(fused)
        data = H.make_data(g, 9 + ci, dtype)
        Lm = np.tril(g.densify(data).astype(np.float64))
        rhs = H.oapi().random_data_array(g.order * nrhs, -1, 1, 37 + ci, dtype=dtype).reshape(nrhs, g.order)
        x = torch_of(rhs)
        g.solve(torch_of(data), x, mode)
        got = x.cpu().numpy().astype(np.float64).T
        b = rhs.astype(np.float64).T
        if mode == bsp.SOLVE_L:
            exp = np.linalg.solve(Lm, b)
        elif mode == bsp.SOLVE_LT:
            exp = np.linalg.solve(Lm.T, b)
        else:
            exp = np.linalg.solve(Lm.T, np.linalg.solve(Lm, b))
        assert np.linalg.norm(got - exp) < H.EPS[dtype][0] * 10
        ref = rhs.copy()
        o.solve(data, ref, mode)
        assert np.abs(got - ref.astype(np.float64).T).max() <= H.ORACLE_RTOL[dtype] * max(1.0, np.abs(ref).max()) * 20


@pytest.mark.parametrize("dtype", DTYPES)
def test_batched_factor_and_solve(dtype):
    """BatchedCudaFactor / BatchedCudaSolve (BatchedCudaFactorTest.cpp:44-301, BatchedCudaSolveTest.cpp:44-326):
    batches of 8 and of random sizes 3..31, every item verified separately"""
    import torch
    rng = np.random.RandomState(3)
    for i, batch in enumerate([8] + list(rng.randint(3, 32, size=3))):
        n_pts = 100 + 20 * i
        sizes, ptrs, inds = H.ba_problem(n_pts, 9 + i, seed=40 + i, window=3)
        g, o = make_pair(sizes, ptrs, inds, [0, n_pts])
        datas = [H.make_data(g, 100 + q, dtype) for q in range(batch)]
        dev = torch.stack([torch_of(d) for d in datas])
        g.factor_batched(dev)
        rhs = [H.oapi().random_data_array(g.order * 2, -1, 1, 300 + q, dtype=dtype).reshape(2, g.order) for q in range(batch)]
        xs = torch.stack([torch_of(r) for r in rhs])
        g.solve_batched(dev, xs)
        for q in range(batch):
            L = np.linalg.cholesky(H.sym_from_lower(g.densify(datas[q])).astype(np.float64))
            assert H.lower_fro_err(g, dev[q].cpu().numpy(), L) < H.eps2(dtype, g.order)
            A = H.sym_from_lower(g.densify(datas[q])).astype(np.float64)
            x = xs[q].cpu().numpy().astype(np.float64).T
            res = np.linalg.norm(A @ x - rhs[q].T) / np.linalg.norm(rhs[q])
            assert res < (1e-12 if dtype == np.float64 else 2e-4)


@pytest.mark.parametrize("dtype", DTYPES)
def test_partial_factor_and_solve(dtype):
    """Partial.{PartialFactor,SplitFactor,PartialSolveL/Lt(+From)} (PartialFactorSolveTest.cpp:48-520) vs the oracle"""
    for i in range(4):
        sizes, ptrs, inds = H.random_problem(i)
        g, o = make_pair(sizes, ptrs, inds, find_sparse_elim_ranges=False)
        n_spans = g.num_spans
        cut_lump = g.num_lumps // 2
        cut = int(g.lumpToSpan[cut_lump])
        data = H.make_data(g, 9 + i, dtype)
        tol = H.ORACLE_RTOL[dtype] * 20
        # factorUpTo, then factorFrom completes it
        d = torch_of(data)
        g.factor(d, 0, cut)
        ref = data.copy()
        o.factor(ref, 0, cut)
        mask = np.tril(g.densify(np.ones_like(data))) > 0
        dg, dr = g.densify(d.cpu().numpy()), o.densify(ref)
        assert np.abs(dg - dr)[mask].max() <= tol * np.abs(ref).max()
        g.factor(d, cut, n_spans)
        o.factor(ref, cut, n_spans)
        full = data.copy()
        o.factor(full)
        dg, dr, df = g.densify(d.cpu().numpy()), o.densify(ref), o.densify(full)
        assert np.abs(dg - dr)[mask].max() <= tol * np.abs(ref).max()
        assert np.abs(dg - df)[mask].max() <= tol * 10 * np.abs(ref).max()
        # partial solves
        rhs = H.oapi().random_data_array(g.order * 3, -1, 1, 77 + i, dtype=dtype).reshape(3, g.order)
        for mode in (bsp.SOLVE_L, bsp.SOLVE_LT):
            for (a, b) in ((0, cut), (cut, n_spans)):
                x = torch_of(rhs)
                g.solve(d, x, mode, a, b)
                xr = rhs.copy()
                o.solve(ref, xr, mode, a, b)
                assert np.abs(x.cpu().numpy() - xr).max() <= tol * max(1.0, np.abs(xr).max())


@pytest.mark.parametrize("dtype", DTYPES)
def test_add_mv_and_pseudo_factor(dtype):
    """CudaPartial.{PartialAddMv,testPseudoFactor} (CudaPartialTest.cpp:48-185) vs the oracle and dense algebra"""
    for i in range(4):
        sizes, ptrs, inds = H.random_problem(i)
        g, o = make_pair(sizes, ptrs, inds, find_sparse_elim_ranges=False)
        cut = int(g.lumpToSpan[g.num_lumps // 3])
        data = H.make_data(g, 9 + i, dtype)
        rhs = H.oapi().random_data_array(g.order * 3, -1, 1, 5 + i, dtype=dtype).reshape(3, g.order)
        out0 = H.oapi().random_data_array(g.order * 3, -1, 1, 6 + i, dtype=dtype).reshape(3, g.order)
        x, y = torch_of(rhs), torch_of(out0)
        g.add_mv_from(torch_of(data), cut, x, y, alpha=0.7)
        A = H.sym_from_lower(g.densify(data)).astype(np.float64)
        r0 = int(g.spanStart[cut])
        exp = out0.astype(np.float64).T.copy()
        exp[r0:] += 0.7 * (A[r0:, r0:] @ rhs.astype(np.float64).T[r0:])
        assert np.linalg.norm(y.cpu().numpy().astype(np.float64).T - exp) / np.linalg.norm(exp) < (1e-13 if dtype == np.float64 else 1e-5)
        d = torch_of(data)
        g.pseudo_factor_from(d, cut)
        ref = data.copy()
        o.pseudo_factor_from(ref, cut)
        assert np.abs(d.cpu().numpy() - ref).max() <= H.ORACLE_RTOL[dtype] * 20 * np.abs(ref).max()


@pytest.mark.parametrize("dtype", DTYPES)
def test_add_mv_over_sparse_elimination_ranges(dtype):
    """addMvFrom started inside / in front of sparse-elimination ranges: the backend's two-launch product over a whole
    range (MatOps.h sparseElimMV; the reference walks the lumps, Solver.cpp:408-446) against dense algebra on the
    densified matrix, from span 0 (both ranges in one go), from the second range, and from the dense part; three
    right-hand sides, alpha != 1, output accumulated onto a non-zero vector."""
    n_pts = 420
    sizes, ptrs, inds = H.ba_problem(n_pts, 14, seed=9, window=4)
    g, o = make_pair(sizes, ptrs, inds, [0, 200, n_pts])
    assert g.num_elim_ranges >= 2
    data = H.make_data(g, 4, dtype)
    A = H.sym_from_lower(g.densify(data)).astype(np.float64)
    rhs = H.oapi().random_data_array(g.order * 3, -1, 1, 15, dtype=dtype).reshape(3, g.order)
    out0 = H.oapi().random_data_array(g.order * 3, -1, 1, 16, dtype=dtype).reshape(3, g.order)
    ranges = g.sparseElimRanges
    for cut_lump in (0, int(ranges[1]), int(ranges[-1])):
        cut = int(g.lumpToSpan[cut_lump])
        x, y = torch_of(rhs), torch_of(out0)
        g.add_mv_from(torch_of(data), cut, x, y, alpha=-1.3)
        r0 = int(g.spanStart[cut])
        exp = out0.astype(np.float64).T.copy()
        exp[r0:] += -1.3 * (A[r0:, r0:] @ rhs.astype(np.float64).T[r0:])
        err = np.linalg.norm(y.cpu().numpy().astype(np.float64).T - exp) / np.linalg.norm(exp)
        assert err < (1e-13 if dtype == np.float64 else 1e-5), (cut_lump, err)


def test_factor_solve_host_end_to_end():
    """the host-buffer entry point bench.py's e2e leg times: H2D, factor, solve, D2H"""
    n_pts = 400
    sizes, ptrs, inds = H.ba_problem(n_pts, 20, seed=5, window=5)
    g, o = make_pair(sizes, ptrs, inds, [0, n_pts])
    data = H.make_data(g, 3, np.float64, 1.2)
    rhs = H.oapi().random_data_array(g.order, -1, 1, 38).reshape(1, g.order)
    x = rhs.copy()
    fac = np.empty_like(data)
    n0 = g.launch_count()
    g.factor_solve_host(data, x, fac)
    assert g.launch_count() > n0
    A = H.sym_from_lower(g.densify(data))
    assert np.linalg.norm(A @ x[0] - rhs[0]) / np.linalg.norm(rhs[0]) < 1e-12
    ref = data.copy()
    o.factor(ref)
    mask = np.tril(g.densify(np.ones_like(data))) > 0
    assert np.abs(g.densify(fac) - o.densify(ref))[mask].max() <= H.TOL_FACTOR * np.abs(ref).max()


def test_host_end_to_end_wide_lump_and_heavy_destinations():
    """BA-shaped problem whose camera part is one 600-wide lump (the host entry point uploads wide diagonal blocks in row
    bands that stop at the diagonal) and whose diagonal camera blocks collect > 512 pair tasks each (the staged,
    whole-CTA gather kernel of the sparse elimination), checked against the oracle and against A x = b."""
    n_pts, n_cams = 14000, 140
    sizes, ptrs, inds = H.ba_problem(n_pts, n_cams, seed=11, window=20)
    g, o = make_pair(sizes, ptrs, inds, [0, n_pts], computation_model=_capi.MODEL_B200)
    assert np.diff(g.lumpStart).max() >= 512
    data = H.make_data(g, 5, np.float64, 1.2)
    rhs = H.oapi().random_data_array(g.order, -1, 1, 38).reshape(1, g.order)
    x = rhs.copy()
    fac = np.empty_like(data)
    g.factor_solve_host(data, x, fac)
    ref, xr = data.copy(), rhs.copy()
    o.factor(ref)
    o.solve(ref, xr)
    mask = np.tril(g.densify(np.ones_like(data))) > 0
    assert np.abs(g.densify(fac) - o.densify(ref))[mask].max() <= H.TOL_FACTOR * np.abs(ref).max()
    assert np.abs(x - xr).max() <= H.TOL_SOLVE * max(1.0, np.abs(xr).max())
    # device-pointer path on the same problem (elimination + blocked dense factorization of the wide lump)
    d = torch_of(data)
    g.factor(d)
    assert np.abs(g.densify(d.cpu().numpy()) - o.densify(ref))[mask].max() <= H.TOL_FACTOR * np.abs(ref).max()


@pytest.mark.parametrize("model", [_capi.MODEL_B200, _capi.MODEL_OPENBLAS_I7])
def test_grid_wide_and_small_lumps(model):
    """GRID family (Bench.cpp:322-343 genGrid): wide supernodes with rows below (blocked trapezoid Cholesky, GEMM+assemble
    updates) mixed with small ones (wavefront kernels), factor AND solve against the oracle - the multi-lump case that the
    single dense-supernode tests do not cover."""
    sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_GRID, [36, 36, 1.0, 2], 6, 6, 37)
    g, o = make_pair(sizes, ptrs, inds, find_sparse_elim_ranges=False, computation_model=model)
    widths = np.diff(g.lumpStart)
    assert widths.max() > 96
    data = H.make_data(g, 37, np.float64, 1.2)
    ref = data.copy()
    o.factor(ref)
    for rep in range(2):  # twice: catches stream-ordering races that a single run can miss
        d = torch_of(data)
        g.factor(d)
        got = d.cpu().numpy()
        mask = np.tril(g.densify(np.ones_like(data))) > 0
        assert np.abs(g.densify(got) - o.densify(ref))[mask].max() <= H.TOL_FACTOR * np.abs(ref).max()
    rhs = H.oapi().random_data_array(g.order * 2, -1, 1, 38).reshape(2, g.order)
    x = torch_of(rhs)
    g.solve(d, x)
    xr = rhs.copy()
    o.solve(ref, xr)
    assert np.abs(x.cpu().numpy() - xr).max() <= H.TOL_SOLVE * max(1.0, np.abs(xr).max())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n_blocks3", [40, 130, 200])
def test_chain_solve_wide_lump(dtype, n_blocks3, monkeypatch):
    """flag-chained dense triangular solve (trsv_chain_kernel: one launch per lump and direction): a dense lump of
    2 / 5 / 7 block rows (the last one partial), nRHS 1 (NR=1 instantiation) and 5 (two groups of the NR=4 one),
    L / Lt / LLt separately, against the oracle, dense triangular solves, and the per-step launch path of the same
    library (BSPB200_CHAIN_SOLVE=0)"""
    sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_FLAT, [n_blocks3, 1.0], 3, 3, 37)
    g, o = make_pair(sizes, ptrs, inds, find_sparse_elim_ranges=False)
    monkeypatch.setenv("BSPB200_CHAIN_SOLVE", "0")
    g_steps = bsp.Solver.create(sizes, ptrs, inds, (), find_sparse_elim_ranges=False,
                                computation_model=_capi.MODEL_CUDA_2080TI)
    monkeypatch.delenv("BSPB200_CHAIN_SOLVE")
    assert np.diff(g.lumpStart).max() > 96
    data = H.make_data(g, 11, dtype, 1.2)
    d = torch_of(data)
    g.factor(d)
    fac = d.cpu().numpy()
    Lm = np.tril(g.densify(fac).astype(np.float64))
    tol = H.ORACLE_RTOL[dtype] * 50
    for nrhs in (1, 5):
        rhs = H.oapi().random_data_array(g.order * nrhs, -1, 1, 38 + nrhs, dtype=dtype).reshape(nrhs, g.order)
        b = rhs.astype(np.float64).T
        for mode in (bsp.SOLVE_L, bsp.SOLVE_LT, bsp.SOLVE_LLT):
            for rep in range(2):  # twice: the second launch runs on advanced epochs / tickets
                x = torch_of(rhs)
                g.solve(d, x, mode)
            got = x.cpu().numpy().astype(np.float64).T
            if mode == bsp.SOLVE_L:
                exp = np.linalg.solve(Lm, b)
            elif mode == bsp.SOLVE_LT:
                exp = np.linalg.solve(Lm.T, b)
            else:
                exp = np.linalg.solve(Lm.T, np.linalg.solve(Lm, b))
            scale = max(1.0, np.abs(exp).max())
            assert np.abs(got - exp).max() <= tol * scale
            xr = rhs.copy()
            o.solve(fac, xr, mode)
            assert np.abs(got - xr.astype(np.float64).T).max() <= tol * scale
            xs = torch_of(rhs)
            g_steps.solve(d, xs, mode)
            assert np.abs(got - xs.cpu().numpy().astype(np.float64).T).max() <= tol * scale


def test_chain_solve_batched_wide_lump():
    """batched solve whose dense lump spans several block rows: more chain CTAs (batch x blocks) than one wave of
    resident CTAs is not needed for correctness - the arrival ticket orders them - checked per item"""
    import torch
    n_pts, n_cams, batch = 3000, 60, 37
    sizes, ptrs, inds = H.ba_problem(n_pts, n_cams, seed=21, window=12)
    g, o = make_pair(sizes, ptrs, inds, [0, n_pts], computation_model=_capi.MODEL_B200)
    assert np.diff(g.lumpStart).max() >= 300
    datas = [H.make_data(g, 500 + q, np.float64, 1.2) for q in range(batch)]
    dev = torch.stack([torch_of(dd) for dd in datas])
    g.factor_batched(dev)
    rhs = [H.oapi().random_data_array(g.order, -1, 1, 900 + q).reshape(1, g.order) for q in range(batch)]
    xs = torch.stack([torch_of(r) for r in rhs])
    g.solve_batched(dev, xs)
    for q in range(0, batch, 6):
        ref, xr = datas[q].copy(), rhs[q].copy()
        o.factor(ref)
        o.solve(ref, xr)
        assert np.abs(xs[q].cpu().numpy() - xr).max() <= H.TOL_SOLVE * max(1.0, np.abs(xr).max())


def test_ref_cuda_baseline_against_oracle():
    """the restated reference CUDA backend (oracle/RefCudaOps.cu: cuSOLVER/cuBLAS per lump + thread-per-pair elimination
    with atomics) - the second GPU baseline bench.py --impl ref_cuda times - computes the same factor and solution as the
    CPU oracle (atomics reorder the sums: tolerance, not bits)"""
    from oracle import refcuda
    cases = []
    for i in range(2):
        sizes, ptrs, inds = H.random_problem(i)
        cases.append((sizes, ptrs, inds, ()))
    sizes, ptrs, inds = H.ba_problem(300, 12, seed=61, window=4)
    cases.append((sizes, ptrs, inds, (0, 300)))
    for ci, (sizes, ptrs, inds, ranges) in enumerate(cases):
        kw = dict(computation_model=_capi.MODEL_CUDA_2080TI)
        r = refcuda.RefCudaSolver.create(sizes, ptrs, inds, list(ranges), **kw)
        o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, list(ranges), backend=_capi.BACKEND_REF, **kw)
        for name in _capi.ARRAY_IDS:
            np.testing.assert_array_equal(r.array(name), o.array(name), err_msg=name)
        data = H.make_data(r, 21 + ci, np.float64)
        rhs = H.oapi().random_data_array(r.order * 2, -1, 1, 70 + ci).reshape(2, r.order)
        d, x = torch_of(data), torch_of(rhs)
        r.factor(d)
        r.solve(d, x)
        ref, xr = data.copy(), rhs.copy()
        o.factor(ref)
        o.solve(ref, xr)
        mask = np.tril(r.densify(np.ones_like(data))) > 0
        assert np.abs(r.densify(d.cpu().numpy()) - o.densify(ref))[mask].max() <= H.TOL_FACTOR * np.abs(ref).max()
        assert np.abs(x.cpu().numpy() - xr).max() <= H.TOL_SOLVE * max(1.0, np.abs(xr).max())


def test_full_size_headline_properties():
    """BASELINE.json's headline configuration at FULL size (BAL-shaped 871 cameras x 527 480 points, 82 M factor entries -
    too large for the dense / oracle checks above), through size-independent properties: the residual of A x = b computed
    with the block-sparse product (addMvFrom), linearity of the solve, and run-to-run determinism of factor and solve (the
    sparse elimination here has no atomics). Same calls and inputs as bench.py's timed step."""
    import torch
    from bench import WORKLOADS
    w = WORKLOADS["bal"]
    api = bsp.api()
    sizes, ptrs, inds = api.gen_pattern_arrays(w["kind"], w["params"], w["bsize"][0], w["bsize"][1], 37)
    s = bsp.Solver.create(sizes, ptrs, inds, [0, w["n_elim"]], computation_model=_capi.MODEL_B200,
                          find_sparse_elim_ranges=w["auto"])
    assert s.order == 527480 * 3 + 871 * 6
    data_h = api.random_data_array(s.data_size, -1, 1, 37)
    s.damp(data_h, 0.0, s.order * 1.2)
    pristine = torch.from_numpy(data_h).cuda()
    b1 = torch.from_numpy(api.random_data_array(s.order, -1, 1, 38).reshape(1, s.order)).cuda()
    b2 = torch.from_numpy(api.random_data_array(s.order, -1, 1, 39).reshape(1, s.order)).cuda()

    fac = pristine.clone()
    s.factor(fac)
    x1 = b1.clone()
    s.solve(fac, x1)
    assert bool(torch.isfinite(x1).all())
    # (1) residual through the block-sparse symmetric product on the UNFACTORED matrix
    y = torch.zeros_like(b1)
    s.add_mv_from(pristine, 0, x1, y)
    torch.cuda.synchronize()
    assert float((y - b1).norm() / b1.norm()) < 1e-12
    # (2) linearity: solve(b1 + b2) == solve(b1) + solve(b2)
    x2, x12 = b2.clone(), (b1 + b2).clone()
    s.solve(fac, x2)
    s.solve(fac, x12)
    assert float((x12 - (x1 + x2)).abs().max() / x12.abs().max()) < 1e-12
    # (3) determinism: a second factorization and solve of the same input reproduce the first bit for bit
    fac_b = pristine.clone()
    s.factor(fac_b)
    x1_b = b1.clone()
    s.solve(fac_b, x1_b)
    torch.cuda.synchronize()
    # (the upper triangles of the diagonal blocks are don't-care entries of the format: equal or both NaN)
    assert bool(((fac == fac_b) | (fac.isnan() & fac_b.isnan())).all())
    assert torch.equal(x1, x1_b)
