"""GPU tests of the contracts around the hot path (pytest -m gpu): several Solvers at once on their own streams and host
threads, the device accessor handed by value to a user kernel, the virtual-base-pointer convention of the ...From entry
points, non-SPD input, and partial factor / solve ranges that cut through sparse-elimination ranges on the fused path."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np
import pytest

import baspacho_b200 as bsp
from baspacho_b200 import _capi
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torch_of(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_two_solvers_on_two_threads_and_streams_match_serial_runs_bitwise():
    """Two Solvers (different problems, each with wide lumps - panel kernels, tile-DAG Cholesky, chained solves - and
    small wavefront lumps) factor and solve concurrently from two host threads on two streams, many times; every result
    must equal the serial result of the same solver bit for bit (the device-side state of the kernels - load counters,
    flags, tickets, workspaces - is per solver / per stream, reference MatOpsCuda.cu:55-76 keeps handles per context)."""
    import torch
    problems = []
    for seed, (w, h) in ((37, (30, 30)), (41, (34, 26))):
        sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_GRID, [w, h, 1.0, 2], 6, 6, seed)
        s = bsp.Solver.create(sizes, ptrs, inds, (), find_sparse_elim_ranges=False, computation_model=_capi.MODEL_B200)
        assert np.diff(s.lumpStart).max() > 96
        data = H.make_data(s, seed, np.float64, 1.2)
        rhs = H.oapi().random_data_array(s.order * 2, -1, 1, seed + 1).reshape(2, s.order)
        problems.append((s, data, rhs))
    serial = []
    for s, data, rhs in problems:
        d, x = torch_of(data), torch_of(rhs)
        s.factor(d)
        s.solve(d, x)
        torch.cuda.synchronize()
        serial.append((d.cpu().numpy(), x.cpu().numpy()))
    errors = []

    def worker(k):
        try:
            s, data, rhs = problems[k]
            st = torch.cuda.Stream()
            s.set_stream(st)
            with torch.cuda.stream(st):
                for rep in range(12):
                    d, x = torch_of(data), torch_of(rhs)
                    st.wait_stream(torch.cuda.default_stream())
                    s.factor(d)
                    s.solve(d, x)
                    st.synchronize()
                    mask = H.flat_lower_mask(s)
                    if not np.array_equal(d.cpu().numpy()[mask], serial[k][0][mask]) or not np.array_equal(x.cpu().numpy(), serial[k][1]):
                        errors.append((k, rep))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_device_accessor_passed_by_value_to_a_user_kernel():
    """Solver::deviceAccessor() (reference Solver.h:48, Accessor.h:110-200): a kernel that is NOT part of the library
    receives the accessor by value and writes every block of the lower block pattern (and every diagonal block) on the
    device; the result equals what the host accessor (block_offset, same semantics) places, entry for entry."""
    import torch
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp"), "libaccessor_kernel.so"])
    lib = C.CDLL(os.path.join(ROOT, "tests", "cpp", "libaccessor_kernel.so"))
    lib.accessor_write_blocks.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.accessor_write_blocks.restype = C.c_int
    lib.accessor_entry_value.argtypes = [C.c_int64] * 4
    lib.accessor_entry_value.restype = C.c_double
    sizes, ptrs, inds = H.random_problem(3)
    s = bsp.Solver.create(sizes, ptrs, inds, (), computation_model=_capi.MODEL_CUDA_2080TI)
    dev_ptrs = (C.c_void_p * 8)()
    s.api.check(s.api.device_accessor(s._h, dev_ptrs))
    assert all(dev_ptrs[i] for i in range(8))
    rows, cols = [], []
    for i in range(len(sizes)):  # user block pattern: CSR lower triangle incl. diagonal
        for j in inds[ptrs[i]:ptrs[i + 1]]:
            rows.append(i), cols.append(int(j))
    rb, cb = torch_of(np.array(rows, np.int64)), torch_of(np.array(cols, np.int64))
    data = torch.zeros(s.data_size, dtype=torch.float64, device="cuda")
    rc = lib.accessor_write_blocks(dev_ptrs, len(rows), rb.data_ptr(), cb.data_ptr(), data.data_ptr(), None)
    torch.cuda.synchronize()
    assert rc == 0
    expect = np.zeros(s.data_size)
    for r, c in zip(rows, cols):
        off, stride, flipped = s.block_offset(r, c)
        nr, nc = int(sizes[r]), int(sizes[c])
        for a in range(nr):
            for b in range(nc if r != c else a + 1):
                v = lib.accessor_entry_value(r, c, a, b)
                if r == c:
                    expect[off + a * stride + b] = v
                elif flipped:
                    expect[off + b * stride + a] = v
                else:
                    expect[off + a * stride + b] = v
    assert np.array_equal(data.cpu().numpy(), expect)


@pytest.mark.parametrize("fused", [True, False])
def test_from_entry_points_honour_virtual_base_pointers(fused):
    """factorFrom / solveLFrom / solveLtFrom / addMvFrom are called by the reference's examples with pointers BEFORE the
    start of the real allocation (`matData.data() - spanMatrixOffset(p)`, `vec - spanVectorOffset(p)`,
    examples/Preconditioner.h:118-135): only data[spanMatrixOffset(p)...] and vec[spanVectorOffset(p)...]
    may be touched. The sub-buffers here are exactly that large and are surrounded by canaries."""
    import torch
    for i in range(3):
        sizes, ptrs, inds = H.random_problem(i)
        g = bsp.Solver.create(sizes, ptrs, inds, (), find_sparse_elim_ranges=False, computation_model=_capi.MODEL_CUDA_2080TI)
        o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, (), backend=_capi.BACKEND_REF, find_sparse_elim_ranges=False,
                                             computation_model=_capi.MODEL_CUDA_2080TI)
        g.set_fused(fused)
        cut = int(g.lumpToSpan[g.num_lumps // 2])
        m_off, v_off = g.span_matrix_offset(cut), g.span_vector_offset(cut)
        data = H.make_data(g, 9 + i, np.float64)
        ref = data.copy()
        o.factor(ref, 0, cut)          # the state a caller holds after factorUpTo(cut)
        full = ref.copy()
        o.factor(full, cut, g.num_spans)
        canary = 7.25
        pad = 64
        sub = torch.full((pad + g.data_size - m_off + pad,), canary, dtype=torch.float64, device="cuda")
        sub[pad:pad + g.data_size - m_off] = torch_of(ref[m_off:])
        base = sub.data_ptr() + pad * 8 - m_off * 8  # virtual base: NOT a valid address below m_off
        g.factor_ptr(_capi.F64, base, cut, -1)
        torch.cuda.synchronize()
        got = sub.cpu().numpy()
        assert (got[:pad] == canary).all() and (got[-pad:] == canary).all()
        mask = H.flat_lower_mask(g)[m_off:]
        assert np.abs(got[pad:-pad] - full[m_off:])[mask].max() <= H.TOL_FACTOR * np.abs(full).max()
        # solveLFrom / solveLtFrom on a vector sub-buffer
        rhs = H.oapi().random_data_array(g.order, -1, 1, 70 + i).reshape(1, g.order)
        n_sub = g.order - v_off
        for mode in (bsp.SOLVE_L, bsp.SOLVE_LT):
            vsub = torch.full((pad + n_sub + pad,), canary, dtype=torch.float64, device="cuda")
            vsub[pad:pad + n_sub] = torch_of(rhs[0, v_off:])
            vbase = vsub.data_ptr() + pad * 8 - v_off * 8
            g.solve_ptr(_capi.F64, mode, base, vbase, g.order, 1, cut, -1)
            torch.cuda.synchronize()
            xr = rhs.copy()
            o.solve(full, xr, mode, cut, g.num_spans)
            v = vsub.cpu().numpy()
            assert (v[:pad] == canary).all() and (v[-pad:] == canary).all()
            assert np.abs(v[pad:-pad] - xr[0, v_off:]).max() <= H.TOL_SOLVE * max(1.0, np.abs(xr).max())
        # addMvFrom on the sub-buffers of the UNFACTORED matrix
        msub = torch.full((pad + g.data_size - m_off + pad,), canary, dtype=torch.float64, device="cuda")
        msub[pad:pad + g.data_size - m_off] = torch_of(data[m_off:])
        xin = torch.full((pad + n_sub + pad,), canary, dtype=torch.float64, device="cuda")
        xin[pad:pad + n_sub] = torch_of(rhs[0, v_off:])
        yout = torch.zeros(pad + n_sub + pad, dtype=torch.float64, device="cuda")
        g.add_mv_from_ptr(_capi.F64, msub.data_ptr() + pad * 8 - m_off * 8, cut, xin.data_ptr() + pad * 8 - v_off * 8, g.order,
                          yout.data_ptr() + pad * 8 - v_off * 8, g.order, 1, 1.0)
        torch.cuda.synchronize()
        A = H.sym_from_lower(g.densify(data)).astype(np.float64)
        exp = A[v_off:, v_off:] @ rhs[0, v_off:]
        y = yout.cpu().numpy()
        assert (y[:pad] == 0).all() and (y[-pad:] == 0).all()
        assert np.linalg.norm(y[pad:-pad] - exp) / np.linalg.norm(exp) < 1e-13


@pytest.mark.parametrize("kind", ["wide_lump", "ba_elim", "grid"])
def test_non_spd_input_gives_non_finite_output_and_never_hangs(kind):
    """Non-SPD input is silent in the reference (cusolver's info is discarded, MatOpsCuda.cu:523-526) and callers detect
    it by an isfinite check on the output (examples/Preconditioner.h:178-183). Here: a negative pivot must surface as
    NaN / Inf in the factor and the solution, and nothing may hang - neither the flag waits of the tile-DAG Cholesky nor
    those of the chained triangular solves (the test would time out)."""
    import torch
    if kind == "wide_lump":
        sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_FLAT, [200, 1.0], 3, 3, 37)
        g = bsp.Solver.create(sizes, ptrs, inds, (), find_sparse_elim_ranges=False, computation_model=_capi.MODEL_B200)
    elif kind == "ba_elim":
        sizes, ptrs, inds = H.ba_problem(3000, 80, seed=21, window=12)
        g = bsp.Solver.create(sizes, ptrs, inds, [0, 3000], computation_model=_capi.MODEL_B200)
    else:
        sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_GRID, [30, 30, 1.0, 2], 6, 6, 37)
        g = bsp.Solver.create(sizes, ptrs, inds, (), find_sparse_elim_ranges=False, computation_model=_capi.MODEL_B200)
    data = H.make_data(g, 5, np.float64, 1.2)
    bad = data.copy()
    g.damp(bad, 0.0, -2.4 * g.order)  # diagonal becomes -1.2 * order: negative definite
    d = torch_of(bad)
    x = torch_of(H.oapi().random_data_array(g.order, -1, 1, 38).reshape(1, g.order))
    g.factor(d)
    g.solve(d, x)
    torch.cuda.synchronize()
    assert not bool(torch.isfinite(d).all())
    assert not bool(torch.isfinite(x).all())
    # the solver is still usable afterwards
    d2 = torch_of(data)
    x2 = torch_of(H.oapi().random_data_array(g.order, -1, 1, 38).reshape(1, g.order))
    g.factor(d2)
    g.solve(d2, x2)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(x2).all())


@pytest.mark.parametrize("fused", [True, False])
def test_partial_ranges_across_sparse_elimination_ranges(fused):
    """factorUpTo / factorFrom and solveL / solveLt UpTo / From on a bundle-adjustment shaped problem with a GIVEN
    elimination range plus automatically found ones: cuts at the end of the given range, at the first dense lump, and
    inside the dense part - the early returns and `continue`s of the fused range drivers (B200Ops.cu fusedFactorRange /
    fusedSolveL / fusedSolveLt re-implement the reference's range rules, Solver.cpp:164-219, 268-397) against the oracle."""
    n_pts = 400
    sizes, ptrs, inds = H.ba_problem(n_pts, 60, seed=9, window=6)
    kw = dict(computation_model=_capi.MODEL_CUDA_2080TI)
    g = bsp.Solver.create(sizes, ptrs, inds, [0, n_pts], **kw)
    o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, [0, n_pts], backend=_capi.BACKEND_REF, **kw)
    for name in _capi.ARRAY_IDS:
        np.testing.assert_array_equal(g.array(name), o.array(name), err_msg=name)
    g.set_fused(fused)
    ranges = list(g.array("sparseElimRanges"))
    dense_from = int(g.lumpToSpan[ranges[-1]]) if ranges else 0
    n_spans = g.num_spans
    inside = int(g.lumpToSpan[(ranges[-1] + g.num_lumps) // 2])
    cuts = sorted({int(g.lumpToSpan[r]) for r in ranges[1:]} | {dense_from, inside})
    cuts = [c for c in cuts if 0 < c < n_spans and c <= g.can_factor_up_to_span]
    assert len(cuts) >= 2
    data = H.make_data(g, 11, np.float64)
    mask = H.flat_lower_mask(g)
    rhs = H.oapi().random_data_array(g.order * 2, -1, 1, 77).reshape(2, g.order)
    for cut in cuts:
        d = torch_of(data)
        ref = data.copy()
        g.factor(d, 0, cut)
        o.factor(ref, 0, cut)
        assert np.abs(d.cpu().numpy() - ref)[mask].max() <= H.TOL_FACTOR * np.abs(ref).max(), ("upTo", cut)
        g.factor(d, cut, n_spans)
        o.factor(ref, cut, n_spans)
        assert np.abs(d.cpu().numpy() - ref)[mask].max() <= H.TOL_FACTOR * np.abs(ref).max(), ("from", cut)
        for mode in (bsp.SOLVE_L, bsp.SOLVE_LT):
            for (a, b) in ((0, cut), (cut, n_spans)):
                x = torch_of(rhs)
                g.solve(d, x, mode, a, b)
                xr = rhs.copy()
                o.solve(ref, xr, mode, a, b)
                assert np.abs(x.cpu().numpy() - xr).max() <= H.TOL_SOLVE * max(1.0, np.abs(xr).max()), (mode, a, b)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fragmented_ops_match_oracle_and_dense(dtype):
    """Partial.PartialFragmented{AddMv,SolveL,SolveLt} (reference tests/PartialFactorSolveTest.cpp:522-720) on the device:
    skeletons whose lumps are single spans, nRHS = 1 - solveLFrom / solveLtFrom / addMvFrom run the level-scheduled block
    kernels (fragmentedMV / SolveL / SolveLt of the backend, reference MatOpsFast.cpp:613-1018), checked against dense
    triangular algebra on the bottom-right corner and against the CPU oracle's fragmented ops; a full-range solve and the
    unfused per-lump sequence of the same library as a second opinion."""
    for i in range(6):
        sk = H.fragmented_skel(i)
        n = len(sk["lump_to_span"]) - 1
        nocross = (7 * i) % 150 + 51
        g = bsp.Solver.from_skel(**sk)
        o = H.oracle_cpu.OracleSolver.from_skel(**sk, backend=_capi.BACKEND_FAST, num_threads=2)
        data = H.oapi().random_data_array(g.data_size, -1, 1, 9 + i, dtype=dtype)
        g.damp(data, 0.0, 5.0)
        d = torch_of(data)
        Lm = np.tril(g.densify(data).astype(np.float64))
        r0 = int(g.spanStart[nocross])
        tol = 1e-12 if dtype == np.float64 else 2e-4
        for j in range(2):
            rhs = H.oapi().random_data_array(g.order, -1, 1, 49 + i + j, dtype=dtype).reshape(1, g.order)
            for mode, mat in ((bsp.SOLVE_L, Lm[r0:, r0:]), (bsp.SOLVE_LT, Lm[r0:, r0:].T)):
                x = torch_of(rhs)
                g.solve(d, x, mode, nocross, n)
                got = x.cpu().numpy()[0].astype(np.float64)
                exp = rhs[0].astype(np.float64).copy()
                exp[r0:] = np.linalg.solve(mat, exp[r0:])
                assert np.linalg.norm(got - exp) / np.linalg.norm(exp) < tol
                xr = rhs.copy()
                o.solve(data, xr, mode, nocross, n)
                assert np.abs(got - xr[0]).max() <= (H.TOL_SOLVE if dtype == np.float64 else 1e-4) * max(1.0, np.abs(xr).max())
            # the whole range, fragmented vs the per-lump op sequence (set_fused(False) keeps the fragmented path: it
            # is the reference's own dispatch; BSPB200-independent second opinion = dense algebra)
            x = torch_of(rhs)
            g.solve(d, x, bsp.SOLVE_LLT)
            exp = np.linalg.solve(Lm.T, np.linalg.solve(Lm, rhs[0].astype(np.float64)))
            assert np.linalg.norm(x.cpu().numpy()[0] - exp) / np.linalg.norm(exp) < tol * 10
            y0 = H.oapi().random_data_array(g.order, -1, 1, 149 + i + j, dtype=dtype).reshape(1, g.order)
            xin, y = torch_of(rhs), torch_of(y0)
            g.add_mv_from(d, nocross, xin, y, alpha=3.5)
            A = H.sym_from_lower(g.densify(data)).astype(np.float64)
            exp = y0[0].astype(np.float64).copy()
            exp[r0:] += 3.5 * (A[r0:, r0:] @ rhs[0, r0:].astype(np.float64))
            assert np.linalg.norm(y.cpu().numpy()[0] - exp) / np.linalg.norm(exp) < tol


def test_fp32_gemm_runs_on_tensor_cores_with_fp32_class_accuracy():
    """The fp32 GEMM of the factorization (reference cublasSgemm, MatOpsCuda.cu:568-590; fp32 instantiations
    Solver.cpp:458-536) is a 3xTF32 tensor-core kernel (mma.sync m16n8k8 tf32, hi/lo split): its error against an fp64
    product must be fp32-class (~1e-6 relative to |A||B|), three orders of magnitude below plain TF32 (~1e-3); checked on
    aligned and unaligned shapes, lower-only and with beta."""
    import torch
    api = bsp.api()
    torch.manual_seed(1)
    for (m, n, k, lower, beta) in ((515, 130, 77, False, 0.0), (1024, 1024, 512, True, 1.0), (96, 64, 16, False, 0.5),
                                   (333, 333, 1001, True, 0.0)):
        a = torch.randn(m, k, device="cuda")
        b = torch.randn(n, k, device="cuda")
        c0 = torch.randn(m, n, device="cuda")
        c = c0.clone()
        api.check(api.dev_gemm_nt(_capi.F32, m, n, k, 1.5, a.data_ptr(), k, b.data_ptr(), k, beta, c.data_ptr(), n, int(lower), None))
        torch.cuda.synchronize()
        ref = 1.5 * (a.double() @ b.double().T) + beta * c0.double()
        scale = (a.double().abs() @ b.double().abs().T).max().item()
        err = (c.double() - ref).abs()
        if lower:
            err = torch.tril(err)
            assert torch.equal(torch.triu(c, 1), torch.triu(c0, 1))  # entries above the diagonal are not touched
        assert err.max().item() / scale < 3e-6, (m, n, k, err.max().item() / scale)


def _oracle_pair(sizes, ptrs, inds, ranges, **kw):
    kw = dict(computation_model=_capi.MODEL_B200, **kw)
    g = bsp.Solver.create(sizes, ptrs, inds, ranges, **kw)
    o = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, ranges, backend=_capi.BACKEND_FAST, num_threads=os.cpu_count(), **kw)
    return g, o


@pytest.mark.parametrize("mode", ["1", "6"])
@pytest.mark.parametrize("batch,cam_size", [(1, 6), (3, 6), (1, 9), (2, 9)])
def test_elimination_gather_variants_match_oracle_and_are_deterministic(mode, batch, cam_size, monkeypatch):
    """The two shipped gathers of the sparse elimination - per-lane SIMT (BSPB200_GATHER=1) and one DMMA per block pair
    (=6, the fp64 default; reference kernel replaced: MatOpsCuda.cu:235-331, atomics there) - on a bundle-adjustment-
    shaped problem with long and short task lists, single and batched: each matches the CPU oracle within the stated
    tolerance and gives the same bits twice (fixed summation order, no atomics)."""
    import torch
    monkeypatch.setenv("BSPB200_GATHER", mode)
    # cam_size 9 = pose + intrinsics (the camera model of the BAL data sets): 2 x 2 DMMA tiles per block pair
    sizes, ptrs, inds = H.ba_problem(6000, 40, seed=11, cam_size=cam_size)
    g, o = _oracle_pair(sizes, ptrs, inds, [0, 6000], find_sparse_elim_ranges=True)
    mask = H.flat_lower_mask(g)
    datas = [H.make_data(g, 50 + b, np.float64, 1.3) for b in range(batch)]
    rhss = [H.oapi().random_data_array(g.order, -1, 1, 70 + b) for b in range(batch)]
    runs = []
    for rep in range(2):
        ds, xs = [torch_of(d) for d in datas], [torch_of(r) for r in rhss]
        if batch == 1:
            g.factor(ds[0])
            g.solve(ds[0], xs[0])
        else:
            g.factor_batched(ds)
            g.solve_batched(ds, xs)
        torch.cuda.synchronize()
        runs.append(([d.cpu().numpy() for d in ds], [x.cpu().numpy() for x in xs]))
    for b in range(batch):
        assert np.array_equal(runs[0][0][b][mask], runs[1][0][b][mask]) and np.array_equal(runs[0][1][b], runs[1][1][b])
        ref_f, ref_x = datas[b].copy(), rhss[b].copy()
        o.factor(ref_f)
        o.solve(ref_f, ref_x)
        assert H.ulp_err(runs[0][0][b], ref_f, mask) <= H.FACTOR_ULPS
        assert H.ulp_err(runs[0][1][b], ref_x) <= H.SOLVE_ULPS


def test_eager_schedule_and_solve_lanes_are_deterministic_and_match_oracle():
    """Supernodal tree with many wide lumps (grid problem, no sparse elimination): the factor runs the eager source-driven
    schedule (per-lump events, background lanes), the solve the lanes with lane-private delta vectors (DESIGN.md §5;
    reference loop: Solver.cpp:198-218 / 268-397, strictly sequential there). Three runs must agree bit for bit - the
    order of the contributions into a lump and of the lanes' deltas is fixed, whatever the streams do - and match the
    CPU oracle; two right-hand sides exercise the nRHS > 1 path of the delta vectors."""
    import torch
    sizes, ptrs, inds = H.oapi().gen_pattern_arrays(H.GEN_GRID, [56, 52, 1.0, 2], 6, 6, 23)
    g, o = _oracle_pair(sizes, ptrs, inds, (), find_sparse_elim_ranges=False)
    widths = np.diff(g.lumpStart)
    assert (widths >= 384).sum() >= 4 and (widths >= 192).sum() >= 4  # tile-DAG lumps, eager schedule, solve lanes eligible
    data = H.make_data(g, 5, np.float64, 1.2)
    rhs = H.oapi().random_data_array(g.order * 2, -1, 1, 6).reshape(2, g.order)
    mask = H.flat_lower_mask(g)
    outs = []
    for rep in range(3):
        d, x = torch_of(data), torch_of(rhs)
        g.factor(d)
        g.solve(d, x)
        torch.cuda.synchronize()
        outs.append((d.cpu().numpy(), x.cpu().numpy()))
    for f, x in outs[1:]:
        assert np.array_equal(f[mask], outs[0][0][mask]) and np.array_equal(x, outs[0][1])
    ref_f = data.copy()
    o.factor(ref_f)
    assert H.ulp_err(outs[0][0], ref_f, mask) <= H.FACTOR_ULPS_LONG_SUMS
    for k in range(2):
        ref_x = rhs[k].copy()
        o.solve(ref_f, ref_x)
        assert H.ulp_err(outs[0][1][k], ref_x) <= H.SOLVE_ULPS


@pytest.mark.parametrize("seg,lag", [("0", "0"), ("3", "0"), ("5", "2")])
def test_tile_dag_cholesky_job_list_variants(seg, lag, monkeypatch):
    """lump_chol_kernel with whole sums per job (default) and with the sums cut into segments applied in place
    (BSPB200_LUMPCHOL_SEG / _LAG: measured slower, kept as a switch): same factor within a few ulps of a dense
    float64 Cholesky, for a square lump and a trapezoid with a partial last block (replaces cusolverDnDpotrf +
    cublasDtrsm on a lump column, MatOpsCuda.cu:508-566)."""
    import torch
    monkeypatch.setenv("BSPB200_LUMPCHOL_SEG", seg)
    monkeypatch.setenv("BSPB200_LUMPCHOL_LAG", lag)
    api = bsp.api()
    torch.manual_seed(3)
    for n, rb in ((1346, 0), (1010, 530)):
        M = torch.randn(n, n, dtype=torch.float64, device="cuda")
        A11 = M @ M.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
        A = torch.cat([A11, torch.randn(rb, n, dtype=torch.float64, device="cuda")]).contiguous()
        W = A.clone()
        st = torch.cuda.current_stream().cuda_stream
        api.check(api.dev_potrf(0, n, rb, W.data_ptr(), n, st))
        torch.cuda.synchronize()
        Lref = torch.linalg.cholesky(A11)
        got = torch.tril(W[:n])
        assert float((got - Lref).abs().max() / Lref.abs().max()) < 64 * H.EPS64
        assert torch.equal(torch.triu(W[:n], 1), torch.triu(A11, 1))  # entries above the diagonal are not touched
        if rb:
            Xref = torch.linalg.solve_triangular(Lref, A[n:].T, upper=False).T
            assert float((W[n:] - Xref).abs().max() / Xref.abs().max()) < 256 * H.EPS64
