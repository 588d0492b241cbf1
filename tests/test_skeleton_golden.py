"""Integer structure pinned by the reference's golden vectors (tests/CoalescedBlockMatrixTest.cpp:48-112) and by
the worked example of SURVEY.md appendix B - bit-for-bit. Runs on CPU through both C-ABI libraries' host code."""
import numpy as np
import pytest

import baspacho_b200 as bsp
from baspacho_b200 import _capi
from tests import helpers as H


def _libs():
    return [("oracle", H.oapi(), _capi.BACKEND_REF), ("product", bsp.api(), _capi.BACKEND_SYMBOLIC_ONLY)]


@pytest.mark.parametrize("which", [0, 1])
def test_golden_skeleton_reference_fixture(which):
    name, api, backend = _libs()[which]
    span_start = [0, 1, 2, 4, 5, 7, 9, 12, 14, 16]
    lump_to_span = [0, 1, 3, 4, 6, 7, 9]
    cols = [[0, 1, 2, 5, 8], [1, 2, 3, 6, 7], [3, 4, 5, 8], [4, 5, 7], [6, 8], [7, 8]]
    col_ptr = np.cumsum([0] + [len(c) for c in cols])
    row_ind = [x for c in cols for x in c]
    s = _capi.SolverHandle.from_skel(api, span_start, lump_to_span, col_ptr, row_ind, backend=backend)
    eq = lambda name, exp: np.testing.assert_array_equal(getattr(s, name), np.array(exp, dtype=np.int64), err_msg=name)
    eq("spanToLump", [0, 1, 1, 2, 3, 3, 4, 5, 5, 6])
    eq("lumpStart", [0, 1, 4, 5, 9, 12, 16])
    eq("chainColPtr", [0, 5, 10, 14, 17, 19, 21])
    eq("chainRowSpan", [0, 1, 2, 5, 8, 1, 2, 3, 6, 7, 3, 4, 5, 8, 4, 5, 7, 6, 8, 7, 8])
    eq("chainData", [0, 1, 2, 4, 6, 8, 11, 17, 20, 29, 35, 36, 38, 40, 42, 50, 58, 66, 75, 81, 89, 97])
    eq("chainRowsTillEnd", [1, 2, 4, 6, 8, 1, 3, 4, 7, 9, 1, 3, 5, 7, 2, 4, 6, 3, 5, 2, 4])
    eq("boardColPtr", [0, 5, 10, 14, 17, 20, 22])
    eq("boardRowLump", [0, 1, 3, 5, -1, 1, 2, 4, 5, -1, 2, 3, 5, -1, 3, 5, -1, 4, 5, -1, 5, -1])
    eq("boardChainColOrd", [0, 1, 3, 4, 5, 0, 2, 3, 4, 5, 0, 1, 3, 4, 0, 2, 3, 0, 1, 2, 0, 2])
    eq("boardRowPtr", [0, 1, 3, 5, 8, 10, 16])
    eq("boardColLump", [0, 0, 1, 1, 2, 0, 2, 3, 1, 4, 0, 1, 2, 3, 4, 5])
    eq("boardColOrd", [0, 1, 0, 1, 0, 2, 1, 0, 2, 0, 3, 3, 2, 1, 1, 0])


@pytest.mark.parametrize("which", [0, 1])
def test_worked_example_factor_fixture(which):
    name, api, backend = _libs()[which]
    s = _capi.SolverHandle.from_skel(api, **H.fixture_skel(), backend=backend)
    eq = lambda name, exp: np.testing.assert_array_equal(getattr(s, name), np.array(exp, dtype=np.int64), err_msg=name)
    eq("spanToLump", [0, 0, 1, 1, 2, 2, 3])
    eq("spanOffsetInLump", [0, 2, 0, 2, 0, 2, 0])
    eq("chainData", [0, 10, 25, 40, 55, 65, 80, 90, 105, 115, 130])
    eq("chainRowsTillEnd", [2, 5, 8, 11, 2, 5, 7, 10, 2, 5])
    eq("boardRowLump", [0, 1, 2, -1, 1, 2, -1, 2, -1])
    eq("boardChainColOrd", [0, 2, 3, 4, 0, 2, 4, 0, 2])
    eq("boardColOrd", [0, 1, 0, 2, 1, 0])
    assert s.data_size == 130 and s.order == 15 and s.elim_temp_size == 25


def test_product_and_oracle_build_identical_skeletons():
    """createSolver() in the product library and in the CPU checker must agree on every index array"""
    for i in range(3):
        sizes, ptrs, inds = H.random_problem(i)
        a = _capi.SolverHandle.create(H.oapi(), sizes, ptrs, inds, backend=_capi.BACKEND_REF)
        b = _capi.SolverHandle.create(bsp.api(), sizes, ptrs, inds, backend=_capi.BACKEND_SYMBOLIC_ONLY)
        for name in _capi.ARRAY_IDS:
            np.testing.assert_array_equal(a.array(name), b.array(name), err_msg=name)


def test_c_abi_exports_every_declared_symbol():
    import ctypes
    import re, os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "baspacho_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(bspb200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= len(_capi.DECLARED_SYMBOLS)
    lib = ctypes.CDLL(bsp.library_path())
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted("bspb200_" + s for s in _capi.DECLARED_SYMBOLS) == declared


def test_product_library_has_no_cpu_numeric_path():
    """without a device the numeric backend must refuse loudly (no fallback); with one this test is vacuous"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    sizes, ptrs, inds = H.random_problem(0)
    with pytest.raises(_capi.BaspachoError):
        bsp.Solver.create(sizes, ptrs, inds, backend=bsp.BACKEND_CUDA)
    for backend in (_capi.BACKEND_REF, _capi.BACKEND_FAST):
        with pytest.raises(_capi.BaspachoError):
            bsp.Solver.create(sizes, ptrs, inds, backend=backend)


def test_ref_cuda_baseline_library_loads():
    """oracle/liboracle_refcuda.so (second GPU baseline, measurement infrastructure) builds, loads and exports the C ABI;
    it needs a device to create a solver (cuBLAS handle) - no compute here"""
    from oracle import refcuda
    a = refcuda.api()
    assert b"ref-cuda" in a.version()
    for sym in ("create_solver", "factor", "solve", "set_stream", "do_elimination"):
        assert hasattr(a.lib, "refcuda_" + sym)
