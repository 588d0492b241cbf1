"""ORACLE / TEST INFRASTRUCTURE ONLY.

Python face of oracle/liboracle_cpu.so: the CPU restatement of the reference's CPU backends (BackendRef = naive
loops, BackendFast = BLAS + threads) behind the same C ABI as the product, operating on HOST numpy buffers.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
"""
import glob
import os
import site
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from baspacho_b200 import _capi  # noqa: E402  (binding class only; no product code runs through it here)

_LIB = os.path.join(_HERE, "liboracle_cpu.so")
_api = None
_blas_path = None


def build(force=False):
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE])
    return _LIB


def _find_blas():
    roots = list(site.getsitepackages()) + [os.path.dirname(np.__file__) + "/.."]
    cands = []
    for r in roots:
        cands += [(p, "", "") for p in glob.glob(os.path.join(r, "opencv_python_headless.libs", "libopenblas*.so*"))]
        cands += [(p, "scipy_", "") for p in glob.glob(os.path.join(r, "scipy.libs", "libscipy_openblas-*.so"))]
    cands += [(p, "", "") for p in glob.glob("/usr/lib/x86_64-linux-gnu/libopenblas*.so*")]
    return cands


def api():
    global _api, _blas_path
    if _api is None:
        build()
        a = _capi.CApi(_LIB, "oracle_")
        a.lib.oracle_load_blas.argtypes = [_capi.C.c_char_p] * 3
        a.lib.oracle_load_blas.restype = _capi.C.c_int
        for path, prefix, suffix in _find_blas():
            if a.lib.oracle_load_blas(path.encode(), prefix.encode(), suffix.encode()) == 0:
                _blas_path = path
                break
        _api = a
    return _api


def blas_path():
    api()
    return _blas_path


class OracleSolver(_capi.SolverHandle):
    """Host-buffer solver over the CPU backends; numeric methods take numpy arrays and work in place."""

    @classmethod
    def create(cls, param_sizes, ss_ptrs, ss_inds, sparse_elim_ranges=(), elim_last_ids=(), *, backend=_capi.BACKEND_REF, **kw):
        if backend == _capi.BACKEND_FAST and blas_path() is None:
            raise RuntimeError("no OpenBLAS found for the oracle's BackendFast")
        return super().create(api(), param_sizes, ss_ptrs, ss_inds, sparse_elim_ranges, elim_last_ids, backend=backend, **kw)

    @classmethod
    def from_skel(cls, span_start, lump_to_span, col_ptr, row_ind, sparse_elim_ranges=(), permutation=None, *,
                  backend=_capi.BACKEND_REF, **kw):
        return super().from_skel(api(), span_start, lump_to_span, col_ptr, row_ind, sparse_elim_ranges, permutation,
                                 backend=backend, **kw)

    @staticmethod
    def _ptr(a):
        assert a.flags.c_contiguous or a.flags.f_contiguous
        return a.ctypes.data

    def factor(self, data, start_span=0, end_span=-1):
        self.factor_ptr(_capi.dtype_code(data.dtype), self._ptr(data), start_span, end_span)

    def solve(self, data, vec, mode=_capi.SOLVE_LLT, start_span=0, end_span=-1):
        """vec: (n_rhs, ld) C-contiguous numpy array == column-major order x n_rhs with leading dimension ld"""
        n_rhs, ld = (1, vec.shape[0]) if vec.ndim == 1 else vec.shape
        self.solve_ptr(_capi.dtype_code(data.dtype), mode, self._ptr(data), self._ptr(vec), ld, n_rhs, start_span, end_span)

    def do_elimination(self, data, range_index):
        self.do_elimination_ptr(_capi.dtype_code(data.dtype), self._ptr(data), range_index)

    def add_mv_from(self, data, span_index, in_vec, out_vec, alpha=1.0, offset_data=0, offset_vec=0):
        n_rhs, ld = (1, in_vec.shape[0]) if in_vec.ndim == 1 else in_vec.shape
        es = data.dtype.itemsize
        self.add_mv_from_ptr(_capi.dtype_code(data.dtype), self._ptr(data) - offset_data * es, span_index,
                             self._ptr(in_vec) - offset_vec * es, ld, self._ptr(out_vec) - offset_vec * es, ld, n_rhs, alpha)

    def pseudo_factor_from(self, data, span_index):
        self.pseudo_factor_from_ptr(_capi.dtype_code(data.dtype), self._ptr(data), span_index)
