"""Host-side mirror of the reference's public API (baspacho/baspacho/Solver.h:34-237) over the C ABI.

`Solver.create(...)` == createSolver(settings, paramSizes, ss, sparseElimRanges, elimLastIds);
factor / solve / solveL / solveLt / factorUpTo / ... take CUDA tensors (torch is used only to own device
memory and streams) and run the hand-written sm_100a kernels of libbaspacho_b200.so in place.
"""
import os
import subprocess

import numpy as np

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libbaspacho_b200.so")
_api = None


def library_path():
    return _LIB


def build_library(force=False):
    """compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)"""
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")])
    return _LIB


def api():
    global _api
    if _api is None:
        if not os.path.exists(_LIB):
            raise RuntimeError(f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        _api = _capi.CApi(_LIB, "bspb200_")
    return _api


def _dev_ptr(t):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("numeric buffers must be CUDA tensors (device pointers, reference Solver.h:184-188)")
    if not t.is_contiguous():
        raise ValueError("numeric buffers must be contiguous")
    return t.data_ptr()


class Solver(_capi.SolverHandle):
    """BaSpaCho::Solver on the B200 backend. Numeric methods work in place on CUDA tensors."""

    @classmethod
    def create(cls, param_sizes, ss_ptrs, ss_inds, sparse_elim_ranges=(), elim_last_ids=(), *,
               backend=_capi.BACKEND_CUDA, **kw):
        return super().create(api(), param_sizes, ss_ptrs, ss_inds, sparse_elim_ranges, elim_last_ids,
                              backend=backend, **kw)

    @classmethod
    def from_skel(cls, span_start, lump_to_span, col_ptr, row_ind, sparse_elim_ranges=(), permutation=None, *,
                  backend=_capi.BACKEND_CUDA, **kw):
        return super().from_skel(api(), span_start, lump_to_span, col_ptr, row_ind, sparse_elim_ranges, permutation,
                                 backend=backend, **kw)

    def set_stream(self, stream):
        """stream: torch.cuda.Stream or a raw cudaStream_t integer"""
        raw = getattr(stream, "cuda_stream", stream)
        self.api.check(self.api.set_stream(self._h, _capi.vp(raw)))

    @staticmethod
    def _dt(t):
        import torch
        return _capi.F64 if t.dtype == torch.float64 else _capi.dtype_code(np.float32 if t.dtype == torch.float32 else None)

    # ---- single matrix
    def factor(self, data, start_span=0, end_span=-1):
        self.factor_ptr(self._dt(data), _dev_ptr(data), start_span, end_span)

    def factor_up_to(self, data, span):
        self.factor(data, 0, span)

    def factor_from(self, data, span, offset=0):
        """offset: elements by which `data` starts after the beginning of the full factor buffer"""
        self.factor_ptr(self._dt(data), _dev_ptr(data) - offset * data.element_size(), span, -1)

    def solve(self, data, vec, mode=_capi.SOLVE_LLT, start_span=0, end_span=-1):
        """vec: (n_rhs, ld) contiguous CUDA tensor == column-major order x n_rhs with leading dimension ld"""
        n_rhs, ld = (1, vec.shape[0]) if vec.dim() == 1 else vec.shape
        self.solve_ptr(self._dt(data), mode, _dev_ptr(data), _dev_ptr(vec), ld, n_rhs, start_span, end_span)

    def solve_l(self, data, vec, **kw):
        self.solve(data, vec, _capi.SOLVE_L, **kw)

    def solve_lt(self, data, vec, **kw):
        self.solve(data, vec, _capi.SOLVE_LT, **kw)

    def do_elimination(self, data, range_index):
        self.do_elimination_ptr(self._dt(data), _dev_ptr(data), range_index)

    def add_mv_from(self, data, span_index, in_vec, out_vec, alpha=1.0, offset_data=0, offset_vec=0):
        n_rhs, ld = (1, in_vec.shape[0]) if in_vec.dim() == 1 else in_vec.shape
        es = data.element_size()
        self.add_mv_from_ptr(self._dt(data), _dev_ptr(data) - offset_data * es, span_index,
                             _dev_ptr(in_vec) - offset_vec * es, ld, _dev_ptr(out_vec) - offset_vec * es, ld, n_rhs, alpha)

    def pseudo_factor_from(self, data, span_index):
        self.pseudo_factor_from_ptr(self._dt(data), _dev_ptr(data), span_index)

    # ---- batched (identical structure): lists of CUDA tensors, or one (batch, dataSize) tensor
    @staticmethod
    def _items(x):
        return [x[i] for i in range(x.shape[0])] if hasattr(x, "shape") and x.dim() == 2 and not isinstance(x, list) else list(x)

    def factor_batched(self, datas, start_span=0, end_span=-1):
        items = self._items(datas)
        self.factor_batched_ptrs(self._dt(items[0]), [_dev_ptr(d) for d in items], start_span, end_span)

    def solve_batched(self, datas, vecs, mode=_capi.SOLVE_LLT, start_span=0, end_span=-1):
        """vecs: list of (n_rhs, ld) tensors or one (batch, n_rhs, ld) tensor"""
        ditems = self._items(datas)
        vitems = [vecs[i] for i in range(len(ditems))]
        v0 = vitems[0]
        n_rhs, ld = (1, v0.shape[0]) if v0.dim() == 1 else v0.shape
        self.solve_batched_ptrs(self._dt(ditems[0]), mode, [_dev_ptr(d) for d in ditems], [_dev_ptr(v) for v in vitems],
                                ld, n_rhs, start_span, end_span)

    # ---- end to end on HOST buffers (numpy / pinned torch CPU tensors): H2D, factor, solve, D2H
    def factor_solve_host(self, data, vec=None, factor_out=None):
        def hp(a):
            return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        dt = _capi.dtype_code(data.dtype) if isinstance(data, np.ndarray) else self._dt(data)
        if vec is None:
            n_rhs, ld, vptr = 0, 0, 0
        else:
            shape = tuple(vec.shape)
            n_rhs, ld = (1, shape[0]) if len(shape) == 1 else shape
            vptr = hp(vec)
        self.factor_solve_host_ptr(dt, hp(data), hp(factor_out) if factor_out is not None else 0, vptr, ld, n_rhs)

    def host_copy_bytes(self):
        """(h2d, d2h) bytes moved by the last factor_solve_host / factor_solve_host_batched call"""
        a, b = _capi.c_i64(), _capi.c_i64()
        self.api.check(self.api.host_copy_bytes(self._h, _capi.C.byref(a), _capi.C.byref(b)))
        return a.value, b.value

    def factor_solve_host_batched(self, datas, vecs):
        """datas: (batch, dataSize), vecs: (batch, n_rhs, ld) HOST buffers (numpy or pinned torch CPU tensors)"""
        def hp(a):
            return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        batch = len(datas)
        dt = _capi.dtype_code(datas[0].dtype) if isinstance(datas[0], np.ndarray) else self._dt(datas[0])
        shape = tuple(vecs[0].shape)
        n_rhs, ld = (1, shape[0]) if len(shape) == 1 else shape
        dp = (_capi.vp * batch)(*[hp(datas[q]) for q in range(batch)])
        vpz = (_capi.vp * batch)(*[hp(vecs[q]) for q in range(batch)])
        self.api.check(self.api.factor_solve_host_batched(self._h, dt, dp, batch, vpz, ld, n_rhs))

    def launch_count(self):
        return int(self.api.launch_count())
