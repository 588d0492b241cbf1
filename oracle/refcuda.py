"""ORACLE / MEASUREMENT INFRASTRUCTURE ONLY.

Python face of oracle/liboracle_refcuda.so: the REFERENCE's CUDA backend (MatOpsCuda.cu) restated on cuBLAS / cuSOLVER
and thread-per-item kernels with atomics (oracle/RefCudaOps.cu), behind the same C ABI as the product, device pointers.
It is the "second GPU baseline" of SURVEY.md §8(c)/(d): bench.py --impl ref_cuda times it on the same B200, and
tests/ check it against the CPU oracle. Only tests/ and bench.py may import this module.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from baspacho_b200 import _capi  # noqa: E402  (binding class only)
from baspacho_b200.solver import Solver as _DeviceSolver  # noqa: E402  (device-pointer marshalling only)

_LIB = os.path.join(_HERE, "liboracle_refcuda.so")
_api = None


def build(force=False):
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "liboracle_refcuda.so"])
    return _LIB


def api():
    global _api
    if _api is None:
        build()
        _api = _capi.CApi(_LIB, "refcuda_")
    return _api


class RefCudaSolver(_DeviceSolver):
    """createSolver(BackendCuda) of the restated reference CUDA backend; numeric methods take CUDA tensors"""

    @classmethod
    def create(cls, param_sizes, ss_ptrs, ss_inds, sparse_elim_ranges=(), elim_last_ids=(), *,
               backend=_capi.BACKEND_CUDA, **kw):
        return _capi.SolverHandle.create.__func__(cls, api(), param_sizes, ss_ptrs, ss_inds, sparse_elim_ranges,
                                                  elim_last_ids, backend=backend, **kw)

    @classmethod
    def from_skel(cls, span_start, lump_to_span, col_ptr, row_ind, sparse_elim_ranges=(), permutation=None, *,
                  backend=_capi.BACKEND_CUDA, **kw):
        return _capi.SolverHandle.from_skel.__func__(cls, api(), span_start, lump_to_span, col_ptr, row_ind,
                                                     sparse_elim_ranges, permutation, backend=backend, **kw)
