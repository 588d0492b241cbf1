"""B200-native supernodal sparse Cholesky behind BaSpaCho's Solver / Ops interface.

Python face of the C ABI in include/baspacho_b200.h (libbaspacho_b200.so, hand-written sm_100a kernels).
There is no CPU numeric path here: numeric calls need the CUDA library and a GPU and fail loudly otherwise.
"""
from . import _capi
from ._capi import (BACKEND_CUDA, BACKEND_SYMBOLIC_ONLY, F32, F64, FILL_COMPLETE, FILL_FOR_AUTO_ELIMS,
                    FILL_FOR_GIVEN_ELIMS, FILL_NONE, MODEL_AUTO, MODEL_B200, SOLVE_L, SOLVE_LLT, SOLVE_LT,
                    BaspachoError)
from .solver import Solver, api, build_library, library_path

__all__ = ["Solver", "api", "build_library", "library_path", "BaspachoError", "BACKEND_CUDA",
           "BACKEND_SYMBOLIC_ONLY", "F32", "F64", "FILL_COMPLETE", "FILL_FOR_AUTO_ELIMS", "FILL_FOR_GIVEN_ELIMS",
           "FILL_NONE", "MODEL_AUTO", "MODEL_B200", "SOLVE_L", "SOLVE_LLT", "SOLVE_LT"]
