"""Shared test helpers: problem families of the reference's tests, dense numpy checkers, oracle access.

The oracle (oracle/, CPU restatement of the reference backends) is used here ONLY as the checker.
"""
import numpy as np

from baspacho_b200 import _capi
from oracle import cpu as oracle_cpu

# reference tolerances: tests/CudaFactorTest.cpp:33-42 (abs Frobenius on the lower triangle)
EPS = {np.float64: (1e-10, 1e-8), np.float32: (1e-5, 5e-5)}
# tolerance against the CPU oracle, elementwise relative to max|L| (stated fp64 tolerance of the parity claim)
ORACLE_RTOL = {np.float64: 5e-13, np.float32: 5e-5}



def eps2(dtype, order):
    """the reference's second tolerance (random families, order ~400), scaled for the larger problems used here"""
    return EPS[dtype][1] * max(1.0, order / (300.0 if dtype == np.float64 else 100.0))


GEN_FLAT, GEN_GRID, GEN_MERIDIANS, GEN_BA, GEN_RANDOM_COLS, GEN_FLAT_SCHUR = range(6)


def oapi():
    return oracle_cpu.api()


def fixture_skel():
    """reference FactorTest.cpp:45-50 / SURVEY appendix B: 6 spans, 3 lumps of width 5"""
    return dict(span_start=[0, 2, 5, 7, 10, 12, 15], lump_to_span=[0, 2, 4, 6], col_ptr=[0, 4, 8, 10],
                row_ind=[0, 1, 3, 5, 2, 3, 4, 5, 4, 5])


def sym_from_lower(dense):
    return np.tril(dense) + np.tril(dense, -1).T


def dense_cholesky(dense_lower):
    a = sym_from_lower(dense_lower.astype(np.float64))
    return np.linalg.cholesky(a)


def random_problem(i, fill=0.037, size=115):
    """randomCols(115, fill, 57+i) with block sizes randomVec(n, 2, 5, 47+i) (reference CudaFactorTest.cpp:81-100)"""
    return oapi().gen_pattern_arrays(GEN_RANDOM_COLS, [size, fill], 2, 5, 57 + i)


def ba_problem(n_pts, n_cams, seed=37, pt_size=3, cam_size=6, mean_extra=3.28, window=40, far=0.1):
    return oapi().gen_pattern_arrays(GEN_BA, [n_pts, n_cams, 2, mean_extra, window, far], pt_size, cam_size, seed)


def make_data(solver, seed, dtype, damp_factor=1.5):
    data = oapi().random_data_array(solver.data_size, -1.0, 1.0, seed, dtype=dtype)
    solver.damp(data, 0.0, solver.order * damp_factor)
    return data


def lower_fro_err(solver, data_a, dense_l):
    got = np.tril(solver.densify(np.ascontiguousarray(data_a)).astype(np.float64))
    return np.linalg.norm(got - np.tril(dense_l))
