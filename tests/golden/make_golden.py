"""Generates the committed golden fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

What the fixtures are and are not:
* `skeleton_*.json` are the REFERENCE's own golden integer vectors, transcribed from
  baspacho/baspacho/tests/CoalescedBlockMatrixTest.cpp:48-112 (every index array of the 9-span fixture) and from the
  worked example of FactorTest.cpp:45-54. They pin the integer structure bit for bit.
* `numeric_*.npz` hold small seeded problems of the reference's test families (FactorTest.cpp:45-54 fixture skeleton,
  CudaFactorTest.cpp:81-100 random columns, a BA-shaped elimination case) with the factor and the solution computed by
  an INDEPENDENT dense route: numpy.linalg.cholesky / numpy.linalg.solve (LAPACK) on the densified matrix - exactly the
  check the reference's own tests apply (FactorTest.cpp:33-41: dense Eigen::LLT of the densified matrix). The reference
  holds no floating-point golden vectors and cannot be built here (Eigen/dispenso are URL downloads), so these are
  known-answer vectors from LAPACK, not outputs of the reference binary ("parity unpinned" at the bit level).
  The oracle's own output is stored beside them (`oracle_factor`, `oracle_x`) so that a drift of the oracle shows up.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baspacho_b200 import _capi  # noqa: E402
from tests import helpers as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def skeleton_fixtures():
    ref = {
        "source": "baspacho/baspacho/tests/CoalescedBlockMatrixTest.cpp:48-112",
        "input": {"span_start": [0, 1, 2, 4, 5, 7, 9, 12, 14, 16], "lump_to_span": [0, 1, 3, 4, 6, 7, 9],
                  "cols": [[0, 1, 2, 5, 8], [1, 2, 3, 6, 7], [3, 4, 5, 8], [4, 5, 7], [6, 8], [7, 8]]},
        "expected": {
            "spanToLump": [0, 1, 1, 2, 3, 3, 4, 5, 5, 6],
            "lumpStart": [0, 1, 4, 5, 9, 12, 16],
            "chainColPtr": [0, 5, 10, 14, 17, 19, 21],
            "chainRowSpan": [0, 1, 2, 5, 8, 1, 2, 3, 6, 7, 3, 4, 5, 8, 4, 5, 7, 6, 8, 7, 8],
            "chainData": [0, 1, 2, 4, 6, 8, 11, 17, 20, 29, 35, 36, 38, 40, 42, 50, 58, 66, 75, 81, 89, 97],
            "chainRowsTillEnd": [1, 2, 4, 6, 8, 1, 3, 4, 7, 9, 1, 3, 5, 7, 2, 4, 6, 3, 5, 2, 4],
            "boardColPtr": [0, 5, 10, 14, 17, 20, 22],
            "boardRowLump": [0, 1, 3, 5, -1, 1, 2, 4, 5, -1, 2, 3, 5, -1, 3, 5, -1, 4, 5, -1, 5, -1],
            "boardChainColOrd": [0, 1, 3, 4, 5, 0, 2, 3, 4, 5, 0, 1, 3, 4, 0, 2, 3, 0, 1, 2, 0, 2],
            "boardRowPtr": [0, 1, 3, 5, 8, 10, 16],
            "boardColLump": [0, 0, 1, 1, 2, 0, 2, 3, 1, 4, 0, 1, 2, 3, 4, 5],
            "boardColOrd": [0, 1, 0, 1, 0, 2, 1, 0, 2, 0, 3, 3, 2, 1, 1, 0]}}
    json.dump(ref, open(os.path.join(OUT, "skeleton_reference_fixture.json"), "w"), indent=1)
    worked = {
        "source": "baspacho/baspacho/tests/FactorTest.cpp:45-54 (SURVEY.md appendix B)",
        "input": H.fixture_skel(),
        "expected": {"spanToLump": [0, 0, 1, 1, 2, 2, 3], "spanOffsetInLump": [0, 2, 0, 2, 0, 2, 0],
                     "chainData": [0, 10, 25, 40, 55, 65, 80, 90, 105, 115, 130],
                     "chainRowsTillEnd": [2, 5, 8, 11, 2, 5, 7, 10, 2, 5],
                     "boardRowLump": [0, 1, 2, -1, 1, 2, -1, 2, -1], "boardChainColOrd": [0, 2, 3, 4, 0, 2, 4, 0, 2],
                     "boardColOrd": [0, 1, 0, 2, 1, 0]},
        "scalars": {"data_size": 130, "order": 15, "elim_temp_size": 25}}
    json.dump(worked, open(os.path.join(OUT, "skeleton_worked_example.json"), "w"), indent=1)


def numeric_case(name, solver, create_args, seed):
    data = H.make_data(solver, seed, np.float64)
    rhs = H.oapi().random_data_array(solver.order * 2, -1, 1, seed + 1).reshape(2, solver.order)
    A = H.sym_from_lower(solver.densify(data))
    L = np.linalg.cholesky(A)                      # LAPACK dpotrf: the independent known answer
    x = np.linalg.solve(A, rhs.T).T
    fac = data.copy()
    solver.factor(fac)
    xo = rhs.copy()
    solver.solve(fac, xo)
    np.savez_compressed(os.path.join(OUT, f"numeric_{name}.npz"), data=data, rhs=rhs, dense_L=L, x=x,
                        oracle_factor=fac, oracle_x=xo, **{k: np.asarray(v, dtype=np.int64) for k, v in create_args.items()})
    print(name, "order", solver.order, "data", solver.data_size, "|L_oracle - L_lapack|_F",
          H.lower_fro_err(solver, fac, L), "|x_oracle - x_lapack|max", np.abs(xo - x).max())


def numeric_fixtures():
    OS = H.oracle_cpu.OracleSolver
    fs = H.fixture_skel()
    s = OS.from_skel(**fs, backend=_capi.BACKEND_REF)
    numeric_case("factor_fixture", s, fs, 37)
    sizes, ptrs, inds = H.random_problem(0, fill=0.05, size=40)
    s = OS.create(sizes, ptrs, inds, (), backend=_capi.BACKEND_REF, computation_model=_capi.MODEL_CUDA_2080TI)
    numeric_case("random_cols", s, dict(sizes=sizes, ptrs=ptrs, inds=inds, ranges=[]), 41)
    sizes, ptrs, inds = H.ba_problem(40, 5, seed=57, window=3)
    s = OS.create(sizes, ptrs, inds, (0, 40), backend=_capi.BACKEND_REF, computation_model=_capi.MODEL_CUDA_2080TI)
    numeric_case("ba_elim", s, dict(sizes=sizes, ptrs=ptrs, inds=inds, ranges=[0, 40]), 43)


if __name__ == "__main__":
    skeleton_fixtures()
    numeric_fixtures()
