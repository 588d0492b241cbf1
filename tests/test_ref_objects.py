"""oracle/_ref: the Eigen-free translation units of the REFERENCE (baspacho/baspacho/Utils.cpp, SparseStructure.cpp,
baspacho/testing/TestingUtils.cpp, TestingMatGen.cpp, MathUtils.h), compiled from /root/reference by `make -C oracle ref`
into oracle/_ref/libref_host.so. Every scenario of oracle/host_scenarios.h is run through the reference's own object
code and through this repo's restatement (csrc/host, csrc/testing, oracle/SmallBlockMath.h); outputs must be identical
bit for bit: integers as they are, floating point by bit pattern. This pins the synthetic inputs of every benchmark
configuration (randomData, genFlat / genGrid / addSchurSet ...), the pattern algebra feeding the skeleton (transpose,
symmetricPermutation, elimination fill) and the per-point small-block Cholesky to the reference itself. Not covered:
fillReducingPermutation (needs SuiteSparse / Eigen; the shim oracle/refshim/amd.h stands in)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cpu as oracle_cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_host.so")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (needs /root/reference)")


def _fn(lib, name):
    f = getattr(lib, name)
    f.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int64), C.c_int64]
    f.restype = C.c_int64
    return f


def _run(f, sid, params):
    p = (C.c_double * len(params))(*[float(x) for x in params])
    n = f(sid, p, len(params), None, 0)
    assert n >= 0, f"scenario {sid} failed"
    out = np.empty(max(n, 1), dtype=np.int64)
    assert f(sid, p, len(params), out.ctypes.data_as(C.POINTER(C.c_int64)), n) == n
    return out[:n]


SCENARIOS = [
    ("randomData<double>", 0, [[1000, -1, 1, 37], [4097, -2.5, 0.5, 38], [0, -1, 1, 1]]),
    ("randomData<float>", 1, [[1000, -1, 1, 37], [333, 0, 10, 5]]),
    ("randomVec", 2, [[500, 2, 5, 47], [64, 0, 0, 3]]),
    ("randomPermutation", 3, [[200, 9], [1, 1]]),
    ("randomPartition", 4, [[115, 1, 4, 57], [1000, 2, 9, 3]]),
    ("randomCols + columnsToCscStruct + transpose + csrStructToColumns", 5, [[115, 0.037, 57 + i] for i in range(5)] + [[400, 0.01, 3]]),
    ("genFlat (+ randomVec on the generator's engine, Bench.cpp:279-288)", 6, [[1000, 0.05, 37], [300, 0.2, 41]]),
    ("genGrid", 7, [[120, 120, 1.0, 2, 37], [36, 36, 1.0, 2, 37], [30, 17, 0.6, 3, 5]]),
    ("genMeridians", 8, [[6, 40, 0.5, 3, 5, 4, 4, 37], [3, 15, 0.8, 2, 4, 1, 2, 7]]),
    ("genFlat + addSchurSet", 9, [[40, 0.2, 600, 0.02, 37], [100, 0.1, 2000, 0.002, 41]]),
    ("genLine", 10, [[200, 0.3, 5, 37]]),
    ("symmetricPermutation / clear / inversePermutation / composePermutations", 11, [[115, 0.05, 57 + i, 99 + i, i % 2] for i in range(4)]),
    ("addIndependentEliminationFill / addFullEliminationFill / extractRightBottom", 12, [[115, 0.04, 57 + i, 10 * i, 60 + 10 * i] for i in range(4)]),
    ("makeIndependentElimSet / naiveAddEliminationEntries / joinColums", 13, [[115, 0.04, 57 + i, 5 * i, 50 + 5 * i] for i in range(3)]),
    ("MathUtils cholesky / solveUpperT / solveUpper, fp64 and fp32", 14, [[3, 3, 37], [6, 6, 38], [9, 12, 39], [12, 12, 40], [1, 1, 41]]),
    ("toOrderedPair / cumSumVec / bisect / rewindVec", 15, [[n, 37 + n, n // 2, 3] for n in (1, 2, 5, 6, 13, 40)]),
]


@pytest.mark.parametrize("name,sid,param_sets", SCENARIOS, ids=[s[0].split(" ")[0] for s in SCENARIOS])
def test_restatement_is_bit_identical_to_reference_objects(name, sid, param_sets):
    ref = _fn(C.CDLL(REF_LIB), "ref_hostcheck")
    ours = _fn(oracle_cpu.api().lib, "oracle_hostcheck")
    for params in param_sets:
        a, b = _run(ref, sid, params), _run(ours, sid, params)
        assert a.shape == b.shape and a.size > 0 or params[0] == 0, (name, params)
        assert np.array_equal(a, b), (name, params, int(np.argmax(a != b)) if a.shape == b.shape else "length")


def test_bench_inputs_come_out_of_the_reference_generators():
    """the exact generator calls behind BASELINE.json's configurations (bench.py WORKLOADS): FLAT 1000 / 0.05,
    GRID 120 x 120 conn 2, FLAT 2000 / 0.03, and the seeded values randomData(n, -1, 1, 37) / (.., 38) for data and rhs"""
    ref = _fn(C.CDLL(REF_LIB), "ref_hostcheck")
    ours = _fn(oracle_cpu.api().lib, "oracle_hostcheck")
    for sid, params in [(6, [1000, 0.05, 37]), (7, [120, 120, 1.0, 2, 37]), (6, [2000, 0.03, 37]),
                        (0, [200000, -1, 1, 37]), (0, [50000, -1, 1, 38])]:
        assert np.array_equal(_run(ref, sid, params), _run(ours, sid, params))
