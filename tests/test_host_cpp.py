"""Native tests of the host layer: tests/cpp/host_tests.cpp restates the reference's host-side gtest cases
(SparseStructureTest, EliminationTreeTest, AccessorTest, CoalescedBlockMatrixTest Densify/Damp, CreateSolverTest) in plain
C++ against csrc/host - golden vectors bit for bit, fills against the naive fill, createSolver with every fill policy -
and the PartialFactorSolveTest cases (partial / split factor, pseudo factor, addMv, partial solves) for the CPU checker's
backends against dense algebra, which pins the oracle on exactly the entry points the GPU partial tests check against it."""
import os
import subprocess

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")


def test_native_host_tests():
    subprocess.check_call(["make", "-s", "-C", HERE])
    env = dict(os.environ)
    from oracle import cpu as ocpu
    blas = ocpu.blas_path()  # the BLAS-backed CPU checker (BackendFast) is exercised too when an OpenBLAS is around
    if blas:
        env["ORACLE_BLAS_PATH"] = blas
        env["ORACLE_BLAS_PREFIX"] = "scipy_" if "scipy_openblas" in blas else ""
    out = subprocess.run([os.path.join(HERE, "host_tests"), "4"], capture_output=True, text=True, env=env)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " 0 failures" in out.stdout
    for case in ("SparseStructure.Transpose", "SparseStructure.SymPermutation", "EliminationTree.Build",
                 "CoalescedBlockMatrix.Densify+Densify2+Damp", "CreateSolver.ElimLast_double", "Partial.*_Ref_float"):
        assert f"ok   {case}" in out.stdout
