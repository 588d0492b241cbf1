// Backend operator interface: Ops -> SymbolicCtx -> SymElimCtx / NumericCtx<T> / SolveCtx<T>.
// THE DROP-IN BOUNDARY: class names, virtual signatures, argument meaning and error behaviour are
// those of reference baspacho/baspacho/MatOps.h (Ops :48-54, SymbolicCtx :65-102, SymElimCtx :105-109,
// NumericCtx :112-136, SolveCtx :139-184, factories :215-221), so a backend written against the
// reference header compiles against this one. Additions (all optional, defaulted): the fused
// whole-range entry points (following the reference's own precedent of optional "fragmented"
// overrides, MatOps.h:168-183) and an explicit stream setter.
#pragma once

#include <cxxabi.h>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <vector>
#include "CoalescedBlockMatrix.h"
#include "Utils.h"

namespace BaSpaCho {

struct Ops;
struct SymbolicCtx;
struct SymElimCtx;
template <typename T> struct NumericCtx;
template <typename T> struct SolveCtx;
using OpsPtr = std::unique_ptr<Ops>;
using SymbolicCtxPtr = std::unique_ptr<SymbolicCtx>;
using SymElimCtxPtr = std::unique_ptr<SymElimCtx>;
template <typename T> using NumericCtxPtr = std::unique_ptr<NumericCtx<T>>;
template <typename T> using SolveCtxPtr = std::unique_ptr<SolveCtx<T>>;

// scalar T -> one matrix; std::vector<T*> -> a batch of identically structured matrices
template <typename T>
struct Batch {
  using BaseType = T;
  static int getSize(const T*) { return 1; }
};
template <typename T>
struct Batch<std::vector<T*>> {
  using BaseType = T;
  static int getSize(const std::vector<T*>* data) { return (int)data->size(); }
};
template <typename T> using BaseType = typename Batch<T>::BaseType;

struct Ops {
  virtual ~Ops() {}
  virtual SymbolicCtxPtr createSymbolicCtx(const CoalescedBlockMatrixSkel& skel,
                                           const std::vector<int64_t>& permutation) = 0;
};

struct NumericCtxBase { virtual ~NumericCtxBase() {} };
struct SolveCtxBase { virtual ~SolveCtxBase() {} };

struct SymbolicCtx {
  virtual ~SymbolicCtx() {}

  virtual SymElimCtxPtr prepareElimination(int64_t lumpsBegin, int64_t lumpsEnd) = 0;
  virtual NumericCtxBase* createNumericCtxForType(std::type_index tIdx, int64_t tempBufSize, int batchSize) = 0;
  virtual SolveCtxBase* createSolveCtxForType(std::type_index tIdx, int nRHS, int batchSize) = 0;
  virtual PermutedCoalescedAccessor deviceAccessor() = 0;

  // (addition) device backends: run every subsequent op on this stream (a cudaStream_t)
  virtual void setStream(void* /*stream*/) {}

  template <typename T> NumericCtxPtr<T> createNumericCtx(int64_t tempBufSize, const T* data);
  template <typename T> SolveCtxPtr<T> createSolveCtx(int nRHS, const T* data);

  mutable OpStat<int, int> potrfStat;
  mutable int64_t potrfBiggestN = 0;
  mutable OpStat<int, int, int> trsmStat;
  mutable OpStat<int, int, int, int> sygeStat;
  mutable int64_t gemmCalls = 0;
  mutable int64_t syrkCalls = 0;
  mutable OpStat<int, int, int> asmblStat;

  mutable OpStat<> solveSparseLStat;
  mutable OpStat<> solveSparseLtStat;
  mutable OpStat<> pseudoFactorStat;
  mutable OpStat<> symmStat;
  mutable OpStat<> solveLStat;
  mutable OpStat<> solveLtStat;
  mutable OpStat<> solveGemvStat;
  mutable OpStat<> solveGemvTStat;
  mutable OpStat<> solveAssVStat;
  mutable OpStat<> solveAssVTStat;
};

struct SymElimCtx {
  virtual ~SymElimCtx() {}
  mutable OpStat<> elimStat;
};

template <typename T>
struct NumericCtx : NumericCtxBase {
  virtual ~NumericCtx() {}

  // per span: factor the span's diagonal block and solve the rows below it (block-Jacobi style)
  virtual void pseudoFactorSpans(T* data, int64_t spanBegin, int64_t spanEnd) = 0;

  // sparse ("Schur") elimination of the independent lumps [lumpsBegin, lumpsEnd)
  virtual void doElimination(const SymElimCtx& elimData, T* data, int64_t lumpsBegin, int64_t lumpsEnd) = 0;

  // in-place Cholesky of the row-major n x n block at offA (lower triangle)
  virtual void potrf(int64_t n, T* data, int64_t offA) = 0;

  // X * tril(A)^T = B, in place on the k x n row-major panel at offB
  virtual void trsm(int64_t n, int64_t k, T* data, int64_t offA, int64_t offB) = 0;

  // temp(n x m) = B(n x k) * A(m x k)^T with A = first m rows of B (upper part of A*A^T is don't-care)
  virtual void saveSyrkGemm(int64_t m, int64_t n, int64_t k, const T* data, int64_t offset) = 0;

  virtual void prepareAssemble(int64_t targetLump) = 0;

  virtual void assemble(T* data, int64_t rectRowBegin, int64_t dstStride, int64_t srcColDataOffset,
                        int64_t srcRectWidth, int64_t numBlockRows, int64_t numBlockCols) = 0;

  // (addition) whole-range factorization in one backend call: eliminate the sparse ranges and
  // factor the dense lumps with sources in [startLump, ...) and targets < upToLump, exactly the
  // work Solver::internalFactorRange sequences through the fine-grained ops above.
  virtual bool hasFusedFactor() { return false; }
  virtual void fusedFactorRange(T* /*data*/, int64_t /*startLump*/, int64_t /*upToLump*/) {
    throw std::runtime_error("fusedFactorRange: not supported");
  }
};

template <typename T>
struct SolveCtx : SolveCtxBase {
  virtual ~SolveCtx() {}

  virtual void sparseElimSolveL(const SymElimCtx& elimData, const T* data, int64_t lumpsBegin,
                                int64_t lumpsEnd, T* C, int64_t ldc) = 0;
  virtual void sparseElimSolveLt(const SymElimCtx& elimData, const T* data, int64_t lumpsBegin,
                                 int64_t lumpsEnd, T* C, int64_t ldc) = 0;
  virtual void symm(const T* data, int64_t offset, int64_t n, const T* C, int64_t offC, int64_t ldc,
                    T* D, int64_t ldd, BaseType<T> alpha) = 0;
  virtual void solveL(const T* data, int64_t offset, int64_t n, T* C, int64_t offC, int64_t ldc) = 0;
  virtual void gemv(const T* data, int64_t offset, int64_t nRows, int64_t nCols, const T* A,
                    int64_t offA, int64_t lda, BaseType<T> alpha) = 0;
  virtual void assembleVec(int64_t chainColPtr, int64_t numColItems, T* C, int64_t ldc) = 0;
  virtual void solveLt(const T* data, int64_t offset, int64_t n, T* C, int64_t offC, int64_t ldc) = 0;
  virtual void gemvT(const T* data, int64_t offset, int64_t nRows, int64_t nCols, T* A, int64_t offA,
                     int64_t lda, BaseType<T> alpha) = 0;
  virtual void assembleVecT(const T* C, int64_t ldc, int64_t chainColPtr, int64_t numColItems) = 0;

  // optional (B200 backend, round 2): out += alpha * A * in restricted to the COLUMNS of one sparse-elimination range -
  // diagonal blocks, the blocks below them and their transposes - in two launches instead of five per lump. The
  // reference's addMvFrom walks every lump ("sparse ops not supported yet", Solver.cpp:408-446).
  virtual bool hasSparseElimMV() { return false; }
  virtual void sparseElimMV(const SymElimCtx&, const T*, const T*, int64_t, T*, int64_t, BaseType<T>) {
    throw std::runtime_error("sparseElimMV: not supported");
  }

  virtual bool hasFragmentedOps() { return false; }
  virtual void fragmentedMV(const T*, const T*, int64_t, int64_t, T*, BaseType<T>) {
    throw std::runtime_error("fragmentedMV: not supported");
  }
  virtual void fragmentedSolveL(const T*, int64_t, int64_t, T*) {
    throw std::runtime_error("fragmentedSolveL: not supported");
  }
  virtual void fragmentedSolveLt(const T*, int64_t, int64_t, T*) {
    throw std::runtime_error("fragmentedSolveLt: not supported");
  }

  // (addition) whole-range triangular solves over lumps [startLump, upToLump)
  virtual bool hasFusedSolve() { return false; }
  virtual void fusedSolveL(const T*, int64_t, int64_t, T*, int64_t) {
    throw std::runtime_error("fusedSolveL: not supported");
  }
  virtual void fusedSolveLt(const T*, int64_t, int64_t, T*, int64_t) {
    throw std::runtime_error("fusedSolveLt: not supported");
  }
};

template <typename T>
std::string prettyTypeName(const T& t) {
  char* s = abi::__cxa_demangle(typeid(t).name(), nullptr, nullptr, nullptr);
  std::string out(s ? s : typeid(t).name());
  free(s);
  return out;
}

template <typename T>
NumericCtxPtr<T> SymbolicCtx::createNumericCtx(int64_t tempBufSize, const T* data) {
  NumericCtxBase* ctx = createNumericCtxForType(std::type_index(typeid(T)), tempBufSize, Batch<T>::getSize(data));
  auto* typed = dynamic_cast<NumericCtx<T>*>(ctx);
  if (!typed) delete ctx;
  BASPACHO_CHECK_NOTNULL(typed);
  return NumericCtxPtr<T>(typed);
}

template <typename T>
SolveCtxPtr<T> SymbolicCtx::createSolveCtx(int nRHS, const T* data) {
  SolveCtxBase* ctx = createSolveCtxForType(std::type_index(typeid(T)), nRHS, Batch<T>::getSize(data));
  auto* typed = dynamic_cast<SolveCtx<T>*>(ctx);
  if (!typed) delete ctx;
  BASPACHO_CHECK_NOTNULL(typed);
  return SolveCtxPtr<T>(typed);
}

// Backend factories. The product library provides b200Ops() (hand-written sm_100a kernels) and
// symbolicOnlyOps() (analysis without a device; every numeric op throws). The CPU backends of the
// reference (simpleOps / fastOps, MatOps.h:215-217) live ONLY in oracle/ as the parity checker.
OpsPtr symbolicOnlyOps();
OpsPtr b200Ops();
inline OpsPtr cudaOps() { return b200Ops(); }  // reference name for the device backend (MatOps.h:219-221)

}  // namespace BaSpaCho
