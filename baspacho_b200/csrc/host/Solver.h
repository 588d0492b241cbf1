// Public solver API + host driver. Same surface as reference baspacho/baspacho/Solver.h:34-237
// (class Solver, BackendType, AddFillPolicy, Settings, createSolver), so callers of the reference
// compile against it. The driver sequences the fine-grained backend ops exactly like the reference
// (Solver.cpp:42-455) and, when the backend offers them, hands whole ranges to the fused entry points.
#pragma once

#include <functional>
#include <memory>
#include <unordered_set>
#include "CoalescedBlockMatrix.h"
#include "MatOps.h"
#include "SparseStructure.h"

namespace BaSpaCho {

class Solver {
 public:
  Solver(CoalescedBlockMatrixSkel&& factorSkel, std::vector<int64_t>&& sparseElimRanges,
         std::vector<int64_t>&& permutation, OpsPtr&& ops, int64_t canFactorUpTo = -1);

  PermutedCoalescedAccessor accessor() const {
    PermutedCoalescedAccessor a;
    a.init(factorSkel.accessor(), permutation.data());
    return a;
  }
  PermutedCoalescedAccessor deviceAccessor() const { return symCtx->deviceAccessor(); }

  void enableStats(bool enabled = true);
  void printStats() const;
  void resetStats();

  template <typename T> void factor(T* data, bool verbose = false) const;
  template <typename T> void solve(const T* matData, T* vecData, int64_t stride, int nRHS) const;
  template <typename T> void solveL(const T* matData, T* vecData, int64_t stride, int nRHS) const;
  template <typename T> void solveLt(const T* matData, T* vecData, int64_t stride, int nRHS) const;
  template <typename T> void factorUpTo(T* data, int64_t spanIndex, bool verbose = false) const;
  template <typename T> void solveLUpTo(const T* data, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const;
  template <typename T> void solveLtUpTo(const T* data, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const;
  template <typename T>
  void addMvFrom(const T* matData, int64_t spanIndex, const T* inVecData, int64_t inStride, T* outVecData,
                 int64_t outStride, int nRHS, BaseType<T> alpha = 1.0) const;
  template <typename T> void pseudoFactorFrom(T* data, int64_t spanIndex, bool verbose = false) const;
  template <typename T> void factorFrom(T* data, int64_t spanIndex, bool verbose = false) const;
  template <typename T> void solveLFrom(const T* data, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const;
  template <typename T> void solveLtFrom(const T* data, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const;

  int64_t order() const { return factorSkel.order(); }
  int64_t dataSize() const { return factorSkel.dataSize(); }
  int64_t canFactorUpToSpan() const { return canFactorUpTo; }
  int64_t spanVectorOffset(int64_t spanIndex) const { return factorSkel.spanVectorOffset(spanIndex); }
  int64_t spanMatrixOffset(int64_t spanIndex) const { return factorSkel.spanMatrixOffset(spanIndex); }
  const CoalescedBlockMatrixSkel& skel() const { return factorSkel; }
  const std::vector<int64_t>& sparseEliminationRanges() const { return sparseElimRanges; }
  const std::vector<int64_t>& paramToSpan() const { return permutation; }

  // testing hooks (same names as the reference)
  SymbolicCtx& internalSymbolicContext() { return *symCtx; }
  SymElimCtx& internalGetElimCtx(size_t i) {
    BASPACHO_CHECK_LT(i, elimCtxs.size());
    return *elimCtxs[i];
  }
  // (addition) run on this stream and force the per-op (unfused) path, for A/B tests
  void setStream(void* stream) { symCtx->setStream(stream); }
  void setUseFusedOps(bool on) { useFusedOps = on; }
  int64_t elimTempSize() const { return maxElimTempSize; }

 private:
  struct ColumnGeom;  // index arithmetic of one lump column (below-diagonal panel etc.)
  ColumnGeom columnGeom(int64_t lump) const;

  void initElimination();
  int64_t boardElimTempSize(int64_t lump, int64_t boardIndexInCol) const;
  template <typename T> void factorLump(NumericCtx<T>& numCtx, T* data, int64_t lump) const;
  template <typename T> void eliminateBoard(NumericCtx<T>& numCtx, T* data, int64_t ptr) const;
  template <typename T>
  void internalFactorRange(T* data, int64_t startSpanIndex, int64_t endSpanIndex, bool verbose = false) const;
  template <typename T>
  void internalSolveLRange(SolveCtx<T>& slvCtx, const T* data, int64_t startSpanIndex, int64_t endSpanIndex,
                           T* vecData, int64_t stride, int nRHS) const;
  template <typename T>
  void internalSolveLtRange(SolveCtx<T>& slvCtx, const T* data, int64_t startSpanIndex, int64_t endSpanIndex,
                            T* vecData, int64_t stride, int nRHS) const;
  void checkSpanRange(int64_t startSpanIndex, int64_t endSpanIndex) const;

  CoalescedBlockMatrixSkel factorSkel;
  std::vector<int64_t> sparseElimRanges;
  std::vector<int64_t> permutation;  // on indices: v'[p[i]] = v[i]
  int64_t canFactorUpTo;

  OpsPtr ops;
  SymbolicCtxPtr symCtx;
  std::vector<SymElimCtxPtr> elimCtxs;
  std::vector<int64_t> startElimRowPtr;
  int64_t maxElimTempSize = 0;
  bool useFusedOps = true;
};

using SolverPtr = std::unique_ptr<Solver>;

enum BackendType {
  BackendRef,   // naive CPU (oracle/ only)
  BackendFast,  // BLAS CPU (oracle/ only)
  BackendCuda,  // device backend: in this repo = the B200 backend
  BackendB200 = BackendCuda,
  BackendSymbolicOnly = 100,  // analysis only, numeric ops throw
};

enum AddFillPolicy {
  AddFillComplete,       // add fill for complete factoring, reorder
  AddFillForAutoElims,   // add fill for given+auto elim-ranges, reorder
  AddFillForGivenElims,  // fill for elimination of the given ranges only, no reorder
  AddFillNone,           // no fill added, no reorder
};

struct ComputationModel;

struct Settings {
  bool findSparseEliminationRanges = true;
  int numThreads = 16;
  // the reference defaults to BackendFast (Solver.h:203); this library has no CPU numeric path, so a default-constructed
  // Settings - `createSolver({}, ...)`, the common call of the reference's examples - selects the device backend
  BackendType backend = BackendCuda;
  AddFillPolicy addFillPolicy = AddFillComplete;
  const ComputationModel* computationModel = nullptr;
};

// CPU backends are not part of the product library; oracle/ registers them here when loaded.
using BackendFactory = std::function<OpsPtr(int numThreads)>;
void registerBackend(BackendType type, BackendFactory factory);
// backend selection used by createSolver (reference Solver.cpp:596-609)
OpsPtr getBackend(const Settings& settings);

SolverPtr createSolver(const Settings& settings, const std::vector<int64_t>& paramSizes, const SparseStructure& ss,
                       const std::vector<int64_t>& sparseElimRanges = {},
                       const std::unordered_set<int64_t>& elimLastIds = {});

}  // namespace BaSpaCho
