// Host driver of the left-looking supernodal factorization and of the triangular solves, and
// createSolver() (symbolic pipeline). Behaviour follows reference baspacho/baspacho/Solver.cpp:
// ctor :24-40, factorLump :42-64, eliminateBoard :66-98, initElimination :117-147,
// internalFactorRange :164-219, internalSolveLRange :268-330, internalSolveLtRange :332-397,
// addMvFrom :399-449, pseudoFactorFrom :451-455, createSolver :611-752. Written from scratch.
#include "Solver.h"
#include <iostream>
#include <map>
#include <numeric>
#include "ComputationModel.h"
#include "DebugMacros.h"
#include "EliminationTree.h"
#include "Utils.h"

namespace BaSpaCho {

using std::vector;

// Geometry of one lump column, derived from the skeleton: where the diagonal block and the
// below-diagonal panel live and how many rows the panel has.
struct Solver::ColumnGeom {
  int64_t start, size;          // first scalar column / width of the lump
  int64_t chainBegin;           // first chain of the column
  int64_t diagOffset;           // data offset of the diagonal block
  int64_t belowChainOrd;        // ordinal of the first chain below the diagonal lump
  int64_t numChains;            // chains in the column
  int64_t belowOffset;          // data offset of the below-diagonal panel
  int64_t rowsBelow;            // rows of the below-diagonal panel
};

Solver::ColumnGeom Solver::columnGeom(int64_t lump) const {
  const auto& sk = factorSkel;
  ColumnGeom g;
  g.start = sk.lumpStart[lump];
  g.size = sk.lumpStart[lump + 1] - g.start;
  g.chainBegin = sk.chainColPtr[lump];
  g.diagOffset = sk.chainData[g.chainBegin];
  int64_t bBegin = sk.boardColPtr[lump], bEnd = sk.boardColPtr[lump + 1];
  g.belowChainOrd = sk.boardChainColOrd[bBegin + 1];  // second board (or the sentinel) starts below the diagonal lump
  g.numChains = sk.boardChainColOrd[bEnd - 1];
  g.belowOffset = sk.chainData[g.chainBegin + g.belowChainOrd];
  g.rowsBelow = sk.chainRowsTillEnd[g.chainBegin + g.numChains - 1] -
                sk.chainRowsTillEnd[g.chainBegin + g.belowChainOrd - 1];
  return g;
}

Solver::Solver(CoalescedBlockMatrixSkel&& factorSkel_, vector<int64_t>&& sparseElimRanges_,
               vector<int64_t>&& permutation_, OpsPtr&& ops_, int64_t canFactorUpTo_)
    : factorSkel(std::move(factorSkel_)),
      sparseElimRanges(std::move(sparseElimRanges_)),
      permutation(std::move(permutation_)),
      canFactorUpTo(canFactorUpTo_ < 0 ? factorSkel.numSpans() : canFactorUpTo_),
      ops(std::move(ops_)) {
  symCtx = ops->createSymbolicCtx(factorSkel, permutation);
  for (size_t r = 0; r + 1 < sparseElimRanges.size(); r++)
    elimCtxs.push_back(symCtx->prepareElimination(sparseElimRanges[r], sparseElimRanges[r + 1]));
  initElimination();
}

template <typename T>
void Solver::factorLump(NumericCtx<T>& numCtx, T* data, int64_t lump) const {
  ColumnGeom g = columnGeom(lump);
  numCtx.potrf(g.size, data, g.diagOffset);
  if (g.rowsBelow > 0) numCtx.trsm(g.size, g.rowsBelow, data, g.diagOffset, g.belowOffset);
}

namespace {
// rows of a source column from board `boardOrd` on: (first chain ordinal, chains in board end, chains in column,
// first row, rows in board, rows to the end)
struct BoardRows {
  int64_t chainOrd, chainOrdBoardEnd, chainOrdColEnd, rowBegin, rowsInBoard, rowsToEnd;
};
BoardRows boardRows(const CoalescedBlockMatrixSkel& sk, int64_t lump, int64_t boardOrd) {
  int64_t cBegin = sk.chainColPtr[lump];
  int64_t bBegin = sk.boardColPtr[lump], bEnd = sk.boardColPtr[lump + 1];
  BoardRows b;
  b.chainOrd = sk.boardChainColOrd[bBegin + boardOrd];
  b.chainOrdBoardEnd = sk.boardChainColOrd[bBegin + boardOrd + 1];
  b.chainOrdColEnd = sk.boardChainColOrd[bEnd - 1];
  b.rowBegin = sk.chainRowsTillEnd[cBegin + b.chainOrd - 1];
  b.rowsInBoard = sk.chainRowsTillEnd[cBegin + b.chainOrdBoardEnd - 1] - b.rowBegin;
  b.rowsToEnd = sk.chainRowsTillEnd[cBegin + b.chainOrdColEnd - 1] - b.rowBegin;
  return b;
}
}  // namespace

template <typename T>
void Solver::eliminateBoard(NumericCtx<T>& numCtx, T* data, int64_t ptr) const {
  const auto& sk = factorSkel;
  int64_t srcLump = sk.boardColLump[ptr], boardOrd = sk.boardColOrd[ptr];
  int64_t srcSize = sk.lumpStart[srcLump + 1] - sk.lumpStart[srcLump];
  int64_t cBegin = sk.chainColPtr[srcLump];
  BoardRows b = boardRows(sk, srcLump, boardOrd);

  // temp(rowsToEnd x rowsInBoard) = panel * (its first rowsInBoard rows)^T
  numCtx.saveSyrkGemm(b.rowsInBoard, b.rowsToEnd, srcSize, data, sk.chainData[cBegin + b.chainOrd]);

  int64_t targetLump = sk.boardRowLump[sk.boardColPtr[srcLump] + boardOrd];
  int64_t targetSize = sk.lumpStart[targetLump + 1] - sk.lumpStart[targetLump];
  numCtx.assemble(data, b.rowBegin, targetSize, cBegin + b.chainOrd, b.rowsInBoard,
                  b.chainOrdColEnd - b.chainOrd, b.chainOrdBoardEnd - b.chainOrd);
}

int64_t Solver::boardElimTempSize(int64_t lump, int64_t boardIndexInCol) const {
  BoardRows b = boardRows(factorSkel, lump, boardIndexInCol);
  return b.rowsInBoard * b.rowsToEnd;
}

void Solver::initElimination() {
  const auto& sk = factorSkel;
  const int64_t nLumps = sk.numLumps();
  const int64_t denseFrom = sparseElimRanges.empty() ? 0 : sparseElimRanges.back();
  startElimRowPtr.assign(nLumps - denseFrom, 0);
  maxElimTempSize = 0;
  for (int64_t l = denseFrom; l < nLumps; l++) {
    // boards of row-lump l, by increasing source column; skip sources handled by sparse elimination
    int64_t r = sk.boardRowPtr[l], rEnd = sk.boardRowPtr[l + 1];
    BASPACHO_CHECK_EQ(sk.boardColLump[rEnd - 1], l);
    while (sk.boardColLump[r] < denseFrom) r++;
    BASPACHO_CHECK_LT(r, rEnd);
    startElimRowPtr[l - denseFrom] = r;
    for (; r < rEnd && sk.boardColLump[r] < l; r++) {
      int64_t src = sk.boardColLump[r], ord = sk.boardColOrd[r];
      BASPACHO_CHECK_LT(ord, sk.boardColPtr[src + 1] - sk.boardColPtr[src]);
      BASPACHO_CHECK_EQ(l, sk.boardRowLump[sk.boardColPtr[src] + ord]);
      maxElimTempSize = std::max(maxElimTempSize, boardElimTempSize(src, ord));
    }
  }
}

void Solver::checkSpanRange(int64_t startSpanIndex, int64_t endSpanIndex) const {
  BASPACHO_CHECK_GE(startSpanIndex, 0);
  BASPACHO_CHECK_LE(startSpanIndex, endSpanIndex);
  BASPACHO_CHECK_LT(endSpanIndex, (int64_t)factorSkel.spanOffsetInLump.size());
  BASPACHO_CHECK_EQ(factorSkel.spanOffsetInLump[startSpanIndex], 0);
  BASPACHO_CHECK_EQ(factorSkel.spanOffsetInLump[endSpanIndex], 0);
}

template <typename T>
void Solver::factor(T* data, bool verbose) const {
  factorUpTo(data, factorSkel.numSpans(), verbose);
}
template <typename T>
void Solver::factorUpTo(T* data, int64_t spanIndex, bool verbose) const {
  internalFactorRange(data, 0, spanIndex, verbose);
}
template <typename T>
void Solver::factorFrom(T* data, int64_t spanIndex, bool verbose) const {
  internalFactorRange(data, spanIndex, factorSkel.numSpans(), verbose);
}

template <typename T>
void Solver::internalFactorRange(T* data, int64_t startSpanIndex, int64_t endSpanIndex, bool verbose) const {
  checkSpanRange(startSpanIndex, endSpanIndex);
  BASPACHO_CHECK_LE(endSpanIndex, canFactorUpTo);
  const auto& sk = factorSkel;
  const int64_t startLump = sk.spanToLump[startSpanIndex], upToLump = sk.spanToLump[endSpanIndex];

  NumericCtxPtr<T> numCtx = symCtx->createNumericCtx<T>(maxElimTempSize, data);

  // validate how the requested range cuts the sparse-elimination ranges (same rules on both paths)
  for (size_t r = 0; r + 1 < sparseElimRanges.size(); r++) {
    if (sparseElimRanges[r + 1] > upToLump) {
      BASPACHO_CHECK_EQ(sparseElimRanges[r], upToLump);
      break;
    }
    if (startLump > sparseElimRanges[r]) BASPACHO_CHECK_GE(startLump, sparseElimRanges[r + 1]);
  }

  if (useFusedOps && numCtx->hasFusedFactor()) {
    numCtx->fusedFactorRange(data, startLump, upToLump);
    return;
  }

  for (size_t r = 0; r + 1 < sparseElimRanges.size(); r++) {
    if (sparseElimRanges[r + 1] > upToLump) return;
    if (startLump > sparseElimRanges[r]) continue;
    if (verbose)
      std::cout << "Elim set: " << r << " (" << sparseElimRanges[r] << ".." << sparseElimRanges[r + 1] << ")" << std::endl;
    numCtx->doElimination(*elimCtxs[r], data, sparseElimRanges[r], sparseElimRanges[r + 1]);
  }

  const int64_t denseFrom = sparseElimRanges.empty() ? 0 : sparseElimRanges.back();
  if (verbose) std::cout << "Block-Fact from: " << denseFrom << std::endl;

  for (int64_t l = std::max(startLump, denseFrom); l < sk.numLumps(); l++) {
    numCtx->prepareAssemble(l);
    // contributions of every already factored source column with a board in row-lump l
    for (int64_t r = startElimRowPtr[l - denseFrom], rEnd = sk.boardRowPtr[l + 1] - 1; r < rEnd; r++) {
      int64_t src = sk.boardColLump[r];
      if (src >= upToLump) break;
      if (src < startLump) continue;
      eliminateBoard(*numCtx, data, r);
    }
    if (l < upToLump) factorLump(*numCtx, data, l);
  }
}

template <typename T>
void Solver::solve(const T* matData, T* vecData, int64_t stride, int nRHS) const {
  SolveCtxPtr<T> slvCtx = symCtx->createSolveCtx<T>(nRHS, matData);
  internalSolveLRange(*slvCtx, matData, 0, factorSkel.numSpans(), vecData, stride, nRHS);
  internalSolveLtRange(*slvCtx, matData, 0, factorSkel.numSpans(), vecData, stride, nRHS);
}
template <typename T>
void Solver::solveL(const T* matData, T* vecData, int64_t stride, int nRHS) const {
  solveLUpTo(matData, factorSkel.numSpans(), vecData, stride, nRHS);
}
template <typename T>
void Solver::solveLt(const T* matData, T* vecData, int64_t stride, int nRHS) const {
  solveLtUpTo(matData, factorSkel.numSpans(), vecData, stride, nRHS);
}
template <typename T>
void Solver::solveLUpTo(const T* matData, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const {
  SolveCtxPtr<T> slvCtx = symCtx->createSolveCtx<T>(nRHS, matData);
  internalSolveLRange(*slvCtx, matData, 0, spanIndex, vecData, stride, nRHS);
}
template <typename T>
void Solver::solveLtUpTo(const T* matData, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const {
  SolveCtxPtr<T> slvCtx = symCtx->createSolveCtx<T>(nRHS, matData);
  internalSolveLtRange(*slvCtx, matData, 0, spanIndex, vecData, stride, nRHS);
}
template <typename T>
void Solver::solveLFrom(const T* matData, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const {
  SolveCtxPtr<T> slvCtx = symCtx->createSolveCtx<T>(nRHS, matData);
  internalSolveLRange(*slvCtx, matData, spanIndex, factorSkel.numSpans(), vecData, stride, nRHS);
}
template <typename T>
void Solver::solveLtFrom(const T* matData, int64_t spanIndex, T* vecData, int64_t stride, int nRHS) const {
  SolveCtxPtr<T> slvCtx = symCtx->createSolveCtx<T>(nRHS, matData);
  internalSolveLtRange(*slvCtx, matData, spanIndex, factorSkel.numSpans(), vecData, stride, nRHS);
}

template <typename T>
void Solver::internalSolveLRange(SolveCtx<T>& slvCtx, const T* matData, int64_t startSpanIndex, int64_t endSpanIndex,
                                 T* vecData, int64_t stride, int nRHS) const {
  checkSpanRange(startSpanIndex, endSpanIndex);
  const auto& sk = factorSkel;
  const int64_t startLump = sk.spanToLump[startSpanIndex], upToLump = sk.spanToLump[endSpanIndex];

  for (size_t r = 0; r + 1 < sparseElimRanges.size(); r++) {
    if (sparseElimRanges[r + 1] > upToLump) {
      BASPACHO_CHECK_EQ(sparseElimRanges[r], upToLump);
      break;
    }
    if (startLump > sparseElimRanges[r]) BASPACHO_CHECK_GE(startLump, sparseElimRanges[r + 1]);
  }
  // a fragmented skeleton (every lump a single span) with one right-hand side goes to the backend's fragmented ops, as
  // in the reference (Solver.cpp:299-301), also when the backend offers a fused range solve: one launch per lump is the
  // worst case for the latter
  const bool fragL = sk.numSpans() == sk.numLumps() && nRHS == 1 && slvCtx.hasFragmentedOps();
  if (useFusedOps && slvCtx.hasFusedSolve() && !fragL) {
    slvCtx.fusedSolveL(matData, startLump, upToLump, vecData, stride);
    return;
  }

  for (size_t r = 0; r + 1 < sparseElimRanges.size(); r++) {
    if (sparseElimRanges[r + 1] > upToLump) return;
    if (startLump > sparseElimRanges[r]) continue;
    slvCtx.sparseElimSolveL(*elimCtxs[r], matData, sparseElimRanges[r], sparseElimRanges[r + 1], vecData, stride);
  }
  const int64_t denseFrom = std::max(startLump, sparseElimRanges.empty() ? int64_t(0) : sparseElimRanges.back());

  if (sk.numSpans() == sk.numLumps() && slvCtx.hasFragmentedOps() && nRHS == 1) {
    BASPACHO_CHECK_EQ(sk.lumpToSpan[denseFrom], denseFrom);
    slvCtx.fragmentedSolveL(matData, denseFrom, upToLump, vecData);
    return;
  }
  for (int64_t l = denseFrom; l < upToLump; l++) {
    ColumnGeom g = columnGeom(l);
    slvCtx.solveL(matData, g.diagOffset, g.size, vecData, g.start, stride);
    if (g.rowsBelow == 0) continue;
    slvCtx.gemv(matData, g.belowOffset, g.rowsBelow, g.size, vecData, g.start, stride, BaseType<T>(-1.0));
    slvCtx.assembleVec(g.chainBegin + g.belowChainOrd, g.numChains - g.belowChainOrd, vecData, stride);
  }
}

template <typename T>
void Solver::internalSolveLtRange(SolveCtx<T>& slvCtx, const T* matData, int64_t startSpanIndex, int64_t endSpanIndex,
                                  T* vecData, int64_t stride, int nRHS) const {
  checkSpanRange(startSpanIndex, endSpanIndex);
  const auto& sk = factorSkel;
  const int64_t startLump = sk.spanToLump[startSpanIndex], upToLump = sk.spanToLump[endSpanIndex];

  const bool fragLt = sk.numSpans() == sk.numLumps() && nRHS == 1 && slvCtx.hasFragmentedOps();
  if (useFusedOps && slvCtx.hasFusedSolve() && !fragLt) {
    for (int64_t r = (int64_t)sparseElimRanges.size() - 2; r >= 0; r--) {
      if (sparseElimRanges[r + 1] > upToLump) {
        BASPACHO_CHECK_LE(sparseElimRanges[r], upToLump);
        continue;
      }
      if (sparseElimRanges[r] < startLump) {
        BASPACHO_CHECK_GE(startLump, sparseElimRanges[r + 1]);
        break;
      }
    }
    slvCtx.fusedSolveLt(matData, startLump, upToLump, vecData, stride);
    return;
  }

  const int64_t denseFrom = std::max(startLump, sparseElimRanges.empty() ? int64_t(0) : sparseElimRanges.back());
  int64_t spansInRange = sk.lumpToSpan[std::max(upToLump, denseFrom)] - sk.lumpToSpan[denseFrom];
  if (spansInRange == upToLump - denseFrom && slvCtx.hasFragmentedOps() && nRHS == 1) {
    BASPACHO_CHECK_EQ(sk.lumpToSpan[denseFrom], denseFrom);
    slvCtx.fragmentedSolveLt(matData, denseFrom, upToLump, vecData);
  } else {
    for (int64_t l = upToLump - 1; l >= denseFrom; l--) {
      ColumnGeom g = columnGeom(l);
      if (g.rowsBelow > 0) {
        slvCtx.assembleVecT(vecData, stride, g.chainBegin + g.belowChainOrd, g.numChains - g.belowChainOrd);
        slvCtx.gemvT(matData, g.belowOffset, g.rowsBelow, g.size, vecData, g.start, stride, BaseType<T>(-1.0));
      }
      slvCtx.solveLt(matData, g.diagOffset, g.size, vecData, g.start, stride);
    }
  }

  for (int64_t r = (int64_t)sparseElimRanges.size() - 2; r >= 0; r--) {
    if (sparseElimRanges[r + 1] > upToLump) {
      BASPACHO_CHECK_LE(sparseElimRanges[r], upToLump);
      continue;
    }
    if (sparseElimRanges[r] < startLump) {
      BASPACHO_CHECK_GE(startLump, sparseElimRanges[r + 1]);
      return;
    }
    slvCtx.sparseElimSolveLt(*elimCtxs[r], matData, sparseElimRanges[r], sparseElimRanges[r + 1], vecData, stride);
  }
}

template <typename T>
void Solver::addMvFrom(const T* matData, int64_t spanIndex, const T* inVecData, int64_t inStride, T* outVecData,
                       int64_t outStride, int nRHS, BaseType<T> alpha) const {
  SolveCtxPtr<T> slvCtx = symCtx->createSolveCtx<T>(nRHS, matData);
  const auto& sk = factorSkel;
  BASPACHO_CHECK_GE(spanIndex, 0);
  BASPACHO_CHECK_LT(spanIndex, (int64_t)sk.spanOffsetInLump.size());
  BASPACHO_CHECK_EQ(sk.spanOffsetInLump[spanIndex], 0);
  const int64_t fromLump = sk.spanToLump[spanIndex], nLumps = sk.numLumps();

  if (sk.lumpToSpan[nLumps] - sk.lumpToSpan[fromLump] == nLumps - fromLump && slvCtx->hasFragmentedOps() && nRHS == 1) {
    BASPACHO_CHECK_EQ(sk.lumpToSpan[fromLump], fromLump);
    slvCtx->fragmentedMV(matData, inVecData, fromLump, nLumps, outVecData, alpha);
    return;
  }
  // whole sparse-elimination ranges at or behind the start: one backend call each when the backend offers it
  int64_t skipBegin = -1, skipEnd = -1;  // the ranges are contiguous from lump sparseElimRanges[0] on: one skipped interval
  if (slvCtx->hasSparseElimMV())
    for (size_t r = 0; r + 1 < sparseElimRanges.size(); r++) {
      if (sparseElimRanges[r] < fromLump) continue;
      slvCtx->sparseElimMV(*elimCtxs[r], matData, inVecData, inStride, outVecData, outStride, alpha);
      if (skipBegin < 0) skipBegin = sparseElimRanges[r];
      skipEnd = sparseElimRanges[r + 1];
    }
  for (int64_t l = fromLump; l < nLumps; l++) {
    if (l >= skipBegin && l < skipEnd) continue;
    ColumnGeom g = columnGeom(l);
    slvCtx->symm(matData, g.diagOffset, g.size, inVecData, g.start, inStride, outVecData, outStride, alpha);
    if (g.rowsBelow == 0) continue;
    int64_t firstChain = g.chainBegin + g.belowChainOrd, nChains = g.numChains - g.belowChainOrd;
    // out[rows below] += alpha * L21 * in[lump]
    slvCtx->gemv(matData, g.belowOffset, g.rowsBelow, g.size, inVecData, g.start, inStride, alpha);
    slvCtx->assembleVec(firstChain, nChains, outVecData, outStride);
    // out[lump] += alpha * L21^T * in[rows below]
    slvCtx->assembleVecT(inVecData, inStride, firstChain, nChains);
    slvCtx->gemvT(matData, g.belowOffset, g.rowsBelow, g.size, outVecData, g.start, outStride, alpha);
  }
}

template <typename T>
void Solver::pseudoFactorFrom(T* data, int64_t spanIndex, bool /*verbose*/) const {
  NumericCtxPtr<T> numCtx = symCtx->createNumericCtx<T>(maxElimTempSize, data);
  numCtx->pseudoFactorSpans(data, spanIndex, factorSkel.numSpans());
}

#define BSP_INSTANTIATE(T)                                                                                   \
  template void Solver::factor<T>(T*, bool) const;                                                           \
  template void Solver::factorUpTo<T>(T*, int64_t, bool) const;                                              \
  template void Solver::factorFrom<T>(T*, int64_t, bool) const;                                              \
  template void Solver::pseudoFactorFrom<T>(T*, int64_t, bool) const;                                        \
  template void Solver::solve<T>(const T*, T*, int64_t, int) const;                                          \
  template void Solver::solveL<T>(const T*, T*, int64_t, int) const;                                         \
  template void Solver::solveLt<T>(const T*, T*, int64_t, int) const;                                        \
  template void Solver::solveLUpTo<T>(const T*, int64_t, T*, int64_t, int) const;                            \
  template void Solver::solveLtUpTo<T>(const T*, int64_t, T*, int64_t, int) const;                           \
  template void Solver::solveLFrom<T>(const T*, int64_t, T*, int64_t, int) const;                            \
  template void Solver::solveLtFrom<T>(const T*, int64_t, T*, int64_t, int) const;                           \
  template void Solver::addMvFrom<T>(const T*, int64_t, const T*, int64_t, T*, int64_t, int, BaseType<T>) const;
BSP_INSTANTIATE(double)
BSP_INSTANTIATE(float)
BSP_INSTANTIATE(vector<double*>)
BSP_INSTANTIATE(vector<float*>)
#undef BSP_INSTANTIATE

void Solver::printStats() const {
  using std::cout;
  using std::endl;
  cout << "Matrix stats:\n  data size......: " << factorSkel.dataSize() << "\n  solve temp data: " << maxElimTempSize << endl;
  if (sparseElimRanges.size() >= 2) cout << "Sparse elimination sets:" << endl;
  for (size_t r = 0; r + 1 < sparseElimRanges.size(); r++)
    cout << "  elim set [" << sparseElimRanges[r] << ".." << sparseElimRanges[r + 1] << "]: " << elimCtxs[r]->elimStat.toString() << endl;
  cout << "Factor timings and call stats:\n  largest node size: " << symCtx->potrfBiggestN
       << "\n  potrf: " << symCtx->potrfStat.toString() << "\n  trsm: " << symCtx->trsmStat.toString()
       << "\n  syrk/gemm(" << symCtx->syrkCalls << "+" << symCtx->gemmCalls << "): " << symCtx->sygeStat.toString()
       << "\n  asmbl: " << symCtx->asmblStat.toString() << endl;
  if (symCtx->solveSparseLStat.numRuns + symCtx->solveSparseLtStat.numRuns + symCtx->solveLStat.numRuns +
          symCtx->solveLtStat.numRuns > 0) {
    cout << "Solve timings and call stats:\n  solveSparseLStat: " << symCtx->solveSparseLStat.toString()
         << "\n  solveSparseLtStat: " << symCtx->solveSparseLtStat.toString()
         << "\n  solveLStat: " << symCtx->solveLStat.toString() << "\n  solveLtStat: " << symCtx->solveLtStat.toString()
         << "\n  solveGemvStat: " << symCtx->solveGemvStat.toString()
         << "\n  solveGemvTStat: " << symCtx->solveGemvTStat.toString()
         << "\n  solveAssVStat: " << symCtx->solveAssVStat.toString()
         << "\n  solveAssVTStat: " << symCtx->solveAssVTStat.toString() << endl;
  }
}

void Solver::enableStats(bool enable) {
  for (auto& e : elimCtxs) e->elimStat.enabled = enable;
  symCtx->potrfStat.enabled = symCtx->trsmStat.enabled = symCtx->sygeStat.enabled = symCtx->asmblStat.enabled = enable;
}

void Solver::resetStats() {
  for (auto& e : elimCtxs) e->elimStat.reset();
  symCtx->potrfBiggestN = symCtx->syrkCalls = symCtx->gemmCalls = 0;
  symCtx->potrfStat.reset();
  symCtx->trsmStat.reset();
  symCtx->sygeStat.reset();
  symCtx->asmblStat.reset();
}

// ---------------------------------------------------------------------------------------------
static std::map<int, BackendFactory>& backendRegistry() {
  static std::map<int, BackendFactory> reg;
  return reg;
}

void registerBackend(BackendType type, BackendFactory factory) { backendRegistry()[(int)type] = std::move(factory); }

OpsPtr getBackend(const Settings& settings) {
  auto it = backendRegistry().find((int)settings.backend);
  if (it != backendRegistry().end()) return it->second(settings.numThreads);
  if (settings.backend == BackendSymbolicOnly) return symbolicOnlyOps();
  if (settings.backend == BackendCuda) return b200Ops();
  throw std::runtime_error(
      "BaSpaCho-B200: CPU backends (BackendRef/BackendFast) are not part of the product library; "
      "they exist only in oracle/ as the parity checker");
}

SolverPtr createSolver(const Settings& settings, const vector<int64_t>& paramSize, const SparseStructure& ssIn,
                       const vector<int64_t>& sparseElimRanges, const std::unordered_set<int64_t>& elimLastIds) {
  BASPACHO_CHECK(settings.addFillPolicy == AddFillComplete || elimLastIds.empty());
  BASPACHO_CHECK((int64_t)sparseElimRanges.size() != 1);
  const int64_t nParams = (int64_t)paramSize.size();
  const int64_t givenElimEnd = sparseElimRanges.empty() ? 0 : sparseElimRanges.back();
  if (!sparseElimRanges.empty()) {
    BASPACHO_CHECK(isStrictlyIncreasing(sparseElimRanges, 0, sparseElimRanges.size()));
    for (int64_t id : elimLastIds) BASPACHO_CHECK_GE(id, givenElimEnd);
  }

  SparseStructure ss = ssIn;
  if (settings.addFillPolicy != AddFillNone)
    for (size_t e = 0; e + 1 < sparseElimRanges.size(); e++)
      ss = ss.addIndependentEliminationFill(sparseElimRanges[e], sparseElimRanges[e + 1]);

  if (settings.addFillPolicy == AddFillNone || settings.addFillPolicy == AddFillForGivenElims) {
    // no reordering, lumps == spans
    vector<int64_t> spanStart(paramSize.begin(), paramSize.end());
    spanStart.push_back(0);
    cumSumVec(spanStart);
    vector<int64_t> lumpToSpan(nParams + 1), identity(nParams);
    std::iota(lumpToSpan.begin(), lumpToSpan.end(), 0);
    std::iota(identity.begin(), identity.end(), 0);
    SparseStructure byCol = ss.transpose();
    CoalescedBlockMatrixSkel skel(spanStart, lumpToSpan, byCol.ptrs, byCol.inds);
    vector<int64_t> ranges = sparseElimRanges;
    return SolverPtr(new Solver(std::move(skel), std::move(ranges), std::move(identity), getBackend(settings),
                                settings.addFillPolicy == AddFillNone ? 0 : givenElimEnd));
  }

  // order the part left after the given eliminations
  SparseStructure ssBottom = ss.extractRightBottom(givenElimEnd);
  vector<int64_t> perm = ssBottom.fillReducingPermutation();
  vector<int64_t> noCrossPoints;
  if (!elimLastIds.empty()) {  // stable partition: requested ids go last, barrier in between
    vector<int64_t> head, tail;
    for (int64_t p : perm) (elimLastIds.count(p + givenElimEnd) ? tail : head).push_back(p);
    noCrossPoints.push_back((int64_t)head.size());
    perm = head;
    perm.insert(perm.end(), tail.begin(), tail.end());
  }
  vector<int64_t> invPerm = inversePermutation(perm);
  SparseStructure sortedBottom = ssBottom.symmetricPermutation(invPerm, false);

  vector<int64_t> sortedBottomSizes(nParams - givenElimEnd);
  for (int64_t i = givenElimEnd; i < nParams; i++) sortedBottomSizes[invPerm[i - givenElimEnd]] = paramSize[i];

  const ComputationModel* model = settings.computationModel ? settings.computationModel
                                  : settings.backend == BackendCuda ? &ComputationModel::model_B200  // this library's device
                                                                    : &ComputationModel::model_OpenBlas_i7_1185g7;

  EliminationTree et(sortedBottomSizes, sortedBottom, model);
  et.buildTree();
  et.processTree(settings.findSparseEliminationRanges, noCrossPoints, settings.addFillPolicy == AddFillForAutoElims);
  et.computeAggregateStruct(settings.addFillPolicy == AddFillForAutoElims);

  // stitch: identity on the given-elimination prefix, etree result (shifted) after it
  vector<int64_t> etInvPerm = composePermutations(et.permInverse, invPerm);
  vector<int64_t> fullInvPerm(nParams);
  std::iota(fullInvPerm.begin(), fullInvPerm.begin() + givenElimEnd, 0);
  for (size_t i = 0; i < etInvPerm.size(); i++) fullInvPerm[givenElimEnd + i] = givenElimEnd + etInvPerm[i];

  vector<int64_t> fullSpanStart(nParams + 1, 0);
  leftPermute(fullSpanStart.begin(), fullInvPerm, paramSize);
  cumSumVec(fullSpanStart);

  vector<int64_t> fullLumpToSpan(givenElimEnd);
  std::iota(fullLumpToSpan.begin(), fullLumpToSpan.end(), 0);
  shiftConcat(fullLumpToSpan, givenElimEnd, et.lumpToSpan.begin(), et.lumpToSpan.end());
  BASPACHO_CHECK_EQ((int64_t)fullSpanStart.size() - 1, fullLumpToSpan.back());

  SparseStructure sortedByCol = ss.symmetricPermutation(fullInvPerm, false).transpose();
  const int64_t prefixEntries = sortedByCol.ptrs[givenElimEnd];
  vector<int64_t> fullColStart(sortedByCol.ptrs.begin(), sortedByCol.ptrs.begin() + givenElimEnd);
  shiftConcat(fullColStart, prefixEntries, et.colStart.begin(), et.colStart.end());
  BASPACHO_CHECK_EQ(fullColStart.size(), fullLumpToSpan.size());
  vector<int64_t> fullRowParam(sortedByCol.inds.begin(), sortedByCol.inds.begin() + prefixEntries);
  shiftConcat(fullRowParam, givenElimEnd, et.rowParam.begin(), et.rowParam.end());
  BASPACHO_CHECK_EQ((int64_t)fullRowParam.size(), fullColStart.back());

  CoalescedBlockMatrixSkel skel(fullSpanStart, fullLumpToSpan, fullColStart, fullRowParam);

  vector<int64_t> fullRanges = sparseElimRanges;
  if (!et.sparseElimRanges.empty())
    shiftConcat(fullRanges, givenElimEnd, et.sparseElimRanges.begin() + (sparseElimRanges.empty() ? 0 : 1),
                et.sparseElimRanges.end());
  if (fullRanges.size() == 1) fullRanges.clear();
  int64_t fullElimEnd = fullRanges.empty() ? 0 : fullRanges.back();

  return SolverPtr(new Solver(std::move(skel), std::move(fullRanges), std::move(fullInvPerm), getBackend(settings),
                              settings.addFillPolicy == AddFillForAutoElims ? fullElimEnd : nParams));
}

}  // namespace BaSpaCho
