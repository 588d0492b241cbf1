// Skeleton construction; array contents follow reference
// baspacho/baspacho/CoalescedBlockMatrix.cpp:17-122 exactly (pinned by golden tests);
// densify :124-170 and damp :172-187 restated without Eigen (row-major dense output).
#include "CoalescedBlockMatrix.h"
#include <algorithm>
#include "Utils.h"

namespace BaSpaCho {

using std::vector;

CoalescedBlockMatrixSkel::CoalescedBlockMatrixSkel(const vector<int64_t>& spanStart_, const vector<int64_t>& lumpToSpan_,
                                                   const vector<int64_t>& colPtr, const vector<int64_t>& rowInd)
    : spanStart(spanStart_), lumpToSpan(lumpToSpan_) {
  BASPACHO_CHECK_GE(spanStart.size(), lumpToSpan.size());
  BASPACHO_CHECK_GE((int64_t)lumpToSpan.size(), 1);
  BASPACHO_CHECK_EQ((int64_t)spanStart.size() - 1, lumpToSpan.back());
  BASPACHO_CHECK_EQ(colPtr.size(), lumpToSpan.size());
  BASPACHO_CHECK(isStrictlyIncreasing(spanStart, 0, spanStart.size()));
  BASPACHO_CHECK(isStrictlyIncreasing(lumpToSpan, 0, lumpToSpan.size()));

  const int64_t nSpans = (int64_t)spanStart.size() - 1;
  const int64_t nLumps = (int64_t)lumpToSpan.size() - 1;
  auto spanRows = [&](int64_t s) { return spanStart[s + 1] - spanStart[s]; };

  // span <-> lump maps
  spanToLump.assign(nSpans + 1, nLumps);
  lumpStart.assign(nLumps + 1, spanStart[nSpans]);
  spanOffsetInLump.assign(nSpans + 1, 0);
  for (int64_t l = 0; l < nLumps; l++) {
    lumpStart[l] = spanStart[lumpToSpan[l]];
    for (int64_t s = lumpToSpan[l]; s < lumpToSpan[l + 1]; s++) {
      spanToLump[s] = l;
      spanOffsetInLump[s] = spanStart[s] - lumpStart[l];
    }
  }

  // chains and boards, column by column
  chainColPtr.assign(nLumps + 1, 0);
  boardColPtr.assign(nLumps + 1, 0);
  int64_t dataCursor = 0;
  for (int64_t l = 0; l < nLumps; l++) {
    const int64_t cBegin = colPtr[l], cEnd = colPtr[l + 1];
    const int64_t nOwnSpans = lumpToSpan[l + 1] - lumpToSpan[l];
    const int64_t width = lumpStart[l + 1] - lumpStart[l];
    BASPACHO_CHECK(isStrictlyIncreasing(rowInd, cBegin, cEnd));
    // the column must open with the lump's own spans (full diagonal block present)
    BASPACHO_CHECK_GE(cEnd - cBegin, nOwnSpans);
    BASPACHO_CHECK_EQ(rowInd[cBegin], lumpToSpan[l]);
    BASPACHO_CHECK_EQ(rowInd[cBegin + nOwnSpans - 1], lumpToSpan[l + 1] - 1);

    chainColPtr[l] = (int64_t)chainRowSpan.size();
    boardColPtr[l] = (int64_t)boardRowLump.size();
    int64_t rowsSoFar = 0, openRowLump = kInvalid;
    for (int64_t i = cBegin; i < cEnd; i++) {
      int64_t s = rowInd[i];
      chainRowSpan.push_back(s);
      chainData.push_back(dataCursor);
      dataCursor += width * spanRows(s);
      rowsSoFar += spanRows(s);
      chainRowsTillEnd.push_back(rowsSoFar);
      if (spanToLump[s] != openRowLump) {  // a new board starts here
        openRowLump = spanToLump[s];
        boardRowLump.push_back(openRowLump);
        boardChainColOrd.push_back(i - cBegin);
      }
    }
    boardRowLump.push_back(kInvalid);
    boardChainColOrd.push_back(cEnd - cBegin);
  }
  chainColPtr[nLumps] = (int64_t)chainRowSpan.size();
  boardColPtr[nLumps] = (int64_t)boardRowLump.size();
  chainData.push_back(dataCursor);

  // row-ordered view of the boards (counting sort by row lump; column order preserved inside a row)
  boardRowPtr.assign(nLumps + 1, 0);
  for (int64_t l = 0; l < nLumps; l++)
    for (int64_t b = boardColPtr[l]; b + 1 < boardColPtr[l + 1]; b++) boardRowPtr[boardRowLump[b]]++;
  int64_t nBoards = cumSumVec(boardRowPtr);
  boardColLump.resize(nBoards);
  boardColOrd.resize(nBoards);
  vector<int64_t> cursor(boardRowPtr.begin(), boardRowPtr.end() - 1);
  for (int64_t l = 0; l < nLumps; l++)
    for (int64_t b = boardColPtr[l]; b + 1 < boardColPtr[l + 1]; b++) {
      int64_t slot = cursor[boardRowLump[b]]++;
      boardColLump[slot] = l;
      boardColOrd[slot] = b - boardColPtr[l];
    }
}

template <typename T>
void CoalescedBlockMatrixSkel::densify(T* dense, const T* data, bool fillUpperHalf, int64_t startSpanIndex) const {
  BASPACHO_CHECK_GE(startSpanIndex, 0);
  BASPACHO_CHECK_LT(startSpanIndex, (int64_t)spanOffsetInLump.size());
  BASPACHO_CHECK_EQ(spanOffsetInLump[startSpanIndex], 0);
  const int64_t base = spanStart[startSpanIndex];
  const int64_t n = order() - base;
  std::fill(dense, dense + n * n, T(0));
  for (int64_t l = spanToLump[startSpanIndex]; l < numLumps(); l++) {
    const int64_t c0 = lumpStart[l] - base, width = lumpSize(l);
    for (int64_t i = chainColPtr[l]; i < chainColPtr[l + 1]; i++) {
      int64_t s = chainRowSpan[i];
      int64_t r0 = spanStart[s] - base, rows = spanStart[s + 1] - spanStart[s];
      const T* src = data + chainData[i];
      for (int64_t r = 0; r < rows; r++)
        std::copy(src + r * width, src + (r + 1) * width, dense + (r0 + r) * n + c0);
    }
  }
  if (fillUpperHalf)
    for (int64_t r = 0; r < n; r++)
      for (int64_t c = 0; c < r; c++) dense[c * n + r] = dense[r * n + c];
}

template <typename T>
vector<T> CoalescedBlockMatrixSkel::densify(const vector<T>& data, bool fillUpperHalf) const {
  BASPACHO_CHECK_EQ(dataSize(), (int64_t)data.size());
  vector<T> dense(order() * order());
  densify(dense.data(), data.data(), fillUpperHalf, 0);
  return dense;
}

template <typename T>
void CoalescedBlockMatrixSkel::damp(T* data, T alpha, T beta) const {
  for (int64_t l = 0; l < numLumps(); l++) {
    int64_t width = lumpSize(l);
    T* diag = data + lumpDataOffset(l);
    for (int64_t i = 0; i < width; i++) {
      T& d = diag[i * (width + 1)];
      d = d * (T(1) + alpha) + beta;
    }
  }
}

template void CoalescedBlockMatrixSkel::densify<double>(double*, const double*, bool, int64_t) const;
template void CoalescedBlockMatrixSkel::densify<float>(float*, const float*, bool, int64_t) const;
template vector<double> CoalescedBlockMatrixSkel::densify<double>(const vector<double>&, bool) const;
template vector<float> CoalescedBlockMatrixSkel::densify<float>(const vector<float>&, bool) const;
template void CoalescedBlockMatrixSkel::damp<double>(double*, double, double) const;
template void CoalescedBlockMatrixSkel::damp<float>(float*, float, float) const;

}  // namespace BaSpaCho
