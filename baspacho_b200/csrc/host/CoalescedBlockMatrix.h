// Index skeleton of the supernodal factor and its numeric data layout.
//   span  = one user parameter block;  lump = supernode = consecutive spans
//   chain = (rows of one span) x (columns of one lump), stored row-major, stride = lump width
//   board = all chains of one lump-column whose row spans belong to the same row lump
// Chains of a lump are consecutive in memory, so a lump column is one row-major
// (totalRows x lumpWidth) matrix whose first lumpWidth rows are the diagonal block.
// Same member names/meaning as reference baspacho/baspacho/CoalescedBlockMatrix.h:38-111; the arrays
// are the bit-for-bit integer contract (golden vectors: reference tests/CoalescedBlockMatrixTest.cpp:48-112).
#pragma once

#include <cstdint>
#include <vector>
#include "Accessor.h"
#include "DebugMacros.h"

namespace BaSpaCho {

constexpr int64_t kInvalid = -1;

struct CoalescedBlockMatrixSkel {
  CoalescedBlockMatrixSkel(const std::vector<int64_t>& spanStart, const std::vector<int64_t>& lumpToSpan,
                           const std::vector<int64_t>& colPtr, const std::vector<int64_t>& rowInd);

  // dense (order-offset)^2 matrix, ROW-major, from span `startSpanIndex` (on a lump boundary) on;
  // only the stored (lower) part is written unless fillUpperHalf
  template <typename T>
  void densify(T* dense, const T* data, bool fillUpperHalf = false, int64_t startSpanIndex = 0) const;

  template <typename T>
  std::vector<T> densify(const std::vector<T>& data, bool fillUpperHalf = false) const;

  // diag <- diag*(1+alpha) + beta on every lump's diagonal block
  template <typename T>
  void damp(T* data, T alpha, T beta) const;
  template <typename T>
  void damp(std::vector<T>& data, T alpha, T beta) const {
    BASPACHO_CHECK_EQ(dataSize(), (int64_t)data.size());
    damp(data.data(), alpha, beta);
  }

  int64_t numSpans() const { return (int64_t)spanStart.size() - 1; }
  int64_t numLumps() const { return (int64_t)lumpStart.size() - 1; }
  int64_t order() const { return spanStart.back(); }
  int64_t dataSize() const { return chainData.back(); }
  int64_t spanVectorOffset(int64_t span) const { return spanStart[span]; }
  int64_t spanMatrixOffset(int64_t span) const {
    BASPACHO_CHECK_EQ(spanOffsetInLump[span], 0);
    return chainData[chainColPtr[spanToLump[span]]];
  }

  // convenience queries used by drivers/backends
  int64_t lumpSize(int64_t l) const { return lumpStart[l + 1] - lumpStart[l]; }
  int64_t lumpTotalRows(int64_t l) const { return chainRowsTillEnd[chainColPtr[l + 1] - 1]; }
  int64_t lumpDataOffset(int64_t l) const { return chainData[chainColPtr[l]]; }

  CoalescedAccessor accessor() const {
    CoalescedAccessor a;
    a.init(spanStart.data(), spanToLump.data(), lumpStart.data(), spanOffsetInLump.data(),
           chainColPtr.data(), chainRowSpan.data(), chainData.data());
    return a;
  }

  std::vector<int64_t> spanStart;         // (+ final)
  std::vector<int64_t> spanToLump;        // (+ final)
  std::vector<int64_t> lumpStart;         // (+ final)
  std::vector<int64_t> lumpToSpan;        // (+ final)
  std::vector<int64_t> spanOffsetInLump;  // (+ final)

  // per chain, column-ordered
  std::vector<int64_t> chainColPtr;       // per lump (+ final)
  std::vector<int64_t> chainRowSpan;
  std::vector<int64_t> chainData;         // numeric offset (+ final = dataSize)
  std::vector<int64_t> chainRowsTillEnd;  // rows of the column up to and including this chain

  // per board, column-ordered; each column closed by a sentinel entry
  std::vector<int64_t> boardColPtr;       // per lump (+ final)
  std::vector<int64_t> boardRowLump;      // sentinel: kInvalid
  std::vector<int64_t> boardChainColOrd;  // first chain (ordinal in column); sentinel: #chains

  // per board, row-ordered (no sentinels)
  std::vector<int64_t> boardRowPtr;   // per row lump (+ final)
  std::vector<int64_t> boardColLump;
  std::vector<int64_t> boardColOrd;   // ordinal of the board inside its column
};

}  // namespace BaSpaCho
