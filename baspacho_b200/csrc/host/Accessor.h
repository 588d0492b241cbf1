// Lightweight (pointer-only, trivially copyable) accessors locating a block inside the flat factor
// data; usable as by-value kernel arguments. Same layout/semantics as reference
// baspacho/baspacho/Accessor.h:18-200 minus the Eigen-typed views (offset/stride/flip only).
#pragma once
#include <cstdint>
#include <tuple>
#include <utility>
#include "Utils.h"

namespace BaSpaCho {

struct CoalescedAccessor {
  void init(const int64_t* spanStart_, const int64_t* spanToLump_, const int64_t* lumpStart_,
            const int64_t* spanOffsetInLump_, const int64_t* chainColPtr_, const int64_t* chainRowSpan_,
            const int64_t* chainData_) {
    spanStart = spanStart_, spanToLump = spanToLump_, lumpStart = lumpStart_;
    spanOffsetInLump = spanOffsetInLump_, chainColPtr = chainColPtr_;
    chainRowSpan = chainRowSpan_, chainData = chainData_;
  }

  BSP_HD int64_t paramSize(int64_t b) const { return spanStart[b + 1] - spanStart[b]; }
  BSP_HD int64_t paramStart(int64_t b) const { return spanStart[b]; }

  // (offset, row stride) of block (rowBlock, colBlock), rowBlock >= colBlock, internal ordering
  BSP_HD std::pair<int64_t, int64_t> blockOffset(int64_t rowBlock, int64_t colBlock) const {
    int64_t lump = spanToLump[colBlock];
    int64_t first = chainColPtr[lump], count = chainColPtr[lump + 1] - first;
    int64_t pos = bisect(chainRowSpan + first, count, rowBlock);
    return {chainData[first + pos] + spanOffsetInLump[colBlock], lumpStart[lump + 1] - lumpStart[lump]};
  }

  BSP_HD std::pair<int64_t, int64_t> diagBlockOffset(int64_t b) const {
    int64_t lump = spanToLump[b];
    int64_t width = lumpStart[lump + 1] - lumpStart[lump];
    return {chainData[chainColPtr[lump]] + spanOffsetInLump[b] * (width + 1), width};
  }

  const int64_t* spanStart;
  const int64_t* spanToLump;
  const int64_t* lumpStart;
  const int64_t* spanOffsetInLump;
  const int64_t* chainColPtr;
  const int64_t* chainRowSpan;
  const int64_t* chainData;
};

struct PermutedCoalescedAccessor {
  void init(const CoalescedAccessor& acc, const int64_t* permutation_) {
    plainAcc = acc;
    permutation = permutation_;
  }
  void init(const int64_t* spanStart_, const int64_t* spanToLump_, const int64_t* lumpStart_,
            const int64_t* spanOffsetInLump_, const int64_t* chainColPtr_, const int64_t* chainRowSpan_,
            const int64_t* chainData_, const int64_t* permutation_) {
    plainAcc.init(spanStart_, spanToLump_, lumpStart_, spanOffsetInLump_, chainColPtr_, chainRowSpan_, chainData_);
    permutation = permutation_;
  }

  BSP_HD int64_t paramSize(int64_t b) const { return plainAcc.paramSize(permutation[b]); }
  BSP_HD int64_t paramStart(int64_t b) const { return plainAcc.paramStart(permutation[b]); }

  // user block indices; flipped=true means the stored block is the transpose of the requested one
  BSP_HD std::tuple<int64_t, int64_t, bool> blockOffset(int64_t rowBlock, int64_t colBlock) const {
    int64_t r = permutation[rowBlock], c = permutation[colBlock];
    bool flipped = r < c;
    auto os = flipped ? plainAcc.blockOffset(c, r) : plainAcc.blockOffset(r, c);
    return {os.first, os.second, flipped};
  }

  BSP_HD std::pair<int64_t, int64_t> diagBlockOffset(int64_t b) const {
    return plainAcc.diagBlockOffset(permutation[b]);
  }

  CoalescedAccessor plainAcc;
  const int64_t* permutation;
};

}  // namespace BaSpaCho
