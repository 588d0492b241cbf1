// Host-side execution plans of the B200 backend, built once per sparsity pattern at createSymbolicCtx /
// prepareElimination time (pure C++, no CUDA: unit-testable on a machine without a GPU).
//
// ElimPlan: the sparse ("Schur") elimination of a range of independent lumps, restated TARGET-major:
// the reference enumerates (column, block pair) tasks and applies them with one atomicAdd per scalar
// (MatOpsCuda.cu:235-331, CudaAtomic.cuh:39-49), which is non-deterministic; here every destination block
// (row span j, col span i) owns the list of the block pairs that update it, so one owner sums them in a fixed
// order and writes the block once - deterministic, no atomics, no device-side bisect.
#pragma once

#include <cstdint>
#include <vector>
#include "../host/CoalescedBlockMatrix.h"

namespace BaSpaCho {
namespace b200 {

struct ElimPlan {
  int64_t lumpsBegin = 0, lumpsEnd = 0;
  int64_t spanRowBegin = 0;  // first row span below the range: lumpToSpan[lumpsEnd]
  int uniformLumpSize = 0;   // > 0 when every lump of the range has this width

  // destination blocks (sorted by data offset) and their pair tasks
  std::vector<int64_t> dstOff;      // offset of the block inside the factor data
  std::vector<int32_t> dstStride;   // row stride (width of the target lump)
  std::vector<int16_t> dstRows;     // rows of the block   (= rows of the B chain)
  std::vector<int16_t> dstCols;     // columns of the block (= rows of the A chain)
  std::vector<int32_t> dstTaskPtr;  // per destination (+1)
  std::vector<uint32_t> taskA;      // offset of the A chain (rows = dstCols, cols = k)
  std::vector<uint32_t> taskB;      // offset of the B chain (rows = dstRows, cols = k)
  std::vector<uint16_t> taskK;      // width of the source lump
  int maxDstElems = 0;
  int uniRows = 0, uniCols = 0, uniK = 0;  // > 0 when every destination / task has these dimensions
  // destinations split by task count: light ones are summed by a few lanes, heavy ones by a whole CTA
  static constexpr int kHeavyTasks = 512;
  std::vector<int32_t> lightList, heavyList;

  // row view for the triangular solves: per row span >= spanRowBegin, the chains found in that row
  std::vector<int32_t> rowPtr;        // per row span - spanRowBegin (+1)
  std::vector<int64_t> rowChainOff;   // data offset of the chain (rows(span) x k)
  std::vector<int32_t> rowChainCol;   // first scalar column of the source lump (lumpStart)
  std::vector<int16_t> rowChainK;     // width of the source lump
  int maxRowSpanSize = 0;

  // algorithmic work (entries / flops), for the roofline report
  double factorEntries = 0, gatherFlops = 0, gatherEntries = 0;

  int64_t numTasks() const { return (int64_t)taskA.size(); }
  int64_t numDst() const { return (int64_t)dstOff.size(); }
};

ElimPlan buildElimPlan(const CoalescedBlockMatrixSkel& skel, int64_t lumpsBegin, int64_t lumpsEnd);

}  // namespace b200
}  // namespace BaSpaCho
