// Plan builders of the B200 backend (host only). See B200Plan.h.
#include "B200Plan.h"
#include <algorithm>
#include <stdexcept>
#include "../host/Utils.h"

namespace BaSpaCho {
namespace b200 {

ElimPlan buildElimPlan(const CoalescedBlockMatrixSkel& sk, int64_t lumpsBegin, int64_t lumpsEnd) {
  ElimPlan p;
  p.lumpsBegin = lumpsBegin, p.lumpsEnd = lumpsEnd;
  p.spanRowBegin = sk.lumpToSpan[lumpsEnd];
  const int64_t nRows = sk.numSpans() - p.spanRowBegin;
  if (sk.dataSize() >= (int64_t(1) << 32))
    throw std::runtime_error("B200 sparse elimination plan: factor data beyond 2^32 entries is not supported yet");

  // the plan stores block dimensions in 16 bits and vector / stride positions in 32: refuse what does not fit
  // instead of truncating (the reference has no such limit; nothing near it occurs in its problem families)
  if (sk.order() >= (int64_t(1) << 31))
    throw std::runtime_error("B200 sparse elimination plan: vector order beyond 2^31 is not supported");
  for (int64_t l = lumpsBegin; l < lumpsEnd; l++)
    if (sk.lumpSize(l) > 32767) throw std::runtime_error("B200 sparse elimination plan: eliminated lump wider than 32767");
  for (int64_t sp = p.spanRowBegin; sp < sk.numSpans(); sp++)
    if (sk.spanStart[sp + 1] - sk.spanStart[sp] > 32767)
      throw std::runtime_error("B200 sparse elimination plan: row span larger than 32767");
  for (int64_t l = lumpsEnd; l < sk.numLumps(); l++)
    if (sk.lumpSize(l) >= (int64_t(1) << 31)) throw std::runtime_error("B200 sparse elimination plan: target lump too wide");

  p.uniformLumpSize = lumpsEnd > lumpsBegin ? (int)sk.lumpSize(lumpsBegin) : 0;
  for (int64_t l = lumpsBegin; l < lumpsEnd; l++) {
    if (sk.lumpSize(l) != p.uniformLumpSize) p.uniformLumpSize = 0;
    p.factorEntries += (double)sk.lumpSize(l) * sk.lumpTotalRows(l);
    p.gatherEntries += (double)sk.lumpSize(l) * (sk.lumpTotalRows(l) - sk.lumpSize(l));
  }

  // ---- row view: per row span, the chains of the range's lumps found there (ascending source lump)
  auto firstBelow = [&](int64_t l) { return sk.chainColPtr[l] + (sk.lumpToSpan[l + 1] - sk.lumpToSpan[l]); };
  std::vector<int64_t> cnt(nRows + 1, 0);
  for (int64_t l = lumpsBegin; l < lumpsEnd; l++)
    for (int64_t c = firstBelow(l); c < sk.chainColPtr[l + 1]; c++) {
      int64_t rel = sk.chainRowSpan[c] - p.spanRowBegin;
      if (rel < 0) throw std::runtime_error("B200 sparse elimination: lumps of the range are not independent");
      cnt[rel]++;
    }
  int64_t tot = cumSumVec(cnt);
  p.rowPtr.assign(cnt.begin(), cnt.end());
  p.rowChainOff.resize(tot);
  p.rowChainCol.resize(tot);
  p.rowChainK.resize(tot);
  std::vector<int64_t> rowChainIdx(tot), rowLump(tot);
  {
    std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
    for (int64_t l = lumpsBegin; l < lumpsEnd; l++)
      for (int64_t c = firstBelow(l); c < sk.chainColPtr[l + 1]; c++) {
        int64_t e = cur[sk.chainRowSpan[c] - p.spanRowBegin]++;
        rowChainIdx[e] = c, rowLump[e] = l;
        p.rowChainOff[e] = sk.chainData[c];
        p.rowChainCol[e] = (int32_t)sk.lumpStart[l];
        p.rowChainK[e] = (int16_t)sk.lumpSize(l);
      }
  }
  for (int64_t r = 0; r < nRows; r++)
    if (p.rowPtr[r] < p.rowPtr[r + 1]) {
      int64_t s = r + p.spanRowBegin;
      p.maxRowSpanSize = std::max<int>(p.maxRowSpanSize, (int)(sk.spanStart[s + 1] - sk.spanStart[s]));
    }

  // ---- destination-major pair tasks: destination (row span b, col span a), a <= b, both below the range
  std::vector<int32_t> perB(nRows, 0), pos(nRows, 0);
  std::vector<int64_t> touched;
  p.dstTaskPtr.push_back(0);
  for (int64_t aRel = 0; aRel < nRows; aRel++) {
    if (p.rowPtr[aRel] == p.rowPtr[aRel + 1]) continue;
    const int64_t spanA = aRel + p.spanRowBegin, target = sk.spanToLump[spanA];
    const int64_t tBegin = sk.chainColPtr[target], tCount = sk.chainColPtr[target + 1] - tBegin;
    const int64_t colOff = sk.spanOffsetInLump[spanA], tWidth = sk.lumpSize(target);
    const int64_t rowsA = sk.spanStart[spanA + 1] - sk.spanStart[spanA];
    touched.clear();
    for (int64_t e = p.rowPtr[aRel]; e < p.rowPtr[aRel + 1]; e++)
      for (int64_t cb = rowChainIdx[e], end = sk.chainColPtr[rowLump[e] + 1]; cb < end; cb++) {
        int64_t bRel = sk.chainRowSpan[cb] - p.spanRowBegin;
        if (perB[bRel]++ == 0) touched.push_back(bRel);
      }
    std::sort(touched.begin(), touched.end());
    int64_t base = (int64_t)p.taskA.size(), run = base;
    for (int64_t bRel : touched) {
      const int64_t spanB = bRel + p.spanRowBegin;
      int64_t at = bisect(sk.chainRowSpan.data() + tBegin, tCount, spanB);
      if (sk.chainRowSpan[tBegin + at] != spanB)
        throw std::runtime_error("B200 sparse elimination: target block missing from the skeleton (no fill?)");
      const int64_t rowsB = sk.spanStart[spanB + 1] - sk.spanStart[spanB];
      p.dstOff.push_back(sk.chainData[tBegin + at] + colOff);
      p.dstStride.push_back((int32_t)tWidth);
      p.dstRows.push_back((int16_t)rowsB);
      p.dstCols.push_back((int16_t)rowsA);
      p.maxDstElems = std::max<int>(p.maxDstElems, (int)(rowsA * rowsB));
      p.gatherEntries += 2.0 * rowsA * rowsB;
      pos[bRel] = (int32_t)run;
      run += perB[bRel];
      p.dstTaskPtr.push_back((int32_t)run);
    }
    p.taskA.resize(run);
    p.taskB.resize(run);
    p.taskK.resize(run);
    for (int64_t e = p.rowPtr[aRel]; e < p.rowPtr[aRel + 1]; e++) {
      const int64_t ca = rowChainIdx[e], l = rowLump[e];
      for (int64_t cb = ca, end = sk.chainColPtr[l + 1]; cb < end; cb++) {
        int64_t bRel = sk.chainRowSpan[cb] - p.spanRowBegin;
        int64_t idx = pos[bRel]++;
        p.taskA[idx] = (uint32_t)sk.chainData[ca];
        p.taskB[idx] = (uint32_t)sk.chainData[cb];
        p.taskK[idx] = (uint16_t)sk.lumpSize(l);
        p.gatherFlops += 2.0 * rowsA * (double)(sk.spanStart[sk.chainRowSpan[cb] + 1] - sk.spanStart[sk.chainRowSpan[cb]]) * sk.lumpSize(l);
      }
    }
    for (int64_t bRel : touched) perB[bRel] = 0;
    if (run >= (int64_t(1) << 31)) throw std::runtime_error("B200 sparse elimination plan: too many block pairs");
  }
  for (int64_t d = 0; d < p.numDst(); d++) {
    const int32_t nt = p.dstTaskPtr[d + 1] - p.dstTaskPtr[d];
    (nt > ElimPlan::kHeavyTasks ? p.heavyList : p.lightList).push_back((int32_t)d);
  }
  if (p.numDst() > 0) {
    p.uniRows = p.dstRows[0], p.uniCols = p.dstCols[0], p.uniK = p.taskK.empty() ? 0 : p.taskK[0];
    for (int64_t d = 0; d < p.numDst(); d++)
      if (p.dstRows[d] != p.uniRows || p.dstCols[d] != p.uniCols) p.uniRows = p.uniCols = 0;
    for (uint16_t k : p.taskK)
      if (k != p.uniK) p.uniK = 0;
  }
  return p;
}

}  // namespace b200
}  // namespace BaSpaCho
