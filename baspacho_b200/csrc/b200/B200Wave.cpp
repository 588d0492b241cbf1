// Wavefront plan builder (host only). See B200Wave.h.
#include "B200Wave.h"
#include <algorithm>
#include <stdexcept>
#include "../host/Utils.h"

namespace BaSpaCho {
namespace b200 {

WavePlan buildWavePlan(const CoalescedBlockMatrixSkel& sk, int64_t firstLump) {
  WavePlan p;
  p.firstLump = firstLump;
  const int64_t nLumps = sk.numLumps();
  if (sk.dataSize() >= (int64_t(1) << 40)) throw std::runtime_error("wave plan: factor too large");

  // level of every dense lump: 1 + max level of its dense sources
  std::vector<int32_t> level(nLumps, 0);
  int32_t maxLevel = -1;
  for (int64_t l = firstLump; l < nLumps; l++) {
    int32_t lv = 0;
    for (int64_t r = sk.boardRowPtr[l], rEnd = sk.boardRowPtr[l + 1] - 1; r < rEnd; r++) {
      int64_t src = sk.boardColLump[r];
      if (src >= firstLump) lv = std::max(lv, level[src] + 1);
    }
    level[l] = lv;
    maxLevel = std::max(maxLevel, lv);
  }
  p.levels.resize(maxLevel + 1);
  std::vector<std::vector<int64_t>> byLevel(maxLevel + 1);
  for (int64_t l = firstLump; l < nLumps; l++) byLevel[level[l]].push_back(l);

  for (int32_t lv = 0; lv <= maxLevel; lv++) {
    WaveLevel& L = p.levels[lv];
    L.tileBegin = (int32_t)p.tiles.size();
    L.panelBegin = (int32_t)p.panels.size();
    for (int64_t t : byLevel[lv]) {
      const int64_t w = sk.lumpSize(t), totalRows = sk.lumpTotalRows(t);
      if (w > WavePlan::kMaxSmallWidth) {
        L.bigLumps.push_back(t);
        p.numBig++;
        continue;
      }
      p.numSmall++;
      WaveTarget tg;
      tg.dataOff = sk.lumpDataOffset(t);
      tg.width = (int32_t)w, tg.totalRows = (int32_t)totalRows;
      tg.srcBegin = (int32_t)p.sources.size();
      const int64_t tBegin = sk.chainColPtr[t], tCount = sk.chainColPtr[t + 1] - tBegin;
      for (int64_t r = sk.boardRowPtr[t], rEnd = sk.boardRowPtr[t + 1] - 1; r < rEnd; r++) {
        const int64_t src = sk.boardColLump[r];
        if (src < firstLump) continue;
        const int64_t ord = sk.boardColOrd[r], cb = sk.chainColPtr[src], bb = sk.boardColPtr[src];
        const int64_t ch0 = sk.boardChainColOrd[bb + ord];
        const int64_t chEnd = sk.boardChainColOrd[sk.boardColPtr[src + 1] - 1];
        const int64_t rowBegin = sk.chainRowsTillEnd[cb + ch0 - 1];
        WaveSource s;
        s.dataOff = sk.chainData[cb + ch0];
        s.k = (int32_t)sk.lumpSize(src);
        s.rows = (int32_t)(sk.chainRowsTillEnd[cb + chEnd - 1] - rowBegin);
        s.mapBegin = (int32_t)p.rowMap.size();
        s.pad = 0;
        for (int64_t c = cb + ch0; c < cb + chEnd; c++) {
          const int64_t span = sk.chainRowSpan[c];
          const int64_t at = bisect(sk.chainRowSpan.data() + tBegin, tCount, span);
          if (sk.chainRowSpan[tBegin + at] != span)
            throw std::runtime_error("wave plan: source row span missing from the target column (no fill?)");
          const int64_t tRow = sk.chainRowsTillEnd[tBegin + at] - (sk.spanStart[span + 1] - sk.spanStart[span]);
          const int64_t nr = sk.chainRowsTillEnd[c] - sk.chainRowsTillEnd[c - 1];
          for (int64_t i = 0; i < nr; i++) p.rowMap.push_back((int32_t)(tRow + i));
        }
        p.sources.push_back(s);
        if (p.rowMap.size() >= (size_t(1) << 31)) throw std::runtime_error("wave plan: row maps too large");
      }
      tg.srcEnd = (int32_t)p.sources.size();
      const int32_t ti = (int32_t)p.targets.size();
      p.targets.push_back(tg);
      if (tg.srcEnd > tg.srcBegin)
        for (int64_t r0 = 0; r0 < totalRows; r0 += WavePlan::kTileRows) p.tiles.push_back(WaveTile{ti, (int32_t)r0});
      const int64_t below = totalRows - w;
      const int64_t slabs = std::max<int64_t>(1, (below + WavePlan::kPanelRows - 1) / WavePlan::kPanelRows);
      for (int64_t sl = 0; sl < slabs; sl++)
        p.panels.push_back(WavePanel{tg.dataOff, (int32_t)w, (int32_t)below, (int32_t)sl, L.numSmall});
      L.numSmall++;
    }
    L.tileEnd = (int32_t)p.tiles.size();
    L.panelEnd = (int32_t)p.panels.size();
  }
  return p;
}

}  // namespace b200
}  // namespace BaSpaCho
