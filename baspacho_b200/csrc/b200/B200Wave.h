// Level-of-tree ("wavefront") plan of the dense part of the factorization, built once per sparsity pattern on the
// host (pure C++). Lumps are grouped by their level in the supernodal dependency graph (a lump depends on every
// earlier lump that has a board in its row range - the left-looking schedule of reference Solver.cpp:198-218); all
// lumps of a level are independent, so a level is TWO launches for the small supernodes:
//   wave_update_kernel : per (target lump, 64-row tile) gather-GEMM of every source board straight into the target
//                        (replaces saveSyrkGemm + prepareAssemble + assemble, reference MatOpsCuda.cu:471-498, 568-590;
//                        fixed source order -> deterministic, no temp buffer, no atomics)
//   panel_kernel (batched over a work list): diagonal Cholesky + triangular solve of every small lump of the level
// Wide lumps (> one 96-column panel) keep the per-lump path (GEMM into the temp + assemble, blocked potrf).
#pragma once

#include <cstdint>
#include <vector>
#include "../host/CoalescedBlockMatrix.h"

namespace BaSpaCho {
namespace b200 {

struct WaveTarget {
  int64_t dataOff;         // offset of the lump column (diagonal block first)
  int32_t width, totalRows;
  int32_t srcBegin, srcEnd;  // range in WavePlan::sources
};
struct WaveSource {
  int64_t dataOff;   // first row of the board inside the source column (row-major, stride k)
  int32_t k;         // width of the source lump
  int32_t rows;      // rows from the board start to the end of the source column
  int32_t mapBegin;  // rowMap[mapBegin + i] = row of source row i inside the target column (monotone increasing)
  int32_t pad;
};
struct WaveTile {
  int32_t target, row0;
};
struct WavePanel {
  int64_t dataOff;
  int32_t n, rows, slab;
  int32_t lumpIdx;  // index of the lump inside its level (selects the load counter of the panel kernel)
};
struct WaveLevel {
  int32_t tileBegin = 0, tileEnd = 0, panelBegin = 0, panelEnd = 0;
  int32_t numSmall = 0;  // small lumps of the level
  std::vector<int64_t> bigLumps;
};

struct WavePlan {
  static constexpr int kTileRows = 64;
  static constexpr int kMaxSmallWidth = 96;
  static constexpr int kPanelRows = 64;
  int64_t firstLump = 0;
  std::vector<WaveTarget> targets;
  std::vector<WaveSource> sources;
  std::vector<int32_t> rowMap;
  std::vector<WaveTile> tiles;
  std::vector<WavePanel> panels;
  std::vector<WaveLevel> levels;
  int64_t numSmall = 0, numBig = 0;
};

// dense lumps [firstLump, numLumps); sources before firstLump were handled by the sparse elimination
WavePlan buildWavePlan(const CoalescedBlockMatrixSkel& skel, int64_t firstLump);

}  // namespace b200
}  // namespace BaSpaCho
