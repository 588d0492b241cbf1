// Elimination tree of a (permuted) block pattern plus the two heuristics layered on it:
// detection of "sparse elimination ranges" (large sets of small same-height leaves) and
// cost-model-driven merging of children into parents (supernode formation).
// Same role / public members as reference baspacho/baspacho/EliminationTree.h:28-97.
#pragma once

#include <tuple>
#include <vector>
#include "ComputationModel.h"
#include "SparseStructure.h"

namespace BaSpaCho {

struct EliminationTree {
  EliminationTree(const std::vector<int64_t>& paramSize, const SparseStructure& ss,
                  const ComputationModel* compMod = nullptr);

  void buildTree();

  void processTree(bool detectSparseElimRanges, const std::vector<int64_t>& noCrossPoints = {},
                   bool findOnlyElims = false);

  void computeAggregateStruct(bool fillOnlyForElims = false);

  std::vector<int64_t> computeSpanStart();

  // internal steps of processTree
  void computeNodeHeights(const std::vector<int64_t>& noCrossPoints);
  void computeSparseElimRanges(const std::vector<int64_t>& noCrossPoints);
  void computeMerges();
  void collapseMergePointers();

  // inputs
  std::vector<int64_t> paramSize;
  const SparseStructure& ss;  // CSR lower triangle (row k lists columns <= k)
  const ComputationModel& compMod;

  // buildTree outputs
  std::vector<int64_t> parent;
  std::vector<int64_t> nodeSize;
  std::vector<int64_t> nodeRows;       // scalar rows strictly below the node's diagonal block
  std::vector<int64_t> nodeRowBlocks;  // same, counted in blocks
  std::vector<std::vector<int64_t>> perColNodes;
  struct NodeStats {  // one (row node, column node) entry of L seen from the row
    int64_t colIdx, rBlocks, rows, rBlocksDown, rowsDown;
  };
  std::vector<std::vector<NodeStats>> perRowNodeStats;
  std::vector<Lin2> sygeCosts;   // per column: syrk/gemm cost, affine in node size
  std::vector<Lin2> asmblCosts;  // per column: assemble cost, affine in #merged nodes

  // processTree outputs
  std::vector<int64_t> sparseElimRanges;
  std::vector<std::tuple<int64_t, int64_t, int64_t>> unmergedHeightNode;  // (height, size, node)
  std::vector<bool> forbidMerge;
  std::vector<int64_t> numMergedNodes;
  std::vector<int64_t> mergeWith;
  int64_t numMerges = 0;

  // aggregate structure
  std::vector<int64_t> permInverse;
  std::vector<int64_t> lumpStart;
  std::vector<int64_t> lumpToSpan;
  std::vector<int64_t> colStart;
  std::vector<int64_t> rowParam;
};

}  // namespace BaSpaCho
