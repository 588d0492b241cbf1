// Behaviour follows reference baspacho/baspacho/EliminationTree.cpp: buildTree :28-95,
// computeNodeHeights :106-131, computeSparseElimRanges :133-180 (thresholds 12 / 50 / 0.8 / x3),
// computeMerges :182-293 (merge iff modelled time of merged node < sum of the two),
// processTree :311-366, computeAggregateStruct :368-402. Written from scratch, Eigen-free.
#include "EliminationTree.h"
#include <algorithm>
#include <queue>
#include "DebugMacros.h"
#include "Utils.h"

namespace BaSpaCho {

using std::vector;

EliminationTree::EliminationTree(const vector<int64_t>& paramSize_, const SparseStructure& ss_,
                                 const ComputationModel* cm)
    : paramSize(paramSize_), ss(ss_), compMod(cm ? *cm : ComputationModel::model_OpenBlas_i7_1185g7) {
  BASPACHO_CHECK_EQ(paramSize.size(), ss.ptrs.size() - 1);
}

void EliminationTree::buildTree() {
  const int64_t n = ss.order();
  parent.assign(n, -1);
  nodeSize = paramSize;
  nodeRows.assign(n, 0);
  nodeRowBlocks.assign(n, 0);
  perColNodes.assign(n, {});

  // row subtrees: row k of L contains every node on the etree paths from the nonzeros of
  // A(k,0:k) up to the first node already reached for this k
  vector<int64_t> stamp(n, -1);
  for (int64_t k = 0; k < n; k++) {
    stamp[k] = k;
    for (int64_t q = ss.ptrs[k]; q < ss.ptrs[k + 1]; q++) {
      for (int64_t i = ss.inds[q]; i < k && stamp[i] != k; i = parent[i]) {
        if (parent[i] < 0) parent[i] = k;
        stamp[i] = k;
        nodeRows[i] += paramSize[k];
        nodeRowBlocks[i]++;
        perColNodes[i].push_back(k);
      }
    }
  }

  sygeCosts.assign(n, Lin2{});
  asmblCosts.assign(n, Lin2{});
  perRowNodeStats.assign(n, {});
  for (int64_t col = 0; col < n; col++) {
    auto& rowsOfCol = perColNodes[col];
    rowsOfCol.push_back(col);
    std::sort(rowsOfCol.begin(), rowsOfCol.end());
    // walk the column bottom-up: each row block r generates one syrk/gemm of (rows from r down) x (r)
    int64_t rowsBelow = 0, blocksBelow = 0;
    for (auto it = rowsOfCol.rbegin(); it != rowsOfCol.rend(); ++it) {
      int64_t r = *it, sz = paramSize[r];
      sygeCosts[col] += compMod.sygeLinEst(double(rowsBelow + sz), double(sz));
      asmblCosts[col] += compMod.asmblLinEst(double(blocksBelow + 1));
      perRowNodeStats[r].push_back(NodeStats{col, 1, sz, blocksBelow, rowsBelow});
      rowsBelow += sz;
      blocksBelow++;
    }
  }
}

void EliminationTree::computeNodeHeights(const vector<int64_t>& noCrossPoints) {
  const int64_t n = ss.order();
  unmergedHeightNode.resize(n);
  forbidMerge.assign(n, false);
  vector<int64_t> height(n, 0);

  int64_t segBegin = 0;
  for (size_t seg = 0; seg <= noCrossPoints.size(); seg++) {
    int64_t segEnd = seg < noCrossPoints.size() ? noCrossPoints[seg] : n;
    for (int64_t k = segBegin; k < segEnd; k++) {
      unmergedHeightNode[k] = std::make_tuple(height[k], nodeSize[k], k);
      int64_t par = parent[k];
      if (par < 0) continue;
      if (par >= segEnd) forbidMerge[k] = true;  // merging would cross the barrier
      height[par] = std::max(height[par], height[k] + 1);
    }
    std::sort(unmergedHeightNode.begin() + segBegin, unmergedHeightNode.begin() + segEnd);
    segBegin = segEnd;
  }
}

static constexpr int64_t kMaxSparseElimNodeSize = 12;
static constexpr int64_t kMinNumSparseElimNodes = 50;

void EliminationTree::computeSparseElimRanges(const vector<int64_t>& noCrossPoints) {
  const int64_t n = ss.order();
  sparseElimRanges.push_back(0);

  int64_t segBegin = 0;
  for (size_t seg = 0; seg <= noCrossPoints.size(); seg++) {
    int64_t segEnd = seg < noCrossPoints.size() ? noCrossPoints[seg] : n;
    int64_t k0 = segBegin;
    while (k0 < segEnd) {
      // maximal run of small nodes with the same height, in the height-sorted order
      const int64_t h = std::get<0>(unmergedHeightNode[k0]);
      int64_t k1 = k0, easyMerges = 0;
      for (; k1 < segEnd && std::get<0>(unmergedHeightNode[k1]) == h &&
             std::get<1>(unmergedHeightNode[k1]) <= kMaxSparseElimNodeSize;
           k1++) {
        // NB (kept from the reference, EliminationTree.cpp:150-156): indexes parent[] by position
        int64_t p = parent[k1];
        if (p >= 0) {
          double fillAfterMerge = double(nodeRows[k1]) / double(nodeRows[p] + nodeSize[p]);
          if (fillAfterMerge > 0.8) easyMerges++;
        }
      }
      int64_t count = k1 - k0;
      if (count < kMinNumSparseElimNodes || count < easyMerges * 3) break;
      for (int64_t k = k0; k < k1; k++) forbidMerge[std::get<2>(unmergedHeightNode[k])] = true;
      sparseElimRanges.push_back(k1);
      k0 = k1;
    }
    if (k0 < segEnd) break;
    segBegin = segEnd;
  }
  if (sparseElimRanges.size() == 1) sparseElimRanges.clear();
}

void EliminationTree::computeMerges() {
  const int64_t n = ss.order();
  numMergedNodes.assign(n, 1);
  mergeWith.assign(n, -1);
  numMerges = 0;

  auto score = [&](int64_t k, int64_t p) { return double(nodeRows[k]) / double(nodeRows[p] + nodeSize[p]); };
  auto nodeCost = [&](double size, double rows, const Lin2& syge, const Lin2& asmbl, double merged) {
    return compMod.potrfEst(size) + compMod.trsmEst(size, rows) + syge.a + syge.b * size + asmbl.a + asmbl.b * merged;
  };

  std::priority_queue<std::tuple<double, int64_t, int64_t>> candidates;
  for (int64_t k = n - 1; k >= 0; k--) {
    if (forbidMerge[k] || parent[k] < 0) continue;
    candidates.emplace(score(k, parent[k]), k, parent[k]);
  }

  vector<NodeStats> mergedStats;
  while (!candidates.empty()) {
    auto [oldScore, k, pOld] = candidates.top();
    (void)oldScore;
    candidates.pop();

    int64_t p = pOld;
    while (mergeWith[p] >= 0) p = mergeWith[p];
    if (p != pOld) {  // parent got merged upward meanwhile: re-score against its root
      candidates.emplace(score(k, p), k, p);
      continue;
    }

    double sk = double(nodeSize[k]), rk = double(nodeRows[k]);
    double sp = double(nodeSize[p]), rp = double(nodeRows[p]);
    double tk = nodeCost(sk, rk, sygeCosts[k], asmblCosts[k], double(numMergedNodes[k]));
    double tp = nodeCost(sp, rp, sygeCosts[p], asmblCosts[p], double(numMergedNodes[p]));
    double tm = nodeCost(sk + sp, rp, sygeCosts[p], asmblCosts[p], double(numMergedNodes[k] + numMergedNodes[p]));
    if (!(tm < tk + tp)) continue;

    const int64_t pSizeBefore = nodeSize[p], pMergedBefore = numMergedNodes[p];
    mergeWith[k] = p;
    nodeSize[p] += nodeSize[k];
    numMergedNodes[p] += numMergedNodes[k];
    numMerges++;

    // Row view of k and p: columns that had entries in row k and/or row p now see a single, taller
    // row block; update those columns' modelled costs and build the merged row list (sorted by column).
    const auto& rowK = perRowNodeStats[k];
    const auto& rowP = perRowNodeStats[p];
    mergedStats.clear();
    size_t ik = 0, ip = 0;
    while (ik < rowK.size() || ip < rowP.size()) {
      bool takeK = ip >= rowP.size() || (ik < rowK.size() && rowK[ik].colIdx < rowP[ip].colIdx);
      bool takeP = ik >= rowK.size() || (ip < rowP.size() && rowP[ip].colIdx < rowK[ik].colIdx);
      if (takeK) {
        if (rowK[ik].colIdx != k) mergedStats.push_back(rowK[ik]);
        ik++;
      } else if (takeP) {
        if (rowP[ip].colIdx != p) mergedStats.push_back(rowP[ip]);
        ip++;
      } else {
        const NodeStats& a = rowK[ik];
        const NodeStats& b = rowP[ip];
        int64_t c = b.colIdx;
        sygeCosts[c] -= compMod.sygeLinEst(double(a.rowsDown + a.rows), double(a.rows));
        asmblCosts[c] -= compMod.asmblLinEst(double(a.rBlocksDown + a.rBlocks));
        sygeCosts[c] -= compMod.sygeLinEst(double(b.rowsDown + b.rows), double(b.rows));
        asmblCosts[c] -= compMod.asmblLinEst(double(b.rBlocksDown + b.rBlocks));
        int64_t rows = a.rows + b.rows, blocks = a.rBlocks + b.rBlocks;
        sygeCosts[c] += compMod.sygeLinEst(double(b.rowsDown + rows), double(rows));
        asmblCosts[c] += compMod.asmblLinEst(double(b.rBlocksDown + blocks));
        mergedStats.push_back(NodeStats{c, blocks, rows, b.rBlocksDown, b.rowsDown});
        ik++, ip++;
      }
    }
    // the merged node's own diagonal entry
    sygeCosts[p] -= compMod.sygeLinEst(double(nodeRows[p] + pSizeBefore), double(pSizeBefore));
    asmblCosts[p] -= compMod.asmblLinEst(double(nodeRowBlocks[p] + pMergedBefore));
    sygeCosts[p] += compMod.sygeLinEst(double(nodeRows[p] + nodeSize[p]), double(nodeSize[p]));
    asmblCosts[p] += compMod.asmblLinEst(double(nodeRowBlocks[p] + numMergedNodes[p]));
    mergedStats.push_back(NodeStats{p, numMergedNodes[p], nodeSize[p], nodeRowBlocks[p], nodeRows[p]});
    perRowNodeStats[p].swap(mergedStats);
  }
}

void EliminationTree::collapseMergePointers() {
  // parents have larger indices, so a descending sweep makes every pointer reach its root
  for (int64_t k = ss.order() - 1; k >= 0; k--) {
    int64_t p = mergeWith[k];
    if (p >= 0 && mergeWith[p] >= 0) mergeWith[k] = mergeWith[p];
  }
}

void EliminationTree::processTree(bool detectSparseElimRanges, const vector<int64_t>& noCrossPoints,
                                  bool findOnlyElims) {
  const int64_t n = ss.order();
  computeNodeHeights(noCrossPoints);
  if (detectSparseElimRanges) computeSparseElimRanges(noCrossPoints);

  if (findOnlyElims) {
    mergeWith.assign(n, -1);
    numMergedNodes.assign(n, 1);
    numMerges = 0;
  } else {
    computeMerges();
    collapseMergePointers();
  }

  // lumps = merge roots, visited in (height, size, index) order
  const int64_t numLumps = n - numMerges;
  lumpStart.assign(numLumps + 1, 0);
  lumpToSpan.assign(numLumps + 1, 0);
  vector<int64_t> rootToLump(n, -1);
  int64_t lump = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t k = std::get<2>(unmergedHeightNode[i]);
    if (mergeWith[k] >= 0) continue;
    rootToLump[k] = lump;
    lumpStart[lump] = nodeSize[k];
    lumpToSpan[lump] = numMergedNodes[k];
    lump++;
  }
  BASPACHO_CHECK_EQ(lump, numLumps);
  cumSumVec(lumpStart);
  cumSumVec(lumpToSpan);

  // spans of a lump keep their relative (index) order
  permInverse.resize(n);
  vector<int64_t> cursor(lumpToSpan.begin(), lumpToSpan.end() - 1);
  for (int64_t i = 0; i < n; i++) {
    int64_t root = mergeWith[i] >= 0 ? mergeWith[i] : i;
    permInverse[i] = cursor[rootToLump[root]]++;
  }
}

void EliminationTree::computeAggregateStruct(bool fillOnlyForElims) {
  const int64_t n = ss.order();
  const int64_t numLumps = n - numMerges;

  SparseStructure filled = ss.symmetricPermutation(permInverse, /*lowerHalf=*/false, /*sortIndices=*/false);
  if (fillOnlyForElims) {
    for (size_t e = 0; e + 1 < sparseElimRanges.size(); e++)
      filled = filled.addIndependentEliminationFill(sparseElimRanges[e], sparseElimRanges[e + 1]);
  } else {
    filled = filled.addFullEliminationFill();
  }
  SparseStructure byCol = filled.transpose();

  // union of the row sets of the spans of each lump
  vector<int64_t> lastLump(n, -1);
  colStart.assign(1, 0);
  rowParam.clear();
  for (int64_t a = 0; a < numLumps; a++) {
    for (int64_t i = byCol.ptrs[lumpToSpan[a]]; i < byCol.ptrs[lumpToSpan[a + 1]]; i++) {
      int64_t r = byCol.inds[i];
      if (lastLump[r] < a) {
        lastLump[r] = a;
        rowParam.push_back(r);
      }
    }
    std::sort(rowParam.begin() + colStart.back(), rowParam.end());
    colStart.push_back((int64_t)rowParam.size());
  }
}

vector<int64_t> EliminationTree::computeSpanStart() {
  vector<int64_t> spanStart(paramSize.size() + 1, 0);
  leftPermute(spanStart.begin(), permInverse, paramSize);
  cumSumVec(spanStart);
  return spanStart;
}

}  // namespace BaSpaCho
