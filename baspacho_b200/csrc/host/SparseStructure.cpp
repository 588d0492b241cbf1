// Pattern algebra for the symbolic phase. Behaviour follows reference
// baspacho/baspacho/SparseStructure.cpp:24-373 (transpose :36-69, clear :72-110,
// symmetricPermutation :112-159, addIndependentEliminationFill :161-222,
// addFullEliminationFill :224-289, extractRightBottom :334-373); written from scratch.
#include "SparseStructure.h"
#include <algorithm>
#include "DebugMacros.h"
#include "Utils.h"

namespace BaSpaCho {

using std::vector;

namespace {
// Generic "count, prefix-sum, fill" builder: calls emit(cb) twice, where cb(bucket, value)
// registers one entry.
template <typename Emit>
SparseStructure bucketBuild(int64_t numBuckets, Emit&& emit) {
  SparseStructure out;
  out.ptrs.assign(numBuckets + 1, 0);
  emit([&](int64_t b, int64_t) { out.ptrs[b]++; });
  out.inds.resize(cumSumVec(out.ptrs));
  vector<int64_t> cursor(out.ptrs.begin(), out.ptrs.end() - 1);
  emit([&](int64_t b, int64_t v) { out.inds[cursor[b]++] = v; });
  return out;
}
}  // namespace

void SparseStructure::sortIndices() {
  for (int64_t i = 0, n = order(); i < n; i++) std::sort(inds.begin() + ptrs[i], inds.begin() + ptrs[i + 1]);
}

SparseStructure SparseStructure::transpose() const {
  int64_t n = order();
  return bucketBuild(n, [&](auto&& put) {
    for (int64_t i = 0; i < n; i++)
      for (int64_t k = ptrs[i]; k < ptrs[i + 1]; k++) {
        BASPACHO_CHECK_LT(inds[k], n);
        put(inds[k], i);
      }
  });
}

SparseStructure SparseStructure::clear(bool lowerHalf) const {
  int64_t n = order();
  return bucketBuild(n, [&](auto&& put) {
    for (int64_t i = 0; i < n; i++)
      for (int64_t k = ptrs[i]; k < ptrs[i + 1]; k++) {
        int64_t j = inds[k];
        BASPACHO_CHECK_LT(j, n);
        bool dropped = (i != j) && ((j > i) == lowerHalf);
        if (!dropped) put(i, j);
      }
  });
}

SparseStructure SparseStructure::symmetricPermutation(const vector<int64_t>& mapPerm, bool lowerHalf,
                                                      bool sortIdx) const {
  int64_t n = order();
  BASPACHO_CHECK_EQ(n, (int64_t)mapPerm.size());
  SparseStructure out = bucketBuild(n, [&](auto&& put) {
    for (int64_t i = 0; i < n; i++) {
      int64_t pi = mapPerm[i];
      BASPACHO_CHECK_LT(pi, n);
      for (int64_t k = ptrs[i]; k < ptrs[i + 1]; k++) {
        BASPACHO_CHECK_LT(inds[k], n);
        int64_t pj = mapPerm[inds[k]];
        BASPACHO_CHECK_LT(pj, n);
        int64_t lo = std::min(pi, pj), hi = std::max(pi, pj);
        if (lowerHalf) put(lo, hi); else put(hi, lo);
      }
    }
  });
  if (sortIdx) out.sortIndices();
  return out;
}

SparseStructure SparseStructure::addIndependentEliminationFill(int64_t elimStart, int64_t elimEnd,
                                                               bool sortIdx) const {
  int64_t n = order();
  if (elimEnd == n) return *this;  // nothing below the eliminated set: no fill

  // column view of the eliminated nodes: rows below each of them, ascending
  SparseStructure cols = transpose();
  for (int64_t c = elimStart; c < elimEnd; c++) std::sort(cols.inds.begin() + cols.ptrs[c], cols.inds.begin() + cols.ptrs[c + 1]);

  SparseStructure out;
  out.ptrs.assign(ptrs.begin(), ptrs.begin() + elimEnd + 1);
  out.inds.assign(inds.begin(), inds.begin() + ptrs[elimEnd]);

  // row k (>= elimEnd) gains every row w<k that shares an eliminated column with it
  vector<int64_t> seenInRow(n, -1);
  for (int64_t k = elimEnd; k < n; k++) {
    seenInRow[k] = k;
    out.inds.push_back(k);
    for (int64_t q = ptrs[k]; q < ptrs[k + 1]; q++) {
      int64_t c = inds[q];
      if (c >= k) continue;
      if (seenInRow[c] != k) {
        seenInRow[c] = k;
        out.inds.push_back(c);
      }
      if (c < elimStart || c >= elimEnd) continue;
      for (int64_t t = cols.ptrs[c]; t < cols.ptrs[c + 1]; t++) {
        int64_t w = cols.inds[t];
        if (w >= k) break;
        if (seenInRow[w] < k) {
          seenInRow[w] = k;
          out.inds.push_back(w);
        }
      }
    }
    out.ptrs.push_back((int64_t)out.inds.size());
  }
  if (sortIdx) out.sortIndices();
  return out;
}

SparseStructure SparseStructure::addFullEliminationFill() const {
  // Row-by-row symbolic factorization: pattern of row k of L = nodes met walking the elimination
  // tree upward from each nonzero of A(k, 0:k) until an already visited node (Liu's row-subtree
  // characterisation; the parent array is discovered on the fly).
  int64_t n = order();
  vector<int64_t> parent(n, -1), visited(n, -1);
  vector<vector<int64_t>> rows(n);
  for (int64_t k = 0; k < n; k++) {
    visited[k] = k;
    rows[k].push_back(k);
    for (int64_t q = ptrs[k]; q < ptrs[k + 1]; q++) {
      int64_t i = inds[q];
      if (i >= k) continue;
      while (visited[i] != k) {
        if (parent[i] < 0) parent[i] = k;
        visited[i] = k;
        rows[k].push_back(i);
        i = parent[i];
      }
    }
  }
  SparseStructure out;
  out.ptrs.assign(n + 1, 0);
  for (int64_t k = 0; k < n; k++) out.ptrs[k] = (int64_t)rows[k].size();
  out.inds.resize(cumSumVec(out.ptrs));
  for (int64_t k = 0; k < n; k++) {
    std::sort(rows[k].begin(), rows[k].end());
    std::copy(rows[k].begin(), rows[k].end(), out.inds.begin() + out.ptrs[k]);
  }
  return out;
}

std::vector<int64_t> SparseStructure::fillReducingPermutation() const {
  return approximateMinimumDegree(order(), ptrs, inds);
}

SparseStructure SparseStructure::extractRightBottom(int64_t startRow) {
  int64_t n = order();
  BASPACHO_CHECK_LE(startRow, n);
  BASPACHO_CHECK_GE(startRow, 0);
  return bucketBuild(n - startRow, [&](auto&& put) {
    for (int64_t i = startRow; i < n; i++)
      for (int64_t k = ptrs[i]; k < ptrs[i + 1]; k++) {
        BASPACHO_CHECK_LT(inds[k], n);
        if (inds[k] >= startRow) put(i - startRow, inds[k] - startRow);
      }
  });
}

}  // namespace BaSpaCho
