// Block sparsity pattern in compressed form (CSR or CSC depending on the caller's convention) and
// the pattern algebra the symbolic analysis needs. Each entry stands for a whole parameter block.
// API mirrors reference baspacho/baspacho/SparseStructure.h:20-56 (same names / argument meaning).
#pragma once

#include <cstdint>
#include <vector>

namespace BaSpaCho {

struct SparseStructure {
  std::vector<int64_t> ptrs;
  std::vector<int64_t> inds;

  SparseStructure() {}
  SparseStructure(std::vector<int64_t>&& p, std::vector<int64_t>&& i) : ptrs(std::move(p)), inds(std::move(i)) {}
  SparseStructure(const std::vector<int64_t>& p, const std::vector<int64_t>& i) : ptrs(p), inds(i) {}

  int64_t order() const { return (int64_t)ptrs.size() - 1; }

  void sortIndices();

  SparseStructure transpose() const;

  // drop the strictly-upper (clearLower=false ... ) see .cpp; keeps the diagonal
  SparseStructure clear(bool clearLower = true) const;

  // Input holds one (any) triangle. Entry (i,j) moves to (mapPerm[i], mapPerm[j]) and is stored in
  // the lower (lowerHalf=true) or upper triangle, column/row-compressed by the smaller/larger index.
  SparseStructure symmetricPermutation(const std::vector<int64_t>& mapPerm, bool lowerHalf = true,
                                       bool sortIndices = true) const;

  // CSR lower-triangular input: add the fill created by eliminating the mutually independent
  // nodes [start,end)
  SparseStructure addIndependentEliminationFill(int64_t start, int64_t end, bool sortIdx = true) const;

  // CSR lower-triangular input: add the fill of a complete symbolic Cholesky
  SparseStructure addFullEliminationFill() const;

  // perm[i] = old index that goes to position i (from-scratch approximate minimum degree)
  std::vector<int64_t> fillReducingPermutation() const;

  SparseStructure extractRightBottom(int64_t start);
};

// approximate-minimum-degree ordering of a symmetric pattern given as any mix of
// lower/upper entries (diagonal ignored). Returns perm (new position -> old index).
std::vector<int64_t> approximateMinimumDegree(int64_t n, const std::vector<int64_t>& ptrs,
                                              const std::vector<int64_t>& inds);

}  // namespace BaSpaCho
