// ORACLE / TEST INFRASTRUCTURE ONLY.
// One set of host-side scenarios compiled TWICE: once against the reference's own sources where they lie under
// /root/reference (oracle/_ref/libref_host.so, built by `make -C oracle ref`) and once against this repo's restatement
// (inside liboracle_cpu.so). Both expose `<prefix>hostcheck(id, params, n, out, cap)`; tests/test_ref_objects.py asks
// both for the same scenario and requires bit-identical output (integers as they are, floating point by bit pattern).
// The including file provides: SparseStructure, testing_utils::*, composePermutations / inversePermutation /
// cumSumVec / rewindVec / bisect, and cholesky / solveUpperT / solveUpper / toOrderedPair - all in namespace BaSpaCho
// under the reference's names (Utils.h, SparseStructure.h, MathUtils.h, testing/TestingUtils.h, testing/TestingMatGen.h).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace hostcheck {

using namespace BaSpaCho;
using namespace BaSpaCho::testing_utils;
using Out = std::vector<int64_t>;

template <typename T>
void putBits(Out& o, const std::vector<T>& v) {
  for (T x : v) {
    int64_t bits = 0;
    std::memcpy(&bits, &x, sizeof(T));
    o.push_back(bits);
  }
}
inline void putInts(Out& o, const std::vector<int64_t>& v) { o.insert(o.end(), v.begin(), v.end()); }
inline void putStruct(Out& o, const SparseStructure& s) {
  o.push_back((int64_t)s.ptrs.size());
  putInts(o, s.ptrs);
  o.push_back((int64_t)s.inds.size());
  putInts(o, s.inds);
}
inline void putColumns(Out& o, const std::vector<std::set<int64_t>>& cols) {
  o.push_back((int64_t)cols.size());
  for (const auto& c : cols) {
    o.push_back((int64_t)c.size());
    o.insert(o.end(), c.begin(), c.end());
  }
}

inline Out run(int id, const double* p, int np) {
  auto P = [&](int i) -> double {
    if (i >= np) throw std::runtime_error("hostcheck: too few parameters");
    return p[i];
  };
  auto I = [&](int i) -> int64_t { return (int64_t)P(i); };
  Out o;
  switch (id) {
    case 0: putBits(o, randomData<double>((size_t)I(0), P(1), P(2), I(3))); break;
    case 1: putBits(o, randomData<float>((size_t)I(0), (float)P(1), (float)P(2), I(3))); break;
    case 2: putInts(o, randomVec((size_t)I(0), I(1), I(2), I(3))); break;
    case 3: putInts(o, randomPermutation((size_t)I(0), I(1))); break;
    case 4: putInts(o, randomPartition(I(0), I(1), I(2), I(3))); break;
    case 5: {  // randomCols -> CSC -> transpose (the problem family of the factor / solve tests)
      auto cols = randomCols(I(0), P(1), I(2));
      putColumns(o, cols);
      SparseStructure csc = columnsToCscStruct(cols);
      putStruct(o, csc);
      putStruct(o, csc.transpose());
      putColumns(o, csrStructToColumns(csc));
      break;
    }
    case 6: { auto g = SparseMatGenerator::genFlat(I(0), P(1), I(2)); putColumns(o, g.columns); putInts(o, randomVec(g.columns.size(), 2, 5, g.gen)); break; }
    case 7: { auto g = SparseMatGenerator::genGrid(I(0), I(1), P(2), I(3), I(4)); putColumns(o, g.columns); break; }
    case 8: { auto g = SparseMatGenerator::genMeridians(I(0), I(1), P(2), I(3), I(4), I(5), I(6), I(7)); putColumns(o, g.columns); break; }
    case 9: { auto g = SparseMatGenerator::genFlat(I(0), P(1), I(4)); g.addSchurSet(I(2), P(3)); putColumns(o, g.columns); break; }
    case 10: { auto g = SparseMatGenerator::genLine(I(0), P(1), I(2), I(3)); putColumns(o, g.columns); break; }
    case 11: {  // symmetricPermutation of a random lower pattern, both halves, sorted and unsorted
      SparseStructure ss = columnsToCscStruct(randomCols(I(0), P(1), I(2))).transpose();
      auto perm = randomPermutation((size_t)I(0), I(3));
      putStruct(o, ss.symmetricPermutation(perm, true, true));
      putStruct(o, ss.symmetricPermutation(perm, false, true));
      putStruct(o, ss.symmetricPermutation(perm, I(4) != 0, false));
      putStruct(o, ss.clear(true));
      putStruct(o, ss.clear(false));
      putInts(o, inversePermutation(perm));
      putInts(o, composePermutations(perm, randomPermutation((size_t)I(0), I(3) + 1)));
      break;
    }
    case 12: {  // elimination fill
      SparseStructure ss = columnsToCscStruct(randomCols(I(0), P(1), I(2))).transpose();
      putStruct(o, ss.addIndependentEliminationFill(I(3), I(4), true));
      putStruct(o, ss.addFullEliminationFill());
      putStruct(o, ss.extractRightBottom(I(3)));
      break;
    }
    case 13: {  // the naive fill the reference's own tests compare against + independent elimination sets
      auto cols = randomCols(I(0), P(1), I(2));
      auto ind = makeIndependentElimSet(cols, I(3), I(4));
      putColumns(o, ind);
      naiveAddEliminationEntries(ind, I(3), I(4));
      putColumns(o, ind);
      auto lumpStart = randomPartition(I(0), 1, 4, I(2));  // sizes -> starts (+ sentinel)
      lumpStart.push_back(0);
      cumSumVec(lumpStart);
      putColumns(o, joinColums(cols, lumpStart));
      break;
    }
    case 14: {  // small-block Cholesky and solves on a seeded SPD block (the per-point work of the sparse elimination)
      const int n = (int)I(0), lda = (int)I(1);
      auto a = randomData<double>((size_t)(lda * n), -1.0, 1.0, I(2));
      for (int i = 0; i < n; i++) a[(size_t)i * lda + i] += 2.0 * n;
      auto v = randomData<double>((size_t)n, -1.0, 1.0, I(2) + 1), w = v;
      cholesky(a.data(), lda, n);
      solveUpperT(a.data(), lda, n, v.data());
      solveUpper(a.data(), lda, n, w.data());
      putBits(o, a), putBits(o, v), putBits(o, w);
      auto af = randomData<float>((size_t)(lda * n), -1.0f, 1.0f, I(2));
      for (int i = 0; i < n; i++) af[(size_t)i * lda + i] += 2.0f * n;
      auto vf = randomData<float>((size_t)n, -1.0f, 1.0f, I(2) + 1), wf = vf;
      cholesky(af.data(), lda, n);
      solveUpperT(af.data(), lda, n, vf.data());
      solveUpper(af.data(), lda, n, wf.data());
      putBits(o, af), putBits(o, vf), putBits(o, wf);
      break;
    }
    case 15: {  // pair enumeration of the elimination kernel + bisect + the cumulative-sum helpers
      const int64_t n = I(0);
      for (int64_t q = 0; q < n * (n + 1) / 2; q++) {
        auto xy = toOrderedPair(n, q);
        o.push_back(xy.first), o.push_back(xy.second);
      }
      auto v = randomVec((size_t)n, 0, 7, I(1));
      o.push_back(cumSumVec(v));
      putInts(o, v);
      for (int64_t needle = -1; needle <= (v.empty() ? 0 : v.back()) + 1; needle++)
        o.push_back(bisect(v.data(), (int64_t)v.size(), needle));
      rewindVec(v, I(2), I(3));
      putInts(o, v);
      break;
    }
    default: throw std::runtime_error("hostcheck: unknown scenario");
  }
  return o;
}

inline int64_t entry(int id, const double* params, int n, int64_t* out, int64_t cap) {
  try {
    Out o = run(id, params, n);
    if (out) std::memcpy(out, o.data(), sizeof(int64_t) * (size_t)std::min<int64_t>(cap, (int64_t)o.size()));
    return (int64_t)o.size();
  } catch (const std::exception&) {
    return -1;
  }
}

}  // namespace hostcheck
