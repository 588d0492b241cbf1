// Native tests of the host layer (symbolic analysis, skeleton, accessors, createSolver) - the cases of the reference's
// own gtest files restated without gtest/Eigen (neither is in this image):
//   SparseStructureTest.cpp:20-152   Transpose / SymPermutation golden vectors, elimination fill against the naive fill,
//                                    fill bound of the fill-reducing permutation
//   EliminationTreeTest.cpp:25-82    tree + merges + aggregate structure contain the original and the filled pattern
//   AccessorTest.cpp:30-152          plain and permuted block accessors address the blocks densify() shows
//   CoalescedBlockMatrixTest.cpp:112-215  densify golden matrices (full and "fill upper half from span 1"), damp
//   CreateSolverTest.cpp:44-156      createSolver with the four fill policies, given elimination ranges and "eliminate
//                                    last" ids: partial factor == dense Cholesky of the leading block + Schur complement
// The numeric backend under createSolver here is the CPU checker's naive one (oracle/CpuOps.cpp, BackendRef).
// Built and run by tests/test_host_cpp.py:  make -C tests/cpp && tests/cpp/host_tests
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../baspacho_b200/csrc/host/CoalescedBlockMatrix.h"
#include "../../baspacho_b200/csrc/host/EliminationTree.h"
#include "../../baspacho_b200/csrc/host/Solver.h"
#include "../../baspacho_b200/csrc/host/SparseStructure.h"
#include "../../baspacho_b200/csrc/host/Utils.h"
#include "../../baspacho_b200/csrc/testing/TestingUtils.h"
#include "../../oracle/BlasLoader.h"
#include <functional>
#include <memory>

using namespace BaSpaCho;
using namespace BaSpaCho::testing_utils;
using std::set;
using std::vector;

namespace BaSpaCho {
OpsPtr oracleRefOps();
OpsPtr oracleFastOps(int numThreads);
OpsPtr b200Ops() { throw std::runtime_error("host tests: no device backend"); }
}  // namespace BaSpaCho

static int g_failures = 0, g_checks = 0;
static std::string g_test;
#define CHECK(cond)                                                                                   \
  do {                                                                                                \
    g_checks++;                                                                                       \
    if (!(cond)) {                                                                                    \
      g_failures++;                                                                                   \
      std::printf("FAIL [%s] %s:%d  %s\n", g_test.c_str(), __FILE__, __LINE__, #cond);                \
      return;                                                                                         \
    }                                                                                                 \
  } while (0)
#define CHECK_VEC(a, ...) CHECK((a) == (vector<int64_t>{__VA_ARGS__}))

// ---------------------------------------------------------------------------------------------- SparseStructure
static void testTranspose() {
  SparseStructure ss({0, 2, 4, 7, 9, 11}, {0, 3, 2, 4, 0, 1, 4, 1, 2, 2, 4});
  SparseStructure t = ss.transpose();
  CHECK_VEC(t.ptrs, 0, 2, 4, 7, 8, 11);
  CHECK_VEC(t.inds, 0, 2, 2, 3, 1, 3, 4, 0, 1, 2, 4);
}

static void testSymPermutation() {
  SparseStructure ss({0, 1, 2, 4, 6, 8, 12}, {0, 1, 0, 1, 1, 3, 2, 4, 0, 1, 4, 5});
  SparseStructure lowerCsr = ss.symmetricPermutation({4, 5, 2, 1, 0, 3}, /*lowerHalf=*/false);
  CHECK_VEC(lowerCsr.ptrs, 0, 1, 2, 3, 5, 8, 12);
  CHECK_VEC(lowerCsr.inds, 0, 1, 0, 0, 3, 2, 3, 4, 1, 2, 3, 5);
  SparseStructure lowerCsc = ss.symmetricPermutation({4, 5, 2, 1, 0, 3}, /*lowerHalf=*/true);
  CHECK_VEC(lowerCsc.ptrs, 0, 3, 5, 7, 10, 11, 12);
  CHECK_VEC(lowerCsc.inds, 0, 2, 3, 1, 5, 4, 5, 3, 4, 5, 4, 5);
}

static void testEliminationFills() {
  int seed = 37;
  for (int64_t size : {10, 20, 30, 40})
    for (double fill : {0.15, 0.23, 0.3}) {
      ColumnSets original = randomCols(size, fill, seed++);
      {  // full fill against the naive column-by-column fill
        ColumnSets cols = original;
        SparseStructure ss = columnsToCscStruct(cols).transpose();
        naiveAddEliminationEntries(cols, 0, size);
        SparseStructure expected = columnsToCscStruct(cols).transpose();
        SparseStructure got = ss.addFullEliminationFill();
        CHECK(expected.ptrs == got.ptrs);
        CHECK(expected.inds == got.inds);
      }
      for (int64_t start = 0; start < size * 2 / 3; start += 3)
        for (int64_t end = start + 3; end < size; end += 3) {
          ColumnSets base = original;
          ColumnSets cols = makeIndependentElimSet(base, start, end);
          SparseStructure ss = columnsToCscStruct(cols).transpose();
          naiveAddEliminationEntries(cols, start, end);
          SparseStructure expected = columnsToCscStruct(cols).transpose();
          SparseStructure got = ss.addIndependentEliminationFill(start, end);
          CHECK(expected.ptrs == got.ptrs);
          CHECK(expected.inds == got.inds);
        }
    }
}

static void testFillReducingPermutation() {
  // 24-node pattern of SparseStructureTest.cpp:117-152 (both halves listed; cleared to the lower half)
  vector<int64_t> ptrs{0, 9, 15, 21, 27, 33, 39, 48, 57, 61, 70, 76, 82, 88, 94, 100, 106, 110, 119, 128, 137, 143, 152, 156, 160};
  vector<int64_t> inds{0, 5, 6, 12, 13, 17, 18, 19, 21, 1, 8, 9, 13, 14, 17, 2, 6, 11, 20, 21, 22, 3, 7, 10, 15, 18, 19,
                       4, 7, 9, 14, 15, 16, 0, 5, 6, 12, 13, 17, 0, 2, 5, 6, 11, 12, 19, 21, 23, 3, 4, 7, 9, 14, 15, 16, 17, 18,
                       1, 8, 9, 14, 1, 4, 7, 8, 9, 13, 14, 17, 18, 3, 10, 18, 19, 20, 21, 2, 6, 11, 12, 21, 23,
                       0, 5, 6, 11, 12, 23, 0, 1, 5, 9, 13, 17, 1, 4, 7, 8, 9, 14, 3, 4, 7, 15, 16, 18, 4, 7, 15, 16,
                       0, 1, 5, 7, 9, 13, 17, 18, 19, 0, 3, 7, 9, 10, 15, 17, 18, 19, 0, 3, 6, 10, 17, 18, 19, 20, 21,
                       2, 10, 19, 20, 21, 22, 0, 2, 6, 10, 11, 19, 20, 21, 22, 2, 20, 21, 22, 6, 11, 12, 23};
  SparseStructure lower = SparseStructure(ptrs, inds).clear();
  vector<int64_t> perm = lower.fillReducingPermutation();
  vector<int64_t> sorted = perm;
  std::sort(sorted.begin(), sorted.end());
  for (int64_t i = 0; i < 24; i++) CHECK(sorted[i] == i);
  SparseStructure filled = lower.symmetricPermutation(inversePermutation(perm), false).addFullEliminationFill();
  CHECK((int64_t)filled.inds.size() <= 130);  // the reference's bound (its AMD reaches 120)
}

// ---------------------------------------------------------------------------------------------- EliminationTree
static void testEliminationTreeBuild() {
  for (int h = 0; h < 200; h++) {
    ColumnSets colsOrig = randomCols(70, 0.05, h + 37);
    SparseStructure ssOrig = columnsToCscStruct(colsOrig).transpose();
    vector<int64_t> invPerm = inversePermutation(ssOrig.fillReducingPermutation());
    SparseStructure ss = ssOrig.symmetricPermutation(invPerm, false);
    vector<int64_t> paramSize(ssOrig.order(), 1);
    EliminationTree et(paramSize, ss);
    const int64_t nocross = (7 * h) % 60 + 5;  // a merge barrier: a lump must start exactly there
    et.buildTree();
    et.processTree(/*detectSparseElimRanges=*/false, {nocross});
    et.computeAggregateStruct();
    CoalescedBlockMatrixSkel skel(et.computeSpanStart(), et.lumpToSpan, et.colStart, et.rowParam);
    CHECK(skel.spanOffsetInLump[nocross] == 0);

    vector<double> ones(skel.dataSize(), 1.0);
    vector<double> mat = skel.densify(ones);
    const int64_t n = skel.order();
    auto at = [&](int64_t r, int64_t c) { return mat[r * n + c]; };
    // every original entry is present after both permutations
    vector<int64_t> idMap = composePermutations(et.permInverse, invPerm);
    for (int64_t i = 0; i < ssOrig.order(); i++)
      for (int64_t q = ssOrig.ptrs[i]; q < ssOrig.ptrs[i + 1]; q++) {
        int64_t a = idMap[i], b = idMap[ssOrig.inds[q]];
        CHECK(at(std::max(a, b), std::min(a, b)) > 0.5);
      }
    // and so is every entry of the filled pattern in the tree's ordering
    SparseStructure filled = ss.symmetricPermutation(et.permInverse, false, true).addFullEliminationFill();
    for (int64_t i = 0; i < filled.order(); i++)
      for (int64_t q = filled.ptrs[i]; q < filled.ptrs[i + 1]; q++) CHECK(at(i, filled.inds[q]) > 0.5);
  }
}

// ---------------------------------------------------------------------------------------------- skeleton fixtures
static CoalescedBlockMatrixSkel nineSpanSkel() {
  vector<int64_t> spanStart{0, 1, 2, 4, 5, 7, 9, 12, 14, 16}, lumpToSpan{0, 1, 3, 4, 6, 7, 9};
  ColumnSets cols{{0, 1, 2, 5, 8}, {1, 2, 3, 6, 7}, {3, 4, 5, 8}, {4, 5, 7}, {6, 8}, {7, 8}};
  SparseStructure s = columnsToCscStruct(cols);
  return CoalescedBlockMatrixSkel(spanStart, lumpToSpan, s.ptrs, s.inds);
}

static const double kDensifyGolden[16][16] = {
    {13, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},          {14, 21, 22, 23, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {15, 24, 25, 26, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},       {16, 27, 28, 29, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 30, 31, 32, 48, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},       {0, 0, 0, 0, 49, 55, 56, 57, 58, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 50, 59, 60, 61, 62, 0, 0, 0, 0, 0, 0, 0},      {17, 0, 0, 0, 51, 63, 64, 65, 66, 0, 0, 0, 0, 0, 0, 0},
    {18, 0, 0, 0, 52, 67, 68, 69, 70, 0, 0, 0, 0, 0, 0, 0},     {0, 33, 34, 35, 0, 0, 0, 0, 0, 79, 80, 81, 0, 0, 0, 0},
    {0, 36, 37, 38, 0, 0, 0, 0, 0, 82, 83, 84, 0, 0, 0, 0},     {0, 39, 40, 41, 0, 0, 0, 0, 0, 85, 86, 87, 0, 0, 0, 0},
    {0, 42, 43, 44, 0, 71, 72, 73, 74, 0, 0, 0, 94, 95, 96, 97}, {0, 45, 46, 47, 0, 75, 76, 77, 78, 0, 0, 0, 98, 99, 100, 101},
    {19, 0, 0, 0, 53, 0, 0, 0, 0, 88, 89, 90, 102, 103, 104, 105},
    {20, 0, 0, 0, 54, 0, 0, 0, 0, 91, 92, 93, 106, 107, 108, 109}};

static void testDensifyAndDamp() {
  CoalescedBlockMatrixSkel skel = nineSpanSkel();
  vector<double> data(skel.dataSize());
  std::iota(data.begin(), data.end(), 13.0);
  vector<double> mat = skel.densify(data);
  for (int r = 0; r < 16; r++)
    for (int c = 0; c < 16; c++) CHECK(mat[r * 16 + c] == kDensifyGolden[r][c]);

  // from span 1 on, upper half mirrored: the 15 x 15 trailing block, symmetric, with the lower half of each diagonal block
  vector<double> sub(15 * 15, -1.0);
  skel.densify(sub.data(), data.data(), /*fillUpperHalf=*/true, /*startSpanIndex=*/1);
  for (int r = 0; r < 15; r++)
    for (int c = 0; c <= r; c++) {
      CHECK(sub[r * 15 + c] == kDensifyGolden[r + 1][c + 1] || (c > r));
      CHECK(sub[c * 15 + r] == sub[r * 15 + c]);
    }
  CHECK(sub[0 * 15 + 1] == 24 && sub[3 * 15 + 4] == 49 && sub[13 * 15 + 14] == 108);  // CoalescedBlockMatrixTest.cpp:168-183

  // damp: diag -> diag * (1 + alpha) + beta, everything else untouched
  vector<double> damped = data;
  skel.damp(damped, 2.0, 100.0);
  vector<double> matD = skel.densify(damped);
  for (int r = 0; r < 16; r++)
    for (int c = 0; c < 16; c++)
      CHECK(matD[r * 16 + c] == (r == c ? mat[r * 16 + c] * 3.0 + 100.0 : mat[r * 16 + c]));
}

// ---------------------------------------------------------------------------------------------- accessors
static void testAccessors() {
  ColumnSets colBlocks{{0, 3, 5}, {1}, {2, 4}, {3}, {4}, {5}};
  SparseStructure ss = columnsToCscStruct(colBlocks).transpose().addFullEliminationFill();
  vector<int64_t> spanStart{0, 2, 5, 7, 10, 12, 15}, lumpToSpan{0, 2, 4, 6};
  SparseStructure grouped = columnsToCscStruct(joinColums(csrStructToColumns(ss), lumpToSpan));
  CoalescedBlockMatrixSkel skel(spanStart, lumpToSpan, grouped.ptrs, grouped.inds);
  const int64_t n = skel.order();

  for (int permuted = 0; permuted < 2; permuted++) {
    vector<int64_t> perm(colBlocks.size());
    std::iota(perm.begin(), perm.end(), 0);
    if (permuted) {
      std::mt19937 g(37);
      std::shuffle(perm.begin(), perm.end(), g);
    }
    vector<int64_t> invP = inversePermutation(perm);
    PermutedCoalescedAccessor acc;
    acc.init(skel.accessor(), perm.data());
    vector<double> data(skel.dataSize(), 0.0), dense(n * n, 0.0);
    int seed = 0;
    for (int64_t pc = 0; pc < (int64_t)colBlocks.size(); pc++) {
      const int64_t c = invP[pc], cSize = acc.paramSize(c), cStart = acc.paramStart(c);
      for (int64_t pr : colBlocks[pc]) {
        const int64_t r = invP[pr], rSize = acc.paramSize(r), rStart = acc.paramStart(r);
        auto [off, stride, flip] = acc.blockOffset(r, c);
        CHECK(flip == (pr < pc));
        vector<double> blk = randomData<double>(rSize * cSize, -1.0, 1.0, seed++);
        for (int64_t a = 0; a < rSize; a++)
          for (int64_t b = 0; b < cSize; b++) {
            const double v = blk[a * cSize + b];
            if (pr == pc) {  // diagonal block: accumulate, as the reference test does
              data[off + a * stride + b] += v;
              dense[(rStart + a) * n + cStart + b] += v;
            } else if (!flip) {
              data[off + a * stride + b] = v;
              dense[(rStart + a) * n + cStart + b] = v;
            } else {  // stored transposed
              data[off + b * stride + a] = v;
              dense[(cStart + b) * n + rStart + a] = v;
            }
          }
      }
      auto [dOff, dStride] = acc.diagBlockOffset(c);
      auto [bOff, bStride, bFlip] = acc.blockOffset(c, c);
      CHECK(dOff == bOff && dStride == bStride && !bFlip);
    }
    vector<double> got = skel.densify(data);
    double err = 0;
    for (int64_t r = 0; r < n; r++)
      for (int64_t c = 0; c <= r; c++) err += std::fabs(got[r * n + c] - dense[r * n + c]);
    CHECK(err < 1e-12);
  }
}

// ---------------------------------------------------------------------------------------------- createSolver
static void denseCholeskyLeading(vector<double>& a, int64_t n, int64_t upTo) {
  // in-place partial Cholesky: columns [0, upTo) eliminated, trailing block = Schur complement (lower triangle)
  for (int64_t j = 0; j < upTo; j++) {
    double d = std::sqrt(a[j * n + j]);
    a[j * n + j] = d;
    for (int64_t i = j + 1; i < n; i++) a[i * n + j] /= d;
    for (int64_t c = j + 1; c < n; c++)
      for (int64_t r = c; r < n; r++) a[r * n + c] -= a[r * n + j] * a[c * n + j];
  }
}

template <typename T>
static void checkSolver(Solver& solver, int seed, const std::unordered_set<int64_t>& elimLastIds, double eps) {
  vector<T> data = randomData<T>(solver.dataSize(), T(-1.0), T(1.0), 9 + seed);
  solver.skel().damp(data, T(0.0), T(solver.order() * 2.0));
  const int64_t n = solver.order();
  vector<T> dense = solver.skel().densify(data);
  vector<double> expect(dense.begin(), dense.end());
  for (int64_t r = 0; r < n; r++)
    for (int64_t c = r + 1; c < n; c++) expect[r * n + c] = 0;
  const int64_t upToSpan = solver.canFactorUpToSpan();
  denseCholeskyLeading(expect, n, solver.skel().spanStart[upToSpan]);
  solver.factorUpTo(data.data(), upToSpan);
  vector<T> got = solver.skel().densify(data);
  double num = 0, den = 0;
  for (int64_t r = 0; r < n; r++)
    for (int64_t c = 0; c <= r; c++) {
      num += (expect[r * n + c] - got[r * n + c]) * (expect[r * n + c] - got[r * n + c]);
      den += expect[r * n + c] * expect[r * n + c];
    }
  CHECK(std::sqrt(num / den) < eps);
  if (!elimLastIds.empty()) {
    const int64_t s = (int64_t)elimLastIds.size();
    CHECK(solver.skel().spanOffsetInLump[solver.skel().numSpans() - s] == 0);
    for (int64_t e : elimLastIds) CHECK(solver.paramToSpan()[e] >= solver.skel().numSpans() - s);
  }
}

template <typename T>
static void testCreateSolver(bool elimSet, bool lastIds, int reps, double eps) {
  for (int i = 0; i < reps; i++) {
    const int numParams = 215;
    ColumnSets colBlocks = randomCols(numParams, 0.03, 57 + i);
    vector<int64_t> ranges;
    if (elimSet) {
      colBlocks = makeIndependentElimSet(colBlocks, 0, 150);
      ranges = {0, 90};
    } else {
      colBlocks = makeIndependentElimSet(colBlocks, 0, 60);
    }
    std::unordered_set<int64_t> last;
    if (lastIds) {
      last = {105, 123, 165, 194, 209, 214};
      if (!elimSet) last.insert({0, 30, 49, 87});
    }
    SparseStructure ss = columnsToCscStruct(colBlocks).transpose();
    vector<int64_t> paramSize = randomVec(ss.order(), 2, 3, 47);
    Settings st;
    st.backend = BackendRef;
    {
      st.addFillPolicy = AddFillComplete;
      auto solver = createSolver(st, paramSize, ss, ranges, last);
      CHECK(solver->canFactorUpToSpan() == numParams);
      checkSolver<T>(*solver, 4 * i + 0, last, eps);
    }
    if (lastIds) continue;
    {
      st.addFillPolicy = AddFillForAutoElims;
      auto solver = createSolver(st, paramSize, ss, ranges);
      CHECK(solver->canFactorUpToSpan() >= (elimSet ? 145 : 55));  // the ranges are found automatically
      checkSolver<T>(*solver, 4 * i + 1, {}, eps);
    }
    {
      st.addFillPolicy = AddFillForGivenElims;
      auto solver = createSolver(st, paramSize, ss, ranges);
      if (elimSet) CHECK(solver->canFactorUpToSpan() == 90);
      checkSolver<T>(*solver, 4 * i + 2, {}, eps);
    }
    {
      st.addFillPolicy = AddFillNone;
      auto solver = createSolver(st, paramSize, ss, ranges);
      CHECK(solver->canFactorUpToSpan() == 0);
      checkSolver<T>(*solver, 4 * i + 3, {}, eps);
    }
  }
}

// ---------------------------------------------------------------------------------------------- partial factor / solve
// PartialFactorSolveTest.cpp:48-560: a skeleton with a sparse-elimination range and a merge barrier at span `nocross`;
// every partial entry point of Solver against dense algebra on the densified matrix, for a CPU checker backend.
struct PartialCase {
  std::unique_ptr<Solver> solver;
  int64_t nocross, order, barrierAt;
};
static PartialCase makePartialCase(int i, const std::function<OpsPtr()>& genOps) {
  const int64_t kMinElim = 50;
  ColumnSets colBlocks = randomCols(215, 0.03, 57 + i);
  colBlocks = makeIndependentElimSet(colBlocks, 0, 150);
  SparseStructure sortedSs = columnsToCscStruct(colBlocks).transpose();
  PartialCase pc;
  pc.nocross = (7 * i) % (210 - kMinElim) + kMinElim + 1;
  vector<int64_t> paramSize = randomVec(sortedSs.order(), 2, 3, 47);
  EliminationTree et(paramSize, sortedSs);
  et.buildTree();
  et.processTree(/*detectSparseElimRanges=*/true, {pc.nocross});
  et.computeAggregateStruct();
  CoalescedBlockMatrixSkel skel(et.computeSpanStart(), et.lumpToSpan, et.colStart, et.rowParam);
  if (skel.spanOffsetInLump[pc.nocross] != 0 || et.sparseElimRanges.size() < 2) return pc;  // solver stays null -> CHECK fails
  pc.order = skel.order();
  pc.barrierAt = skel.spanStart[pc.nocross];
  vector<int64_t> ranges = et.sparseElimRanges;
  pc.solver.reset(new Solver(std::move(skel), std::move(ranges), {}, genOps()));
  return pc;
}
static double relLowerDiff(const vector<double>& a, const vector<double>& b, int64_t n) {
  double num = 0, den = 0;
  for (int64_t r = 0; r < n; r++)
    for (int64_t c = 0; c <= r; c++) num += (a[r * n + c] - b[r * n + c]) * (a[r * n + c] - b[r * n + c]), den += a[r * n + c] * a[r * n + c];
  return std::sqrt(num / den);
}
static double relDiff(const vector<double>& a, const vector<double>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); i++) num += (a[i] - b[i]) * (a[i] - b[i]), den += a[i] * a[i];
  return std::sqrt(num / den);
}

template <typename T>
static void testPartial(const std::function<OpsPtr()>& genOps, int reps, double eps) {
  for (int i = 0; i < reps; i++) {
    // ---- PartialFactor / SplitFactor / testPseudoFactor (matrix strongly damped)
    {
      PartialCase pc = makePartialCase(i, genOps);
      CHECK(pc.solver != nullptr);
      const int64_t n = pc.order;
      vector<T> data = randomData<T>(pc.solver->dataSize(), T(-1), T(1), 9 + i);
      pc.solver->skel().damp(data, T(0), T(n * 2.0));
      vector<T> d0 = pc.solver->skel().densify(data);
      vector<double> lower(d0.begin(), d0.end());
      for (int64_t r = 0; r < n; r++)
        for (int64_t c = r + 1; c < n; c++) lower[r * n + c] = 0;

      vector<double> marginal = lower, full = lower;
      denseCholeskyLeading(marginal, n, pc.barrierAt);
      denseCholeskyLeading(full, n, n);
      vector<T> part = data;
      pc.solver->factorUpTo(part.data(), pc.nocross);
      vector<T> gotPart = pc.solver->skel().densify(part);
      CHECK(relLowerDiff(marginal, vector<double>(gotPart.begin(), gotPart.end()), n) < eps);
      pc.solver->factorFrom(part.data(), pc.nocross);
      vector<T> gotFull = pc.solver->skel().densify(part);
      CHECK(relLowerDiff(full, vector<double>(gotFull.begin(), gotFull.end()), n) < eps);

      // pseudo factor: per span, Cholesky of the span's diagonal block and the rows below solved against it
      vector<double> pseudo = lower;
      const auto& sk = pc.solver->skel();
      for (int64_t j = 0; j < sk.numSpans(); j++) {
        const int64_t b = sk.spanStart[j], e = sk.spanStart[j + 1];
        for (int64_t c = b; c < e; c++) {  // in-block Cholesky, then rows below: x L^T = row
          double d = pseudo[c * n + c];
          for (int64_t q = b; q < c; q++) d -= pseudo[c * n + q] * pseudo[c * n + q];
          d = std::sqrt(d);
          pseudo[c * n + c] = d;
          for (int64_t r = c + 1; r < n; r++) {
            double v = pseudo[r * n + c];
            for (int64_t q = b; q < c; q++) v -= pseudo[r * n + q] * pseudo[c * n + q];
            pseudo[r * n + c] = v / d;
          }
        }
      }
      vector<T> ps = data;
      pc.solver->pseudoFactorFrom(ps.data(), 0);
      vector<T> gotPs = sk.densify(ps);
      CHECK(relLowerDiff(pseudo, vector<double>(gotPs.begin(), gotPs.end()), n) < eps);

      // addMvFrom: out[after] += sym(A[after, after]) in[after]
      const int nRHS = 3;
      vector<T> vin = randomData<T>(n * nRHS, T(-1), T(1), 49 + i), vout = vin;
      vector<double> ref(vout.begin(), vout.end());
      for (int c = 0; c < nRHS; c++)
        for (int64_t r = pc.barrierAt; r < n; r++) {
          double acc = 0;
          for (int64_t q = pc.barrierAt; q < n; q++) acc += (q <= r ? lower[r * n + q] : lower[q * n + r]) * vin[c * n + q];
          ref[c * n + r] += acc;
        }
      pc.solver->addMvFrom(data.data(), pc.nocross, vin.data(), n, vout.data(), n, nRHS);
      CHECK(relDiff(ref, vector<double>(vout.begin(), vout.end())) < eps);
    }
    // ---- PartialSolveL / Lt (UpTo) and (From): the data is used AS the factor (mildly damped, as the reference does)
    {
      PartialCase pc = makePartialCase(i, genOps);
      CHECK(pc.solver != nullptr);
      const int64_t n = pc.order, bar = pc.barrierAt;
      vector<T> data = randomData<T>(pc.solver->dataSize(), T(-1), T(1), 9 + i);
      pc.solver->skel().damp(data, T(0), T(3.0));
      vector<T> d0 = pc.solver->skel().densify(data);
      vector<double> L(d0.begin(), d0.end());
      const int nRHS = 3;
      for (int j = 0; j < 2; j++) {
        vector<T> v0 = randomData<T>(n * nRHS, T(-1), T(1), 49 + j + i);
        vector<double> b(v0.begin(), v0.end());
        // solveLUpTo: top = L11^-1 top ; bottom -= L21 top
        vector<double> ref = b;
        for (int c = 0; c < nRHS; c++) {
          double* x = ref.data() + c * n;
          for (int64_t r = 0; r < bar; r++) {
            double s = x[r];
            for (int64_t q = 0; q < r; q++) s -= L[r * n + q] * x[q];
            x[r] = s / L[r * n + r];
          }
          for (int64_t r = bar; r < n; r++)
            for (int64_t q = 0; q < bar; q++) x[r] -= L[r * n + q] * x[q];
        }
        vector<T> v = v0;
        pc.solver->solveLUpTo(data.data(), pc.nocross, v.data(), n, nRHS);
        CHECK(relDiff(ref, vector<double>(v.begin(), v.end())) < eps);
        // solveLtUpTo: top -= L21^T bottom ; top = L11^-T top
        ref = b;
        for (int c = 0; c < nRHS; c++) {
          double* x = ref.data() + c * n;
          for (int64_t q = 0; q < bar; q++)
            for (int64_t r = bar; r < n; r++) x[q] -= L[r * n + q] * x[r];
          for (int64_t r = bar - 1; r >= 0; r--) {
            double s = x[r];
            for (int64_t q = r + 1; q < bar; q++) s -= L[q * n + r] * x[q];
            x[r] = s / L[r * n + r];
          }
        }
        v = v0;
        pc.solver->solveLtUpTo(data.data(), pc.nocross, v.data(), n, nRHS);
        CHECK(relDiff(ref, vector<double>(v.begin(), v.end())) < eps);
        // solveLFrom / solveLtFrom: only the trailing block
        ref = b;
        for (int c = 0; c < nRHS; c++) {
          double* x = ref.data() + c * n;
          for (int64_t r = bar; r < n; r++) {
            double s = x[r];
            for (int64_t q = bar; q < r; q++) s -= L[r * n + q] * x[q];
            x[r] = s / L[r * n + r];
          }
        }
        v = v0;
        pc.solver->solveLFrom(data.data(), pc.nocross, v.data(), n, nRHS);
        CHECK(relDiff(ref, vector<double>(v.begin(), v.end())) < eps);
        ref = b;
        for (int c = 0; c < nRHS; c++) {
          double* x = ref.data() + c * n;
          for (int64_t r = n - 1; r >= bar; r--) {
            double s = x[r];
            for (int64_t q = r + 1; q < n; q++) s -= L[q * n + r] * x[q];
            x[r] = s / L[r * n + r];
          }
        }
        v = v0;
        pc.solver->solveLtFrom(data.data(), pc.nocross, v.data(), n, nRHS);
        CHECK(relDiff(ref, vector<double>(v.begin(), v.end())) < eps);
      }
    }
  }
}

int main(int argc, char** argv) {
  registerBackend(BackendRef, [](int) { return oracleRefOps(); });
  registerBackend(BackendFast, [](int n) { return oracleFastOps(n); });
  const int reps = argc > 1 ? std::atoi(argv[1]) : 6;  // the reference runs 20 random problems per CreateSolver case
  struct Case {
    const char* name;
    std::function<void()> fn;
  };
  vector<Case> cases = {
      {"SparseStructure.Transpose", testTranspose},
      {"SparseStructure.SymPermutation", testSymPermutation},
      {"SparseStructure.IndependentEliminationFill+FullEliminationFill", testEliminationFills},
      {"SparseStructure.FillReducingPermutation", testFillReducingPermutation},
      {"EliminationTree.Build", testEliminationTreeBuild},
      {"CoalescedBlockMatrix.Densify+Densify2+Damp", testDensifyAndDamp},
      {"Accessor.CoalescedAccessor+PermutedCoalescedAccessor", testAccessors},
      {"CreateSolver.Plain_double", [&] { testCreateSolver<double>(false, false, reps, 1e-9); }},
      {"CreateSolver.Plain_float", [&] { testCreateSolver<float>(false, false, reps, 2e-5); }},
      {"CreateSolver.Elim_double", [&] { testCreateSolver<double>(true, false, reps, 1e-9); }},
      {"CreateSolver.Elim_float", [&] { testCreateSolver<float>(true, false, reps, 2e-5); }},
      {"CreateSolver.Last_double", [&] { testCreateSolver<double>(false, true, reps, 1e-9); }},
      {"CreateSolver.ElimLast_double", [&] { testCreateSolver<double>(true, true, reps, 1e-9); }},
      {"Partial.{PartialFactor,SplitFactor,PseudoFactor,PartialAddMv,PartialSolveL/Lt(+From)}_Ref_double",
       [&] { testPartial<double>([] { return oracleRefOps(); }, reps, 1e-9); }},
      {"Partial.*_Ref_float", [&] { testPartial<float>([] { return oracleRefOps(); }, reps, 5e-5); }},
  };
  // BackendFast needs a BLAS: tests/test_host_cpp.py passes the OpenBLAS the python side found
  if (const char* blas = getenv("ORACLE_BLAS_PATH")) {
    std::string err;
    const char* prefix = getenv("ORACLE_BLAS_PREFIX");
    if (oracle_blas::load(blas, prefix ? prefix : "", "", &err)) {
      cases.push_back({"Partial.*_Blas_double", [&] { testPartial<double>([] { return oracleFastOps(4); }, reps, 1e-9); }});
      cases.push_back({"Partial.*_Blas_float", [&] { testPartial<float>([] { return oracleFastOps(4); }, reps, 5e-5); }});
      cases.push_back({"CreateSolver.Plain_Blas_double", [&] {
                         registerBackend(BackendRef, [](int) { return oracleFastOps(4); });
                         testCreateSolver<double>(false, false, reps, 1e-9);
                         registerBackend(BackendRef, [](int) { return oracleRefOps(); });
                       }});
    } else {
      std::printf("skip BackendFast cases: %s\n", err.c_str());
    }
  }
  for (auto& c : cases) {
    g_test = c.name;
    const int before = g_failures;
    try {
      c.fn();
    } catch (const std::exception& e) {
      g_failures++;
      std::printf("FAIL [%s] exception: %s\n", c.name, e.what());
    }
    std::printf("%s %s\n", g_failures == before ? "ok  " : "FAIL", c.name);
  }
  std::printf("%d checks, %d failures\n", g_checks, g_failures);
  return g_failures ? 1 : 0;
}
