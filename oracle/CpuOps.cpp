// ============================================================================================
// ORACLE / TEST INFRASTRUCTURE ONLY.
// CPU restatement of the reference's two CPU backends behind the same Ops interface:
//   * refOps()  - naive single-thread loops  (reference baspacho/baspacho/MatOpsRef.cpp:33-357 +
//                 MatOpsCpuBase.h:28-434; the reference uses Eigen LLT / triangular solves / products)
//   * fastOps() - BLAS + thread pool         (reference baspacho/baspacho/MatOpsFast.cpp:24-1134)
// Nothing here is linked into, or called from, the product library (baspacho_b200/csrc). Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
// Parity pin: integer skeleton vs the reference's golden vectors (tests/test_skeleton.py), numerics vs
// dense LAPACK Cholesky / triangular solves exactly as the reference's own tests do (FactorTest.cpp,
// SolveTest.cpp); the reference holds no floating-point golden vectors.
// ============================================================================================
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include "../baspacho_b200/csrc/host/DebugMacros.h"
#include "../baspacho_b200/csrc/host/MatOps.h"
#include "../baspacho_b200/csrc/host/Solver.h"
#include "BlasLoader.h"
#include "ThreadPool.h"

namespace BaSpaCho {
namespace {

using std::vector;

// ---------------------------------------------------------------- tiny dense kernels (row-major)
// in-place lower Cholesky of the n x n block (cf. reference MathUtils.h:36-63 / Eigen::LLT use in
// MatOpsCpuBase.h:124-131)
template <typename T>
void denseCholesky(T* A, int64_t n, int64_t lda) {
  for (int64_t j = 0; j < n; j++) {
    T* rj = A + j * lda;
    T d = rj[j];
    for (int64_t q = 0; q < j; q++) d -= rj[q] * rj[q];
    d = std::sqrt(d);
    rj[j] = d;
    for (int64_t i = j + 1; i < n; i++) {
      T* ri = A + i * lda;
      T v = ri[j];
      for (int64_t q = 0; q < j; q++) v -= ri[q] * rj[q];
      ri[j] = v / d;
    }
  }
}

// X * tril(L)^T = B in place on k rows (cf. MatOpsCpuBase.h:133-141)
template <typename T>
void solveRowsLowerT(const T* L, int64_t n, int64_t ldl, T* B, int64_t k, int64_t ldb) {
  for (int64_t r = 0; r < k; r++) {
    T* x = B + r * ldb;
    for (int64_t j = 0; j < n; j++) {
      const T* lj = L + j * ldl;
      T v = x[j];
      for (int64_t q = 0; q < j; q++) v -= x[q] * lj[q];
      x[j] = v / lj[j];
    }
  }
}

// vectors: column-major n x nRHS, leading dimension ldc.  tril(L) * X = C in place
template <typename T>
void solveColsLower(const T* L, int64_t n, T* C, int64_t ldc, int nRHS) {
  for (int c = 0; c < nRHS; c++) {
    T* x = C + c * ldc;
    for (int64_t i = 0; i < n; i++) {
      const T* li = L + i * n;
      T v = x[i];
      for (int64_t q = 0; q < i; q++) v -= li[q] * x[q];
      x[i] = v / li[i];
    }
  }
}

// tril(L)^T * X = C in place
template <typename T>
void solveColsLowerT(const T* L, int64_t n, T* C, int64_t ldc, int nRHS) {
  for (int c = 0; c < nRHS; c++) {
    T* x = C + c * ldc;
    for (int64_t i = n - 1; i >= 0; i--) {
      T v = x[i];
      for (int64_t q = i + 1; q < n; q++) v -= L[q * n + i] * x[q];
      x[i] = v / L[i * n + i];
    }
  }
}

// ---------------------------------------------------------------- elimination plan (row view)
// For the rectangle (span rows >= lumpToSpan[lumpsEnd]) x (lumps in [lumpsBegin, lumpsEnd)):
// per row span, the chains found there (source lump + ordinal of the chain in its column).
// Restates CpuBaseSymElimCtx / prepareElimination, reference MatOpsCpuBase.h:28-117.
struct CpuSymElimCtx : SymElimCtx {
  int64_t spanRowBegin = 0;
  int64_t maxBufferSize = 0;
  vector<int64_t> rowPtr, colLump, chainColOrd;
};

struct CpuSymbolicCtx : SymbolicCtx {
  CpuSymbolicCtx(const CoalescedBlockMatrixSkel& s, int nThreads, bool blas)
      : skel(s), useBlas(blas), pool(nThreads) {}

  PermutedCoalescedAccessor deviceAccessor() override {
    throw std::runtime_error("no device accessor can be created from a cpu-only backend");
  }

  SymElimCtxPtr prepareElimination(int64_t lumpsBegin, int64_t lumpsEnd) override {
    auto* e = new CpuSymElimCtx;
    e->spanRowBegin = skel.lumpToSpan[lumpsEnd];
    int64_t nRows = skel.numSpans() - e->spanRowBegin;
    e->rowPtr.assign(nRows + 1, 0);
    auto forEachChain = [&](auto&& f) {
      for (int64_t l = lumpsBegin; l < lumpsEnd; l++)
        for (int64_t i = skel.chainColPtr[l]; i < skel.chainColPtr[l + 1]; i++) {
          int64_t s = skel.chainRowSpan[i];
          if (s >= e->spanRowBegin) f(s - e->spanRowBegin, l, i - skel.chainColPtr[l]);
        }
    };
    forEachChain([&](int64_t r, int64_t, int64_t) { e->rowPtr[r]++; });
    int64_t tot = cumSumVec(e->rowPtr);
    e->colLump.resize(tot);
    e->chainColOrd.resize(tot);
    vector<int64_t> cur(e->rowPtr.begin(), e->rowPtr.end() - 1);
    forEachChain([&](int64_t r, int64_t l, int64_t ord) {
      e->colLump[cur[r]] = l;
      e->chainColOrd[cur[r]++] = ord;
    });
    for (int64_t r = 0; r < nRows; r++)
      for (int64_t i = e->rowPtr[r]; i < e->rowPtr[r + 1]; i++) {
        int64_t first = skel.chainColPtr[e->colLump[i]] + e->chainColOrd[i];
        int64_t rowsChain = skel.chainRowsTillEnd[first] - skel.chainRowsTillEnd[first - 1];
        int64_t rowsOnward = skel.chainRowsTillEnd[skel.chainColPtr[e->colLump[i] + 1] - 1];
        e->maxBufferSize = std::max(e->maxBufferSize, rowsOnward * rowsChain);
      }
    return SymElimCtxPtr(e);
  }

  NumericCtxBase* createNumericCtxForType(std::type_index tIdx, int64_t tempBufSize, int batchSize) override;
  SolveCtxBase* createSolveCtxForType(std::type_index tIdx, int nRHS, int batchSize) override;

  const CoalescedBlockMatrixSkel& skel;
  bool useBlas;
  oracle::ThreadPool pool;
  // which row-chain variant doElimination runs: 0 = the reference's dispatch rule, 1 = eliminateRowChain,
  // 2 = eliminateVerySparseRowChain (ORACLE_ELIM_VARIANT, read when the context is created; tests force both)
  int elimVariant = getenv("ORACLE_ELIM_VARIANT") ? atoi(getenv("ORACLE_ELIM_VARIANT")) : 0;
};

// ---------------------------------------------------------------- BLAS shims (row-major views)
template <typename T> struct Blas;
template <> struct Blas<double> {
  static void potrf(int64_t n, double* A) {
    int nn = (int)n, info = 0;
    oracle_blas::api().dpotrf("U", &nn, A, &nn, &info);  // col-major upper == row-major lower
  }
  static void trsm(int64_t n, int64_t k, const double* A, double* B) {
    int nn = (int)n, kk = (int)k;
    double one = 1.0;
    oracle_blas::api().dtrsm("L", "U", "C", "N", &nn, &kk, &one, A, &nn, B, &nn);
  }
  static void syrk(int64_t m, int64_t k, const double* A, double* C) {
    int mm = (int)m, kk = (int)k;
    double one = 1.0, zero = 0.0;
    oracle_blas::api().dsyrk("U", "C", &mm, &kk, &one, A, &kk, &zero, C, &mm);
  }
  // C(col-major m x n, ld m) = A^T(m x k) * B(k x n): A,B col-major with ld k
  static void gemmTN(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                     int64_t ldb, double beta, double* C, int64_t ldc) {
    int mm = (int)m, nn = (int)n, kk = (int)k, la = (int)lda, lb = (int)ldb, lc = (int)ldc;
    oracle_blas::api().dgemm("C", "N", &mm, &nn, &kk, &alpha, A, &la, B, &lb, &beta, C, &lc);
  }
};
template <> struct Blas<float> {
  static void potrf(int64_t n, float* A) {
    int nn = (int)n, info = 0;
    oracle_blas::api().spotrf("U", &nn, A, &nn, &info);
  }
  static void trsm(int64_t n, int64_t k, const float* A, float* B) {
    int nn = (int)n, kk = (int)k;
    float one = 1.0f;
    oracle_blas::api().strsm("L", "U", "C", "N", &nn, &kk, &one, A, &nn, B, &nn);
  }
  static void syrk(int64_t m, int64_t k, const float* A, float* C) {
    int mm = (int)m, kk = (int)k;
    float one = 1.0f, zero = 0.0f;
    oracle_blas::api().ssyrk("U", "C", &mm, &kk, &one, A, &kk, &zero, C, &mm);
  }
  static void gemmTN(int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float beta, float* C, int64_t ldc) {
    int mm = (int)m, nn = (int)n, kk = (int)k, la = (int)lda, lb = (int)ldb, lc = (int)ldc;
    oracle_blas::api().sgemm("C", "N", &mm, &nn, &kk, &alpha, A, &la, B, &lb, &beta, C, &lc);
  }
};

// ---------------------------------------------------------------- numeric context
template <typename T>
struct CpuNumericCtx : NumericCtx<T> {
  CpuNumericCtx(const CpuSymbolicCtx& s, int64_t bufSize)
      : sym(s), skel(s.skel), tempBuffer(bufSize), spanToChainOffset(s.skel.numSpans()) {}

  // potrf + trsm of one whole lump column (reference MatOpsCpuBase.h:162-186 factorLump)
  void factorLumpColumn(T* data, int64_t lump) const {
    int64_t w = skel.lumpSize(lump);
    T* diag = data + skel.lumpDataOffset(lump);
    denseCholesky(diag, w, w);
    int64_t rowsBelow = skel.lumpTotalRows(lump) - w;
    if (rowsBelow > 0) solveRowsLowerT(diag, w, w, diag + w * w, rowsBelow, w);
  }

  // same, restricted to the columns of one span (reference MatOpsCpuBase.h:188-217 factorSpan)
  void factorSpanColumn(T* data, int64_t span) const {
    int64_t lump = skel.spanToLump[span], w = skel.lumpSize(lump);
    int64_t sz = skel.spanStart[span + 1] - skel.spanStart[span];
    int64_t ordInLump = span - skel.lumpToSpan[lump];
    int64_t first = skel.chainColPtr[lump];
    T* diag = data + skel.chainData[first + ordInLump] + skel.spanOffsetInLump[span];
    denseCholesky(diag, sz, w);
    int64_t rowsBelow = skel.lumpTotalRows(lump) - skel.chainRowsTillEnd[first + ordInLump];
    if (rowsBelow > 0) {
      T* below = data + skel.chainData[first + ordInLump + 1] + skel.spanOffsetInLump[span];
      solveRowsLowerT(diag, sz, w, below, rowsBelow, w);
    }
  }

  void pseudoFactorSpans(T* data, int64_t spanBegin, int64_t spanEnd) override {
    auto timer = sym.pseudoFactorStat.instance();
    const_cast<oracle::ThreadPool&>(sym.pool).parallelFor(spanBegin, spanEnd, 1, [&](int64_t b, int64_t e, int) {
      for (int64_t s = b; s < e; s++) factorSpanColumn(data, s);
    });
  }

  // Target-row-major ("gather") elimination of one row span: for every chain found in this row, subtract
  // (rows from the chain downward) * (chain)^T from the column of this span inside its own lump.
  // Restates eliminateRowChain, reference MatOpsCpuBase.h:267-319 (deterministic, no atomics).
  void eliminateRow(const CpuSymElimCtx& elim, T* data, int64_t sRel, vector<int64_t>& chainOffsetOfSpan) const {
    if (elim.rowPtr[sRel] == elim.rowPtr[sRel + 1]) return;
    const int64_t s = sRel + elim.spanRowBegin;
    const int64_t target = skel.spanToLump[s], tw = skel.lumpSize(target);
    const int64_t colInTarget = skel.spanStart[s] - skel.lumpStart[target];
    for (int64_t i = skel.chainColPtr[target]; i < skel.chainColPtr[target + 1]; i++)
      chainOffsetOfSpan[skel.chainRowSpan[i]] = skel.chainData[i];

    for (int64_t i = elim.rowPtr[sRel]; i < elim.rowPtr[sRel + 1]; i++) {
      const int64_t src = elim.colLump[i], k = skel.lumpSize(src);
      const int64_t first = skel.chainColPtr[src] + elim.chainColOrd[i], end = skel.chainColPtr[src + 1];
      BASPACHO_CHECK_EQ(skel.chainRowSpan[first], s);
      const int64_t m = skel.chainRowsTillEnd[first] - skel.chainRowsTillEnd[first - 1];
      const T* A = data + skel.chainData[first];  // m x k
      for (int64_t c = first; c < end; c++) {
        const int64_t rows = skel.chainRowsTillEnd[c] - skel.chainRowsTillEnd[c - 1];
        const T* B = data + skel.chainData[c];    // rows x k
        T* dst = data + chainOffsetOfSpan[skel.chainRowSpan[c]] + colInTarget;
        for (int64_t r = 0; r < rows; r++) {
          int64_t jEnd = (c == first) ? r + 1 : m;  // diagonal block: lower triangle only
          for (int64_t j = 0; j < jEnd; j++) {
            T acc = dst[r * tw + j];
            for (int64_t q = 0; q < k; q++) acc -= B[r * k + q] * A[j * k + q];
            dst[r * tw + j] = acc;
          }
        }
      }
    }
  }

  // The variant the reference picks when the rows of the elimination rectangle hold few chains each
  // (eliminateVerySparseRowChain, reference MatOpsCpuBase.h:321-373): per chain found in the row the WHOLE product
  // (rows from the chain downward) x (chain)^T goes to a reusable buffer, then every block of it is subtracted from
  // the target column; the target chain is located by bisection over the target column's chainRowSpan (no
  // span -> chain table to fill), and the diagonal block is subtracted as a full square (its upper triangle is a
  // don't-care region of the format).
  void eliminateVerySparseRow(const CpuSymElimCtx& elim, T* data, int64_t sRel, vector<T>& prod) const {
    if (elim.rowPtr[sRel] == elim.rowPtr[sRel + 1]) return;
    const int64_t s = sRel + elim.spanRowBegin;
    const int64_t target = skel.spanToLump[s], tw = skel.lumpSize(target);
    const int64_t colInTarget = skel.spanStart[s] - skel.lumpStart[target];
    const int64_t bisectStart = skel.chainColPtr[target], bisectEnd = skel.chainColPtr[target + 1];
    for (int64_t i = elim.rowPtr[sRel]; i < elim.rowPtr[sRel + 1]; i++) {
      const int64_t src = elim.colLump[i], k = skel.lumpSize(src);
      BASPACHO_CHECK_GE(elim.chainColOrd[i], 1);  // there must be a diagonal block
      const int64_t first = skel.chainColPtr[src] + elim.chainColOrd[i], end = skel.chainColPtr[src + 1];
      BASPACHO_CHECK_EQ(skel.chainRowSpan[first], s);
      const int64_t rowsAbove = skel.chainRowsTillEnd[first - 1];
      const int64_t m = skel.chainRowsTillEnd[first] - rowsAbove;
      const int64_t rowsOnward = skel.chainRowsTillEnd[end - 1] - rowsAbove;
      const T* A = data + skel.chainData[first];  // (rowsOnward x k) from the chain downward, first m rows = chain
      prod.resize((size_t)(rowsOnward * m));
      for (int64_t r = 0; r < rowsOnward; r++)
        for (int64_t j = 0; j < m; j++) {
          T acc = T(0);
          for (int64_t q = 0; q < k; q++) acc += A[r * k + q] * A[j * k + q];
          prod[r * m + j] = acc;
        }
      for (int64_t c = first; c < end; c++) {
        const int64_t relRow = skel.chainRowsTillEnd[c - 1] - rowsAbove;
        const int64_t rows = skel.chainRowsTillEnd[c] - rowsAbove - relRow;
        const int64_t pos = bisect(skel.chainRowSpan.data() + bisectStart, bisectEnd - bisectStart, skel.chainRowSpan[c]);
        T* dst = data + skel.chainData[bisectStart + pos] + colInTarget;
        for (int64_t r = 0; r < rows; r++)
          for (int64_t j = 0; j < m; j++) dst[r * tw + j] -= prod[(relRow + r) * m + j];
      }
    }
  }

  void doElimination(const SymElimCtx& elimData, T* data, int64_t lumpsBegin, int64_t lumpsEnd) override {
    const auto* elim = dynamic_cast<const CpuSymElimCtx*>(&elimData);
    BASPACHO_CHECK_NOTNULL(elim);
    auto timer = elim->elimStat.instance();
    auto& pool = const_cast<oracle::ThreadPool&>(sym.pool);
    pool.parallelFor(lumpsBegin, lumpsEnd, 5, [&](int64_t b, int64_t e, int) {
      for (int64_t l = b; l < e; l++) factorLumpColumn(data, l);
    });
    int64_t nRows = (int64_t)elim->rowPtr.size() - 1;
    // dispatch of the reference's BLAS backend (MatOpsFast.cpp:105-147): rows holding more than 3 chains on average
    // take the span -> chain table variant, sparser rectangles the bisect + product-buffer one; its naive backend
    // always runs the former (MatOpsRef.cpp:64-82)
    const int variant = sym.elimVariant;  // 0 = the reference's rule, 1 / 2 = force (tests)
    const bool dense = variant == 1 || (variant == 0 && (!sym.useBlas || elim->colLump.size() > 3 * (elim->rowPtr.size() - 1)));
    if (dense) {
      vector<vector<int64_t>> scratch(pool.numThreads());
      pool.parallelFor(0, nRows, 5, [&](int64_t b, int64_t e, int slot) {
        auto& map = scratch[slot];
        if (map.empty()) map.resize(skel.numSpans());
        for (int64_t r = b; r < e; r++) eliminateRow(*elim, data, r, map);
      });
    } else {
      vector<vector<T>> scratch(pool.numThreads());
      pool.parallelFor(0, nRows, 5, [&](int64_t b, int64_t e, int slot) {
        for (int64_t r = b; r < e; r++) eliminateVerySparseRow(*elim, data, r, scratch[slot]);
      });
    }
  }

  void potrf(int64_t n, T* data, int64_t offA) override {
    auto timer = sym.potrfStat.instance(sizeof(T), n);
    sym.potrfBiggestN = std::max(sym.potrfBiggestN, n);
    if (sym.useBlas) Blas<T>::potrf(n, data + offA); else denseCholesky(data + offA, n, n);
  }

  void trsm(int64_t n, int64_t k, T* data, int64_t offA, int64_t offB) override {
    auto timer = sym.trsmStat.instance(sizeof(T), n, k);
    if (!sym.useBlas) {
      solveRowsLowerT(data + offA, n, n, data + offB, k, n);
      return;
    }
    // The reference replaces OpenBLAS trsm by an Eigen solve chunked over 16-row slabs on its pool
    // (MatOpsFast.cpp:257-280); here: ?trsm of the BLAS per slab, slabs spread over the pool.
    auto& pool = const_cast<oracle::ThreadPool&>(sym.pool);
    const T* A = data + offA;
    T* B = data + offB;
    int64_t slab = std::max<int64_t>(16, (k + pool.numThreads() * 4 - 1) / (pool.numThreads() * 4));
    pool.parallelFor(0, k, slab, [&](int64_t b, int64_t e, int) { Blas<T>::trsm(n, e - b, A, B + b * n); });
  }

  void saveSyrkGemm(int64_t m, int64_t n, int64_t k, const T* data, int64_t offset) override {
    auto timer = sym.sygeStat.instance(sizeof(T), m, n, k);
    BASPACHO_CHECK_LE(m * n, (int64_t)tempBuffer.size());
    const T* AB = data + offset;
    T* C = tempBuffer.data();
    if (!sym.useBlas) {
      for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
          T acc = 0;
          for (int64_t q = 0; q < k; q++) acc += AB[i * k + q] * AB[j * k + q];
          C[i * m + j] = acc;
        }
      sym.gemmCalls++;
      return;
    }
    // syrk on the top m x m + gemm on the rest when worthwhile (reference MatOpsFast.cpp:307-333)
    bool doSyrk = (m == n) || (m + n + k > 150);
    bool doGemm = !(doSyrk && m == n);
    if (doSyrk) {
      Blas<T>::syrk(m, k, AB, C);
      sym.syrkCalls++;
    }
    if (doGemm) {
      int64_t skip = doSyrk ? m : 0;
      Blas<T>::gemmTN(m, n - skip, k, T(1), AB, k, AB + skip * k, k, T(0), C + skip * m, m);
      sym.gemmCalls++;
    }
  }

  void prepareAssemble(int64_t targetLump) override {
    for (int64_t i = skel.chainColPtr[targetLump]; i < skel.chainColPtr[targetLump + 1]; i++)
      spanToChainOffset[skel.chainRowSpan[i]] = skel.chainData[i];
  }

  // target[rowSpan r][colSpan c] -= temp block, for block rows r of the panel and block cols c <= r, c < numBlockCols
  // (reference MatOpsRef.cpp:144-175, MatOpsFast.cpp:168-226)
  void assemble(T* data, int64_t rectRowBegin, int64_t dstStride, int64_t srcColDataOffset, int64_t srcRectWidth,
                int64_t numBlockRows, int64_t numBlockCols) override {
    auto timer = sym.asmblStat.instance(sizeof(T), numBlockRows, numBlockCols);
    const int64_t* rowsTillEnd = skel.chainRowsTillEnd.data() + srcColDataOffset;
    const int64_t* toSpan = skel.chainRowSpan.data() + srcColDataOffset;
    const T* temp = tempBuffer.data();
    auto rowBlock = [&](int64_t r) {
      int64_t rBegin = rowsTillEnd[r - 1] - rectRowBegin, rSize = rowsTillEnd[r] - rowsTillEnd[r - 1];
      int64_t rowOffset = spanToChainOffset[toSpan[r]];
      for (int64_t c = 0, cEnd = std::min(numBlockCols, r + 1); c < cEnd; c++) {
        int64_t cBegin = rowsTillEnd[c - 1] - rectRowBegin, cSize = rowsTillEnd[c] - rowsTillEnd[c - 1];
        T* dst = data + rowOffset + skel.spanOffsetInLump[toSpan[c]];
        const T* src = temp + rBegin * srcRectWidth + cBegin;
        for (int64_t i = 0; i < rSize; i++)
          for (int64_t j = 0; j < cSize; j++) dst[i * dstStride + j] -= src[i * srcRectWidth + j];
      }
    };
    if (sym.useBlas && sym.pool.numThreads() > 1 && numBlockRows > 8) {
      const_cast<oracle::ThreadPool&>(sym.pool).parallelFor(0, numBlockRows, 3, [&](int64_t b, int64_t e, int) {
        for (int64_t r = b; r < e; r++) rowBlock(r);
      });
    } else {
      for (int64_t r = 0; r < numBlockRows; r++) rowBlock(r);
    }
  }

  const CpuSymbolicCtx& sym;
  const CoalescedBlockMatrixSkel& skel;
  vector<T> tempBuffer;
  vector<int64_t> spanToChainOffset;
};

// ---------------------------------------------------------------- solve context
// Restates CpuBaseSolveCtx (MatOpsCpuBase.h:376-434) + SimpleSolveCtx (MatOpsRef.cpp:189-327).
template <typename T>
struct CpuSolveCtx : SolveCtx<T> {
  CpuSolveCtx(const CpuSymbolicCtx& s, int nRHS_) : sym(s), skel(s.skel), nRHS(nRHS_), tmp(s.skel.order() * nRHS_) {}

  void sparseElimSolveL(const SymElimCtx& elimData, const T* data, int64_t lumpsBegin, int64_t lumpsEnd, T* C,
                        int64_t ldc) override {
    auto timer = sym.solveSparseLStat.instance();
    const auto* elim = dynamic_cast<const CpuSymElimCtx*>(&elimData);
    BASPACHO_CHECK_NOTNULL(elim);
    auto& pool = const_cast<oracle::ThreadPool&>(sym.pool);
    pool.parallelFor(lumpsBegin, lumpsEnd, 16, [&](int64_t b, int64_t e, int) {
      for (int64_t l = b; l < e; l++)
        solveColsLower(data + skel.lumpDataOffset(l), skel.lumpSize(l), C + skel.lumpStart[l], ldc, nRHS);
    });
    // row-major pass: each row span gathers from the chains present in its row
    int64_t nRows = (int64_t)elim->rowPtr.size() - 1;
    pool.parallelFor(0, nRows, 16, [&](int64_t b, int64_t e, int) {
      for (int64_t sRel = b; sRel < e; sRel++) {
        int64_t span = sRel + elim->spanRowBegin;
        int64_t r0 = skel.spanStart[span], rows = skel.spanStart[span + 1] - r0;
        for (int64_t i = elim->rowPtr[sRel]; i < elim->rowPtr[sRel + 1]; i++) {
          int64_t l = elim->colLump[i], w = skel.lumpSize(l), c0 = skel.lumpStart[l];
          const T* blk = data + skel.chainData[skel.chainColPtr[l] + elim->chainColOrd[i]];
          for (int c = 0; c < nRHS; c++)
            for (int64_t r = 0; r < rows; r++) {
              T acc = C[c * ldc + r0 + r];
              for (int64_t q = 0; q < w; q++) acc -= blk[r * w + q] * C[c * ldc + c0 + q];
              C[c * ldc + r0 + r] = acc;
            }
        }
      }
    });
  }

  void sparseElimSolveLt(const SymElimCtx&, const T* data, int64_t lumpsBegin, int64_t lumpsEnd, T* C,
                         int64_t ldc) override {
    auto timer = sym.solveSparseLtStat.instance();
    const_cast<oracle::ThreadPool&>(sym.pool).parallelFor(lumpsBegin, lumpsEnd, 16, [&](int64_t b, int64_t e, int) {
      for (int64_t l = b; l < e; l++) {
        int64_t w = skel.lumpSize(l), c0 = skel.lumpStart[l];
        for (int64_t i = skel.chainColPtr[l] + 1; i < skel.chainColPtr[l + 1]; i++) {
          int64_t span = skel.chainRowSpan[i];
          int64_t r0 = skel.spanStart[span], rows = skel.spanStart[span + 1] - r0;
          const T* blk = data + skel.chainData[i];
          for (int c = 0; c < nRHS; c++)
            for (int64_t q = 0; q < w; q++) {
              T acc = C[c * ldc + c0 + q];
              for (int64_t r = 0; r < rows; r++) acc -= blk[r * w + q] * C[c * ldc + r0 + r];
              C[c * ldc + c0 + q] = acc;
            }
        }
        solveColsLowerT(data + skel.lumpDataOffset(l), w, C + c0, ldc, nRHS);
      }
    });
  }

  void symm(const T* data, int64_t offM, int64_t n, const T* C, int64_t offC, int64_t ldc, T* D, int64_t ldd,
            T alpha) override {
    auto timer = sym.symmStat.instance();
    const T* M = data + offM;
    for (int c = 0; c < nRHS; c++)
      for (int64_t i = 0; i < n; i++) {
        T acc = 0;
        for (int64_t j = 0; j < n; j++) acc += (j <= i ? M[i * n + j] : M[j * n + i]) * C[offC + c * ldc + j];
        D[offC + c * ldd + i] += alpha * acc;
      }
  }

  void solveL(const T* data, int64_t offM, int64_t n, T* C, int64_t offC, int64_t ldc) override {
    auto timer = sym.solveLStat.instance();
    solveColsLower(data + offM, n, C + offC, ldc, nRHS);
  }

  void solveLt(const T* data, int64_t offM, int64_t n, T* C, int64_t offC, int64_t ldc) override {
    auto timer = sym.solveLtStat.instance();
    solveColsLowerT(data + offM, n, C + offC, ldc, nRHS);
  }

  // tmp(nRows x nRHS, row-major) = alpha * M(nRows x nCols) * A(nCols x nRHS, col-major lda)
  void gemv(const T* data, int64_t offM, int64_t nRows, int64_t nCols, const T* A, int64_t offA, int64_t lda,
            T alpha) override {
    auto timer = sym.solveGemvStat.instance();
    const T* M = data + offM;
    if (sym.useBlas && nRows * nCols > 4096) {
      // tmp^T (col-major nRHS x nRows) = alpha * A^T (nRHS x nCols) * M^T (col-major nCols x nRows)
      Blas<T>::gemmTN(nRHS, nRows, nCols, alpha, A + offA, lda, M, nCols, T(0), tmp.data(), nRHS);
      return;
    }
    for (int64_t r = 0; r < nRows; r++)
      for (int c = 0; c < nRHS; c++) {
        T acc = 0;
        for (int64_t q = 0; q < nCols; q++) acc += M[r * nCols + q] * A[offA + c * lda + q];
        tmp[r * nRHS + c] = alpha * acc;
      }
  }

  // A(nCols x nRHS) += alpha * M^T * tmp
  void gemvT(const T* data, int64_t offM, int64_t nRows, int64_t nCols, T* A, int64_t offA, int64_t lda,
             T alpha) override {
    auto timer = sym.solveGemvTStat.instance();
    const T* M = data + offM;
    for (int c = 0; c < nRHS; c++)
      for (int64_t q = 0; q < nCols; q++) {
        T acc = 0;
        for (int64_t r = 0; r < nRows; r++) acc += M[r * nCols + q] * tmp[r * nRHS + c];
        A[offA + c * lda + q] += alpha * acc;
      }
  }

  void assembleVec(int64_t chainColPtr, int64_t numColItems, T* C, int64_t ldc) override {
    auto timer = sym.solveAssVStat.instance();
    const int64_t* rowsTillEnd = skel.chainRowsTillEnd.data() + chainColPtr;
    int64_t startRow = rowsTillEnd[-1];
    for (int64_t i = 0; i < numColItems; i++) {
      int64_t rowOff = rowsTillEnd[i - 1] - startRow, span = skel.chainRowSpan[chainColPtr + i];
      int64_t r0 = skel.spanStart[span], rows = skel.spanStart[span + 1] - r0;
      for (int64_t r = 0; r < rows; r++)
        for (int c = 0; c < nRHS; c++) C[c * ldc + r0 + r] += tmp[(rowOff + r) * nRHS + c];
    }
  }

  void assembleVecT(const T* C, int64_t ldc, int64_t chainColPtr, int64_t numColItems) override {
    auto timer = sym.solveAssVTStat.instance();
    const int64_t* rowsTillEnd = skel.chainRowsTillEnd.data() + chainColPtr;
    int64_t startRow = rowsTillEnd[-1];
    for (int64_t i = 0; i < numColItems; i++) {
      int64_t rowOff = rowsTillEnd[i - 1] - startRow, span = skel.chainRowSpan[chainColPtr + i];
      int64_t r0 = skel.spanStart[span], rows = skel.spanStart[span + 1] - r0;
      for (int64_t r = 0; r < rows; r++)
        for (int c = 0; c < nRHS; c++) tmp[(rowOff + r) * nRHS + c] = C[c * ldc + r0 + r];
    }
  }

  // ---- "fragmented" whole-range ops of the reference's BLAS backend (MatOpsFast.cpp:613-1018; its naive backend has
  // none, MatOps.h:168-183): used by the Solver when every lump is a single span and nRHS == 1. Restated from the
  // sequential branches (the threaded ones compute the same sums chunk-wise).
  bool hasFragmentedOps() override { return sym.useBlas; }

  // y[spans >= spanBegin] += alpha * sym(A) x over the block columns [spanBegin, spanEnd)  (MatOpsFast.cpp:615-770)
  void fragmentedMV(const T* data, const T* x, int64_t spanBegin, int64_t spanEnd, T* y, T alpha) override {
    for (int64_t s = spanBegin; s < spanEnd; s++) {
      const int64_t s0 = skel.spanStart[s], sn = skel.spanStart[s + 1] - s0, cp = skel.chainColPtr[s];
      const T* D = data + skel.chainData[cp];  // sn x sn, lower triangle meaningful
      for (int64_t i = 0; i < sn; i++) {
        T acc = T(0);
        for (int64_t j = 0; j <= i; j++) acc += D[i * sn + j] * x[s0 + j];
        for (int64_t j = i + 1; j < sn; j++) acc += D[j * sn + i] * x[s0 + j];
        y[s0 + i] += alpha * acc;
      }
      for (int64_t pp = cp + 1; pp < skel.chainColPtr[s + 1]; pp++) {
        const int64_t r = skel.chainRowSpan[pp], r0 = skel.spanStart[r], rn = skel.spanStart[r + 1] - r0;
        const T* B = data + skel.chainData[pp];  // rn x sn
        for (int64_t i = 0; i < rn; i++) {
          T acc = T(0);
          for (int64_t j = 0; j < sn; j++) acc += B[i * sn + j] * x[s0 + j];
          y[r0 + i] += alpha * acc;
        }
        for (int64_t j = 0; j < sn; j++) {
          T acc = T(0);
          for (int64_t i = 0; i < rn; i++) acc += B[i * sn + j] * x[r0 + i];
          y[s0 + j] += alpha * acc;
        }
      }
    }
  }

  // forward substitution over the block columns [spanBegin, spanEnd), updating every row below (MatOpsFast.cpp:772-921)
  void fragmentedSolveL(const T* data, int64_t spanBegin, int64_t spanEnd, T* y) override {
    for (int64_t s = spanBegin; s < spanEnd; s++) {
      const int64_t s0 = skel.spanStart[s], sn = skel.spanStart[s + 1] - s0, cp = skel.chainColPtr[s];
      solveColsLower(data + skel.chainData[cp], sn, y + s0, sn, 1);
      for (int64_t pp = cp + 1; pp < skel.chainColPtr[s + 1]; pp++) {
        const int64_t r = skel.chainRowSpan[pp], r0 = skel.spanStart[r], rn = skel.spanStart[r + 1] - r0;
        const T* B = data + skel.chainData[pp];
        for (int64_t i = 0; i < rn; i++) {
          T acc = y[r0 + i];
          for (int64_t j = 0; j < sn; j++) acc -= B[i * sn + j] * y[s0 + j];
          y[r0 + i] = acc;
        }
      }
    }
  }

  // backward substitution, block columns spanEnd-1 .. spanBegin (MatOpsFast.cpp:923-1018)
  void fragmentedSolveLt(const T* data, int64_t spanBegin, int64_t spanEnd, T* y) override {
    for (int64_t s = spanEnd - 1; s >= spanBegin; s--) {
      const int64_t s0 = skel.spanStart[s], sn = skel.spanStart[s + 1] - s0, cp = skel.chainColPtr[s];
      for (int64_t pp = cp + 1; pp < skel.chainColPtr[s + 1]; pp++) {
        const int64_t r = skel.chainRowSpan[pp], r0 = skel.spanStart[r], rn = skel.spanStart[r + 1] - r0;
        const T* B = data + skel.chainData[pp];
        for (int64_t j = 0; j < sn; j++) {
          T acc = y[s0 + j];
          for (int64_t i = 0; i < rn; i++) acc -= B[i * sn + j] * y[r0 + i];
          y[s0 + j] = acc;
        }
      }
      solveColsLowerT(data + skel.chainData[cp], sn, y + s0, sn, 1);
    }
  }

  const CpuSymbolicCtx& sym;
  const CoalescedBlockMatrixSkel& skel;
  int nRHS;
  vector<T> tmp;
};

NumericCtxBase* CpuSymbolicCtx::createNumericCtxForType(std::type_index tIdx, int64_t tempBufSize, int batchSize) {
  BASPACHO_CHECK_EQ(batchSize, 1);  // CPU backends are not batched (reference MatOpsFast.cpp:1110, MatOpsRef.cpp:335)
  if (tIdx == std::type_index(typeid(double))) return new CpuNumericCtx<double>(*this, tempBufSize);
  if (tIdx == std::type_index(typeid(float))) return new CpuNumericCtx<float>(*this, tempBufSize);
  return nullptr;
}

SolveCtxBase* CpuSymbolicCtx::createSolveCtxForType(std::type_index tIdx, int nRHS, int batchSize) {
  BASPACHO_CHECK_EQ(batchSize, 1);
  if (tIdx == std::type_index(typeid(double))) return new CpuSolveCtx<double>(*this, nRHS);
  if (tIdx == std::type_index(typeid(float))) return new CpuSolveCtx<float>(*this, nRHS);
  return nullptr;
}

struct CpuOps : Ops {
  CpuOps(int nThreads_, bool blas_) : nThreads(nThreads_), blas(blas_) {}
  SymbolicCtxPtr createSymbolicCtx(const CoalescedBlockMatrixSkel& skel, const vector<int64_t>&) override {
    if (blas && !oracle_blas::api().loaded)
      throw std::runtime_error("oracle fastOps: no BLAS loaded (call oracle_load_blas first)");
    if (blas && oracle_blas::api().set_num_threads) oracle_blas::api().set_num_threads(nThreads);
    return SymbolicCtxPtr(new CpuSymbolicCtx(skel, blas ? nThreads : 1, blas));
  }
  int nThreads;
  bool blas;
};

}  // namespace

OpsPtr oracleRefOps() { return OpsPtr(new CpuOps(1, false)); }
OpsPtr oracleFastOps(int numThreads) { return OpsPtr(new CpuOps(numThreads, true)); }

}  // namespace BaSpaCho
