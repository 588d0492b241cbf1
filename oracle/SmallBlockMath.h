// ORACLE / TEST INFRASTRUCTURE ONLY.
// Restatement of the reference's small-block helpers (baspacho/baspacho/MathUtils.h:17-97): the pair enumeration of
// its CUDA sparse-elimination kernel and the in-place Cholesky / triangular solves one thread runs on a point block.
// Same signatures, so that oracle/host_scenarios.h compiles against either this file or the reference's header;
// tests/test_ref_objects.py holds the two bit-identical (oracle/_ref = the reference's own object code).
#pragma once
#include <cmath>
#include <cstdint>
#include <utility>

namespace BaSpaCho {

// p in [0, n(n+1)/2) -> (x, y) with 0 <= x <= y < n; for a fixed row of the folded rectangle x runs sequentially
// (reference MathUtils.h:17-33)
inline std::pair<int64_t, int64_t> toOrderedPair(int64_t n, int64_t p) {
  const int64_t odd = n & 1, width = n + 1 - odd;
  int64_t x = p % width, y = n - 1 - p / width;
  if (x > y) {  // the triangle above the diagonal of the rectangle folds back next to the origin
    x -= y + 1;
    y = n - 1 - odd - y;
  }
  return {x, y};
}

// in-place right-looking Cholesky, column i scaled then the trailing rows updated (reference MathUtils.h:36-63;
// the data is the lower triangle of a row-major block = upper triangle of the column-major view)
template <typename T>
inline void cholesky(T* A, int lda, int n) {
  for (int i = 0; i < n; i++) {
    T* diag = A + (int64_t)i * lda + i;
    const T d = sqrt(*diag);
    *diag = d;
    for (int j = i + 1; j < n; j++) {
      T* rowJ = A + (int64_t)j * lda;
      const T c = rowJ[i] / d;
      rowJ[i] = c;
      for (int k = i + 1; k <= j; k++) rowJ[k] -= c * A[(int64_t)k * lda + i];
    }
  }
}

// v <- tril(A)^-1 v (reference MathUtils.h:66-79 `solveUpperT`)
template <typename T>
inline void solveUpperT(const T* A, int lda, int n, T* v) {
  for (int i = 0; i < n; i++) {
    const T* row = A + (int64_t)i * lda;
    T x = v[i];
    for (int j = 0; j < i; j++) x -= row[j] * v[j];
    v[i] = x / row[i];
  }
}

// v <- tril(A)^-T v (reference MathUtils.h:82-97 `solveUpper`)
template <typename T>
inline void solveUpper(const T* A, int lda, int n, T* v) {
  for (int i = n - 1; i >= 0; i--) {
    T x = v[i];
    for (int j = i + 1; j < n; j++) x -= A[(int64_t)j * lda + i] * v[j];
    v[i] = x / A[(int64_t)i * lda + i];
  }
}

}  // namespace BaSpaCho
