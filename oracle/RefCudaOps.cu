// ORACLE / MEASUREMENT INFRASTRUCTURE ONLY - never linked into the product library.
//
// "Second GPU baseline" of SURVEY.md §8(c)/(d): an Eigen-free restatement of the REFERENCE's own CUDA backend
// (baspacho/baspacho/MatOpsCuda.cu) behind the same operator interface, so that the algorithm the reference runs on
// a GPU - cuSOLVER potrf + cuBLAS trsm/gemm per lump, one thread per eliminated lump, one thread per block pair with
// fp64 atomics, one thread per block for the assemble scatter, a synchronous H2D copy of the span-to-chain table per
// target lump - can be timed on the same B200 beside the hand-written kernels. What is kept from the reference:
//   call shapes / launch shapes (32-thread CTAs everywhere), int64 indices, the atomics, the per-lump host sync copy
//   (MatOpsCuda.cu:148-186 factor_lumps_kernel, :235-331 pair kernel, :370-406 assemble_kernel, :471-481
//   prepareAssemble, :508-590 potrf/trsm/gemm, :836-1181 solve context).
// What differs: Eigen::Map block products are plain loops; ops run on the solver's stream (the reference uses stream
// 0 only); double precision, single matrix only (what the baseline measurement needs).
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <algorithm>
#include <sstream>
#include <stdexcept>
#include <typeindex>
#include <vector>

#include "../baspacho_b200/csrc/host/DebugMacros.h"
#include "../baspacho_b200/csrc/host/MatOps.h"

namespace BaSpaCho {
namespace {

using std::vector;

#define RC_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      std::stringstream ss_;                                                                       \
      ss_ << "[ref_cuda " << __LINE__ << "] " << #call << ": " << cudaGetErrorString(e_);          \
      throw std::runtime_error(ss_.str());                                                         \
    }                                                                                              \
  } while (0)
#define RC_LIB(call)                                                                               \
  do {                                                                                             \
    int s_ = (int)(call);                                                                          \
    if (s_ != 0) {                                                                                 \
      std::stringstream ss_;                                                                       \
      ss_ << "[ref_cuda " << __LINE__ << "] " << #call << " failed with status " << s_;            \
      throw std::runtime_error(ss_.str());                                                         \
    }                                                                                              \
  } while (0)

template <typename T>
struct Dev {  // grow-only device array (role of the reference's DevMirror, CudaDefs.h:75-128)
  T* ptr = nullptr;
  size_t cap = 0;
  ~Dev() {
    if (ptr) cudaFree(ptr);
  }
  void atLeast(size_t n) {
    if (n <= cap) return;
    if (ptr) RC_CUDA(cudaFree(ptr));
    RC_CUDA(cudaMalloc((void**)&ptr, std::max<size_t>(n, 1) * sizeof(T)));
    cap = n;
  }
  void load(const vector<T>& v) {
    atLeast(v.size());
    if (!v.empty()) RC_CUDA(cudaMemcpy(ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
};

struct Idx {  // device skeleton arrays
  const int64_t *lumpStart, *lumpToSpan, *spanStart, *spanToLump, *spanOffsetInLump, *chainColPtr, *chainRowSpan, *chainData,
      *chainRowsTillEnd, *boardColPtr, *boardChainColOrd;
};

// ---- small dense helpers on row-major blocks (role of MathUtils.h:36-97)
__device__ inline void rcCholesky(double* A, int64_t ld, int64_t n) {
  for (int64_t j = 0; j < n; j++) {
    double d = A[j * ld + j];
    for (int64_t q = 0; q < j; q++) d -= A[j * ld + q] * A[j * ld + q];
    d = sqrt(d);
    A[j * ld + j] = d;
    for (int64_t i = j + 1; i < n; i++) {
      double s = A[i * ld + j];
      for (int64_t q = 0; q < j; q++) s -= A[i * ld + q] * A[j * ld + q];
      A[i * ld + j] = s / d;
    }
  }
}
// x L^T = b in place on one row (forward substitution with lower-triangular L)
__device__ inline void rcSolveRow(const double* L, int64_t ld, int64_t n, double* x, int64_t xs) {
  for (int64_t j = 0; j < n; j++) {
    double s = x[j * xs];
    for (int64_t q = 0; q < j; q++) s -= L[j * ld + q] * x[q * xs];
    x[j * xs] = s / L[j * ld + j];
  }
}
// L^T x = b in place (backward substitution)
__device__ inline void rcSolveRowT(const double* L, int64_t ld, int64_t n, double* x, int64_t xs) {
  for (int64_t j = n - 1; j >= 0; j--) {
    double s = x[j * xs];
    for (int64_t q = j + 1; q < n; q++) s -= L[q * ld + j] * x[q * xs];
    x[j * xs] = s / L[j * ld + j];
  }
}
__device__ inline int64_t rcBisect(const int64_t* a, int64_t n, int64_t needle) {  // last position with a[pos] <= needle
  int64_t lo = 0;
  while (n > 1) {
    int64_t h = n / 2;
    if (needle >= a[lo + h]) lo += h, n -= h;
    else n = h;
  }
  return lo;
}

// one thread per eliminated lump: Cholesky of the diagonal block, then every row below solved against it
__global__ void rc_factor_lumps_kernel(Idx ix, double* data, int64_t lumpsBegin, int64_t lumpsEnd) {
  const int64_t lump = lumpsBegin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lump >= lumpsEnd) return;
  const int64_t n = ix.lumpStart[lump + 1] - ix.lumpStart[lump];
  const int64_t c0 = ix.chainColPtr[lump];
  double* diag = data + ix.chainData[c0];
  rcCholesky(diag, n, n);
  const int64_t b0 = ix.boardColPtr[lump], b1 = ix.boardColPtr[lump + 1];
  const int64_t chFirst = ix.boardChainColOrd[b0 + 1], chEnd = ix.boardChainColOrd[b1 - 1];
  const int64_t rows = ix.chainRowsTillEnd[c0 + chEnd - 1] - ix.chainRowsTillEnd[c0 + chFirst - 1];
  double* row = data + ix.chainData[c0 + chFirst];
  for (int64_t r = 0; r < rows; r++, row += n) rcSolveRow(diag, n, n, row, 1);
}

// one thread per pair (i <= j) of below-diagonal blocks of an eliminated lump: target(j, i) -= B_j B_i^T with atomics
__global__ void rc_elim_pairs_kernel(Idx ix, double* data, int64_t lumpsBegin, int64_t lumpsEnd, const int64_t* pairPtr,
                                     int64_t numPairs) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= numPairs) return;
  const int64_t pos = rcBisect(pairPtr, lumpsEnd - lumpsBegin, t);
  const int64_t l = lumpsBegin + pos;
  const int64_t first = ix.chainColPtr[l] + 1, nb = ix.chainColPtr[l + 1] - first;
  // ordered pair (di <= dj) number `local` in row-major enumeration of the upper triangle of nb x nb
  int64_t local = t - pairPtr[pos], di = 0;
  while (local >= nb - di) local -= nb - di, di++;
  const int64_t dj = di + local;
  const int64_t k = ix.lumpStart[l + 1] - ix.lumpStart[l];
  const int64_t ci = first + di, cj = first + dj;
  const int64_t si = ix.chainRowSpan[ci], sj = ix.chainRowSpan[cj];
  const int64_t ni = ix.spanStart[si + 1] - ix.spanStart[si], nj = ix.spanStart[sj + 1] - ix.spanStart[sj];
  const double* Bi = data + ix.chainData[ci];
  const double* Bj = data + ix.chainData[cj];
  const int64_t tl = ix.spanToLump[si];
  const int64_t t0 = ix.chainColPtr[tl], t1 = ix.chainColPtr[tl + 1];
  const int64_t tw = ix.lumpStart[tl + 1] - ix.lumpStart[tl];
  const int64_t hit = rcBisect(ix.chainRowSpan + t0, t1 - t0, sj);
  double* dst = data + ix.chainData[t0 + hit] + ix.spanOffsetInLump[si];
  for (int64_t r = 0; r < nj; r++)
    for (int64_t c = 0; c < ni; c++) {
      double s = 0;
      for (int64_t q = 0; q < k; q++) s += Bj[r * k + q] * Bi[c * k + q];
      atomicAdd(dst + r * tw + c, -s);
    }
}

// one thread per (block row r, block column c <= r) of the product rectangle
__global__ void rc_assemble_kernel(int64_t numBlockRows, int64_t numBlockCols, int64_t rectRowBegin, int64_t srcWidth,
                                   int64_t dstStride, const int64_t* rowsTillEnd, const int64_t* toSpan,
                                   const int64_t* spanToChainOffset, const int64_t* spanOffsetInLump, const double* rect,
                                   double* data) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numBlockRows * numBlockCols) return;
  const int64_t r = i % numBlockRows, c = i / numBlockRows;
  if (c > r) return;
  const int64_t rBegin = rowsTillEnd[r - 1] - rectRowBegin, rSize = rowsTillEnd[r] - rectRowBegin - rBegin;
  const int64_t cBegin = rowsTillEnd[c - 1] - rectRowBegin, cSize = rowsTillEnd[c] - rectRowBegin - cBegin;
  double* dst = data + spanToChainOffset[toSpan[r]] + spanOffsetInLump[toSpan[c]];
  const double* src = rect + rBegin * srcWidth + cBegin;
  for (int64_t a = 0; a < rSize; a++)
    for (int64_t b = 0; b < cSize; b++) dst[a * dstStride + b] -= src[a * srcWidth + b];
}

__global__ void rc_factor_spans_kernel(Idx ix, double* data, int64_t spanBegin, int64_t spanEnd) {
  const int64_t span = spanBegin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (span >= spanEnd) return;
  const int64_t lump = ix.spanToLump[span], off = ix.spanOffsetInLump[span], inLump = span - ix.lumpToSpan[lump];
  const int64_t size = ix.spanStart[span + 1] - ix.spanStart[span];
  const int64_t w = ix.lumpStart[lump + 1] - ix.lumpStart[lump], c0 = ix.chainColPtr[lump];
  double* diag = data + ix.chainData[c0 + inLump] + off;
  rcCholesky(diag, w, size);
  const int64_t chEnd = ix.boardChainColOrd[ix.boardColPtr[lump + 1] - 1];
  const int64_t rows = ix.chainRowsTillEnd[c0 + chEnd - 1] - ix.chainRowsTillEnd[c0 + inLump];
  double* row = data + ix.chainData[c0 + inLump + 1] + off;
  for (int64_t r = 0; r < rows; r++, row += w) rcSolveRow(diag, w, size, row, 1);
}

// ---- solve kernels (thread per lump / per chain, atomics on shared rows)
__global__ void rc_elim_diag_solve_kernel(Idx ix, const double* data, double* v, int64_t ldc, int nRHS, int64_t lumpsBegin,
                                          int64_t lumpsEnd, bool transposed) {
  const int64_t lump = lumpsBegin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lump >= lumpsEnd) return;
  const int64_t s = ix.lumpStart[lump], n = ix.lumpStart[lump + 1] - s;
  const double* diag = data + ix.chainData[ix.chainColPtr[lump]];
  for (int c = 0; c < nRHS; c++) {
    if (transposed) rcSolveRowT(diag, n, n, v + s + ldc * c, 1);
    else rcSolveRow(diag, n, n, v + s + ldc * c, 1);
  }
}
__global__ void rc_elim_sub_mult_kernel(Idx ix, const double* data, double* v, int64_t ldc, int nRHS, int64_t lumpsBegin,
                                        int64_t lumpsEnd) {
  const int64_t lump = lumpsBegin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lump >= lumpsEnd) return;
  const int64_t s = ix.lumpStart[lump], n = ix.lumpStart[lump + 1] - s;
  for (int64_t ch = ix.chainColPtr[lump] + 1; ch < ix.chainColPtr[lump + 1]; ch++) {
    const int64_t span = ix.chainRowSpan[ch], r0 = ix.spanStart[span], rn = ix.spanStart[span + 1] - r0;
    const double* B = data + ix.chainData[ch];
    for (int c = 0; c < nRHS; c++)
      for (int64_t r = 0; r < rn; r++) {
        double acc = 0;
        for (int64_t q = 0; q < n; q++) acc += B[r * n + q] * v[s + q + ldc * c];
        atomicAdd(v + r0 + r + ldc * c, -acc);  // rows are shared between lumps
      }
  }
}
__global__ void rc_elim_sub_mult_t_kernel(Idx ix, const double* data, double* v, int64_t ldc, int nRHS, int64_t lumpsBegin,
                                          int64_t lumpsEnd) {
  const int64_t lump = lumpsBegin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lump >= lumpsEnd) return;
  const int64_t s = ix.lumpStart[lump], n = ix.lumpStart[lump + 1] - s;
  for (int64_t ch = ix.chainColPtr[lump] + 1; ch < ix.chainColPtr[lump + 1]; ch++) {
    const int64_t span = ix.chainRowSpan[ch], r0 = ix.spanStart[span], rn = ix.spanStart[span + 1] - r0;
    const double* B = data + ix.chainData[ch];
    for (int c = 0; c < nRHS; c++)
      for (int64_t q = 0; q < n; q++) {
        double acc = 0;
        for (int64_t r = 0; r < rn; r++) acc += B[r * n + q] * v[r0 + r + ldc * c];
        v[s + q + ldc * c] -= acc;  // the lump's own rows: no other thread writes them
      }
  }
}
// thread per chain: C[span rows] += tmp rows (tmp row-major rows x nRHS), and the gather mirror
__global__ void rc_assemble_vec_kernel(const int64_t* rowsTillEnd, const int64_t* toSpan, const int64_t* spanStart,
                                       double* tmp, int64_t numChains, double* C, int64_t ldc, int nRHS, bool gather) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numChains) return;
  const int64_t rowOff = rowsTillEnd[i - 1] - rowsTillEnd[-1];
  const int64_t span = toSpan[i], s0 = spanStart[span], sn = spanStart[span + 1] - s0;
  for (int64_t r = 0; r < sn; r++)
    for (int c = 0; c < nRHS; c++) {
      if (gather) tmp[(rowOff + r) * nRHS + c] = C[s0 + r + ldc * c];
      else C[s0 + r + ldc * c] += tmp[(rowOff + r) * nRHS + c];
    }
}

struct RefCudaSymElimCtx : SymElimCtx {
  int64_t numPairs = 0;
  Dev<int64_t> pairPtr;
};

struct RefCudaSymbolicCtx : SymbolicCtx {
  RefCudaSymbolicCtx(const CoalescedBlockMatrixSkel& s, const vector<int64_t>& permutation) : skel(s) {
    RC_LIB(cublasCreate(&cublasH));
    RC_LIB(cusolverDnCreate(&cusolverH));
    dLumpStart.load(s.lumpStart), dLumpToSpan.load(s.lumpToSpan), dSpanStart.load(s.spanStart);
    dSpanToLump.load(s.spanToLump), dSpanOffsetInLump.load(s.spanOffsetInLump), dChainColPtr.load(s.chainColPtr);
    dChainRowSpan.load(s.chainRowSpan), dChainData.load(s.chainData), dChainRowsTillEnd.load(s.chainRowsTillEnd);
    dBoardColPtr.load(s.boardColPtr), dBoardChainColOrd.load(s.boardChainColOrd), dPermutation.load(permutation);
    ix = Idx{dLumpStart.ptr,   dLumpToSpan.ptr,  dSpanStart.ptr,        dSpanToLump.ptr,  dSpanOffsetInLump.ptr, dChainColPtr.ptr,
             dChainRowSpan.ptr, dChainData.ptr,  dChainRowsTillEnd.ptr, dBoardColPtr.ptr, dBoardChainColOrd.ptr};
  }
  ~RefCudaSymbolicCtx() override {
    if (cublasH) cublasDestroy(cublasH);
    if (cusolverH) cusolverDnDestroy(cusolverH);
  }
  void setStream(void* s) override {
    stream = (cudaStream_t)s;
    RC_LIB(cublasSetStream(cublasH, stream));
    RC_LIB(cusolverDnSetStream(cusolverH, stream));
  }
  PermutedCoalescedAccessor deviceAccessor() override {
    PermutedCoalescedAccessor a;
    a.init(dSpanStart.ptr, dSpanToLump.ptr, dLumpStart.ptr, dSpanOffsetInLump.ptr, dChainColPtr.ptr, dChainRowSpan.ptr,
           dChainData.ptr, dPermutation.ptr);
    return a;
  }
  SymElimCtxPtr prepareElimination(int64_t lumpsBegin, int64_t lumpsEnd) override {
    auto* e = new RefCudaSymElimCtx;
    vector<int64_t> ptr(lumpsEnd - lumpsBegin + 1, 0);
    for (int64_t l = lumpsBegin; l < lumpsEnd; l++) {
      const int64_t nb = skel.chainColPtr[l + 1] - skel.chainColPtr[l] - 1;
      ptr[l - lumpsBegin + 1] = ptr[l - lumpsBegin] + nb * (nb + 1) / 2;
    }
    e->numPairs = ptr.back();
    e->pairPtr.load(ptr);
    return SymElimCtxPtr(e);
  }
  NumericCtxBase* createNumericCtxForType(std::type_index tIdx, int64_t tempBufSize, int batchSize) override;
  SolveCtxBase* createSolveCtxForType(std::type_index tIdx, int nRHS, int batchSize) override;

  const CoalescedBlockMatrixSkel& skel;
  cudaStream_t stream = nullptr;
  cublasHandle_t cublasH = nullptr;
  cusolverDnHandle_t cusolverH = nullptr;
  Dev<int64_t> dLumpStart, dLumpToSpan, dSpanStart, dSpanToLump, dSpanOffsetInLump, dChainColPtr, dChainRowSpan, dChainData,
      dChainRowsTillEnd, dBoardColPtr, dBoardChainColOrd, dPermutation;
  Idx ix;
};

inline int groups(int64_t n) { return (int)((n + 31) / 32); }

struct RefCudaNumericCtx : NumericCtx<double> {
  RefCudaNumericCtx(RefCudaSymbolicCtx& s, int64_t bufSize) : sym(s), spanToChainOffset(s.skel.numSpans()) {
    temp.atLeast(bufSize);
    dSpanToChainOffset.atLeast(spanToChainOffset.size());
  }
  void pseudoFactorSpans(double* data, int64_t spanBegin, int64_t spanEnd) override {
    if (spanEnd <= spanBegin) return;
    rc_factor_spans_kernel<<<groups(spanEnd - spanBegin), 32, 0, sym.stream>>>(sym.ix, data, spanBegin, spanEnd);
  }
  void doElimination(const SymElimCtx& elimData, double* data, int64_t lumpsBegin, int64_t lumpsEnd) override {
    const auto* e = dynamic_cast<const RefCudaSymElimCtx*>(&elimData);
    BASPACHO_CHECK_NOTNULL(e);
    if (lumpsEnd <= lumpsBegin) return;
    rc_factor_lumps_kernel<<<groups(lumpsEnd - lumpsBegin), 32, 0, sym.stream>>>(sym.ix, data, lumpsBegin, lumpsEnd);
    if (e->numPairs > 0)
      rc_elim_pairs_kernel<<<groups(e->numPairs), 32, 0, sym.stream>>>(sym.ix, data, lumpsBegin, lumpsEnd, e->pairPtr.ptr,
                                                                        e->numPairs);
    RC_CUDA(cudaGetLastError());
  }
  // row-major lower == column-major upper: the reference's trick (MatOpsCuda.cu:508-566)
  void potrf(int64_t n, double* data, int64_t offA) override {
    int lwork = 0;
    RC_LIB(cusolverDnDpotrf_bufferSize(sym.cusolverH, CUBLAS_FILL_MODE_UPPER, (int)n, data + offA, (int)n, &lwork));
    work.atLeast(lwork);
    info.atLeast(1);
    RC_LIB(cusolverDnDpotrf(sym.cusolverH, CUBLAS_FILL_MODE_UPPER, (int)n, data + offA, (int)n, work.ptr, lwork, info.ptr));
  }
  void trsm(int64_t n, int64_t k, double* data, int64_t offA, int64_t offB) override {
    const double one = 1.0;
    RC_LIB(cublasDtrsm(sym.cublasH, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)n, (int)k,
                       &one, data + offA, (int)n, data + offB, (int)n));
  }
  void saveSyrkGemm(int64_t m, int64_t n, int64_t k, const double* data, int64_t offset) override {
    const double one = 1.0, zero = 0.0;
    BASPACHO_CHECK_LE(m * n, (int64_t)temp.cap);
    RC_LIB(cublasDgemm(sym.cublasH, CUBLAS_OP_T, CUBLAS_OP_N, (int)m, (int)n, (int)k, &one, data + offset, (int)k, data + offset,
                       (int)k, &zero, temp.ptr, (int)m));
  }
  void prepareAssemble(int64_t targetLump) override {
    const auto& sk = sym.skel;
    for (int64_t i = sk.chainColPtr[targetLump]; i < sk.chainColPtr[targetLump + 1]; i++)
      spanToChainOffset[sk.chainRowSpan[i]] = sk.chainData[i];
    // the reference's synchronous whole-table copy per target lump (MatOpsCuda.cu:471-481); ordered with the stream
    RC_CUDA(cudaStreamSynchronize(sym.stream));
    RC_CUDA(cudaMemcpy(dSpanToChainOffset.ptr, spanToChainOffset.data(), spanToChainOffset.size() * sizeof(int64_t),
                       cudaMemcpyHostToDevice));
  }
  void assemble(double* data, int64_t rectRowBegin, int64_t dstStride, int64_t srcColDataOffset, int64_t srcRectWidth,
                int64_t numBlockRows, int64_t numBlockCols) override {
    if (numBlockRows * numBlockCols <= 0) return;
    rc_assemble_kernel<<<groups(numBlockRows * numBlockCols), 32, 0, sym.stream>>>(
        numBlockRows, numBlockCols, rectRowBegin, srcRectWidth, dstStride, sym.dChainRowsTillEnd.ptr + srcColDataOffset,
        sym.dChainRowSpan.ptr + srcColDataOffset, dSpanToChainOffset.ptr, sym.dSpanOffsetInLump.ptr, temp.ptr, data);
  }
  RefCudaSymbolicCtx& sym;
  Dev<double> temp, work;
  Dev<int> info;
  Dev<int64_t> dSpanToChainOffset;
  vector<int64_t> spanToChainOffset;
};

struct RefCudaSolveCtx : SolveCtx<double> {
  RefCudaSolveCtx(RefCudaSymbolicCtx& s, int nRHS_) : sym(s), nRHS(nRHS_) { buf.atLeast((size_t)s.skel.order() * nRHS_); }
  void sparseElimSolveL(const SymElimCtx&, const double* data, int64_t b, int64_t e, double* C, int64_t ldc) override {
    if (e <= b) return;
    rc_elim_diag_solve_kernel<<<groups(e - b), 32, 0, sym.stream>>>(sym.ix, data, C, ldc, nRHS, b, e, false);
    rc_elim_sub_mult_kernel<<<groups(e - b), 32, 0, sym.stream>>>(sym.ix, data, C, ldc, nRHS, b, e);
  }
  void sparseElimSolveLt(const SymElimCtx&, const double* data, int64_t b, int64_t e, double* C, int64_t ldc) override {
    if (e <= b) return;
    rc_elim_sub_mult_t_kernel<<<groups(e - b), 32, 0, sym.stream>>>(sym.ix, data, C, ldc, nRHS, b, e);
    rc_elim_diag_solve_kernel<<<groups(e - b), 32, 0, sym.stream>>>(sym.ix, data, C, ldc, nRHS, b, e, true);
  }
  void symm(const double* data, int64_t offM, int64_t n, const double* C, int64_t offC, int64_t ldc, double* D, int64_t ldd,
            double alpha) override {
    const double one = 1.0;
    RC_LIB(cublasDsymm(sym.cublasH, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, (int)n, nRHS, &alpha, data + offM, (int)n,
                       C + offC, (int)ldc, &one, D + offC, (int)ldd));
  }
  void solveL(const double* data, int64_t offM, int64_t n, double* C, int64_t offC, int64_t ldc) override {
    const double one = 1.0;
    RC_LIB(cublasDtrsm(sym.cublasH, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)n, nRHS,
                       &one, data + offM, (int)n, C + offC, (int)ldc));
  }
  void solveLt(const double* data, int64_t offM, int64_t n, double* C, int64_t offC, int64_t ldc) override {
    const double one = 1.0;
    RC_LIB(cublasDtrsm(sym.cublasH, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, (int)n, nRHS,
                       &one, data + offM, (int)n, C + offC, (int)ldc));
  }
  void gemv(const double* data, int64_t offM, int64_t nRows, int64_t nCols, const double* A, int64_t offA, int64_t lda,
            double alpha) override {
    const double zero = 0.0;
    RC_LIB(cublasDgemm(sym.cublasH, CUBLAS_OP_T, CUBLAS_OP_N, nRHS, (int)nRows, (int)nCols, &alpha, A + offA, (int)lda,
                       data + offM, (int)nCols, &zero, buf.ptr, nRHS));
  }
  void gemvT(const double* data, int64_t offM, int64_t nRows, int64_t nCols, double* A, int64_t offA, int64_t lda,
             double alpha) override {
    const double one = 1.0;
    RC_LIB(cublasDgemm(sym.cublasH, CUBLAS_OP_N, CUBLAS_OP_T, (int)nCols, nRHS, (int)nRows, &alpha, data + offM, (int)nCols,
                       buf.ptr, nRHS, &one, A + offA, (int)lda));
  }
  void assembleVec(int64_t chainColPtr, int64_t numColItems, double* C, int64_t ldc) override {
    if (numColItems <= 0) return;
    rc_assemble_vec_kernel<<<groups(numColItems), 32, 0, sym.stream>>>(sym.dChainRowsTillEnd.ptr + chainColPtr,
                                                                        sym.dChainRowSpan.ptr + chainColPtr, sym.dSpanStart.ptr,
                                                                        buf.ptr, numColItems, C, ldc, nRHS, false);
  }
  void assembleVecT(const double* C, int64_t ldc, int64_t chainColPtr, int64_t numColItems) override {
    if (numColItems <= 0) return;
    rc_assemble_vec_kernel<<<groups(numColItems), 32, 0, sym.stream>>>(
        sym.dChainRowsTillEnd.ptr + chainColPtr, sym.dChainRowSpan.ptr + chainColPtr, sym.dSpanStart.ptr, buf.ptr, numColItems,
        const_cast<double*>(C), ldc, nRHS, true);
  }
  RefCudaSymbolicCtx& sym;
  int nRHS;
  Dev<double> buf;
};

NumericCtxBase* RefCudaSymbolicCtx::createNumericCtxForType(std::type_index tIdx, int64_t tempBufSize, int batchSize) {
  if (tIdx == std::type_index(typeid(double)) && batchSize == 1) return new RefCudaNumericCtx(*this, tempBufSize);
  return nullptr;  // the baseline covers the double, single-matrix path only
}
SolveCtxBase* RefCudaSymbolicCtx::createSolveCtxForType(std::type_index tIdx, int nRHS, int batchSize) {
  if (tIdx == std::type_index(typeid(double)) && batchSize == 1) return new RefCudaSolveCtx(*this, nRHS);
  return nullptr;
}

struct RefCudaOps : Ops {
  SymbolicCtxPtr createSymbolicCtx(const CoalescedBlockMatrixSkel& skel, const vector<int64_t>& permutation) override {
    return SymbolicCtxPtr(new RefCudaSymbolicCtx(skel, permutation));
  }
};

}  // namespace

OpsPtr refCudaOps() { return OpsPtr(new RefCudaOps); }

}  // namespace BaSpaCho
