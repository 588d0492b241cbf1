// Dense building blocks of the supernodal factorization, hand-written for sm_100a:
//   gemmNT         - C = alpha A B^T + beta C; fp64 on the tensor pipe (DMMA mma.sync.m8n8k4.f64, the only fp64
//                    MMA on Blackwell: tcgen05 has no f64 kind), cp.async multi-stage smem pipeline; fp32 SIMT.
//                    Replaces cublas<t>gemm in the reference (MatOpsCuda.cu:568-590) and feeds the blocked
//                    potrf/trsm below (reference: cusolverDn<t>potrf :508-548, cublas<t>trsm :550-566).
//   potrfBlock     - one-CTA shared-memory Cholesky of a diagonal block (<= maxBlockDim)
//   trsmBlock      - X L^T = B by forward substitution, one thread per row, L in shared memory
//   potrfTrapezoid - recursive blocked Cholesky of a whole lump column (diagonal block + rows below)
//   trsmAny        - blocked X L^T = B for any size
#include <algorithm>
#include "B200Kernels.h"

namespace BaSpaCho {
namespace b200 {

std::atomic<int64_t>& launchCounter() {
  static std::atomic<int64_t> c{0};
  return c;
}

namespace {

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES>
__device__ __forceinline__ void cpAsync(uint32_t dst, const void* src, int srcBytes) {
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(srcBytes));
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(srcBytes));
  }
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cpAsyncWait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

struct GemmShape {
  int64_t m, n, k, lda, ldb, ldc;
  int lowerOnly;
};

// ------------------------------------------------------------------------------------------------ fp64 DMMA GEMM
// CTA tile BM x BN x BK, warp tile WM x WN built from m8n8k4 DMMA tiles; A and B tiles are both K-contiguous
// ([row][k] with a 4-double pad: the quad-strided fragment loads are bank-conflict free).
template <int BM, int BN, int BK, int WM, int WN, int STAGES, int VEC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
    gemm_nt_f64_kernel(GemmShape s, double alpha, Operand<double> Aop, Operand<double> Bop, double beta,
                       Operand<double> Cop) {
  constexpr int NWN = BN / WN;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LDS = BK + 4;
  constexpr int TM = WM / 8, TN = WN / 8;
  extern __shared__ __align__(16) double smemD[];
  double* As = smemD;
  double* Bs = smemD + STAGES * BM * LDS;

  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  if (s.lowerOnly && n0 > m0 + BM - 1) return;  // tile strictly above the diagonal
  const int b = blockIdx.z;
  const double* __restrict__ A = Aop.at(b);
  const double* __restrict__ B = Bop.at(b);
  double* __restrict__ C = Cop.at(b);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp / NWN) * WM, wn = (warp % NWN) * WN;
  const int g = lane >> 2, t = lane & 3;

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (int)((s.k + BK - 1) / BK);

  auto loadTile = [&](int stage, int kt) {
    const int64_t k0 = (int64_t)kt * BK;
    constexpr int CPR = BK / VEC;  // chunks per row
#pragma unroll
    for (int i = tid; i < BM * CPR; i += NT) {
      int r = i / CPR, kc = (i % CPR) * VEC;
      int64_t gr = m0 + r, gk = k0 + kc;
      bool ok = gr < s.m && gk < s.k;
      const double* src = ok ? A + gr * s.lda + gk : A;
      int bytes = ok ? (int)min((int64_t)VEC, s.k - gk) * 8 : 0;
      cpAsync<VEC * 8>(smemAddr(&As[(stage * BM + r) * LDS + kc]), src, bytes);
    }
#pragma unroll
    for (int i = tid; i < BN * CPR; i += NT) {
      int r = i / CPR, kc = (i % CPR) * VEC;
      int64_t gr = n0 + r, gk = k0 + kc;
      bool ok = gr < s.n && gk < s.k;
      const double* src = ok ? B + gr * s.ldb + gk : B;
      int bytes = ok ? (int)min((int64_t)VEC, s.k - gk) * 8 : 0;
      cpAsync<VEC * 8>(smemAddr(&Bs[(stage * BN + r) * LDS + kc]), src, bytes);
    }
  };

#pragma unroll
  for (int st = 0; st < STAGES - 1; st++) {
    if (st < KT) loadTile(st, st);
    cpAsyncCommit();
  }

  for (int kt = 0; kt < KT; kt++) {
    cpAsyncWait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) loadTile(nk % STAGES, nk);
      cpAsyncCommit();
    }
    const int st = kt % STAGES;
    const double* as = As + (st * BM + wm + g) * LDS + t;
    const double* bs = Bs + (st * BN + wn + g) * LDS + t;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) af[i] = as[i * 8 * LDS + kk];
#pragma unroll
      for (int j = 0; j < TN; j++) bf[j] = bs[j * 8 * LDS + kk];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cpAsyncWait<0>();

  // epilogue: each thread owns 2 consecutive columns of every 8x8 tile
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int64_t row = m0 + wm + i * 8 + g;
    if (row >= s.m) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      int64_t col = n0 + wn + j * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int64_t c = col + e;
        if (c >= s.n || (s.lowerOnly && c > row)) continue;
        double* dst = C + row * s.ldc + c;
        double v = alpha * acc[i][j][e];
        if (beta != 0.0) v += beta * *dst;
        *dst = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ SIMT GEMM (fp32 / generic)
template <typename T, int BM, int BN, int BK>
__global__ void __launch_bounds__(256) gemm_nt_simt_kernel(GemmShape s, T alpha, Operand<T> Aop, Operand<T> Bop,
                                                           T beta, Operand<T> Cop) {
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ T As[BK][BM + 1];
  __shared__ T Bs[BK][BN + 1];
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  if (s.lowerOnly && n0 > m0 + BM - 1) return;
  const int b = blockIdx.z;
  const T* __restrict__ A = Aop.at(b);
  const T* __restrict__ B = Bop.at(b);
  T* __restrict__ C = Cop.at(b);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = T(0);

  for (int64_t k0 = 0; k0 < s.k; k0 += BK) {
    for (int i = tid; i < BM * BK; i += 256) {
      int r = i / BK, kc = i % BK;
      int64_t gr = m0 + r, gk = k0 + kc;
      As[kc][r] = (gr < s.m && gk < s.k) ? A[gr * s.lda + gk] : T(0);
    }
    for (int i = tid; i < BN * BK; i += 256) {
      int r = i / BK, kc = i % BK;
      int64_t gr = n0 + r, gk = k0 + kc;
      Bs[kc][r] = (gr < s.n && gk < s.k) ? B[gr * s.ldb + gk] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      T a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; j++) bb[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] += a[i] * bb[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int64_t row = m0 + ty + 16 * i;
    if (row >= s.m) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      int64_t col = n0 + tx + 16 * j;
      if (col >= s.n || (s.lowerOnly && col > row)) continue;
      T* dst = C + row * s.ldc + col;
      T v = alpha * acc[i][j];
      if (beta != T(0)) v += beta * *dst;
      *dst = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ one-CTA Cholesky
// Right-looking, column by column, whole block resident in shared memory (odd row stride).
template <typename T>
__global__ void __launch_bounds__(512) potrf_block_kernel(int n, Operand<T> Aop, int64_t lda) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* S = reinterpret_cast<T*>(smemRaw);
  const int lds = n | 1;
  T* __restrict__ A = Aop.at(blockIdx.z);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n * n; i += nt) {
    int r = i / n, c = i - r * n;
    S[r * lds + c] = (c <= r) ? A[(int64_t)r * lda + c] : T(0);
  }
  __syncthreads();
  for (int j = 0; j < n; j++) {
    const T d = sqrt(S[j * lds + j]);
    const T inv = T(1) / d;
    __syncthreads();  // everyone has read the pivot before it is overwritten
    for (int i = j + 1 + tid; i < n; i += nt) S[i * lds + j] *= inv;
    if (tid == 0) S[j * lds + j] = d;
    __syncthreads();
    // trailing update of the lower triangle: rows i > j, cols j < c <= i
    const int m = n - j - 1;
    for (int e = tid; e < m * m; e += nt) {
      int ri = e / m, ci = e - ri * m;
      if (ci > ri) continue;
      int i = j + 1 + ri, c = j + 1 + ci;
      S[i * lds + c] -= S[i * lds + j] * S[c * lds + j];
    }
    __syncthreads();
  }
  for (int i = tid; i < n * n; i += nt) {
    int r = i / n, c = i - r * n;
    if (c <= r) A[(int64_t)r * lda + c] = S[r * lds + c];
  }
}

// ------------------------------------------------------------------------------------------------ X L^T = B, L <= block
// CTA = ROWS rows of B (one thread per row, row kept in shared memory with odd stride), L packed lower in smem.
template <typename T, int ROWS>
__global__ void __launch_bounds__(ROWS) trsm_block_kernel(int n, int64_t rows, Operand<T> Lop, int64_t ldl,
                                                          Operand<T> Bop, int64_t ldb) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* Ls = reinterpret_cast<T*>(smemRaw);  // packed lower: (j, q) -> j(j+1)/2 + q
  const int ldx = n | 1;
  T* Xs = Ls + (n * (n + 1) / 2 + 1);
  const T* __restrict__ L = Lop.at(blockIdx.z);
  T* __restrict__ B = Bop.at(blockIdx.z);
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * ROWS;
  const int nr = (int)min((int64_t)ROWS, rows - r0);
  for (int i = tid; i < n * n; i += ROWS) {
    int r = i / n, c = i - r * n;
    if (c <= r) Ls[r * (r + 1) / 2 + c] = L[(int64_t)r * ldl + c];
  }
  for (int i = tid; i < nr * n; i += ROWS) {
    int r = i / n, c = i - r * n;
    Xs[r * ldx + c] = B[(r0 + r) * ldb + c];
  }
  __syncthreads();
  if (tid < nr) {
    T* x = Xs + tid * ldx;
    for (int j = 0; j < n; j++) {
      const T* lj = Ls + j * (j + 1) / 2;
      T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      int q = 0;
      for (; q + 4 <= j; q += 4) {
        s0 += x[q] * lj[q];
        s1 += x[q + 1] * lj[q + 1];
        s2 += x[q + 2] * lj[q + 2];
        s3 += x[q + 3] * lj[q + 3];
      }
      for (; q < j; q++) s0 += x[q] * lj[q];
      x[j] = (x[j] - ((s0 + s1) + (s2 + s3))) / lj[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < nr * n; i += ROWS) {
    int r = i / n, c = i - r * n;
    B[(r0 + r) * ldb + c] = Xs[r * ldx + c];
  }
}

template <typename KernelT>
void setSmem(KernelT kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES>
void launchGemmF64(cudaStream_t st, int batch, const GemmShape& s, double alpha, Operand<double> A, Operand<double> B,
                   double beta, Operand<double> C, bool aligned16) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  size_t smem = (size_t)STAGES * (BM + BN) * (BK + 4) * sizeof(double);
  dim3 grid(ceilDiv(s.n, BN), ceilDiv(s.m, BM), batch);
  if (aligned16) {
    auto kern = gemm_nt_f64_kernel<BM, BN, BK, WM, WN, STAGES, 2>;
    static bool once = (setSmem(kern, smem), true);
    (void)once;
    kern<<<grid, NT, smem, st>>>(s, alpha, A, B, beta, C);
  } else {
    auto kern = gemm_nt_f64_kernel<BM, BN, BK, WM, WN, STAGES, 1>;
    static bool once = (setSmem(kern, smem), true);
    (void)once;
    kern<<<grid, NT, smem, st>>>(s, alpha, A, B, beta, C);
  }
  B200_LAUNCH_CHECK();
}

}  // namespace

template <>
void gemmNT<double>(cudaStream_t st, int batch, int64_t m, int64_t n, int64_t k, double alpha, Operand<double> A,
                    int64_t lda, Operand<double> B, int64_t ldb, double beta, Operand<double> C, int64_t ldc,
                    bool lowerOnly) {
  if (m <= 0 || n <= 0) return;
  GemmShape s{m, n, k, lda, ldb, ldc, lowerOnly ? 1 : 0};
  // 16-byte cp.async needs even element offsets/strides and 16B-aligned bases (cudaMalloc bases are; batch pointers
  // supplied by the caller are only guaranteed 8B aligned -> 8-byte copies there)
  bool aligned16 = !A.many && !B.many && (lda % 2 == 0) && (ldb % 2 == 0) && (A.off % 2 == 0) && (B.off % 2 == 0) &&
                   (A.bstride % 2 == 0) && (B.bstride % 2 == 0) && ((uintptr_t)A.base % 16 == 0) &&
                   ((uintptr_t)B.base % 16 == 0);
  if (m >= 96 && n >= 96)
    launchGemmF64<128, 128, 16, 64, 32, 4>(st, batch, s, alpha, A, B, beta, C, aligned16);
  else
    launchGemmF64<64, 64, 16, 32, 16, 4>(st, batch, s, alpha, A, B, beta, C, aligned16);
}

template <>
void gemmNT<float>(cudaStream_t st, int batch, int64_t m, int64_t n, int64_t k, float alpha, Operand<float> A,
                   int64_t lda, Operand<float> B, int64_t ldb, float beta, Operand<float> C, int64_t ldc,
                   bool lowerOnly) {
  if (m <= 0 || n <= 0) return;
  GemmShape s{m, n, k, lda, ldb, ldc, lowerOnly ? 1 : 0};
  dim3 grid(ceilDiv(n, 64), ceilDiv(m, 64), batch);
  gemm_nt_simt_kernel<float, 64, 64, 16><<<grid, 256, 0, st>>>(s, alpha, A, B, beta, C);
  B200_LAUNCH_CHECK();
}

template <>
int maxBlockDim<double>() {
  return 96;
}
template <>
int maxBlockDim<float>() {
  return 128;
}

template <typename T>
void potrfBlock(cudaStream_t st, int batch, int n, Operand<T> A, int64_t lda) {
  if (n <= 0) return;
  if (n > maxBlockDim<T>()) throw std::runtime_error("potrfBlock: block too large");
  size_t smem = (size_t)n * (n | 1) * sizeof(T);
  static bool once = (setSmem(potrf_block_kernel<T>, (size_t)maxBlockDim<T>() * (maxBlockDim<T>() | 1) * sizeof(T)), true);
  (void)once;
  int threads = n <= 16 ? 64 : n <= 32 ? 128 : n <= 64 ? 256 : 512;
  potrf_block_kernel<T><<<dim3(1, 1, batch), threads, smem, st>>>(n, A, lda);
  B200_LAUNCH_CHECK();
}

template <typename T>
void trsmBlock(cudaStream_t st, int batch, int n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B, int64_t ldb) {
  if (n <= 0 || rows <= 0) return;
  if (n > maxBlockDim<T>()) throw std::runtime_error("trsmBlock: block too large");
  constexpr int ROWS = 64;
  auto smemFor = [](int nn) { return ((size_t)nn * (nn + 1) / 2 + 1 + (size_t)ROWS * (nn | 1)) * sizeof(T); };
  static bool once = (setSmem(trsm_block_kernel<T, ROWS>, smemFor(maxBlockDim<T>())), true);
  (void)once;
  trsm_block_kernel<T, ROWS><<<dim3(ceilDiv(rows, ROWS), 1, batch), ROWS, smemFor(n), st>>>(n, rows, L, ldl, B, ldb);
  B200_LAUNCH_CHECK();
}

// Recursive blocked Cholesky of the (n + rowsBelow) x n trapezoid (row-major, ld). Columns [c0, c0 + w):
//   w small  : potrfBlock on the diagonal block + trsmBlock on every row below it
//   otherwise: left half; trailing update of the right half (lower-only GEMM, K = left width); right half
template <typename T>
static void potrfRec(cudaStream_t st, int batch, int64_t totalRows, int64_t c0, int64_t w, Operand<T> A, int64_t ld) {
  const int nb = maxBlockDim<T>();
  if (w <= nb) {
    Operand<T> diag = shifted(A, c0 * ld + c0);
    potrfBlock<T>(st, batch, (int)w, diag, ld);
    int64_t below = totalRows - (c0 + w);
    if (below > 0) trsmBlock<T>(st, batch, (int)w, below, diag, ld, shifted(A, (c0 + w) * ld + c0), ld);
    return;
  }
  int64_t blocks = (w + nb - 1) / nb;  // >= 2
  int64_t w1 = (blocks / 2) * nb;      // left width: a multiple of the block size, < w
  potrfRec<T>(st, batch, totalRows, c0, w1, A, ld);
  // rows below the left diagonal block: panel P = A[c0+w1 .., c0 .. c0+w1); update A[c0+w1.., c0+w1 .. c0+w) -= P P1^T
  int64_t r0 = c0 + w1, m = totalRows - r0, n = w - w1;
  Operand<T> P = shifted(A, r0 * ld + c0);
  gemmNT<T>(st, batch, m, n, w1, T(-1), P, ld, P, ld, T(1), shifted(A, r0 * ld + r0), ld, true);
  potrfRec<T>(st, batch, totalRows, r0, n, A, ld);
}

template <typename T>
void potrfTrapezoid(cudaStream_t st, int batch, int64_t n, int64_t rowsBelow, Operand<T> A, int64_t ld) {
  if (n <= 0) return;
  potrfRec<T>(st, batch, n + rowsBelow, 0, n, A, ld);
}

template <typename T>
void trsmAny(cudaStream_t st, int batch, int64_t n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B, int64_t ldb) {
  if (n <= 0 || rows <= 0) return;
  const int nb = maxBlockDim<T>();
  for (int64_t j0 = 0; j0 < n; j0 += nb) {
    int64_t jb = std::min<int64_t>(nb, n - j0);
    if (j0 > 0)  // B[:, j0:j0+jb] -= B[:, 0:j0] * L[j0:j0+jb, 0:j0]^T
      gemmNT<T>(st, batch, rows, jb, j0, T(-1), B, ldb, shifted(L, j0 * ldl), ldl, T(1), shifted(B, j0), ldb, false);
    trsmBlock<T>(st, batch, (int)jb, rows, shifted(L, j0 * ldl + j0), ldl, shifted(B, j0), ldb);
  }
}

template void potrfBlock<double>(cudaStream_t, int, int, Operand<double>, int64_t);
template void potrfBlock<float>(cudaStream_t, int, int, Operand<float>, int64_t);
template void trsmBlock<double>(cudaStream_t, int, int, int64_t, Operand<double>, int64_t, Operand<double>, int64_t);
template void trsmBlock<float>(cudaStream_t, int, int, int64_t, Operand<float>, int64_t, Operand<float>, int64_t);
template void potrfTrapezoid<double>(cudaStream_t, int, int64_t, int64_t, Operand<double>, int64_t);
template void potrfTrapezoid<float>(cudaStream_t, int, int64_t, int64_t, Operand<float>, int64_t);
template void trsmAny<double>(cudaStream_t, int, int64_t, int64_t, Operand<double>, int64_t, Operand<double>, int64_t);
template void trsmAny<float>(cudaStream_t, int, int64_t, int64_t, Operand<float>, int64_t, Operand<float>, int64_t);

}  // namespace b200
}  // namespace BaSpaCho
