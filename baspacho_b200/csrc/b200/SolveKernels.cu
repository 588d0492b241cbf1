// Dense-lump pieces of the triangular solves (sm_100a): block triangular solve in shared memory, row-wise and
// column-wise matrix-vector products (HBM-bound: the factor is streamed once, coalesced along rows), symmetric
// block product. Replace cublas<t>trsm / gemm / symm of the reference solve path (MatOpsCuda.cu:1093-1181).
#include <algorithm>
#include "B200Kernels.h"

namespace BaSpaCho {
namespace b200 {
namespace {

constexpr int kTrsvWarps = 4;

// one CTA: L (n x n lower) in shared memory, one warp per right-hand side
template <typename T>
__global__ void __launch_bounds__(kTrsvWarps * 32) trsv_block_kernel(int n, Operand<T> Lop, int64_t ldl, Operand<T> Cop,
                                                                     int64_t ldc, int nRHS, bool transposed) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* Ls = reinterpret_cast<T*>(smemRaw);
  const int lds = n | 1;
  T* xs = Ls + n * lds;
  const T* __restrict__ L = Lop.at(blockIdx.z);
  T* C = Cop.at(blockIdx.z);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n * n; i += blockDim.x) {
    int r = i / n, c = i - r * n;
    if (c <= r) Ls[r * lds + c] = L[(int64_t)r * ldl + c];
  }
  __syncthreads();
  T* x = xs + warp * n;
  for (int rhs = warp; rhs < nRHS; rhs += kTrsvWarps) {
    T* c = C + (int64_t)rhs * ldc;
    for (int i = lane; i < n; i += 32) x[i] = c[i];
    __syncwarp();
    if (!transposed) {
      for (int j = 0; j < n; j++) {
        const T xj = x[j] / Ls[j * lds + j];
        __syncwarp();
        if (lane == 0) x[j] = xj;
        for (int i = j + 1 + lane; i < n; i += 32) x[i] -= Ls[i * lds + j] * xj;
        __syncwarp();
      }
    } else {
      for (int j = n - 1; j >= 0; j--) {
        const T xj = x[j] / Ls[j * lds + j];
        __syncwarp();
        if (lane == 0) x[j] = xj;
        for (int i = lane; i < j; i += 32) x[i] -= Ls[j * lds + i] * xj;
        __syncwarp();
      }
    }
    for (int i = lane; i < n; i += 32) c[i] = x[i];
    __syncwarp();
  }
}

// warp per row of M, up to 4 right-hand sides per pass
template <typename T>
__global__ void __launch_bounds__(256) gemv_rows_warp_kernel(int64_t rows, int64_t cols, T alpha, Operand<T> Mop,
                                                             int64_t ldm, Operand<T> Xop, int64_t ldx, Operand<T> Oop,
                                                             int64_t ors, int64_t ocs, int nRHS, bool accumulate) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const T* __restrict__ m = Mop.at(blockIdx.z) + row * ldm;
  const T* __restrict__ X = Xop.at(blockIdx.z);
  T* out = Oop.at(blockIdx.z) + row * ors;
  for (int c0 = 0; c0 < nRHS; c0 += 4) {
    T acc[4] = {0, 0, 0, 0};
    const int nc = min(4, nRHS - c0);
    for (int64_t q = lane; q < cols; q += 32) {
      T mv = m[q];
#pragma unroll
      for (int cc = 0; cc < 4; cc++)
        if (cc < nc) acc[cc] += mv * X[(int64_t)(c0 + cc) * ldx + q];
    }
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
      T v = acc[cc];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && cc < nc) {
        T* dst = out + (int64_t)(c0 + cc) * ocs;
        *dst = accumulate ? *dst + alpha * v : alpha * v;
      }
    }
  }
}

// thread per row (narrow M: cols <= 16)
template <typename T>
__global__ void __launch_bounds__(256) gemv_rows_thread_kernel(int64_t rows, int64_t cols, T alpha, Operand<T> Mop,
                                                               int64_t ldm, Operand<T> Xop, int64_t ldx,
                                                               Operand<T> Oop, int64_t ors, int64_t ocs, int nRHS,
                                                               bool accumulate) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const T* __restrict__ m = Mop.at(blockIdx.z) + row * ldm;
  const T* __restrict__ X = Xop.at(blockIdx.z);
  T* out = Oop.at(blockIdx.z) + row * ors;
  for (int c = 0; c < nRHS; c++) {
    T acc = 0;
    for (int q = 0; q < cols; q++) acc += m[q] * X[(int64_t)c * ldx + q];
    T* dst = out + (int64_t)c * ocs;
    *dst = accumulate ? *dst + alpha * acc : alpha * acc;
  }
}

// CTA = 32 columns x 8 row groups; each group strides over the rows, fixed-order reduction in shared memory
template <typename T>
__global__ void __launch_bounds__(256) gemv_cols_t_kernel(int64_t rows, int64_t cols, T alpha, Operand<T> Mop,
                                                          int64_t ldm, Operand<T> Iop, int64_t irs, int64_t ics,
                                                          Operand<T> Xop, int64_t ldx, int nRHS) {
  __shared__ T red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 32 + tx;
  const T* __restrict__ M = Mop.at(blockIdx.z);
  const T* __restrict__ in = Iop.at(blockIdx.z);
  T* X = Xop.at(blockIdx.z);
  for (int c = 0; c < nRHS; c++) {
    T acc = 0;
    if (q < cols)
      for (int64_t r = ty; r < rows; r += 8) acc += M[r * ldm + q] * in[r * irs + (int64_t)c * ics];
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && q < cols) {
      T tot = 0;
#pragma unroll
      for (int g = 0; g < 8; g++) tot += red[g][tx];
      X[(int64_t)c * ldx + q] += alpha * tot;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(128) symm_lower_kernel(int64_t n, T alpha, Operand<T> Mop, Operand<T> Xop,
                                                         int64_t ldx, Operand<T> Yop, int64_t ldy, int nRHS) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * nRHS) return;
  const int64_t i = gid % n;
  const int c = (int)(gid / n);
  const T* __restrict__ M = Mop.at(blockIdx.z);
  const T* __restrict__ x = Xop.at(blockIdx.z) + (int64_t)c * ldx;
  T acc = 0;
  for (int64_t j = 0; j <= i; j++) acc += M[i * n + j] * x[j];
  for (int64_t j = i + 1; j < n; j++) acc += M[j * n + i] * x[j];
  Yop.at(blockIdx.z)[(int64_t)c * ldy + i] += alpha * acc;
}

template <typename T>
void trsvBlock(cudaStream_t st, int batch, int n, Operand<T> L, int64_t ldl, Operand<T> C, int64_t ldc, int nRHS,
               bool transposed) {
  auto smemFor = [](int nn) { return ((size_t)nn * (nn | 1) + (size_t)kTrsvWarps * nn) * sizeof(T); };
  static bool once = [&] {
    size_t mx = smemFor(maxBlockDim<T>());
    if (mx > 48 * 1024)
      B200_CUDA(cudaFuncSetAttribute(trsv_block_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
    return true;
  }();
  (void)once;
  trsv_block_kernel<T><<<dim3(1, 1, batch), kTrsvWarps * 32, smemFor(n), st>>>(n, L, ldl, C, ldc, nRHS, transposed);
  B200_LAUNCH_CHECK();
}

}  // namespace

template <typename T>
void gemvRows(cudaStream_t st, int batch, int64_t rows, int64_t cols, T alpha, Operand<T> M, int64_t ldm, Operand<T> X,
              int64_t ldx, Operand<T> out, int64_t outRowStride, int64_t outColStride, int nRHS, bool accumulate) {
  if (rows <= 0 || nRHS <= 0) return;
  if (cols <= 16)
    gemv_rows_thread_kernel<T><<<dim3(ceilDiv(rows, 256), 1, batch), 256, 0, st>>>(
        rows, cols, alpha, M, ldm, X, ldx, out, outRowStride, outColStride, nRHS, accumulate);
  else
    gemv_rows_warp_kernel<T><<<dim3(ceilDiv(rows, 8), 1, batch), 256, 0, st>>>(
        rows, cols, alpha, M, ldm, X, ldx, out, outRowStride, outColStride, nRHS, accumulate);
  B200_LAUNCH_CHECK();
}

template <typename T>
void gemvColsT(cudaStream_t st, int batch, int64_t rows, int64_t cols, T alpha, Operand<T> M, int64_t ldm,
               Operand<T> in, int64_t inRowStride, int64_t inColStride, Operand<T> X, int64_t ldx, int nRHS) {
  if (rows <= 0 || cols <= 0 || nRHS <= 0) return;
  gemv_cols_t_kernel<T><<<dim3(ceilDiv(cols, 32), 1, batch), 256, 0, st>>>(rows, cols, alpha, M, ldm, in, inRowStride,
                                                                          inColStride, X, ldx, nRHS);
  B200_LAUNCH_CHECK();
}

template <typename T>
void symmLower(cudaStream_t st, int batch, int64_t n, T alpha, Operand<T> M, Operand<T> X, int64_t ldx, Operand<T> Y,
               int64_t ldy, int nRHS) {
  if (n <= 0 || nRHS <= 0) return;
  symm_lower_kernel<T><<<dim3(ceilDiv(n * nRHS, 128), 1, batch), 128, 0, st>>>(n, alpha, M, X, ldx, Y, ldy, nRHS);
  B200_LAUNCH_CHECK();
}

template <typename T>
void trsvAny(cudaStream_t st, int batch, int64_t n, Operand<T> L, int64_t ldl, Operand<T> C, int64_t ldc, int nRHS,
             bool transposed) {
  if (n <= 0 || nRHS <= 0) return;
  const int nb = maxBlockDim<T>();
  if (!transposed) {
    for (int64_t j0 = 0; j0 < n; j0 += nb) {
      int64_t jb = std::min<int64_t>(nb, n - j0), rb = n - j0 - jb;
      trsvBlock<T>(st, batch, (int)jb, shifted(L, j0 * ldl + j0), ldl, shifted(C, j0), ldc, nRHS, false);
      if (rb > 0)  // x[below] -= L[below, block] * x[block]
        gemvRows<T>(st, batch, rb, jb, T(-1), shifted(L, (j0 + jb) * ldl + j0), ldl, shifted(C, j0), ldc,
                    shifted(C, j0 + jb), 1, ldc, nRHS, true);
    }
  } else {
    int64_t j0 = ((n - 1) / nb) * nb;
    for (; j0 >= 0; j0 -= nb) {
      int64_t jb = std::min<int64_t>(nb, n - j0);
      trsvBlock<T>(st, batch, (int)jb, shifted(L, j0 * ldl + j0), ldl, shifted(C, j0), ldc, nRHS, true);
      if (j0 > 0)  // x[before] -= L[block, before]^T * x[block]
        gemvColsT<T>(st, batch, jb, j0, T(-1), shifted(L, j0 * ldl), ldl, shifted(C, j0), 1, ldc, C, ldc, nRHS);
    }
  }
}

#define B200_INSTANTIATE_SOLVE(T)                                                                                       \
  template void gemvRows<T>(cudaStream_t, int, int64_t, int64_t, T, Operand<T>, int64_t, Operand<T>, int64_t,          \
                            Operand<T>, int64_t, int64_t, int, bool);                                                   \
  template void gemvColsT<T>(cudaStream_t, int, int64_t, int64_t, T, Operand<T>, int64_t, Operand<T>, int64_t, int64_t, \
                             Operand<T>, int64_t, int);                                                                 \
  template void symmLower<T>(cudaStream_t, int, int64_t, T, Operand<T>, Operand<T>, int64_t, Operand<T>, int64_t, int); \
  template void trsvAny<T>(cudaStream_t, int, int64_t, Operand<T>, int64_t, Operand<T>, int64_t, int, bool);
B200_INSTANTIATE_SOLVE(double)
B200_INSTANTIATE_SOLVE(float)

}  // namespace b200
}  // namespace BaSpaCho
