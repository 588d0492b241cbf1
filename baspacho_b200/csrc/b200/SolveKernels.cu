// Dense-lump pieces of the triangular solves (sm_100a): block triangular solve in shared memory, row-wise and
// column-wise matrix-vector products (HBM-bound: the factor is streamed once, coalesced along rows), symmetric
// block product. Replace cublas<t>trsm / gemm / symm of the reference solve path (MatOpsCuda.cu:1093-1181).
#include <algorithm>
#include "B200Kernels.h"

namespace BaSpaCho {
namespace b200 {
namespace {

constexpr int kTrsvWarps = 4;

// one CTA: L (n x n lower, n <= 96) in shared memory (odd stride: row and column walks are both conflict free),
// one warp per right-hand side with the vector in registers (3 entries per lane); the pivot entry is broadcast
// with a shuffle, so a column step costs one shuffle + one multiply + up to 3 FMAs and no barrier.
constexpr int kTB = 96, kTLD = 97;
template <typename T>
__global__ void __launch_bounds__(kTrsvWarps * 32) trsv_block_kernel(int n, Operand<T> Lop, int64_t ldl, Operand<T> Cop,
                                                                     int64_t ldc, int nRHS, bool transposed) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* Ls = reinterpret_cast<T*>(smemRaw);
  T* invd = Ls + kTB * kTLD;
  const T* __restrict__ L = Lop.at(blockIdx.z);
  T* C = Cop.at(blockIdx.z);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    T tmp[(kTB / kTrsvWarps) * 3];
#pragma unroll
    for (int a = 0; a < kTB / kTrsvWarps; a++)
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int r = warp + kTrsvWarps * a, c = lane + 32 * u;
        tmp[a * 3 + u] = (c <= r && r < n) ? L[(int64_t)r * ldl + c] : T(0);
      }
#pragma unroll
    for (int a = 0; a < kTB / kTrsvWarps; a++)
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int r = warp + kTrsvWarps * a, c = lane + 32 * u;
        if (c <= r && r < n) Ls[r * kTLD + c] = tmp[a * 3 + u];
      }
  }
  __syncthreads();
  if (tid < n) invd[tid] = T(1) / Ls[tid * kTLD + tid];
  __syncthreads();
  for (int rhs = warp; rhs < nRHS; rhs += kTrsvWarps) {
    T* c = C + (int64_t)rhs * ldc;
    T x[3];
#pragma unroll
    for (int u = 0; u < 3; u++) x[u] = (lane + 32 * u < n) ? c[lane + 32 * u] : T(0);
    if (!transposed) {
#pragma unroll
      for (int uj = 0; uj < 3; uj++)
        for (int jj = 0; jj < 32 && 32 * uj + jj < n; jj++) {
          const int j = 32 * uj + jj;
          const T xj = __shfl_sync(0xffffffffu, x[uj], jj) * invd[j];
          if (lane == jj) x[uj] = xj;
#pragma unroll
          for (int u = uj; u < 3; u++) {
            const int i = lane + 32 * u;
            if (i > j && i < n) x[u] -= Ls[i * kTLD + j] * xj;
          }
        }
    } else {
#pragma unroll
      for (int uj = 2; uj >= 0; uj--)
        for (int jj = 31; jj >= 0; jj--) {
          const int j = 32 * uj + jj;
          if (j >= n) continue;
          const T xj = __shfl_sync(0xffffffffu, x[uj], jj) * invd[j];
          if (lane == jj) x[uj] = xj;
#pragma unroll
          for (int u = 0; u <= uj; u++) {
            const int i = lane + 32 * u;
            if (i < j) x[u] -= Ls[j * kTLD + i] * xj;
          }
        }
    }
#pragma unroll
    for (int u = 0; u < 3; u++)
      if (lane + 32 * u < n) c[lane + 32 * u] = x[u];
  }
}

// warp per row of M, up to 4 right-hand sides per pass
template <typename T>
__global__ void __launch_bounds__(256) gemv_rows_warp_kernel(int64_t rows, int64_t cols, T alpha, Operand<T> Mop,
                                                             int64_t ldm, Operand<T> Xop, int64_t ldx, Operand<T> Oop,
                                                             int64_t ors, int64_t ocs, int nRHS, bool accumulate,
                                                             const int64_t* __restrict__ rowMap) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const T* __restrict__ m = Mop.at(blockIdx.z) + row * ldm;
  const T* __restrict__ X = Xop.at(blockIdx.z);
  T* out = Oop.at(blockIdx.z) + (rowMap ? rowMap[row] : row) * ors;  // rowMap: scatter (fused assembleVec)
  for (int c0 = 0; c0 < nRHS; c0 += 4) {
    T acc[4] = {0, 0, 0, 0};
    const int nc = min(4, nRHS - c0);
    int64_t q = lane;
    for (; q + 224 < cols; q += 256) {  // 8 independent row loads in flight per lane (the loop is latency bound)
      T mv[8];
#pragma unroll
      for (int e = 0; e < 8; e++) mv[e] = m[q + 32 * e];
#pragma unroll
      for (int cc = 0; cc < 4; cc++)
        if (cc < nc) {
          const T* xc = X + (int64_t)(c0 + cc) * ldx + q;
          T p0 = mv[0] * xc[0] + mv[1] * xc[32], p1 = mv[2] * xc[64] + mv[3] * xc[96];
          T p2 = mv[4] * xc[128] + mv[5] * xc[160], p3 = mv[6] * xc[192] + mv[7] * xc[224];
          acc[cc] += (p0 + p1) + (p2 + p3);
        }
    }
    for (; q < cols; q += 32) {
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
                                                               bool accumulate, const int64_t* __restrict__ rowMap) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const T* __restrict__ m = Mop.at(blockIdx.z) + row * ldm;
  const T* __restrict__ X = Xop.at(blockIdx.z);
  T* out = Oop.at(blockIdx.z) + (rowMap ? rowMap[row] : row) * ors;
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
                                                          Operand<T> Xop, int64_t ldx, int nRHS,
                                                          const int64_t* __restrict__ rowMap) {
  __shared__ T red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 32 + tx;
  const T* __restrict__ M = Mop.at(blockIdx.z);
  const T* __restrict__ in = Iop.at(blockIdx.z);
  T* X = Xop.at(blockIdx.z);
  for (int c = 0; c < nRHS; c++) {
    T acc = 0;
    if (q < cols)
      for (int64_t r = ty; r < rows; r += 8)
        acc += M[r * ldm + q] * in[(rowMap ? rowMap[r] : r) * irs + (int64_t)c * ics];  // rowMap: fused assembleVecT
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

// the same product split over row chunks (grid.y) so that tall panels fill the machine: pass 1 writes the partial
// column sums of every chunk, pass 2 adds them in chunk order (deterministic) into X
template <typename T>
__global__ void __launch_bounds__(256) gemv_cols_t_part_kernel(int64_t rows, int64_t cols, Operand<T> Mop, int64_t ldm,
                                                               Operand<T> Iop, int64_t irs, int64_t ics, Operand<T> Pop,
                                                               int nRHS, int64_t rowsPerChunk,
                                                               const int64_t* __restrict__ rowMap) {
  __shared__ T red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 32 + tx;
  const T* __restrict__ M = Mop.at(blockIdx.z);
  const T* __restrict__ in = Iop.at(blockIdx.z);
  T* part = Pop.at(blockIdx.z) + (int64_t)blockIdx.y * nRHS * cols;
  const int64_t r0 = (int64_t)blockIdx.y * rowsPerChunk, r1 = min(rows, r0 + rowsPerChunk);
  for (int c = 0; c < nRHS; c++) {
    T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (q < cols) {
      int64_t r = r0 + ty;
      for (; r + 24 < r1; r += 32) {
        const T m0 = M[r * ldm + q], m1 = M[(r + 8) * ldm + q], m2 = M[(r + 16) * ldm + q], m3 = M[(r + 24) * ldm + q];
        const int64_t i0 = rowMap ? rowMap[r] : r, i1 = rowMap ? rowMap[r + 8] : r + 8;
        const int64_t i2 = rowMap ? rowMap[r + 16] : r + 16, i3 = rowMap ? rowMap[r + 24] : r + 24;
        a0 += m0 * in[i0 * irs + (int64_t)c * ics], a1 += m1 * in[i1 * irs + (int64_t)c * ics];
        a2 += m2 * in[i2 * irs + (int64_t)c * ics], a3 += m3 * in[i3 * irs + (int64_t)c * ics];
      }
      for (; r < r1; r += 8) a0 += M[r * ldm + q] * in[(rowMap ? rowMap[r] : r) * irs + (int64_t)c * ics];
    }
    red[ty][tx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ty == 0 && q < cols) {
      T tot = 0;
#pragma unroll
      for (int g = 0; g < 8; g++) tot += red[g][tx];
      part[(int64_t)c * cols + q] = tot;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gemv_cols_t_reduce_kernel(int64_t cols, T alpha, Operand<T> Pop, int chunks,
                                                                 Operand<T> Xop, int64_t ldx, int nRHS) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cols * nRHS) return;
  const int64_t q = i % cols, c = i / cols;
  const T* __restrict__ part = Pop.at(blockIdx.z);
  T tot = 0;
  for (int k = 0; k < chunks; k++) tot += part[((int64_t)k * nRHS + c) * cols + q];
  Xop.at(blockIdx.z)[c * ldx + q] += alpha * tot;
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
  if (n > kTB) throw std::runtime_error("trsvBlock: block too large");
  auto smemFor = [](int) { return ((size_t)kTB * kTLD + kTB) * sizeof(T); };
  ensureDynSmem((const void*)trsv_block_kernel<T>, smemFor(maxBlockDim<T>()));
  ProfScope prof(st, KC_SOLVE_DENSE, (double)n * n * nRHS * batch, (double)n * (n + 1) / 2 * sizeof(T) * batch);
  trsv_block_kernel<T><<<dim3(1, 1, batch), kTrsvWarps * 32, smemFor(n), st>>>(n, L, ldl, C, ldc, nRHS, transposed);
  B200_LAUNCH_CHECK();
}

}  // namespace

template <typename T>
void gemvRows(cudaStream_t st, int batch, int64_t rows, int64_t cols, T alpha, Operand<T> M, int64_t ldm, Operand<T> X,
              int64_t ldx, Operand<T> out, int64_t outRowStride, int64_t outColStride, int nRHS, bool accumulate,
              const int64_t* rowMap) {
  if (rows <= 0 || nRHS <= 0) return;
  ProfScope prof(st, KC_SOLVE_DENSE, 2.0 * rows * cols * nRHS * batch, (double)rows * cols * sizeof(T) * batch);
  if (cols <= 16)
    gemv_rows_thread_kernel<T><<<dim3(ceilDiv(rows, 256), 1, batch), 256, 0, st>>>(
        rows, cols, alpha, M, ldm, X, ldx, out, outRowStride, outColStride, nRHS, accumulate, rowMap);
  else
    gemv_rows_warp_kernel<T><<<dim3(ceilDiv(rows, 8), 1, batch), 256, 0, st>>>(
        rows, cols, alpha, M, ldm, X, ldx, out, outRowStride, outColStride, nRHS, accumulate, rowMap);
  B200_LAUNCH_CHECK();
}

template <typename T>
void gemvColsT(cudaStream_t st, int batch, int64_t rows, int64_t cols, T alpha, Operand<T> M, int64_t ldm,
               Operand<T> in, int64_t inRowStride, int64_t inColStride, Operand<T> X, int64_t ldx, int nRHS,
               Operand<T> part, int64_t partCapacity, const int64_t* rowMap) {
  if (rows <= 0 || cols <= 0 || nRHS <= 0) return;
  ProfScope prof(st, KC_SOLVE_DENSE, 2.0 * rows * cols * nRHS * batch, (double)rows * cols * sizeof(T) * batch);
  // tall panel: split the rows over grid.y (about 4 CTAs per SM in total), partial sums through `part`
  const int64_t colTiles = ceilDiv(cols, 32);
  int64_t chunks = std::min<int64_t>({(592 + colTiles * batch - 1) / (colTiles * batch), (rows + 63) / 64,
                                      part.base ? partCapacity / (cols * nRHS) : 0});
  if (chunks >= 2) {
    const int64_t rpc = ((rows + chunks - 1) / chunks + 7) / 8 * 8;
    chunks = (rows + rpc - 1) / rpc;
    gemv_cols_t_part_kernel<T><<<dim3((unsigned)colTiles, (unsigned)chunks, batch), 256, 0, st>>>(
        rows, cols, M, ldm, in, inRowStride, inColStride, part, nRHS, rpc, rowMap);
    B200_LAUNCH_CHECK();
    gemv_cols_t_reduce_kernel<T><<<dim3(ceilDiv(cols * nRHS, 256), 1, batch), 256, 0, st>>>(cols, alpha, part, (int)chunks,
                                                                                            X, ldx, nRHS);
    B200_LAUNCH_CHECK();
    return;
  }
  gemv_cols_t_kernel<T><<<dim3(ceilDiv(cols, 32), 1, batch), 256, 0, st>>>(rows, cols, alpha, M, ldm, in, inRowStride,
                                                                          inColStride, X, ldx, nRHS, rowMap);
  B200_LAUNCH_CHECK();
}

template <typename T>
void symmLower(cudaStream_t st, int batch, int64_t n, T alpha, Operand<T> M, Operand<T> X, int64_t ldx, Operand<T> Y,
               int64_t ldy, int nRHS) {
  if (n <= 0 || nRHS <= 0) return;
  symm_lower_kernel<T><<<dim3(ceilDiv(n * nRHS, 128), 1, batch), 128, 0, st>>>(n, alpha, M, X, ldx, Y, ldy, nRHS);
  B200_LAUNCH_CHECK();
}

// One block step of the dense triangular solve in ONE launch: every CTA solves the diagonal block (redundantly, one
// warp per right-hand side, vector in registers) and then applies its share of the update with the solved block:
//   forward : y[row] -= L[row, block] . x_block      for 64 rows below the block per CTA (warp per row)
//   backward: y[col] -= L[block, col]^T . x_block    for 128 columns before the block per CTA (thread per column)
// The solved block goes to a scratch vector (the input block must stay intact for the other CTAs).
constexpr int kStepRows = 64, kStepCols = 128;
// the CTA's share of the update with the solved block xs (see solve_step_kernel)
template <typename T, bool TR>
__device__ __forceinline__ void stepUpdate(int64_t n, int64_t j0, int jb, const T* __restrict__ L, int64_t ldl, T* C,
                                           int64_t ldc, const T* xs, int g0, int ng) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!TR) {
      const int64_t rbeg = j0 + jb + (int64_t)blockIdx.x * kStepRows;
      constexpr int RPW = kStepRows / kTrsvWarps;  // rows per warp
      constexpr int RB = 8;                        // rows in flight per warp
      for (int rr0 = 0; rr0 < RPW; rr0 += RB) {
        const int64_t rowBase = rbeg + warp * RPW + rr0;
        if (rowBase >= n) break;
        T mv[RB][3];
#pragma unroll
        for (int rr = 0; rr < RB; rr++)
#pragma unroll
          for (int u = 0; u < 3; u++)
            mv[rr][u] = (rowBase + rr < n && lane + 32 * u < jb) ? L[(rowBase + rr) * ldl + j0 + lane + 32 * u] : T(0);
        for (int q = 0; q < ng; q++) {
          const T x0 = xs[q * kTB + lane], x1 = xs[q * kTB + lane + 32], x2 = xs[q * kTB + lane + 64];
          T v[RB];
#pragma unroll
          for (int rr = 0; rr < RB; rr++) v[rr] = mv[rr][0] * x0 + mv[rr][1] * x1 + mv[rr][2] * x2;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int rr = 0; rr < RB; rr++) v[rr] += __shfl_xor_sync(0xffffffffu, v[rr], o);
#pragma unroll
          for (int rr = 0; rr < RB; rr++)
            if (lane == rr && rowBase + rr < n) C[(int64_t)(g0 + q) * ldc + rowBase + rr] -= v[rr];
        }
      }
    } else {
      const int64_t col = (int64_t)blockIdx.x * kStepCols + tid;
      if (col < j0) {
        T acc[kTrsvWarps];
#pragma unroll
        for (int q = 0; q < kTrsvWarps; q++) acc[q] = T(0);
        const T* m = L + j0 * ldl + col;
#pragma unroll 8
        for (int r = 0; r < jb; r++) {
          const T mv = m[(int64_t)r * ldl];
#pragma unroll
          for (int q = 0; q < kTrsvWarps; q++) acc[q] += mv * xs[q * kTB + r];
        }
#pragma unroll
        for (int q = 0; q < kTrsvWarps; q++)
          if (q < ng) C[(int64_t)(g0 + q) * ldc + col] -= acc[q];
      }
    }
}

template <typename T, bool TR>
__global__ void __launch_bounds__(kTrsvWarps * 32)
    solve_step_kernel(int64_t n, int64_t j0, int jb, Operand<T> Lop, int64_t ldl, Operand<T> Cop, int64_t ldc,
                      Operand<T> Xop, int64_t ldx, int nRHS, Operand<T> Wop, bool useInv) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* Ls = reinterpret_cast<T*>(smemRaw);
  T* invd = Ls + kTB * kTLD;
  T* xs = invd + kTB;  // [kTrsvWarps][kTB]
  const T* __restrict__ L = Lop.at(blockIdx.z);
  T* C = Cop.at(blockIdx.z);
  T* X = Xop.at(blockIdx.z);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (useInv) {
    // diagonal block solve = product with the precomputed inverse of the block (no serial column chain):
    // forward x = W b reads W^T, backward x = W^T b reads W, both coalesced over the output index
    const T* __restrict__ Wm = Wop.at(blockIdx.z) + (j0 / kTB) * (int64_t)(2 * kTB * kTB) + (TR ? 0 : kTB * kTB);
    T* bs = Ls;  // [kTrsvWarps][kTB] right-hand sides of the group
    for (int g0 = 0; g0 < nRHS; g0 += kTrsvWarps) {
      const int ng = min(kTrsvWarps, nRHS - g0);
      for (int i = tid; i < kTrsvWarps * kTB; i += kTrsvWarps * 32) {
        const int q = i / kTB, r = i % kTB;
        bs[i] = (q < ng && r < jb) ? C[(int64_t)(g0 + q) * ldc + j0 + r] : T(0);
      }
      __syncthreads();
      if (tid < kTB) {
        T acc[kTrsvWarps];
#pragma unroll
        for (int q = 0; q < kTrsvWarps; q++) acc[q] = T(0);
#pragma unroll 8
        for (int r = 0; r < kTB; r++) {
          const T wv = Wm[r * kTB + tid];
#pragma unroll
          for (int q = 0; q < kTrsvWarps; q++) acc[q] += wv * bs[q * kTB + r];
        }
#pragma unroll
        for (int q = 0; q < kTrsvWarps; q++) {
          xs[q * kTB + tid] = acc[q];
          if (blockIdx.x == 0 && q < ng && tid < jb) X[(int64_t)(g0 + q) * ldx + j0 + tid] = acc[q];
        }
      }
      __syncthreads();
      stepUpdate<T, TR>(n, j0, jb, L, ldl, C, ldc, xs, g0, ng);
      __syncthreads();
    }
    return;
  }
  for (int i = tid; i < kTB * kTLD + kTB; i += kTrsvWarps * 32) Ls[i] = T(0);  // zero padding (block < 96, invd)
  __syncthreads();
  {
    const T* D = L + j0 * ldl + j0;
    T tmp[(kTB / kTrsvWarps) * 3];
#pragma unroll
    for (int a = 0; a < kTB / kTrsvWarps; a++)
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int r = warp + kTrsvWarps * a, c = lane + 32 * u;
        tmp[a * 3 + u] = (c <= r && r < jb) ? D[(int64_t)r * ldl + c] : T(0);
      }
#pragma unroll
    for (int a = 0; a < kTB / kTrsvWarps; a++)
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int r = warp + kTrsvWarps * a, c = lane + 32 * u;
        if (c <= r && r < jb) Ls[r * kTLD + c] = tmp[a * 3 + u];
      }
  }
  __syncthreads();
  if (tid < jb) invd[tid] = T(1) / Ls[tid * kTLD + tid];
  __syncthreads();
  for (int g0 = 0; g0 < nRHS; g0 += kTrsvWarps) {
    const int rhs = g0 + warp;
    if (rhs < nRHS) {
      const T* c = C + (int64_t)rhs * ldc + j0;
      T x[3];
#pragma unroll
      for (int u = 0; u < 3; u++) x[u] = (lane + 32 * u < jb) ? c[lane + 32 * u] : T(0);
      T dinv[3];
#pragma unroll
      for (int u = 0; u < 3; u++) dinv[u] = invd[lane + 32 * u];  // zero beyond jb
      if (!TR) {
#pragma unroll
        for (int uj = 0; uj < 3; uj++) {
          if (32 * uj >= jb) break;
#pragma unroll 8
          for (int jj = 0; jj < 32; jj++) {
            const int j = 32 * uj + jj;
            // the pivot entry is scaled by its own lane before the broadcast (one shuffle on the critical path)
            const T xj = __shfl_sync(0xffffffffu, x[uj] * dinv[uj], jj);
            if (lane == jj) x[uj] = xj;
#pragma unroll
            for (int u = uj; u < 3; u++) {
              const int i = lane + 32 * u;
              if (i > j) x[u] -= Ls[i * kTLD + j] * xj;
            }
          }
        }
      } else {
#pragma unroll
        for (int uj = 2; uj >= 0; uj--) {
          if (32 * uj >= jb) continue;
#pragma unroll 8
          for (int jj = 31; jj >= 0; jj--) {
            const int j = 32 * uj + jj;
            const T xj = __shfl_sync(0xffffffffu, x[uj] * dinv[uj], jj);
            if (lane == jj) x[uj] = xj;
#pragma unroll
            for (int u = 0; u <= uj; u++) {
              const int i = lane + 32 * u;
              if (i < j) x[u] -= Ls[j * kTLD + i] * xj;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int i = lane + 32 * u;
        xs[warp * kTB + i] = (i < jb) ? x[u] : T(0);
        if (blockIdx.x == 0 && i < jb) X[(int64_t)rhs * ldx + j0 + i] = x[u];
      }
    }
    __syncthreads();
    const int ng = min(kTrsvWarps, nRHS - g0);
    stepUpdate<T, TR>(n, j0, jb, L, ldl, C, ldc, xs, g0, ng);
    __syncthreads();
  }
}

// Inverse-based block step, latency optimised: every global operand of the step (the block inverse, the CTA's slice
// of the factor panel, the right-hand-side block) is requested up front and held in registers, so a step costs about
// one memory latency + two barriers instead of a chain of dependent loads.
//   forward : CTA = 64 rows below the block (warp = 8 rows, lanes over the 96 block columns)
//   backward: CTA = 128 columns before the block (thread = column, two halves of the 96 block rows)
constexpr int kInvThreads = 256, kInvWarps = 8, kInvRHS = 4;
template <typename T, bool TR>
__global__ void __launch_bounds__(kInvThreads, 1)
    solve_step_inv_kernel(int64_t n, int64_t j0, int jb, Operand<T> Lop, int64_t ldl, Operand<T> Cop, int64_t ldc,
                          Operand<T> Xop, int64_t ldx, int nRHS, Operand<T> Wop) {
  __shared__ T bs[kInvRHS][kTB];             // right-hand sides of the block
  __shared__ T part[kInvWarps][kInvRHS][kTB];  // partial products per warp / per half
  __shared__ T xs[kInvRHS][kTB];             // solved block
  const T* __restrict__ L = Lop.at(blockIdx.z);
  T* C = Cop.at(blockIdx.z);
  T* X = Xop.at(blockIdx.z);
  const T* __restrict__ Wm = Wop.at(blockIdx.z) + (j0 / kTB) * (int64_t)(2 * kTB * kTB) + (TR ? 0 : kTB * kTB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- all loads of the step, issued back to back
  constexpr int RW = kTB / kInvWarps;  // 12 rows of the inverse per warp
  T wv[RW][3];
#pragma unroll
  for (int a = 0; a < RW; a++)
#pragma unroll
    for (int u = 0; u < 3; u++) wv[a][u] = Wm[(warp * RW + a) * kTB + lane + 32 * u];
  constexpr int UF = kStepRows / kInvWarps;  // forward: 8 rows per warp
  constexpr int UB = kTB / 2;               // backward: 48 block rows per half
  T mv[TR ? UB : UF * 3];
  const int64_t rowBase = j0 + jb + (int64_t)blockIdx.x * kStepRows + warp * UF;  // forward
  const int64_t col = (int64_t)blockIdx.x * kStepCols + (tid & (kStepCols - 1));    // backward
  const int half = tid / kStepCols;
  if (!TR) {
#pragma unroll
    for (int rr = 0; rr < UF; rr++)
#pragma unroll
      for (int u = 0; u < 3; u++)
        mv[rr * 3 + u] = (rowBase + rr < n && lane + 32 * u < jb) ? L[(rowBase + rr) * ldl + j0 + lane + 32 * u] : T(0);
  } else {
#pragma unroll
    for (int r = 0; r < UB; r++) {
      const int br = half * UB + r;
      mv[r] = (col < j0 && br < jb) ? L[(j0 + br) * ldl + col] : T(0);
    }
  }

  // Programmatic dependent launch: everything above reads only the factor and the block inverses, which no kernel of
  // the solve writes, so a step launched with the PDL attribute runs its prologue while the previous step is still
  // executing; from here on it touches the right-hand sides, which the previous step updates -> wait for it, then let
  // the next step start its own prologue. (Without the launch attribute both instructions are no-ops.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  for (int g0 = 0; g0 < nRHS; g0 += kInvRHS) {
    const int ng = min(kInvRHS, nRHS - g0);
    for (int i = tid; i < kInvRHS * kTB; i += kInvThreads) {
      const int q = i / kTB, r = i % kTB;
      bs[q][r] = (q < ng && r < jb) ? C[(int64_t)(g0 + q) * ldc + j0 + r] : T(0);
    }
    __syncthreads();
    // x = W b: every warp covers 12 rows of the sum, lanes cover the outputs
#pragma unroll
    for (int q = 0; q < kInvRHS; q++) {
      T a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
      for (int a = 0; a < RW; a++) {
        const T b = bs[q][warp * RW + a];
        a0 += wv[a][0] * b, a1 += wv[a][1] * b, a2 += wv[a][2] * b;
      }
      part[warp][q][lane] = a0, part[warp][q][lane + 32] = a1, part[warp][q][lane + 64] = a2;
    }
    __syncthreads();
    for (int i = tid; i < kInvRHS * kTB; i += kInvThreads) {
      const int q = i / kTB, r = i % kTB;
      T v = 0;
#pragma unroll
      for (int w = 0; w < kInvWarps; w++) v += part[w][q][r];
      xs[q][r] = v;
      if (blockIdx.x == 0 && q < ng && r < jb) X[(int64_t)(g0 + q) * ldx + j0 + r] = v;
    }
    __syncthreads();
    // update with the solved block
    if (!TR) {
      for (int q = 0; q < ng; q++) {
        const T x0 = xs[q][lane], x1 = xs[q][lane + 32], x2 = xs[q][lane + 64];
        T v[UF];
#pragma unroll
        for (int rr = 0; rr < UF; rr++) v[rr] = mv[rr * 3] * x0 + mv[rr * 3 + 1] * x1 + mv[rr * 3 + 2] * x2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int rr = 0; rr < UF; rr++) v[rr] += __shfl_xor_sync(0xffffffffu, v[rr], o);
#pragma unroll
        for (int rr = 0; rr < UF; rr++)
          if (lane == rr && rowBase + rr < n) C[(int64_t)(g0 + q) * ldc + rowBase + rr] -= v[rr];
      }
    } else {
      T acc[kInvRHS];
#pragma unroll
      for (int q = 0; q < kInvRHS; q++) acc[q] = T(0);
#pragma unroll
      for (int r = 0; r < UB; r++)
#pragma unroll
        for (int q = 0; q < kInvRHS; q++) acc[q] += mv[r] * xs[q][half * UB + r];
      // reuse `part` for the two halves: part[half][q][column within the CTA]  (kStepCols == 128 > kTB: split rows)
      T* red = &part[0][0][0];  // 8*4*96 = 3072 entries >= 2*4*128
#pragma unroll
      for (int q = 0; q < kInvRHS; q++) red[(half * kInvRHS + q) * kStepCols + (tid & (kStepCols - 1))] = acc[q];
      __syncthreads();
      if (half == 0 && col < j0) {
#pragma unroll
        for (int q = 0; q < kInvRHS; q++)
          if (q < ng)
            C[(int64_t)(g0 + q) * ldc + col] -= red[q * kStepCols + tid] + red[(kInvRHS + q) * kStepCols + tid];
      }
    }
    __syncthreads();
  }
}

// W = L_bb^-1 for every 96 x 96 diagonal block b of a dense lower-triangular matrix (one CTA per block): thread c
// computes column c of W by forward substitution (4 accumulators), W is written twice to the scratch, row-major
// (slot 0) and transposed (slot 1), zero padded to 96 x 96 - the operands of the inverse-based block steps.
template <typename T>
__device__ __forceinline__ void invertBlockBody(const T* __restrict__ D, int64_t ldl, int jb, T* W) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* Ls = reinterpret_cast<T*>(smemRaw);  // [kTB][kTLD]
  T* Ws = Ls + kTB * kTLD;                // [kTB][kTLD]
  T* dinv = Ws + kTB * kTLD;              // [kTB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 2 * kTB * kTLD + kTB; i += 128) Ls[i] = T(0);
  __syncthreads();
  {
    T tmp[(kTB / 4) * 3];
#pragma unroll
    for (int a = 0; a < kTB / 4; a++)
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int r = warp + 4 * a, c = lane + 32 * u;
        tmp[a * 3 + u] = (c <= r && r < jb) ? D[(int64_t)r * ldl + c] : T(0);
      }
#pragma unroll
    for (int a = 0; a < kTB / 4; a++)
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const int r = warp + 4 * a, c = lane + 32 * u;
        if (c <= r && r < jb) Ls[r * kTLD + c] = tmp[a * 3 + u];
      }
  }
  __syncthreads();
  if (tid < jb) dinv[tid] = T(1) / Ls[tid * kTLD + tid];
  __syncthreads();
  if (tid < jb) {
    const int c = tid;
    Ws[c * kTLD + c] = dinv[c];
    for (int i = c + 1; i < jb; i++) {
      const T* li = Ls + i * kTLD;
      T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      int q = c;
      for (; q + 4 <= i; q += 4) {
        a0 += li[q] * Ws[q * kTLD + c];
        a1 += li[q + 1] * Ws[(q + 1) * kTLD + c];
        a2 += li[q + 2] * Ws[(q + 2) * kTLD + c];
        a3 += li[q + 3] * Ws[(q + 3) * kTLD + c];
      }
      for (; q < i; q++) a0 += li[q] * Ws[q * kTLD + c];
      Ws[i * kTLD + c] = -((a0 + a1) + (a2 + a3)) * dinv[i];
    }
  }
  __syncthreads();
  for (int i = tid; i < kTB * kTB; i += 128) {
    const int r = i / kTB, c = i % kTB;
    W[i] = Ws[r * kTLD + c];                // row-major W
    W[kTB * kTB + i] = Ws[c * kTLD + r];    // W^T
  }
}

template <typename T>
__global__ void __launch_bounds__(128) invert_blocks_kernel(int64_t n, Operand<T> Lop, int64_t ldl, Operand<T> Wop) {
  const int64_t j0 = (int64_t)blockIdx.x * kTB;
  invertBlockBody<T>(Lop.at(blockIdx.z) + j0 * ldl + j0, ldl, (int)min((int64_t)kTB, n - j0),
                     Wop.at(blockIdx.z) + (int64_t)blockIdx.x * (2 * kTB * kTB));
}

// the same for the diagonal blocks of MANY lumps in one launch (work list built once per skeleton)
template <typename T>
__global__ void __launch_bounds__(128) invert_blocks_list_kernel(const InvBlockDesc* __restrict__ list, Operand<T> Dop,
                                                                   Operand<T> Wop) {
  const InvBlockDesc d = list[blockIdx.x];
  invertBlockBody<T>(Dop.at(blockIdx.z) + d.dataOff, d.ld, d.jb, Wop.at(blockIdx.z) + d.wOff);
}

// ---- chained dense triangular solve: ONE launch per lump and direction instead of one launch per 96-column step.
// CTA i owns block row i (forward) / block column i (backward) of the triangle. It streams its off-diagonal 96 x 96
// tiles in the order their solved blocks become available (register double buffer: the next tile is in flight while
// the CTA waits), accumulates tile * x_j, and as soon as the last neighbour is published finishes
// x_i = W_i (b_i - sum) with the precomputed block inverse (staged in shared memory at kernel start) and publishes it:
// store x_i, fence, release-store of the epoch into flag[i]. Consumers acquire-poll the flag and read x_i with ld.cg
// (L1 may hold the pre-solve values of the same addresses from a CTA that ran earlier on the SM).
// The critical path per block is flag round trip + one tile product + one 96 x 96 matvec instead of a kernel boundary.
// CTAs take their block from an arrival ticket, so a CTA only ever waits for CTAs that arrived before it (no deadlock
// when the grid exceeds the resident capacity, e.g. batched solves). Flags are never reset: every launch (and every
// group of NR right-hand sides inside it) uses a fresh epoch value.
// Measured alternatives (B200, BAL-shaped 5226-wide lump, 55 blocks per direction): this flag protocol 3.4 us per block;
// self-validating {value, epoch} slots polled by every lane (NCCL-LL style, no fence) 4.3 us - twelve polled sectors per
// warp instead of one; the same with __nanosleep back-off for the CTAs that are not next in the chain 9.8 us (the sleep
// granularity makes the far CTAs fall behind and become the critical path). See profiles/README.md.
constexpr int kChThreads = 256, kChWarps = 8, kChRows = kTB / kChWarps;  // 12 tile rows per warp

__device__ __forceinline__ unsigned ldAcquireU32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stReleaseU32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T>
__device__ __forceinline__ void chainLoadTile(T (&t)[kChRows][3], const T* __restrict__ L, int64_t ldl, int64_t r0,
                                              int64_t c0, int64_t rowEnd, int64_t colEnd, int warp, int lane) {
#pragma unroll
  for (int a = 0; a < kChRows; a++) {
    const int64_t row = r0 + warp * kChRows + a;
#pragma unroll
    for (int u = 0; u < 3; u++)
      t[a][u] = (row < rowEnd && c0 + lane + 32 * u < colEnd) ? L[row * ldl + c0 + lane + 32 * u] : T(0);
  }
}

template <typename T, bool TR, int NR>
__global__ void __launch_bounds__(kChThreads, 1)
    trsv_chain_kernel(int64_t n, int nbk, Operand<T> Lop, int64_t ldl, Operand<T> Cop, int64_t ldc, int nRHS,
                      Operand<T> Wop, unsigned* flags, int flagsPerItem, unsigned* ticket, unsigned ticketBase,
                      unsigned epoch0, int64_t rowsBelow, const int64_t* __restrict__ rowMap, Operand<T> Vop) {
  // forward only, rowsBelow > 0: the lump's rows below the diagonal block (L21, contiguous after it, same ld) ride
  // along - extra CTAs of 96 rows each consume the solved blocks as they are published and finish the
  // C[rowMap[r]] -= L21[r, :] x update (the gemv + assembleVec of the reference sequence) right behind the chain
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* Wsm = reinterpret_cast<T*>(smemRaw);  // [kTB][kTB] block inverse (transposed copy for the backward solve)
  T* ys = Wsm + kTB * kTB;                 // [NR][kTB]  b_i - sum
  T* part = ys + NR * kTB;                 // backward: [kChWarps][NR][kTB] partial column sums
  __shared__ unsigned slotS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) slotS = atomicAdd(ticket, 1u) - ticketBase;
  __syncthreads();
  const unsigned slot = slotS;
  const int perItem = nbk + (TR ? 0 : (int)((rowsBelow + kTB - 1) / kTB));
  const int item = (int)(slot / (unsigned)perItem), ord = (int)(slot % (unsigned)perItem);
  const bool below = !TR && ord >= nbk;
  const int i = TR ? nbk - 1 - ord : ord;
  const T* __restrict__ L = Lop.at(item);
  T* C = Cop.at(item);
  const T* __restrict__ Wm = Wop.at(item) + (int64_t)i * (2 * kTB * kTB) + (TR ? kTB * kTB : 0);
  unsigned* flg = flags + (int64_t)item * flagsPerItem;
  const int64_t i0 = (int64_t)i * kTB;
  const int ib = (int)min((int64_t)kTB, n - i0);
  // off-diagonal tiles of this CTA; the s-th one belongs to block (TR ? nbk - 1 - s : s)
  const int cnt = below ? nbk : ord;
  const int64_t rowOrg = below ? n + (int64_t)(ord - nbk) * kTB : i0;  // forward: first tile row of this CTA
  const int64_t rowEnd = below ? n + rowsBelow : n;
  if (!below)
    for (int idx = tid; idx < kTB * kTB; idx += kChThreads) Wsm[idx] = Wm[idx];

  int grp = 0;
  for (int g0 = 0; g0 < nRHS; g0 += NR, grp++) {
    const unsigned epoch = epoch0 + (unsigned)grp;
    const int ng = min(NR, nRHS - g0);
    T acc[NR];     // forward: lane a < 12 of warp w holds the sum of row 12 w + a
    T pc[NR][3];   // backward: partial sums of columns lane + 32 u over the warp's rows
#pragma unroll
    for (int q = 0; q < NR; q++) acc[q] = pc[q][0] = pc[q][1] = pc[q][2] = T(0);
    // this block's right-hand sides, requested now: off the critical path between the last neighbour and x_i
    T bpre[TR ? (NR * kTB + kChThreads - 1) / kChThreads : NR];
    int64_t belowDst = -1;  // below CTA: vector row this lane's tile row updates
    if (!TR) {
      const int row = warp * kChRows + lane;
      if (below) {
        if (lane < kChRows && rowOrg + row < rowEnd) belowDst = rowMap[rowOrg + row - n];
#pragma unroll
        for (int q = 0; q < NR; q++)
          bpre[q] = (belowDst >= 0 && q < ng) ? Vop.at(item)[(int64_t)(g0 + q) * ldc + belowDst] : T(0);
      } else {
#pragma unroll
        for (int q = 0; q < NR; q++)
          bpre[q] = (lane < kChRows && q < ng && row < ib) ? C[(int64_t)(g0 + q) * ldc + i0 + row] : T(0);
      }
    } else {
#pragma unroll
      for (int e = 0; e < (NR * kTB + kChThreads - 1) / kChThreads; e++) {
        const int idx = tid + e * kChThreads, q = idx / kTB, c = idx % kTB;
        bpre[e] = (idx < NR * kTB && q < ng && c < ib) ? C[(int64_t)(g0 + q) * ldc + i0 + c] : T(0);
      }
    }

    auto consume = [&](const T(&t)[kChRows][3], int s) {
      const int j = TR ? nbk - 1 - s : s;
      const int64_t j0 = (int64_t)j * kTB;
      // ">= epoch" in wrap-around arithmetic: with several right-hand-side groups the producer may already be groups ahead
      while ((int)(ldAcquireU32(flg + j) - epoch) < 0) {
      }
      if (!TR) {
        T x[NR][3];
#pragma unroll
        for (int q = 0; q < NR; q++)
#pragma unroll
          for (int u = 0; u < 3; u++)
            x[q][u] = (q < ng && j0 + lane + 32 * u < n) ? __ldcg(&C[(int64_t)(g0 + q) * ldc + j0 + lane + 32 * u]) : T(0);
#pragma unroll
        for (int q = 0; q < NR; q++) {
          T sr[kChRows];
#pragma unroll
          for (int a = 0; a < kChRows; a++) sr[a] = t[a][0] * x[q][0] + t[a][1] * x[q][1] + t[a][2] * x[q][2];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int a = 0; a < kChRows; a++) sr[a] += __shfl_xor_sync(0xffffffffu, sr[a], o);
#pragma unroll
          for (int a = 0; a < kChRows; a++)
            if (lane == a) acc[q] += sr[a];
        }
      } else {
#pragma unroll
        for (int q = 0; q < NR; q++) {
          T xv[kChRows];
#pragma unroll
          for (int a = 0; a < kChRows; a++) {
            const int64_t row = j0 + warp * kChRows + a;
            xv[a] = (q < ng && row < n) ? __ldcg(&C[(int64_t)(g0 + q) * ldc + row]) : T(0);
          }
#pragma unroll
          for (int a = 0; a < kChRows; a++)
#pragma unroll
            for (int u = 0; u < 3; u++) pc[q][u] += t[a][u] * xv[a];
        }
      }
    };
    auto load = [&](T(&t)[kChRows][3], int s) {
      const int j = TR ? nbk - 1 - s : s;
      if (!TR)
        chainLoadTile<T>(t, L, ldl, rowOrg, (int64_t)j * kTB, rowEnd, n, warp, lane);
      else
        chainLoadTile<T>(t, L, ldl, (int64_t)j * kTB, i0, n, n, warp, lane);
    };

    {
      T tA[kChRows][3], tB[kChRows][3];
      if (cnt > 0) load(tA, 0);
      for (int s = 0; s < cnt; s += 2) {
        if (s + 1 < cnt) load(tB, s + 1);
        consume(tA, s);
        if (s + 1 < cnt) {
          if (s + 2 < cnt) load(tA, s + 2);
          consume(tB, s + 1);
        }
      }
    }

    if (below) {  // uniform per CTA: no block to solve, only the update of the rows below
      if (belowDst >= 0) {
#pragma unroll
        for (int q = 0; q < NR; q++)
          if (q < ng) Vop.at(item)[(int64_t)(g0 + q) * ldc + belowDst] = bpre[q] - acc[q];
      }
      continue;
    }
    // y = b_i - sum
    if (!TR) {
      if (lane < kChRows) {
        const int row = warp * kChRows + lane;
#pragma unroll
        for (int q = 0; q < NR; q++)
          ys[q * kTB + row] = (q < ng && row < ib) ? bpre[q] - acc[q] : T(0);
      }
    } else {
#pragma unroll
      for (int q = 0; q < NR; q++)
#pragma unroll
        for (int u = 0; u < 3; u++) part[(warp * NR + q) * kTB + lane + 32 * u] = pc[q][u];
      __syncthreads();
#pragma unroll
      for (int e = 0; e < (NR * kTB + kChThreads - 1) / kChThreads; e++) {
        const int idx = tid + e * kChThreads, q = idx / kTB, c = idx % kTB;
        if (idx < NR * kTB) {
          T v = T(0);
#pragma unroll
          for (int w = 0; w < kChWarps; w++) v += part[(w * NR + q) * kTB + c];
          ys[idx] = (q < ng && c < ib) ? bpre[e] - v : T(0);
        }
      }
    }
    __syncthreads();  // ys (and, first time round, Wsm) complete
    // x_i = W y  (backward: Wsm holds W^T, so the same row-times-vector product gives W^T y)
#pragma unroll
    for (int q = 0; q < NR; q++) {
      if (q < ng) {
        const T y0 = ys[q * kTB + lane], y1 = ys[q * kTB + lane + 32], y2 = ys[q * kTB + lane + 64];
        T sr[kChRows];
#pragma unroll
        for (int a = 0; a < kChRows; a++) {
          const T* wr = Wsm + (warp * kChRows + a) * kTB;
          sr[a] = wr[lane] * y0 + wr[lane + 32] * y1 + wr[lane + 64] * y2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int a = 0; a < kChRows; a++) sr[a] += __shfl_xor_sync(0xffffffffu, sr[a], o);
        T xo = T(0);
#pragma unroll
        for (int a = 0; a < kChRows; a++)
          if (lane == a) xo = sr[a];
        const int row = warp * kChRows + lane;
        if (lane < kChRows && row < ib) C[(int64_t)(g0 + q) * ldc + i0 + row] = xo;
      }
    }
    // no fence.sc before the barrier: the barrier orders the CTA's stores before thread 0's release store, which is
    // cumulative at gpu scope (the explicit fence cost a second memory barrier on every hop of the chain)
    __syncthreads();  // (also: ys / part are reused by the next group of right-hand sides)
    if (tid == 0) stReleaseU32(flg + i, epoch);
  }
}

template <typename T>
__global__ void copy_vec_kernel(int64_t n, int nRHS, Operand<T> Xop, int64_t ldx, Operand<T> Cop, int64_t ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nRHS) return;
  const int64_t r = i % n, c = i / n;
  Cop.at(blockIdx.z)[c * ldc + r] = Xop.at(blockIdx.z)[c * ldx + r];
}

// launch of one inverse-based block step; pdl: programmatic dependent launch on the previous kernel of the stream
// (only for a step whose predecessor in the stream is the previous step of the same solve)
template <typename T, bool TR>
static void launchStepInv(cudaStream_t st, dim3 grid, bool pdl, int64_t n, int64_t j0, int jb, Operand<T> L, int64_t ldl,
                          Operand<T> C, int64_t ldc, Operand<T> X, int64_t ldx, int nRHS, Operand<T> W) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(kInvThreads, 1, 1), cfg.dynamicSmemBytes = 0, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  B200_CUDA(cudaLaunchKernelEx(&cfg, solve_step_inv_kernel<T, TR>, n, j0, jb, L, ldl, C, ldc, X, ldx, nRHS, W));
}

template <typename T>
void invertBlockList(cudaStream_t st, int batch, const InvBlockDesc* list, int64_t count, Operand<T> data,
                     Operand<T> invScratch) {
  if (count <= 0) return;
  const size_t ismem = ((size_t)2 * kTB * kTLD + kTB) * sizeof(T);
  ensureDynSmem((const void*)invert_blocks_list_kernel<T>, ismem);
  ProfScope prof(st, KC_SOLVE_DENSE, 0, (double)count * kTB * kTB / 2 * sizeof(T) * batch);
  invert_blocks_list_kernel<T><<<dim3((unsigned)count, 1, batch), 128, ismem, st>>>(list, data, invScratch);
  B200_LAUNCH_CHECK();
}

template <typename T, bool TR, int NR>
static void launchChain(cudaStream_t st, int batch, int64_t n, Operand<T> L, int64_t ldl, Operand<T> C, int64_t ldc,
                        int nRHS, Operand<T> W, ChainSync* cs, int64_t rowsBelow, const int64_t* rowMap,
                        Operand<T> vec) {
  const int nbk = ceilDiv(n, kTB);
  const int perItem = nbk + (TR ? 0 : ceilDiv(rowsBelow, kTB));
  const size_t smem = ((size_t)kTB * kTB + (size_t)NR * kTB + (TR ? (size_t)kChWarps * NR * kTB : 0)) * sizeof(T);
  ensureDynSmem((const void*)trsv_chain_kernel<T, TR, NR>, smem);
  trsv_chain_kernel<T, TR, NR><<<dim3((unsigned)perItem * batch, 1, 1), kChThreads, smem, st>>>(
      n, nbk, L, ldl, C, ldc, nRHS, W, cs->flags, cs->flagsPerItem, cs->ticket, cs->ticketBase, cs->epoch + 1,
      TR ? 0 : rowsBelow, rowMap, vec);
  B200_LAUNCH_CHECK();
  cs->ticketBase += (unsigned)perItem * batch;
  cs->epoch += (unsigned)ceilDiv(nRHS, NR);
}

template <typename T>
bool trsvAny(cudaStream_t st, int batch, int64_t n, Operand<T> L, int64_t ldl, Operand<T> C, int64_t ldc, int nRHS,
             bool transposed, Operand<T> scratch, Operand<T> invScratch, bool inversesReady, ChainSync* chain,
             int64_t rowsBelow, const int64_t* rowMap, Operand<T> vec) {
  if (n <= 0 || nRHS <= 0) return false;
  const int nb = kTB;
  // BSPB200_PDL=0 disables the programmatic dependent launches; profiling (events between the steps) does too
  static const bool pdlEnv = !(getenv("BSPB200_PDL") && atoi(getenv("BSPB200_PDL")) == 0);
  const bool pdl = pdlEnv && !profileEnabled();
  if (n <= nb) {  // a single block: solved in place
    trsvBlock<T>(st, batch, (int)n, L, ldl, C, ldc, nRHS, transposed);
    return false;
  }
  const size_t smem = ((size_t)kTB * kTLD + kTB + (size_t)kTrsvWarps * kTB) * sizeof(T);
  ensureDynSmem((const void*)solve_step_kernel<T, false>, smem);
  ensureDynSmem((const void*)solve_step_kernel<T, true>, smem);
  const int64_t ldx = n;
  // inverse-based diagonal steps when the caller provided room for the block inverses (2 x 96 x 96 per block)
  const bool useInv = invScratch.base != nullptr;
  if (useInv && !inversesReady) {
    const size_t ismem = ((size_t)2 * kTB * kTLD + kTB) * sizeof(T);
    ensureDynSmem((const void*)invert_blocks_kernel<T>, ismem);
    ProfScope prof(st, KC_SOLVE_DENSE, 0, (double)n * kTB / 2 * sizeof(T) * batch);
    invert_blocks_kernel<T><<<dim3(ceilDiv(n, nb), 1, batch), 128, ismem, st>>>(n, L, ldl, invScratch);
    B200_LAUNCH_CHECK();
  }
  if (useInv && chain && chain->flags && ceilDiv(n, nb) <= chain->flagsPerItem) {
    // the rows below ride along in the forward chain when the caller passed their row table
    const bool fuseBelow = !transposed && rowMap != nullptr && rowsBelow > 0;
    const int64_t rb = fuseBelow ? rowsBelow : 0;
    ProfScope prof(st, KC_SOLVE_DENSE, ((double)n * n + 2.0 * rb * n) * nRHS * batch,
                   ((double)n * (n + 1) / 2 + (double)rb * n) * sizeof(T) * batch);
    if (nRHS == 1) {
      if (transposed) launchChain<T, true, 1>(st, batch, n, L, ldl, C, ldc, nRHS, invScratch, chain, 0, nullptr, vec);
      else launchChain<T, false, 1>(st, batch, n, L, ldl, C, ldc, nRHS, invScratch, chain, rb, rowMap, vec);
    } else {
      if (transposed) launchChain<T, true, 4>(st, batch, n, L, ldl, C, ldc, nRHS, invScratch, chain, 0, nullptr, vec);
      else launchChain<T, false, 4>(st, batch, n, L, ldl, C, ldc, nRHS, invScratch, chain, rb, rowMap, vec);
    }
    return fuseBelow;
  }
  if (!transposed) {
    for (int64_t j0 = 0; j0 < n; j0 += nb) {
      int64_t jb = std::min<int64_t>(nb, n - j0), rb = n - j0 - jb;
      ProfScope prof(st, KC_SOLVE_DENSE, (double)(jb * jb + 2.0 * rb * jb) * nRHS * batch,
                     (double)(jb * (jb + 1) / 2 + rb * jb) * sizeof(T) * batch);
      if (useInv)
        launchStepInv<T, false>(st, dim3(std::max(1, ceilDiv(rb, kStepRows)), 1, batch), pdl && j0 > 0, n, j0, (int)jb, L,
                                ldl, C, ldc, scratch, ldx, nRHS, invScratch);
      else
        solve_step_kernel<T, false><<<dim3(std::max(1, ceilDiv(rb, kStepRows)), 1, batch), kTrsvWarps * 32, smem, st>>>(
            n, j0, (int)jb, L, ldl, C, ldc, scratch, ldx, nRHS, invScratch, false);
      B200_LAUNCH_CHECK();
    }
  } else {
    for (int64_t j0 = ((n - 1) / nb) * nb; j0 >= 0; j0 -= nb) {
      int64_t jb = std::min<int64_t>(nb, n - j0);
      ProfScope prof(st, KC_SOLVE_DENSE, (double)(jb * jb + 2.0 * j0 * jb) * nRHS * batch,
                     (double)(jb * (jb + 1) / 2 + j0 * jb) * sizeof(T) * batch);
      if (useInv)
        launchStepInv<T, true>(st, dim3(std::max(1, ceilDiv(j0, kStepCols)), 1, batch), pdl && j0 + nb < n, n, j0, (int)jb,
                               L, ldl, C, ldc, scratch, ldx, nRHS, invScratch);
      else
        solve_step_kernel<T, true><<<dim3(std::max(1, ceilDiv(j0, kStepCols)), 1, batch), kTrsvWarps * 32, smem, st>>>(
            n, j0, (int)jb, L, ldl, C, ldc, scratch, ldx, nRHS, invScratch, false);
      B200_LAUNCH_CHECK();
    }
  }
  copy_vec_kernel<T><<<dim3(ceilDiv(n * nRHS, 256), 1, batch), 256, 0, st>>>(n, nRHS, scratch, ldx, C, ldc);
  B200_LAUNCH_CHECK();
  return false;
}


// v[i + rhs * ldc] += sum over the lanes (in lane order) of delta_lane[i + rhs * ldc]; the deltas are zeroed for the next
// solve. `delta` points at lane 0's entry of row 0 of the lump; laneStride = distance between the lanes' vectors.
template <typename T>
__global__ void __launch_bounds__(128) gather_lane_deltas_kernel(int64_t n, int nRHS, int64_t ldc, Operand<T> delta,
                                                                 int64_t laneStride, int nLanes, Operand<T> v) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * nRHS) return;
  const int64_t idx = e % n + (e / n) * ldc;
  T* d = delta.at(blockIdx.z) + idx;
  T sum = 0;
  for (int k = 0; k < nLanes; k++) {
    sum += d[(int64_t)k * laneStride];
    d[(int64_t)k * laneStride] = T(0);
  }
  v.at(blockIdx.z)[idx] += sum;
}
template <typename T>
void gatherLaneDeltas(cudaStream_t st, int batch, int64_t n, int nRHS, int64_t ldc, Operand<T> delta, int64_t laneStride,
                      int nLanes, Operand<T> v) {
  if (n <= 0) return;
  gather_lane_deltas_kernel<T><<<dim3((unsigned)ceilDiv(n * nRHS, 128), 1, batch), 128, 0, st>>>(n, nRHS, ldc, delta, laneStride,
                                                                                                  nLanes, v);
  B200_LAUNCH_CHECK();
}

#define B200_INSTANTIATE_SOLVE(T)                                                                                       \
  template void gemvRows<T>(cudaStream_t, int, int64_t, int64_t, T, Operand<T>, int64_t, Operand<T>, int64_t,          \
                            Operand<T>, int64_t, int64_t, int, bool, const int64_t*);                                   \
  template void gemvColsT<T>(cudaStream_t, int, int64_t, int64_t, T, Operand<T>, int64_t, Operand<T>, int64_t, int64_t, \
                             Operand<T>, int64_t, int, Operand<T>, int64_t, const int64_t*);                            \
  template void symmLower<T>(cudaStream_t, int, int64_t, T, Operand<T>, Operand<T>, int64_t, Operand<T>, int64_t, int); \
  template bool trsvAny<T>(cudaStream_t, int, int64_t, Operand<T>, int64_t, Operand<T>, int64_t, int, bool, Operand<T>, \
                           Operand<T>, bool, ChainSync*, int64_t, const int64_t*, Operand<T>);                          \
  template void invertBlockList<T>(cudaStream_t, int, const InvBlockDesc*, int64_t, Operand<T>, Operand<T>);      \
  template void gatherLaneDeltas<T>(cudaStream_t, int, int64_t, int, int64_t, Operand<T>, int64_t, int, Operand<T>);
B200_INSTANTIATE_SOLVE(double)
B200_INSTANTIATE_SOLVE(float)

}  // namespace b200
}  // namespace BaSpaCho
