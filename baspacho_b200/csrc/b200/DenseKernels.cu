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
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include "B200Kernels.h"
#include "B200Wave.h"

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

// entry of a kernel of the programmatic-dependent-launch chain (see launchChain)
#define B200_CHAIN_ENTRY()                                           \
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    \
  asm volatile("griddepcontrol.wait;" ::: "memory");

struct GemmShape {
  int64_t m, n, k, lda, ldb, ldc;
  int lowerOnly;
};

// ------------------------------------------------------------------------------------------------ fp64 DMMA GEMM
// CTA tile BM x BN x BK, warp tile WM x WN built from m8n8k4 DMMA tiles; A and B tiles are both K-contiguous
// ([row][k] with a 4-double pad: the quad-strided fragment loads are bank-conflict free).
template <int BM, int BN, int BK, int WM, int WN, int STAGES, int VEC, int MINB = 1>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
    gemm_nt_f64_kernel(GemmShape s, double alpha, Operand<double> Aop, Operand<double> Bop, double beta,
                       Operand<double> Cop) {
  constexpr int NWN = BN / WN;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LDS = BK + 4;
  constexpr int TM = WM / 8, TN = WN / 8;
  extern __shared__ __align__(16) double smemD[];
  B200_CHAIN_ENTRY()
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

  if (beta != 0.0) {  // pull the C tile towards L2 while the main loop runs (one 128-byte line per request)
    constexpr int LPR = BN * 8 / 128;  // lines per tile row
    for (int i = tid; i < BM * LPR; i += NT) {
      const int64_t row = m0 + i / LPR, col = n0 + (i % LPR) * 16;
      if (row < s.m && col < s.n && !(s.lowerOnly && col > row))
        asm volatile("prefetch.global.L2 [%0];" ::"l"(C + row * s.ldc + col));
    }
  }

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

  // epilogue: each thread owns 2 consecutive columns of every 8x8 tile. With beta != 0 the old C values of tile row
  // i + 1 are loaded before tile row i is stored (software pipelined: the read-modify-write never waits on a
  // load it has just issued); the C tile was prefetched into L2 at kernel start.
  auto valid = [&](int i, int j, int e, int64_t& off) {
    const int64_t row = m0 + wm + i * 8 + g, c = n0 + wn + j * 8 + 2 * t + e;
    off = row * s.ldc + c;
    return row < s.m && c < s.n && !(s.lowerOnly && c > row);
  };
  double cold[2][TN][2];
  auto loadRow = [&](int i, double (*dst)[2]) {
#pragma unroll
    for (int j = 0; j < TN; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int64_t off;
        dst[j][e] = valid(i, j, e, off) ? C[off] : 0.0;
      }
  };
  if (beta != 0.0) loadRow(0, cold[0]);
#pragma unroll
  for (int i = 0; i < TM; i++) {
    if (beta != 0.0 && i + 1 < TM) loadRow(i + 1, cold[(i + 1) & 1]);
#pragma unroll
    for (int j = 0; j < TN; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int64_t off;
        if (!valid(i, j, e, off)) continue;
        double v = alpha * acc[i][j][e];
        if (beta != 0.0) v += beta * cold[i & 1][j][e];
        C[off] = v;
      }
  }
}

// ------------------------------------------------------------------------------------------------ fp32 on tensor cores
// C = alpha A B^T + beta C for float with the 3xTF32 split: every operand x is cut into hi = tf32(x) and lo = tf32(x - hi)
// and the product is accumulated as hi*hi + hi*lo + lo*hi in fp32 (mma.sync.m16n8k8.tf32, HMMA.1688.F32.TF32 in SASS):
// ~2^-21 relative error per product instead of TF32's 2^-11, i.e. fp32-class results at 1/3 of the TF32 rate - the
// tolerances the reference's fp32 tests use (1e-5 / 5e-5, CudaFactorTest.cpp:33-42) hold. Replaces the SIMT kernel
// below for every fp32 GEMM of the factorization (reference instantiations Solver.cpp:458-536, cublasSgemm :568-590).
// Tile BM x BN x BK, warp tile WM x WN from m16n8k8 tiles, cp.async pipeline, [row][k] smem with a 4-float pad
// (conflict-free fragment loads: row stride 20 words).
__device__ __forceinline__ uint32_t tf32Of(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma1688(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int VEC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, 2)
    gemm_nt_tf32x3_kernel(GemmShape s, float alpha, Operand<float> Aop, Operand<float> Bop, float beta,
                          Operand<float> Cop) {
  constexpr int NWN = BN / WN;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int LDS = BK + 4;
  constexpr int TM = WM / 16, TN = WN / 8;
  extern __shared__ __align__(16) float smemF[];
  float* As = smemF;
  float* Bs = smemF + STAGES * BM * LDS;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  if (s.lowerOnly && n0 > m0 + BM - 1) return;  // tile strictly above the diagonal
  const int b = blockIdx.z;
  const float* __restrict__ A = Aop.at(b);
  const float* __restrict__ B = Bop.at(b);
  float* __restrict__ C = Cop.at(b);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp / NWN) * WM, wn = (warp % NWN) * WN;
  const int g = lane >> 2, t = lane & 3;

  float acc[TM][TN][4];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[i][j][e] = 0.f;

  const int KT = (int)((s.k + BK - 1) / BK);
  auto loadTile = [&](int stage, int kt) {
    const int64_t k0 = (int64_t)kt * BK;
    constexpr int CPR = BK / VEC;
#pragma unroll
    for (int i = tid; i < BM * CPR; i += NT) {
      const int r = i / CPR, kc = (i % CPR) * VEC;
      const int64_t gr = m0 + r, gk = k0 + kc;
      const bool ok = gr < s.m && gk < s.k;
      const float* src = ok ? A + gr * s.lda + gk : A;
      const int bytes = ok ? (int)min((int64_t)VEC, s.k - gk) * 4 : 0;
      if constexpr (VEC == 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smemAddr(&As[(stage * BM + r) * LDS + kc])), "l"(src), "r"(bytes));
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smemAddr(&As[(stage * BM + r) * LDS + kc])), "l"(src), "r"(bytes));
    }
#pragma unroll
    for (int i = tid; i < BN * CPR; i += NT) {
      const int r = i / CPR, kc = (i % CPR) * VEC;
      const int64_t gr = n0 + r, gk = k0 + kc;
      const bool ok = gr < s.n && gk < s.k;
      const float* src = ok ? B + gr * s.ldb + gk : B;
      const int bytes = ok ? (int)min((int64_t)VEC, s.k - gk) * 4 : 0;
      if constexpr (VEC == 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smemAddr(&Bs[(stage * BN + r) * LDS + kc])), "l"(src), "r"(bytes));
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smemAddr(&Bs[(stage * BN + r) * LDS + kc])), "l"(src), "r"(bytes));
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
      const int nk = kt + STAGES - 1;
      if (nk < KT) loadTile(nk % STAGES, nk);
      cpAsyncCommit();
    }
    const int st = kt % STAGES;
    const float* as = As + (st * BM + wm + g) * LDS + t;
    const float* bs = Bs + (st * BN + wn + g) * LDS + t;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      uint32_t ah[TM][4], al[TM][4], bh[TN][2], bl[TN][2];
#pragma unroll
      for (int i = 0; i < TM; i++) {
        const float v[4] = {as[(i * 16) * LDS + kk], as[(i * 16 + 8) * LDS + kk], as[(i * 16) * LDS + kk + 4],
                            as[(i * 16 + 8) * LDS + kk + 4]};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          ah[i][e] = tf32Of(v[e]);
          al[i][e] = tf32Of(v[e] - __uint_as_float(ah[i][e]));
        }
      }
#pragma unroll
      for (int j = 0; j < TN; j++) {
        const float v[2] = {bs[(j * 8) * LDS + kk], bs[(j * 8) * LDS + kk + 4]};
#pragma unroll
        for (int e = 0; e < 2; e++) {
          bh[j][e] = tf32Of(v[e]);
          bl[j][e] = tf32Of(v[e] - __uint_as_float(bh[j][e]));
        }
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) {
          mma1688(acc[i][j], al[i], bh[j]);  // small terms first
          mma1688(acc[i][j], ah[i], bl[j]);
          mma1688(acc[i][j], ah[i], bh[j]);
        }
    }
  }
  cpAsyncWait<0>();
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int64_t row = m0 + wm + i * 16 + g + (e >> 1) * 8, col = n0 + wn + j * 8 + 2 * t + (e & 1);
        if (row >= s.m || col >= s.n || (s.lowerOnly && col > row)) continue;
        float* dst = C + row * s.ldc + col;
        float v = alpha * acc[i][j][e];
        if (beta != 0.f) v += beta * *dst;
        *dst = v;
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

// ------------------------------------------------------------------------------------------------ panel kernel
// Diagonal block (n <= kNB) + the rows below it, in ONE launch:
//   DO_POTRF: every CTA loads the n x n diagonal block and factors it (redundantly: the other SMs would idle anyway,
//   and it removes a dependent launch + a reload of L); CTA 0 writes the factor back.
//   Then CTA b solves X L^T = B for rows [64 b, 64 b + 64): 4 adjacent lanes share one row, each owns 8 columns of
//   the current 32-column chunk in registers.
// Cholesky phase: register resident, right looking. Thread (warp w, lane l) owns the entries (i, c) with
//   i = l + 32 a (a < 3), c = w + 8 u (u < 12). Per column its owner warp publishes it (unscaled) to a double-buffered
//   shared column, ONE barrier, then every thread updates its own registers with 1/pivot folded into the row operand.
//   The column loop is unrolled over u so that every register index is static.
// Solve phase: L is kept TRANSPOSED in shared memory, Lt[q][c] = L[c][q], each 8-column piece padded to 10 entries:
//   the 4 lanes of a row read 4 x 64 contiguous-ish bytes per q with 16-byte loads and no bank conflict.
constexpr int kNB = 96;            // max diagonal block
constexpr int kLDS = 100;          // row stride of the row-major staging of the diagonal block
constexpr int kLDT = 120;          // row stride of Lt: 12 pieces x (8 + 2 pad)
constexpr int kLDX = 98;           // row stride of the row slab (even: 16-byte aligned rows, conflict-free 8-row walks)
constexpr int kPanelThreads = 256;
constexpr int kPanelG = 4;         // lanes per row in the solve phase
constexpr int kPanelRows = kPanelThreads / kPanelG;

__device__ __forceinline__ int ltPos(int c) { return (c >> 3) * 10 + (c & 7); }  // column -> padded position in an Lt row

template <typename T>
struct Vec2 {
  T v[2];
};
__device__ __forceinline__ Vec2<double> load2(const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  return {{a.x, a.y}};
}
__device__ __forceinline__ Vec2<float> load2(const float* p) {
  const float2 a = *reinterpret_cast<const float2*>(p);
  return {{a.x, a.y}};
}

template <typename T, bool DO_POTRF>
__global__ void __launch_bounds__(kPanelThreads, 1) panel_kernel(int n, int64_t rows, Operand<T> Lop, int64_t ldl,
                                                             Operand<T> Bop, int64_t ldb, const WavePanel* work,
                                                             int* counters, int lumpsInLaunch, long long* clk) {
  // optional phase time stamps of CTA 0 (diagnostics: BSPB200_PANEL_CLK=1, read with bspb200_debug_read(0, ...))
#define B200_PCLK(i) \
  if (clk && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.z == 0) clk[i] = clock64();
  B200_PCLK(0)
  constexpr int NW = kPanelThreads / 32;  // 8 warps
  constexpr int G = kPanelG, R = kPanelRows, W = 32 / G;
  constexpr int LA = kNB / NW, LU = kNB / 32;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* Lt = reinterpret_cast<T*>(smemRaw);  // [kNB][kLDT] transposed factor (first used as row-major staging [kNB][kLDS])
  T* invd = Lt + kNB * kLDT;              // [kNB]
  T* Xs = invd + kNB;                     // [R][kLDX]  (also the 2 x kNB column buffer of the Cholesky phase)
  T* __restrict__ L = Lop.at(blockIdx.z);
  T* __restrict__ B = Bop.at(blockIdx.z);
  int slab = blockIdx.x, lumpIdx = 0;
  if (work) {  // batched over a work list (wavefront): one item = (lump column, 64-row slab)
    const WavePanel w = work[blockIdx.x];
    n = w.n, rows = w.rows, slab = w.slab, lumpIdx = w.lumpIdx;
    L += w.dataOff;
    B = L + (int64_t)n * n;
    ldl = ldb = n;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (DO_POTRF) {
    constexpr int RA = kNB / 32, CU = kNB / NW;  // 3 row slots, 12 column slots per thread
    T* S = Lt;
    for (int i = tid; i < kNB * kLDS; i += kPanelThreads) S[i] = T(0);
    __syncthreads();
    B200_PCLK(1)
    {  // lower triangle -> smem (coalesced), all loads in flight at once
      T tmp[LA * LU];
#pragma unroll
      for (int a = 0; a < LA; a++)
#pragma unroll
        for (int u = 0; u < LU; u++) {
          const int r = warp + NW * a, c = lane + 32 * u;
          tmp[a * LU + u] = (c <= r && r < n) ? L[(int64_t)r * ldl + c] : T(0);
        }
#pragma unroll
      for (int a = 0; a < LA; a++)
#pragma unroll
        for (int u = 0; u < LU; u++) {
          const int r = warp + NW * a, c = lane + 32 * u;
          if (c <= r && r < n) S[r * kLDS + c] = tmp[a * LU + u];
        }
    }
    __syncthreads();
    B200_PCLK(2)
    // Every CTA of a lump column factors the diagonal block from the ORIGINAL values, so the factor may only be written
    // back once every CTA of that column has loaded the block: the CTAs count themselves in after the load, and the
    // one that arrives last (all others provably hold their copy) is the writer. No CTA ever waits on another, so the
    // protocol is independent of how many CTAs are co-resident; the writer resets the counter for the next launch.
    __shared__ int writerFlag;
    if (tid == 0) {
      const int slabs = max(1, (int)((rows + R - 1) / R));
      int* ctr = counters + (int64_t)blockIdx.z * lumpsInLaunch + lumpIdx;
      __threadfence();
      const int old = atomicAdd(ctr, 1);
      writerFlag = (old == slabs - 1);
      if (writerFlag) *ctr = 0;
    }
    __syncthreads();
    const bool writer = writerFlag != 0;
    B200_PCLK(3)
    T* colbuf = Xs;                // [4][kNB]  raw (not yet scaled) columns of the current 4-column group
    T* ybuf = Xs + 4 * kNB;        // [kNB][4]  finished rows of the group: L[i][j0 .. j0+3]
    T reg[RA][CU];
#pragma unroll
    for (int a = 0; a < RA; a++)
#pragma unroll
      for (int u = 0; u < CU; u++) reg[a][u] = S[(lane + 32 * a) * kLDS + warp + NW * u];
    // Four columns per step (two barriers per step): the owners publish the raw columns, every thread factors the
    // 4x4 pivot block redundantly in registers and solves its own rows against it, three warps publish the finished
    // rows, then the rank-4 update of the trailing register tiles runs with all operands loaded up front.
#pragma unroll
    for (int u = 0; u < CU; u++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int j0 = NW * u + 4 * h;
        if (j0 < n) {
          const int q = warp - 4 * h;  // owner warps of the group: q in [0, 4)
          if (q >= 0 && q < 4) {
#pragma unroll
            for (int a = 0; a < RA; a++)
              if (lane + 32 * a >= j0) colbuf[q * kNB + lane + 32 * a] = reg[a][u];
          }
          __syncthreads();
          // pivot block d[r][c] = entry (j0 + r, j0 + c), r >= c; columns beyond n act as identity
          T d[4][4], raw[RA][4];
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = c; r < 4; r++) d[r][c] = colbuf[c * kNB + j0 + r];
#pragma unroll
          for (int a = 0; a < RA; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) raw[a][c] = colbuf[c * kNB + lane + 32 * a];
#pragma unroll
          for (int c = 0; c < 4; c++)
            if (j0 + c >= n) d[c][c] = T(1);
          T rs[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int k = 0; k < c; k++) d[c][c] -= d[c][k] * d[c][k];
            rs[c] = rsqrt(d[c][c]);
#pragma unroll
            for (int r = c + 1; r < 4; r++) {
#pragma unroll
              for (int k = 0; k < c; k++) d[r][c] -= d[r][k] * d[c][k];
              d[r][c] *= rs[c];
            }
          }
          // own rows: y = raw * D^-T (entries above the diagonal of the pivot block and finished rows -> 0)
          T y[RA][4];
#pragma unroll
          for (int a = 0; a < RA; a++) {
            const int t = lane + 32 * a - j0;  // row index relative to the group
#pragma unroll
            for (int c = 0; c < 4; c++) {
              T v = raw[a][c];
#pragma unroll
              for (int k = 0; k < c; k++) v -= y[a][k] * d[c][k];
              y[a][c] = (t >= c) ? v * rs[c] : T(0);
            }
          }
#pragma unroll
          for (int a = 0; a < RA; a++)
            if (warp == a) {
#pragma unroll
              for (int c = 0; c < 4; c++) ybuf[(lane + 32 * a) * 4 + c] = y[a][c];
            }
          __syncthreads();
          // rank-4 update of the columns after the group
          T yc[CU][4];
#pragma unroll
          for (int u2 = u; u2 < CU; u2++) {
            const Vec2<T> lo = load2(ybuf + (warp + NW * u2) * 4), hi = load2(ybuf + (warp + NW * u2) * 4 + 2);
            yc[u2][0] = lo.v[0], yc[u2][1] = lo.v[1], yc[u2][2] = hi.v[0], yc[u2][3] = hi.v[1];
          }
          if (!(h == 0 && warp >= 4)) {  // this slot's column is inside (or before) the group: no update
#pragma unroll
            for (int c = 0; c < 4; c++) yc[u][c] = T(0);
          }
#pragma unroll
          for (int a = 0; a < RA; a++) {
            if (j0 + 3 < 32 * a + 31) {  // some row of the slot is below the group (warp uniform)
#pragma unroll
              for (int u2 = u; u2 < CU; u2++)
                if (NW * u2 <= 32 * a + 31) {  // static: the tile touches the lower triangle
#pragma unroll
                  for (int c = 0; c < 4; c++) reg[a][u2] -= y[a][c] * yc[u2][c];
                }
            }
          }
          if (q >= 0 && q < 4) {
#pragma unroll
            for (int a = 0; a < RA; a++)
              if (lane + 32 * a >= j0) reg[a][u] = q == 0 ? y[a][0] : (q == 1 ? y[a][1] : (q == 2 ? y[a][2] : y[a][3]));
          }
        }
      }
    }
    __syncthreads();  // column buffers and the row-major staging are dead
    B200_PCLK(4)
    for (int i = tid; i < kNB * kLDT + kNB; i += kPanelThreads) Lt[i] = T(0);
    __syncthreads();
    B200_PCLK(5)
#pragma unroll
    for (int a = 0; a < RA; a++)
#pragma unroll
      for (int u = 0; u < CU; u++) {
        const int i = lane + 32 * a, c = warp + NW * u;
        if (c <= i && i < n) {
          Lt[c * kLDT + ltPos(i)] = reg[a][u];
          if (writer) L[(int64_t)i * ldl + c] = reg[a][u];
        }
      }
  } else {
    for (int i = tid; i < kNB * kLDT + kNB; i += kPanelThreads) Lt[i] = T(0);
    __syncthreads();
    // Lt[c][r] = L[r][c]: lanes walk down a column (strided global reads, L2 resident; conflict-free smem writes)
#pragma unroll 4
    for (int c = warp; c < n; c += NW)
#pragma unroll
      for (int u = 0; u < LU; u++) {
        const int r = lane + 32 * u;
        if (r >= c && r < n) Lt[c * kLDT + ltPos(r)] = L[(int64_t)r * ldl + c];
      }
  }
  B200_PCLK(6)
  if (rows <= 0) return;
  __syncthreads();
  if (tid < n) invd[tid] = T(1) / Lt[tid * kLDT + ltPos(tid)];

  const int64_t r0 = (int64_t)slab * R;
  const int nr = (int)min((int64_t)R, rows - r0);
  {  // row slab -> smem, all loads in flight at once
    constexpr int RW = R / NW;
    T tmp[RW * LU];
#pragma unroll
    for (int a = 0; a < RW; a++)
#pragma unroll
      for (int u = 0; u < LU; u++) {
        const int r = warp + NW * a, c = lane + 32 * u;
        tmp[a * LU + u] = (r < nr && c < n) ? B[(r0 + r) * ldb + c] : T(0);
      }
#pragma unroll
    for (int a = 0; a < RW; a++)
#pragma unroll
      for (int u = 0; u < LU; u++) Xs[(warp + NW * a) * kLDX + lane + 32 * u] = tmp[a * LU + u];
  }
  __syncthreads();
  B200_PCLK(7)
  {
    const int row = tid / G, g = tid % G;  // G adjacent lanes share a row
    T* x = Xs + row * kLDX;
    for (int c0 = 0; c0 < n; c0 += 32) {
      const int cg = c0 + g * W;           // first column of this thread's piece
      const int pg = ltPos(cg);            // its padded position in an Lt row
      T acc[W];
#pragma unroll
      for (int k = 0; k < W; k++) acc[k] = x[cg + k];
#pragma unroll 2
      for (int q0 = 0; q0 < c0; q0 += 2) {
        const Vec2<T> xq = load2(x + q0);
        Vec2<T> l0[W / 2], l1[W / 2];
#pragma unroll
        for (int k = 0; k < W / 2; k++) l0[k] = load2(Lt + q0 * kLDT + pg + 2 * k);
#pragma unroll
        for (int k = 0; k < W / 2; k++) l1[k] = load2(Lt + (q0 + 1) * kLDT + pg + 2 * k);
#pragma unroll
        for (int k = 0; k < W / 2; k++) {
          acc[2 * k] -= xq.v[0] * l0[k].v[0];
          acc[2 * k + 1] -= xq.v[0] * l0[k].v[1];
        }
#pragma unroll
        for (int k = 0; k < W / 2; k++) {
          acc[2 * k] -= xq.v[1] * l1[k].v[0];
          acc[2 * k + 1] -= xq.v[1] * l1[k].v[1];
        }
      }
      // triangular part of the chunk, piece by piece: piece p is finished by its owner lane, broadcast with shuffles
      // to the lanes of the same row that own later pieces
#pragma unroll
      for (int p = 0; p < G; p++) {
        if (g == p) {
#pragma unroll
          for (int k = 0; k < W; k++) {
            acc[k] *= invd[cg + k];
#pragma unroll
            for (int k2 = k + 1; k2 < W; k2++) acc[k2] -= acc[k] * Lt[(cg + k) * kLDT + pg + k2];
          }
        }
        if (p + 1 < G) {
#pragma unroll
          for (int m = 0; m < W; m++) {
            const T xp = __shfl_sync(0xffffffffu, acc[m], (lane / G) * G + p);
            if (g > p) {
              const T* lrow = Lt + (c0 + p * W + m) * kLDT + pg;
#pragma unroll
              for (int k = 0; k < W; k++) acc[k] -= xp * lrow[k];
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < W; k++) x[cg + k] = acc[k];
      __syncwarp();
    }
  }
  __syncthreads();
  B200_PCLK(8)
  for (int r = warp; r < nr; r += NW)
#pragma unroll
    for (int u = 0; u < LU; u++) {
      const int c = lane + 32 * u;
      if (c < n) B[(r0 + r) * ldb + c] = Xs[r * kLDX + c];
    }
  B200_PCLK(9)
#undef B200_PCLK
}

// ------------------------------------------------------------------------------------------------ panel kernel v2
// Same job as panel_kernel<T, true> (diagonal block + 64 rows below it per CTA), but the triangular solve no longer
// runs as a second phase: the 64 slab rows are appended to the register-resident right-looking Cholesky as two more
// row slots per thread (the factorization of a 160 x 96 trapezoid), so they are solved against every 4-column pivot
// group in the same step, with the same two barriers, as the rows of the diagonal block. Measured on B200 (probe2,
// profiles/): v1 spent 31.6k cycles in the Cholesky phase and 27.0k in the separate solve phase per panel; the slab
// rows add FMA throughput only (no latency) to the former. The slab is prefetched with cp.async while the diagonal
// block loads, the load counter of the writer protocol is not waited for, and both results leave through
// shared-memory staging as coalesced row stores (v1: the writer CTA stored the factor column-wise, 7k cycles).
// 1 / sqrt(x) without the slow-path branch of rsqrt(double): hardware seed (MUFU.RSQ64H, ~2^-22 relative error over the
// whole exponent range) + two Newton steps y <- y (1.5 - 0.5 x y^2), straight-line code on the critical path of every
// pivot. A non-positive pivot yields NaN / Inf, which propagates into the factor exactly like sqrt() of it would.
__device__ __forceinline__ double rsqrtFast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}
__device__ __forceinline__ float rsqrtFast(float x) { return rsqrtf(x); }

constexpr int kP2Slots = 5;                  // row slots per thread: 3 (diagonal block) + 2 (slab)
constexpr int kP2Rows = kNB + kPanelRows;    // 160
constexpr int kP2LD = 97;                    // odd smem row stride: a warp walking down rows is conflict free

template <int BYTES>
__device__ __forceinline__ void cpAsyncZfill(uint32_t dst, const void* src, int srcBytes) {
  if constexpr (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(srcBytes));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(srcBytes));
}

template <typename T, int UF>
__global__ void __launch_bounds__(kPanelThreads, 1)
    panel2_kernel(int n, int64_t rows, Operand<T> Lop, int64_t ldl, Operand<T> Bop, int64_t ldb, const WavePanel* work,
                  int* counters, int lumpsInLaunch, long long* clk) {
#define B200_PCLK(i) \
  if (clk && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.z == 0) clk[i] = clock64();
  B200_CHAIN_ENTRY()
  B200_PCLK(0)
  constexpr int NW = kPanelThreads / 32;  // 8 warps
  constexpr int R = kPanelRows;           // 64 slab rows per CTA
  constexpr int RA = kP2Slots, CU = kNB / NW, LA = kNB / NW, LU = kNB / 32;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* S = reinterpret_cast<T*>(smemRaw);  // [kNB][kP2LD] diagonal block staging (in and out)
  T* Xs = S + kNB * kP2LD;               // [R][kP2LD]   slab staging (in and out)
  T* colbuf = Xs + R * kP2LD;            // [4][kP2Rows] raw columns of the current 4-column group
  T* ybuf = colbuf + 4 * kP2Rows;        // [kNB][4]     finished rows of the group: L[i][j0 .. j0+3]
  __shared__ int writerFlag;
  T* __restrict__ L = Lop.at(blockIdx.z);
  T* __restrict__ B = Bop.at(blockIdx.z);
  int slab = blockIdx.x, lumpIdx = 0;
  if (work) {  // batched over a work list (wavefront): one item = (lump column, 64-row slab)
    const WavePanel w = work[blockIdx.x];
    n = w.n, rows = w.rows, slab = w.slab, lumpIdx = w.lumpIdx;
    L += w.dataOff;
    B = L + (int64_t)n * n;
    ldl = ldb = n;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t r0 = (int64_t)slab * R;
  const int nr = (int)max((int64_t)0, min((int64_t)R, rows - r0));

  // slab rows -> Xs (zero filled outside nr x n), in flight while the diagonal block is fetched
#pragma unroll
  for (int j = 0; j < R * kNB / kPanelThreads; j++) {
    const int i = tid + kPanelThreads * j, r = i / kNB, c = i - r * kNB;
    const bool ok = r < nr && c < n;
    cpAsyncZfill<sizeof(T)>(smemAddr(Xs + r * kP2LD + c), ok ? B + (r0 + r) * ldb + c : B, ok ? (int)sizeof(T) : 0);
  }
  cpAsyncCommit();
  {  // lower triangle -> S (coalesced, every load in flight at once; S is zeroed meanwhile)
    T tmp[LA * LU];
#pragma unroll
    for (int a = 0; a < LA; a++)
#pragma unroll
      for (int u = 0; u < LU; u++) {
        const int r = warp + NW * a, c = lane + 32 * u;
        tmp[a * LU + u] = (c <= r && r < n) ? L[(int64_t)r * ldl + c] : T(0);
      }
    for (int i = tid; i < kNB * kP2LD; i += kPanelThreads) S[i] = T(0);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < LA; a++)
#pragma unroll
      for (int u = 0; u < LU; u++) {
        const int r = warp + NW * a, c = lane + 32 * u;
        if (c <= r && r < n) S[r * kP2LD + c] = tmp[a * LU + u];
      }
  }
  cpAsyncWait<0>();
  __syncthreads();
  B200_PCLK(1)
  // Writer protocol as in panel_kernel: every CTA of a lump column factors the diagonal block from the ORIGINAL values,
  // so the CTA that counts itself in last (every other one provably holds its copy) writes the factor back. The
  // result is only needed at the end of the kernel: nobody waits for the atomic here.
  if (tid == 0) {
    const int slabs = max(1, (int)((rows + R - 1) / R));
    int* ctr = counters + (int64_t)blockIdx.z * lumpsInLaunch + lumpIdx;
    __threadfence();
    const int old = atomicAdd(ctr, 1);
    writerFlag = (old == slabs - 1);
    if (old == slabs - 1) *ctr = 0;
  }
  // thread (warp w, lane l): rows l + 32 a (a < 3: diagonal block, a = 3, 4: slab rows l + 32 (a - 3)), cols w + 8 u
  T reg[RA][CU];
#pragma unroll
  for (int a = 0; a < RA; a++)
#pragma unroll
    for (int u = 0; u < CU; u++)
      reg[a][u] = a < 3 ? S[(lane + 32 * a) * kP2LD + warp + NW * u] : Xs[(lane + 32 * (a - 3)) * kP2LD + warp + NW * u];
  B200_PCLK(2)
  // Column loop: UF column slots (2 UF four-column steps) are unrolled per iteration, then the finished columns go to
  // the shared-memory staging and the remaining slots shift down by UF, so that register indices stay static while
  // the loop body stays small. (Fully unrolled - UF = 12, 340 KB of code - the kernel spent 55 % of its issue slots
  // waiting for instruction fetch: ncu stall_no_inst, profiles/.) During iteration u0, slot s holds column
  // warp + 8 (u0 + s).
#pragma unroll 1
  for (int u0 = 0; u0 < CU; u0 += UF) {
    const int rem = CU - u0;  // live slots
#pragma unroll
    for (int uu = 0; uu < UF; uu++) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int j0 = NW * (u0 + uu) + 4 * h;
        if (j0 < n) {
          const int q = warp - 4 * h;  // owner warps of the group: q in [0, 4)
          if (q >= 0 && q < 4) {
#pragma unroll
            for (int a = 0; a < RA; a++)
              if (a >= 3 || lane + 32 * a >= j0) colbuf[q * kP2Rows + lane + 32 * a] = reg[a][uu];
          }
          __syncthreads();
          // pivot block d[r][c] = entry (j0 + r, j0 + c), r >= c; columns beyond n act as identity
          T d[4][4], raw[RA][4];
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = c; r < 4; r++) d[r][c] = colbuf[c * kP2Rows + j0 + r];
#pragma unroll
          for (int a = 0; a < RA; a++)
#pragma unroll
            for (int c = 0; c < 4; c++)
              raw[a][c] = (UF == CU && a < 3 && 32 * a + 31 < j0) ? T(0) : colbuf[c * kP2Rows + lane + 32 * a];
#pragma unroll
          for (int c = 0; c < 4; c++)
            if (j0 + c >= n) d[c][c] = T(1);
          T rs[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int k = 0; k < c; k++) d[c][c] -= d[c][k] * d[c][k];
            rs[c] = rsqrtFast(d[c][c]);
#pragma unroll
            for (int r = c + 1; r < 4; r++) {
#pragma unroll
              for (int k = 0; k < c; k++) d[r][c] -= d[r][k] * d[c][k];
              d[r][c] *= rs[c];
            }
          }
          // own rows: y = raw * D^-T (entries above the diagonal of the pivot block and finished rows -> 0)
          T y[RA][4];
#pragma unroll
          for (int a = 0; a < RA; a++) {
            const int t = a < 3 ? lane + 32 * a - j0 : kP2Rows;  // row index relative to the group (slab rows: below)
            if (UF == CU && a < 3 && 32 * a + 31 < j0) {  // every row of the slot is finished (static when unrolled)
#pragma unroll
              for (int c = 0; c < 4; c++) y[a][c] = T(0);
              continue;
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
              T v = raw[a][c];
#pragma unroll
              for (int k = 0; k < c; k++) v -= y[a][k] * d[c][k];
              y[a][c] = (t >= c) ? v * rs[c] : T(0);
            }
          }
#pragma unroll
          for (int a = 0; a < 3; a++)
            if (warp == a) {
#pragma unroll
              for (int c = 0; c < 4; c++) ybuf[(lane + 32 * a) * 4 + c] = y[a][c];
            }
          __syncthreads();
          // rank-4 update of the columns after the group (slot sl = column warp + 8 (u0 + sl))
#pragma unroll
          for (int sl = uu; sl < CU; sl++) {
            if (UF == CU || sl < rem) {
              const int cb = NW * (u0 + sl);  // first column of the slot's 8-column block
              const Vec2<T> lo = load2(ybuf + (warp + cb) * 4), hi = load2(ybuf + (warp + cb) * 4 + 2);
              T yc[4] = {lo.v[0], lo.v[1], hi.v[0], hi.v[1]};
              if (sl == uu && !(h == 0 && warp >= 4)) {  // this slot's column is inside (or before) the group: no update
#pragma unroll
                for (int c = 0; c < 4; c++) yc[c] = T(0);
              }
#pragma unroll
              for (int a = 0; a < RA; a++) {
                // a < 3: some row of the slot is below the group (warp uniform) and the tile touches the lower triangle
                if (a >= 3 || (j0 + 3 < 32 * a + 31 && cb <= 32 * a + 31)) {
#pragma unroll
                  for (int c = 0; c < 4; c++) reg[a][sl] -= y[a][c] * yc[c];
                }
              }
            }
          }
          if (q >= 0 && q < 4) {
#pragma unroll
            for (int a = 0; a < RA; a++)
              if (a >= 3 || lane + 32 * a >= j0)
                reg[a][uu] = q == 0 ? y[a][0] : (q == 1 ? y[a][1] : (q == 2 ? y[a][2] : y[a][3]));
          }
        }
      }
    }
    // columns warp + 8 (u0 .. u0 + UF - 1) are final: stage them (lanes walk down rows: conflict free with the odd
    // stride), then shift the slots
#pragma unroll
    for (int uu = 0; uu < UF; uu++)
#pragma unroll
      for (int a = 0; a < RA; a++) {
        if (a < 3)
          S[(lane + 32 * a) * kP2LD + warp + NW * (u0 + uu)] = reg[a][uu];
        else
          Xs[(lane + 32 * (a - 3)) * kP2LD + warp + NW * (u0 + uu)] = reg[a][uu];
      }
    if (UF < CU) {
#pragma unroll
      for (int sl = 0; sl + UF < CU; sl++)
        if (sl + UF < rem) {
#pragma unroll
          for (int a = 0; a < RA; a++) reg[a][sl] = reg[a][sl + UF];
        }
    }
  }
  B200_PCLK(3)
  const bool writer = writerFlag != 0;
  __syncthreads();
  B200_PCLK(4)
  if (writer) {
#pragma unroll 4
    for (int r = warp; r < n; r += NW)
#pragma unroll
      for (int u = 0; u < LU; u++) {
        const int c = lane + 32 * u;
        if (c <= r) L[(int64_t)r * ldl + c] = S[r * kP2LD + c];
      }
  }
  for (int r = warp; r < nr; r += NW)
#pragma unroll
    for (int u = 0; u < LU; u++) {
      const int c = lane + 32 * u;
      if (c < n) B[(r0 + r) * ldb + c] = Xs[r * kP2LD + c];
    }
  B200_PCLK(5)
#undef B200_PCLK
}

// algorithmic flops of C = A B^T (lower-only: entries with col <= row)
double gemmFlops(int64_t m, int64_t n, int64_t k, bool lowerOnly) {
  double elems = (double)m * n;
  if (lowerOnly) {
    double nn = (double)std::min(m, n);
    elems = nn * (nn + 1) / 2 + (double)std::max<int64_t>(0, m - n) * n;
  }
  return 2.0 * elems * k;
}

template <typename KernelT>
void setSmem(KernelT kernel, size_t bytes) {
  ensureDynSmem((const void*)kernel, bytes);  // per (kernel, device), see B200Defs.h
}

// Programmatic dependent launch of the dense-factorization chain (panel -> trailing GEMM -> panel ...): every kernel
// of the chain releases its dependents at once and then waits for its predecessor (griddepcontrol at the top of the
// kernel), so the next kernel's CTAs are already resident when the last wave of the current one drains; the data
// dependence is carried by the wait. Opt-in (see chainPdl); an active profiler (events between launches) turns it off.
static bool chainPdl() {
  // measured on the BAL-shaped benchmark (round 1): factor 5.92 ms without, 6.04 ms with the chain -> off unless
  // BSPB200_PDL_CHAIN=1 (the early-resident CTAs of the next kernel take SM slots from the running one); the
  // solve steps, whose prologue is independent of the previous step, keep their PDL (SolveKernels.cu)
  static const bool env = getenv("BSPB200_PDL_CHAIN") && atoi(getenv("BSPB200_PDL_CHAIN")) != 0;
  return env && !profileEnabled();
}
template <typename... KArgs, typename... Args>
static void launchChain(void (*kernel)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(threads, 1, 1), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = chainPdl() ? 1 : 0;
  B200_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB = 1>
void launchGemmF64(cudaStream_t st, int batch, const GemmShape& s, double alpha, Operand<double> A, Operand<double> B,
                   double beta, Operand<double> C, bool aligned16) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  size_t smem = (size_t)STAGES * (BM + BN) * (BK + 4) * sizeof(double);
  dim3 grid(ceilDiv(s.n, BN), ceilDiv(s.m, BM), batch);
  if (aligned16) {
    auto kern = gemm_nt_f64_kernel<BM, BN, BK, WM, WN, STAGES, 2, MINB>;
    setSmem(kern, smem);
    launchChain(kern, grid, NT, smem, st, s, alpha, A, B, beta, C);
  } else {
    auto kern = gemm_nt_f64_kernel<BM, BN, BK, WM, WN, STAGES, 1, MINB>;
    setSmem(kern, smem);
    launchChain(kern, grid, NT, smem, st, s, alpha, A, B, beta, C);
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
  ProfScope prof(st, KC_GEMM, gemmFlops(m, n, k, lowerOnly) * batch, 0);
  // 16-byte cp.async needs even element offsets/strides and 16B-aligned bases (cudaMalloc bases are; batch pointers
  // supplied by the caller are only guaranteed 8B aligned -> 8-byte copies there)
  bool aligned16 = !A.many && !B.many && (lda % 2 == 0) && (ldb % 2 == 0) && (A.off % 2 == 0) && (B.off % 2 == 0) &&
                   (A.bstride % 2 == 0) && (B.bstride % 2 == 0) && ((uintptr_t)A.base % 16 == 0) &&
                   ((uintptr_t)B.base % 16 == 0);
  // big tiles when they fill the machine, small ones for skinny / small products
  static const int cfg = getenv("BSPB200_GEMM_CFG") ? atoi(getenv("BSPB200_GEMM_CFG")) : 1;
  // 128 x 64 tiles that are actually computed (lower-only launches skip the tiles above the diagonal)
  auto bigTiles = [&]() -> int64_t {
    const int64_t tm = ceilDiv(m, 128), tn = ceilDiv(n, 64);
    if (!lowerOnly) return tm * tn;
    int64_t t = 0;
    for (int64_t i = 0; i < tm; i++) t += std::min<int64_t>(tn, (i * 128 + 127) / 64 + 1);
    return t;
  };
  // the big tiles need about 1.5 CTAs per SM to keep the DMMA pipes busy (one 4-warp CTA alone on an SM stalls on its
  // own fragment loads: 1194^2 x 1248 lower-only = 105 big tiles ran at 9 TF/s, probe2); below that the 64 x 64 tiles
  // (8 warps) spread the same work over all SMs
  static const int64_t minBig = getenv("BSPB200_GEMM_MINBIG") ? atoi(getenv("BSPB200_GEMM_MINBIG")) : 222;
  if (m >= 96 && n >= 64 && bigTiles() * batch >= minBig) {
    // Same tile with 2 stages and registers capped for THREE CTAs per SM (444 slots instead of 296): measured better
    // exactly when it makes every tile resident at once (2442^2 x 1152 lower-only, 419 tiles: 0.355 -> 0.298 ms;
    // 4266 x 672 x 288: 0.107 -> 0.089 ms) and worse otherwise (3594^2 x 1632, 1653 tiles: 0.731 -> 0.780 ms).
    const int64_t tiles = bigTiles() * batch;
    if (cfg == 3 || (cfg == 1 && tiles > 2 * 148 && tiles <= 3 * 148))
      launchGemmF64<128, 64, 16, 64, 32, 2, 3>(st, batch, s, alpha, A, B, beta, C, aligned16);
    else if (cfg == 1)  // 128 x 64 tiles, 4 warps, 3 stages: two CTAs per SM (finer tail, epilogue/main-loop overlap)
      launchGemmF64<128, 64, 16, 64, 32, 3>(st, batch, s, alpha, A, B, beta, C, aligned16);
    else
      launchGemmF64<128, 128, 16, 64, 32, 4>(st, batch, s, alpha, A, B, beta, C, aligned16);
  }
  else {
    // Opt-in experiment for the next round (BSPB200_GEMM_SMALL=1, unmeasured): 64 x 32 tiles (4 warps) when the 64 x 64
    // grid is between one and two tiles per SM - e.g. the 5000 x 96 x 96 updates of the blocked Cholesky put 158 tiles
    // on 148 SMs, so ten SMs run two tiles back to back and set the launch's duration; half-size tiles even that out.
    static const bool smallTiles = getenv("BSPB200_GEMM_SMALL") && atoi(getenv("BSPB200_GEMM_SMALL")) != 0;
    const int64_t t64 = (int64_t)ceilDiv(m, 64) * ceilDiv(n, 64) * batch;
    if (smallTiles && t64 > 148 && t64 < 2 * 148)
      launchGemmF64<64, 32, 16, 32, 16, 4>(st, batch, s, alpha, A, B, beta, C, aligned16);
    else
      launchGemmF64<64, 64, 16, 32, 16, 4>(st, batch, s, alpha, A, B, beta, C, aligned16);
  }
}

template <>
void gemmNT<float>(cudaStream_t st, int batch, int64_t m, int64_t n, int64_t k, float alpha, Operand<float> A,
                   int64_t lda, Operand<float> B, int64_t ldb, float beta, Operand<float> C, int64_t ldc,
                   bool lowerOnly) {
  if (m <= 0 || n <= 0) return;
  GemmShape s{m, n, k, lda, ldb, ldc, lowerOnly ? 1 : 0};
  ProfScope prof(st, KC_GEMM, gemmFlops(m, n, k, lowerOnly) * batch, 0);
  // BSPB200_F32_GEMM=simt: the plain SIMT kernel (kept as the independent check of the tensor-core path)
  static const bool simt = getenv("BSPB200_F32_GEMM") && std::string(getenv("BSPB200_F32_GEMM")) == "simt";
  if (simt) {
    dim3 grid(ceilDiv(n, 64), ceilDiv(m, 64), batch);
    gemm_nt_simt_kernel<float, 64, 64, 16><<<grid, 256, 0, st>>>(s, alpha, A, B, beta, C);
    B200_LAUNCH_CHECK();
    return;
  }
  // 16-byte cp.async needs 4-float aligned rows and bases
  const bool aligned16 = !A.many && !B.many && (lda % 4 == 0) && (ldb % 4 == 0) && (A.off % 4 == 0) && (B.off % 4 == 0) &&
                         (A.bstride % 4 == 0) && (B.bstride % 4 == 0) && ((uintptr_t)A.base % 16 == 0) &&
                         ((uintptr_t)B.base % 16 == 0);
  const bool big = (int64_t)ceilDiv(m, 128) * ceilDiv(n, 64) * batch >= 148;
  auto launch = [&](auto kern, int BM, int BN, int threads) {
    const size_t smem = (size_t)3 * (BM + BN) * (16 + 4) * sizeof(float);
    setSmem(kern, smem);
    kern<<<dim3(ceilDiv(n, BN), ceilDiv(m, BM), batch), threads, smem, st>>>(s, alpha, A, B, beta, C);
  };
  if (big) {
    if (aligned16) launch(gemm_nt_tf32x3_kernel<128, 64, 16, 64, 32, 3, 4>, 128, 64, 128);
    else launch(gemm_nt_tf32x3_kernel<128, 64, 16, 64, 32, 3, 1>, 128, 64, 128);
  } else {
    if (aligned16) launch(gemm_nt_tf32x3_kernel<64, 64, 16, 32, 32, 3, 4>, 64, 64, 128);
    else launch(gemm_nt_tf32x3_kernel<64, 64, 16, 32, 32, 3, 1>, 64, 64, 128);
  }
  B200_LAUNCH_CHECK();
}

template <>
int maxBlockDim<double>() {
  return kNB;
}
template <>
int maxBlockDim<float>() {
  return kNB;
}

// Zero-initialised load counters of the panel kernels' writer election (the kernels leave them zero again), one buffer
// per (device, stream): launches on one stream run in order, so they can share counters; launches on different streams -
// the lanes of concurrent lump columns, other Solvers, other host threads - and on different devices never do.
// (The reference keeps its library handles per symbolic context, MatOpsCuda.cu:55-76; keying by the stream a context
// was given covers direct bspb200_dev_* calls as well.)
static int* panelCounters(cudaStream_t st, int64_t needed) {
  struct Buf {
    int* ptr = nullptr;
    int64_t cap = 0;
  };
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, Buf> bufs;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  Buf& b = bufs[{dev, st}];
  if (needed > b.cap) {
    B200_CUDA(cudaStreamSynchronize(st));
    if (b.ptr) cudaFree(b.ptr);
    b.cap = std::max<int64_t>(needed, 1 << 12);
    B200_CUDA(cudaMalloc((void**)&b.ptr, b.cap * sizeof(int)));
    B200_CUDA(cudaMemsetAsync(b.ptr, 0, b.cap * sizeof(int), st));
  }
  return b.ptr;
}

// phase time stamps of the panel kernel (diagnostics, off unless BSPB200_PANEL_CLK is set)
static long long* panelClockBuf() {
  static long long* buf = [] {
    long long* q = nullptr;
    if (getenv("BSPB200_PANEL_CLK") && atoi(getenv("BSPB200_PANEL_CLK")) != 0) {
      B200_CUDA(cudaMalloc((void**)&q, 64 * sizeof(long long)));
      B200_CUDA(cudaMemset(q, 0, 64 * sizeof(long long)));
    }
    return q;
  }();
  return buf;
}
int64_t debugRead(int what, void* out, int64_t bytes) {
  if (what == 1) return lumpCholDebugRead(out, bytes);
  if (what == 0 && panelClockBuf()) {
    const int64_t n = std::min<int64_t>(bytes, 64 * sizeof(long long));
    B200_CUDA(cudaDeviceSynchronize());
    B200_CUDA(cudaMemcpy(out, panelClockBuf(), n, cudaMemcpyDeviceToHost));
    return n;
  }
  return 0;
}

// BSPB200_PANEL=1 selects the two-phase panel kernel (v1) for the fused potrf + trsm launches; default v2
static int panelVersion() {
  static const int v = getenv("BSPB200_PANEL") ? atoi(getenv("BSPB200_PANEL")) : 2;
  return v;
}
template <typename T>
static size_t panel2Smem() {
  return ((size_t)kNB * kP2LD + (size_t)kPanelRows * kP2LD + 4 * kP2Rows + 4 * kNB) * sizeof(T);
}

// column slots unrolled per loop iteration of panel2_kernel (BSPB200_PANEL_UF: 1, 2, 3, 4, 6 or 12 = fully unrolled)
static int panelUF() {
  static const int v = getenv("BSPB200_PANEL_UF") ? atoi(getenv("BSPB200_PANEL_UF")) : 12;
  return v;
}
template <typename T, int UF>
static void launchPanel2UF(cudaStream_t st, dim3 grid, int n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B,
                           int64_t ldb, const WavePanel* work, int* counters, int lumpsInLaunch, long long* clk) {
  setSmem(panel2_kernel<T, UF>, panel2Smem<T>());
  launchChain(panel2_kernel<T, UF>, grid, kPanelThreads, panel2Smem<T>(), st, n, rows, L, ldl, B, ldb, work, counters,
              lumpsInLaunch, clk);
  B200_LAUNCH_CHECK();
}
template <typename T>
static void launchPanel2(cudaStream_t st, dim3 grid, int n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B,
                         int64_t ldb, const WavePanel* work, int* counters, int lumpsInLaunch, long long* clk) {
  switch (panelUF()) {
    case 1: return launchPanel2UF<T, 1>(st, grid, n, rows, L, ldl, B, ldb, work, counters, lumpsInLaunch, clk);
    case 2: return launchPanel2UF<T, 2>(st, grid, n, rows, L, ldl, B, ldb, work, counters, lumpsInLaunch, clk);
    case 3: return launchPanel2UF<T, 3>(st, grid, n, rows, L, ldl, B, ldb, work, counters, lumpsInLaunch, clk);
    case 4: return launchPanel2UF<T, 4>(st, grid, n, rows, L, ldl, B, ldb, work, counters, lumpsInLaunch, clk);
    case 6: return launchPanel2UF<T, 6>(st, grid, n, rows, L, ldl, B, ldb, work, counters, lumpsInLaunch, clk);
    default: return launchPanel2UF<T, 12>(st, grid, n, rows, L, ldl, B, ldb, work, counters, lumpsInLaunch, clk);
  }
}

template <typename T, bool DO_POTRF>
static void launchPanel(cudaStream_t st, int batch, int n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B,
                        int64_t ldb) {
  if (n > kNB) throw std::runtime_error("panel kernel: block too large");
  int ctas = std::max(1, ceilDiv(rows, kPanelRows));
  if (DO_POTRF && panelVersion() == 2) {
    launchPanel2<T>(st, dim3(ctas, 1, batch), n, rows, L, ldl, B, ldb, nullptr,
                    panelCounters(st, batch), 1, panelClockBuf());
    return;
  }
  size_t smem = ((size_t)kNB * kLDT + kNB + (size_t)kPanelRows * kLDX) * sizeof(T);
  setSmem(panel_kernel<T, DO_POTRF>, smem);
  panel_kernel<T, DO_POTRF><<<dim3(ctas, 1, batch), kPanelThreads, smem, st>>>(n, rows, L, ldl, B, ldb, nullptr,
                                                                               panelCounters(st, batch),
                                                                               1, panelClockBuf());
  B200_LAUNCH_CHECK();
}

// batched over a device work list (one CTA per (lump, slab) item): the small supernodes of one tree level
template <typename T>
void potrfTrsmPanelBatch(cudaStream_t st, int batch, Operand<T> data, const WavePanel* work, int64_t count,
                         int numLumps, double flops) {
  if (count <= 0) return;
  static_assert(WavePlan::kPanelRows == kPanelRows, "slab size of the plan and of the kernel differ");
  ProfScope prof(st, KC_POTRF_BLOCK, flops * batch, 0);
  if (panelVersion() == 2) {
    launchPanel2<T>(st, dim3((unsigned)count, 1, batch), 0, 0, data, 0, data, 0, work,
                    panelCounters(st, (int64_t)batch * numLumps), numLumps, nullptr);
    return;
  }
  size_t smem = ((size_t)kNB * kLDT + kNB + (size_t)kPanelRows * kLDX) * sizeof(T);
  setSmem(panel_kernel<T, true>, smem);
  panel_kernel<T, true><<<dim3((unsigned)count, 1, batch), kPanelThreads, smem, st>>>(
      0, 0, data, 0, data, 0, work, panelCounters(st, (int64_t)batch * numLumps), numLumps, nullptr);
  B200_LAUNCH_CHECK();
}
template void potrfTrsmPanelBatch<double>(cudaStream_t, int, Operand<double>, const WavePanel*, int64_t, int, double);
template void potrfTrsmPanelBatch<float>(cudaStream_t, int, Operand<float>, const WavePanel*, int64_t, int, double);

template <typename T>
void potrfBlock(cudaStream_t st, int batch, int n, Operand<T> A, int64_t lda) {
  if (n <= 0) return;
  ProfScope prof(st, KC_POTRF_BLOCK, (double)n * n * n / 3 * batch, 0);
  launchPanel<T, true>(st, batch, n, 0, A, lda, A, lda);
}

template <typename T>
void trsmBlock(cudaStream_t st, int batch, int n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B, int64_t ldb) {
  if (n <= 0 || rows <= 0) return;
  ProfScope prof(st, KC_TRSM_BLOCK, (double)rows * n * n * batch, 0);
  launchPanel<T, false>(st, batch, n, rows, L, ldl, B, ldb);
}

// diagonal block + rows below in one launch (leaf of the blocked factorization)
template <typename T>
static void potrfTrsmPanel(cudaStream_t st, int batch, int n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B,
                           int64_t ldb) {
  ProfScope prof(st, KC_POTRF_BLOCK, ((double)n * n * n / 3 + (double)rows * n * n) * batch, 0);
  launchPanel<T, true>(st, batch, n, rows, L, ldl, B, ldb);
}

// Recursive blocked Cholesky of the (n + rowsBelow) x n trapezoid (row-major, ld). Columns [c0, c0 + w):
//   w small  : potrfBlock on the diagonal block + trsmBlock on every row below it
//   otherwise: left half; trailing update of the right half (lower-only GEMM, K = left width); right half
template <typename T>
static void potrfRec(cudaStream_t st, int batch, int64_t totalRows, int64_t c0, int64_t w, Operand<T> A, int64_t ld) {
  const int nb = maxBlockDim<T>();
  if (w <= nb) {
    Operand<T> diag = shifted(A, c0 * ld + c0);
    potrfTrsmPanel<T>(st, batch, (int)w, totalRows - (c0 + w), diag, ld, shifted(A, (c0 + w) * ld + c0), ld);
    return;
  }
  int64_t blocks = (w + nb - 1) / nb;  // >= 2
  int64_t w1 = (blocks / 2) * nb;      // left width: a multiple of the block size, < w
  {
    // wave-aware split: the trailing update runs one 128x128 tile per CTA, one CTA per SM; among the splits near
    // the middle pick the one whose tile count fills whole waves of the 148 SMs best
    auto tiles = [&](int64_t left) {
      int64_t m = totalRows - (c0 + left), n2 = w - left;
      int64_t tn = (n2 + 127) / 128, tm = (m + 127) / 128;
      return tn * (tn + 1) / 2 + (tm - tn) * tn;
    };
    double best = -1;
    for (int64_t b = std::max<int64_t>(1, blocks * 3 / 10); b <= std::min(blocks - 1, blocks * 7 / 10); b++) {
      int64_t t = tiles(b * nb) * batch;
      if (t < 148) continue;
      double eff = (double)t / (148.0 * ((t + 147) / 148));
      eff -= 0.002 * std::abs((double)(2 * b - blocks));  // mild preference for balanced halves
      if (eff > best) best = eff, w1 = b * nb;
    }
  }
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
  if constexpr (std::is_same<T, double>::value) {
    // wide single-matrix lumps: one persistent tile-DAG kernel instead of the launch sequence below
    if (batch == 1 && !A.many && lumpCholesky(st, n, rowsBelow, A.at(0), ld)) return;
  }
  // (A right-looking depth-1 lookahead on a priority stream was measured in round 1 and removed: a panel launched behind
  // a GEMM that owns every SM waits for whole SMs to drain. The tile-DAG kernel above overlaps the chain instead.)
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
