// Blocked Cholesky of ONE wide lump column (diagonal block n x n + the rows below it) as a single persistent kernel:
// a left-looking tile DAG with device-side dependency flags, hand-written for sm_100a.
//
// Replaces, for wide supernodes, the reference's cusolverDn<t>potrf + cublas<t>trsm on a lump column
// (MatOpsCuda.cu:508-566) and this backend's own recursive schedule (DenseKernels.cu potrfRec: ~55 panel launches +
// ~54 GEMM launches in series for the 5226-wide camera lump of the BAL-shaped problem, nothing overlapping).
//
// The (n + rowsBelow) x n trapezoid is cut into 96 x 96 tiles. Tile (i, c), i > c, is finished in one go:
//     L(i,c) = ( A(i,c) - sum_{k<c} L(i,k) L(c,k)^T ) W_c^T ,     W_c = L(c,c)^-1 ,
// the sum running over the already finished tiles of block rows i and c (left-looking: every tile is read-modify-
// written exactly once, its accumulator lives in registers for the whole sum). Jobs are handed out by an arrival
// ticket in an order in which a job only ever waits for jobs with smaller tickets, which are finished or being
// executed by resident CTAs: no deadlock whatever the number of co-resident CTAs. A finished tile is published with
// store -> barrier -> release-store of the launch's epoch into its flag; the loader polls the flags of the two tiles
// of a K block with acquire loads before requesting them.
//
// The critical path of a blocked Cholesky is the chain of diagonal blocks. Here ONE CTA (the first to arrive) owns that
// chain: for block d only   L(d,d-1) = M1 W^T  ->  D = P - L(d,d-1) L(d,d-1)^T  ->  potrf(D)  ->  W_d   run there, with
// W_{d-1} still in its shared memory; the sums M1 = A(d,d-1) - sum, P = A(d,d) - sum are accumulated ahead of it by jobs
// of the other CTAs and handed over through global memory. Everything else - the bulk of the flops - fills the other
// SMs in the shadow of that chain.
//
// Data path: operand tiles travel HBM/L2 -> shared memory by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, one
// elected producer thread, full/empty mbarrier ring of 4 stages), the contraction runs on the fp64 tensor pipe
// (mma.sync.m8n8k4.f64 = DMMA; tcgen05.mma has no f64 kind), accumulators in registers. The diagonal 96 x 96
// Cholesky is blocked over 8-column panels: panel solve and trailing update on DMMA, the 8 x 8 pivot tile factored by one
// warp with every lane on its own register copy (potrfTile / factorTile8 below).
//
// Determinism: the summation order of every entry is fixed (k ascending), independent of the schedule.
// Non-SPD input: a non-positive pivot yields NaN, which propagates; flags are still published, nothing hangs.
// Every spin loop is bounded: on timeout the kernel raises the abort flag, stops waiting and poisons A[0] with NaN.
#include <cuda.h>
#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <tuple>
#include <vector>
#include "B200Kernels.h"

namespace BaSpaCho {
namespace b200 {
namespace {

constexpr int TB = 96;                 // tile edge
constexpr int BK = 16;                 // K per pipeline stage: 16 doubles = 128 bytes = the swizzle span
constexpr int NST = 4;                 // pipeline stages
constexpr int kConsumers = 256;        // 8 MMA warps; lane 0 of warp 0 doubles as the TMA producer (a ninth warp would
constexpr int kThreads = kConsumers;   // round the CTA up to 12 warps of registers: 168 per thread instead of 255)
constexpr int kTileBytes = TB * BK * 8;    // 12288: one 96-row operand tile of a stage
constexpr int kStageBytes = 3 * kTileBytes;  // A (96 rows) + B (up to 192 rows)
constexpr int LDE = 100;               // row stride (doubles) of the epilogue operands: [row][k], conflict-free fragments
constexpr int LDQ = LDE;               // row stride of the diagonal block staging (DMMA fragments read it too)
constexpr int kSmemE0 = 0;                            // [96][LDE]  M / L(d,d-1)            76800 B
constexpr int kSmemE1 = TB * LDE * 8;                 // [96][LDE]  W^T operand / [96][LDQ] diagonal block
constexpr int kSmemCol = 2 * TB * LDE * 8;            // [4][96] raw columns of the potrf step
constexpr int kSmemY = kSmemCol + 4 * TB * 8;         // [96][4] finished rows of the potrf step
constexpr int kSmemT = kSmemY + 4 * TB * 8;           // [2][32][LDE] scratch of the blocked inversion      51200 B
constexpr int kSmemBar = kSmemT + 2 * 32 * LDE * 8;   // mbarriers
constexpr int kSmemBytes = kSmemBar + 128 + 1024;     // + alignment slack
static_assert(NST * kStageBytes <= kSmemCol, "the pipeline stages alias the epilogue buffers");
static_assert(kSmemE1 == kSmemE0 + TB * LDE * 8, "the chain CTA ping-pongs between E0 and E1");
constexpr int kPkDoubles = 78 * 64;                    // packed lower 8 x 8 tiles of a diagonal block
constexpr int kMbufDoubles = TB * LDE + kPkDoubles;    // accumulated operands of one diagonal block in global memory

struct LcParams {
  double* A;
  int64_t ld;
  int n, rows;          // lump width, total rows (n + rowsBelow)
  int nbc, nbr;         // block columns / block rows
  int numJobs;          // entries of `jobs`
  const int4* jobs;     // the job list in ticket order: {i, c, k0 | seg << 20, k1 | type << 28 | last << 30}
  unsigned* done;       // [nbr * nbc] tile flags (epoch valued)
  unsigned* seg;        // [nbr * nbc] segments of the tile's sum applied so far (epoch * 256 + count)
  unsigned* wdone;      // [nbc] inverse flags
  unsigned* mdone;      // [nbc] flags of the accumulated operands of the diagonal blocks (mbuf): M1 ..
  unsigned* pdone;      // .. and P
  double* mbuf;         // [nbc][kMbufDoubles]: M1 = A(d,d-1) - sum ([96][LDE]) and P = A(d,d) - sum (78 packed 8 x 8 tiles)
  unsigned* ctr;        // [0] job ticket, [2] arrival counter (never reset)
  unsigned ticketBase, arriveBase;
  unsigned* abortFlag;  // holds the epoch of the launch that timed out (never reset)
  double* wbuf;         // [nbc][96 * LDE] block inverses, row-major, padded rows: the epilogue's operand layout as it is
  unsigned epoch;
  long long* dbg;       // diagnostics (BSPB200_LUMPCHOL_DBG=1): [64][16] clock64 stamps of the diagonal jobs, then
                        // [gridDim.x][4] per-CTA cycle totals (main loop, epilogue, flag waits of the producer, jobs)
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smemU32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbarArriveExpectTx(uint32_t bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
  while (!mbarTryWait(bar, parity)) {
  }
}
__device__ __forceinline__ void tmaLoad2D(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulkLoad(uint32_t dst, const void* src, int bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async;" ::: "memory"); }
// shared -> global bulk copy (async proxy, bulk-group completion): the issuing thread commits and later waits
__device__ __forceinline__ void bulkStore(void* dst, uint32_t src, int bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulkWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ unsigned ldAcquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stRelease(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void consumerBar() { __syncthreads(); }
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double rsqrtNewton(double x) {  // see DenseKernels.cu rsqrtFast
  // rsqrt.approx.f64 is good to ~20 bits; two Newton steps reach full precision. The dependent chain is MUFU (17.5
  // cycles) + 6 fp64 operations of 9 cycles (profiles/r02_ubench_fp64_pipe.txt): ~72 cycles per pivot.
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

// bounded wait for a flag to reach `value` (abort word compared with `epoch`)
__device__ __forceinline__ bool waitFlag(const unsigned* flag, unsigned value, unsigned* abortFlag, unsigned epoch) {
  unsigned spins = 0;
  while (ldAcquire(flag) != value) {
    if ((++spins & 63u) == 0) {
      if (*(volatile unsigned*)abortFlag == epoch) return false;
      if (spins > (1u << 24)) {
        atomicExch(abortFlag, epoch);
        return false;
      }
    }
  }
  return true;
}
// bounded wait for a flag to reach `epoch`; returns false after the abort flag was raised (by us or anybody)
__device__ __forceinline__ bool waitFlag(const unsigned* flag, unsigned epoch, unsigned* abortFlag) {
  unsigned spins = 0;
  while (ldAcquire(flag) != epoch) {
    if ((++spins & 63u) == 0) {
      if (*(volatile unsigned*)abortFlag == epoch) return false;
      if (spins > (1u << 24)) {
        atomicExch(abortFlag, epoch);
        return false;
      }
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ diagonal block
// 8 x 8 diagonal tile, factored by ONE warp with the tile spread over its lanes in the DMMA accumulator layout: lane
// (g, t) holds the entries (g, 2t) and (g, 2t+1) in v0 / v1 - exactly what the trailing-update DMMA leaves in the
// registers, so the tile goes from the update into the factorization without touching shared memory. Right-looking; per
// column: the pivot is broadcast by a shuffle, every lane takes its rsqrt, the owners scale the column, three shuffles
// hand every lane the two column entries its rank-1 update needs. The serial chain per column is
// shuffle -> rsqrt -> scale -> shuffle -> FMA (~200 cycles); the same factorization with every lane holding the whole tile
// in registers measured 2.2 k cycles per tile (the 110 independent updates of a tile compete with the chain for the
// issue slots of the one warp).
// The tile is factored in place in D (row stride LDQ, lower triangle) and the inverse of the factor (row-major 8 x 8) goes to
// invOut[64] for the panel solves. Round 2c: the 32 lanes all run the whole 8 x 8 factorization redundantly on their own copy (the
// lane-distributed version spent 2.2 k cycles per tile in the shuffles of its 8 pivot steps; the warp has nothing
// else to do there). e < 8: columns >= e act as identity; rows >= e of the tile ride along (they come out as M L^-T).
__device__ __forceinline__ void factorTile8(double* D, int e, double* invOut, int lane) {
  // every lane holds the whole lower triangle (broadcast loads, static register indices): no shuffle on the pivot chain
  double a[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c <= (r | 1); c += 2) {
      const double2 v = *reinterpret_cast<const double2*>(D + r * LDQ + c);
      a[r][c] = v.x, a[r][c + 1] = v.y;
    }
  // The inverse of the factor for the panel solves (X = M L^-T on the tensor pipe) is built alongside: lane j < 8 owns
  // column j (the other lanes repeat the work of lane j & 7); its entry c only needs row c of the factor, which is
  // final as soon as column c has been scaled - so the substitution runs in the shadow of the next pivot's rsqrt chain
  // instead of after the last one. Rows >= e count as identity rows.
  const int j = lane & 7;
  double rs[8], w[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    if (c < e) {
      rs[c] = rsqrtNewton(a[c][c]);
#pragma unroll
      for (int r = c; r < 8; r++) a[r][c] *= rs[c];
#pragma unroll
      for (int jj = c + 1; jj < 8; jj++)
#pragma unroll
        for (int r = jj; r < 8; r++) a[r][jj] -= a[r][c] * a[jj][c];
      double v = (c == j) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < c; q++) v -= a[c][q] * w[q];
      w[c] = v * rs[c];
    } else {  // identity column
      rs[c] = 1.0;
#pragma unroll
      for (int r = c; r < 8; r++) a[r][c] = (r == c) ? 1.0 : 0.0;
      w[c] = (c == j) ? 1.0 : 0.0;
    }
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; i++) invOut[i * 8 + j] = w[i];
  }
  // lane 0 writes the tile back (one predicate for straight-line vector stores: a store per lane-and-entry predicate
  // compiled into a divergent jump table, 10 k cycles). The entry right of the diagonal of an even row rides along.
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c <= (r | 1); c += 2) *reinterpret_cast<double2*>(D + r * LDQ + c) = make_double2(a[r][c], a[r][c + 1]);
  }
}

// In-place Cholesky of the lower triangle of S ([96][LDQ], zero outside the valid region) by the 256 threads of the CTA,
// blocked over 8-column panels with a one-panel lookahead:
//   (b) the rows below the diagonal tile are solved against its factor: X = M L^-T with the tile's 8 x 8 inverse (from
//       `fac2`, a by-product of factorTile8), one 8 x 8 tile = two DMMA at a time, tiles over the warps;
//   (c) the trailing update  S22 -= X X^T  runs on the tensor pipe, one 8 x 8 tile (K = 8: two DMMA) at a time: warp 0
//       updates the NEXT diagonal tile first and factors it right away (factorTile8) while the warps 1, 2, 3, 5, 6, 7
//       work through the other lower tiles - the serial pivot chain of panel p + 1 hides behind the update of panel p.
// Two barriers per panel; 26 k cycles for the 96 x 96 block. (Measured on the way: the register-resident scheme of
// panel2_kernel - four columns per step, all 256 threads updating their slots with DFMA - 57 k cycles; this blocked
// scheme without the lookahead 35 k; with a lane-distributed pivot tile (shuffles on the pivot chain) and row solves by
// substitution 36.8 k; redundant per-lane pivot tile 31.6 k; DMMA row solves 27.6 k; inverse inside the pivot loop 26.0 k.)
// nd < 96: columns >= nd act as identity; rows >= nd with entries in columns < nd (rows below a partial last diagonal
// block) ride along as the extra rows of a trapezoid and come out as M L^-T.
__device__ __forceinline__ void potrfTile(double* S, int nd, double* fac2 /* [2][72] */, int tid, int warp, int lane,
                                          long long* dbgp = nullptr /* diagnostics: 16 stamps of the first two panels */) {
  const int g = lane >> 2, t = lane & 3;
  if (warp == 0) factorTile8(S, min(8, nd), fac2, lane);
  __syncthreads();
#pragma unroll 1
  for (int j0 = 0, pb = 0; j0 < nd; j0 += 8, pb ^= 1) {
    const int e = min(8, nd - j0);
    const int m = (TB - j0 - 8) / 8;  // tile rows of the trailing matrix
    if (m <= 0) break;
    const double* fac = fac2 + pb * 72;
#define LC_PSTAMP(i) \
  if (dbgp && tid == 0 && j0 < 16) dbgp[(j0 >> 3) * 8 + (i)] = clock64();
    LC_PSTAMP(0)
    // (b) rows below the diagonal tile: X = M L^-T, one 8 x 8 tile (two DMMA) at a time, tiles over the warps. (A thread
    // per row by substitution measured ~800 cycles per panel: 36 dependent fp64 operations.)
    {
      const double b0 = fac[g * 8 + t], b1 = fac[g * 8 + 4 + t];  // "col" fragment of L^-T: B[k][n] = Linv[n][k]
      for (int tile = warp; tile < m; tile += 8) {
        double* xt = S + (j0 + 8 + 8 * tile + g) * LDQ + j0;
        double c0 = 0.0, c1 = 0.0;
        dmma(c0, c1, xt[t], b0);
        dmma(c0, c1, xt[4 + t], b1);
        __syncwarp();  // every lane has read its fragments of the tile
        *reinterpret_cast<double2*>(xt + 2 * t) = make_double2(c0, c1);
      }
    }
    LC_PSTAMP(1)
    __syncthreads();
    LC_PSTAMP(2)
    // (c) trailing update, lower tiles tt = ti (ti + 1) / 2 + tj of the m x m tile grid
    const double* X = S + (j0 + 8) * LDQ + j0;  // panel below the diagonal tile: [8 m][8]
    double* C = S + (j0 + 8) * LDQ + j0 + 8;
    const int nt = m * (m + 1) / 2;
    if (warp == 0) {
      const double* xa = X + g * LDQ + t;
      double c0 = 0.0, c1 = 0.0;
      dmma(c0, c1, xa[0], xa[0]);
      dmma(c0, c1, xa[4], xa[4]);
      double2* cp = reinterpret_cast<double2*>(C + g * LDQ + 2 * t);
      double2 cv = *cp;
      cv.x -= c0, cv.y -= c1;
      LC_PSTAMP(3)
      *cp = cv;
      if (j0 + 8 < nd) {
        __syncwarp();
        factorTile8(C, min(8, nd - j0 - 8), fac2 + (pb ^ 1) * 72, lane);
      }
      LC_PSTAMP(4)
    } else if (warp != 4) {
      // tiles 1 .. nt-1 over the warps 1, 2, 3, 5, 6, 7, two tiles in flight per warp. Warp 4 sits out: it shares the SM
      // sub-partition - and with it the fp64 / DMMA pipe - with warp 0, whose serial pivot chain is the critical path of
      // the panel (with warp 4 issuing DMMA the 8 x 8 factorization measured 10 k cycles instead of < 1 k)
      const int wslot = warp < 4 ? warp - 1 : warp - 2;  // 0 .. 5
      int tt = 1 + wslot;  // first tile of this warp
      int ti = (int)((sqrtf(8.0f * tt + 1.0f) - 1.0f) * 0.5f);
      ti += ((ti + 1) * (ti + 2) / 2 <= tt) ? 1 : 0;
      ti -= (ti * (ti + 1) / 2 > tt) ? 1 : 0;
      int tj = tt - ti * (ti + 1) / 2;
#pragma unroll 1
      for (; tt < nt; tt += 12) {
        int ti2 = ti, tj2 = tj + 6;
        while (tj2 > ti2) tj2 -= ti2 + 1, ti2++;
        const bool has2 = tt + 6 < nt;
        const double* xa = X + (8 * ti + g) * LDQ + t;
        const double* xb = X + (8 * tj + g) * LDQ + t;
        const double* ya = X + (8 * (has2 ? ti2 : ti) + g) * LDQ + t;
        const double* yb = X + (8 * (has2 ? tj2 : tj) + g) * LDQ + t;
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
        dmma(c0, c1, xa[0], xb[0]);
        dmma(d0, d1, ya[0], yb[0]);
        dmma(c0, c1, xa[4], xb[4]);
        dmma(d0, d1, ya[4], yb[4]);
        double2* cp = reinterpret_cast<double2*>(C + (8 * ti + g) * LDQ + 8 * tj + 2 * t);
        double2 cv = *cp;
        cv.x -= c0, cv.y -= c1;
        *cp = cv;
        if (has2) {
          double2* dp = reinterpret_cast<double2*>(C + (8 * ti2 + g) * LDQ + 8 * tj2 + 2 * t);
          double2 dv = *dp;
          dv.x -= d0, dv.y -= d1;
          *dp = dv;
        }
        // advance (ti, tj) by 12 tiles in the row-major lower-triangular enumeration
        tj += 12;
        while (tj > ti) tj -= ti + 1, ti++;
      }
    }
    __syncthreads();
    LC_PSTAMP(5)
#undef LC_PSTAMP
  }
}

// One warp: dst(8 x 8, ldd) = sign * X(8 x K, ldx) * Y(K x 8, ldy), K a multiple of 4, all in shared memory (DMMA m8n8k4;
// X row-major = the A fragment [g][t], Y row-major read as the "col" fragment: lane (g, t) takes Y[k0 + t][g]).
__device__ __forceinline__ void tileMma(double* dst, int ldd, const double* X, int ldx, const double* Y, int ldy, int K,
                                        double sign, int g, int t) {
  double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 4) dmma(c0, c1, X[g * ldx + k0 + t], Y[(k0 + t) * ldy + g]);
  dst[g * ldd + 2 * t] = sign * c0;
  dst[g * ldd + 2 * t + 1] = sign * c1;
}

// W = L^-1 for the lower-triangular L in S ([96][LDQ], identity beyond nd) -> Wm ([96][LDE], zero above the diagonal).
// Built bottom-up from block inverses, every product on the tensor pipe (this sits on the critical chain of the
// factorization: ~5 k cycles; a thread-per-column substitution + SIMT block products measured 54 k):
//   8 x 8 diagonal blocks by forward substitution (12 blocks x 8 columns = 96 threads, reciprocal diagonals), then for
//   block sizes b = 8, 16, 32:  W21 = -W22 (L21 W11)  doubles the inverted diagonal blocks to 2b, and finally the 3 x 3
//   arrangement of 32-blocks: W21, W32 as before, W31 = -W33 ([L31 L32] [W11; W21]).
// T ([64][LDE]) is scratch for the inner products.
__device__ __forceinline__ void invertTile(const double* S, double* Wm, double* T, int tid, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  double* dinv = T + 63 * LDE;  // last scratch row: reciprocals of the diagonal
  for (int i = tid; i < TB * LDE; i += kConsumers) Wm[i] = 0.0;
  if (tid < TB) dinv[tid] = 1.0 / S[tid * LDQ + tid];
  consumerBar();
  if (tid < TB) {  // 8 x 8 diagonal blocks: thread -> (block tid / 8, column tid % 8)
    const int b0 = (tid >> 3) * 8, c = tid & 7;
    double w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < i; q++) s -= S[(b0 + i) * LDQ + b0 + q] * w[q];
      w[i] = (i >= c) ? s * dinv[b0 + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) Wm[(b0 + i) * LDE + b0 + c] = w[i];
  }
  consumerBar();
  // 8 -> 16: six 16-blocks, one tile each: T = L21 W11, W21 = -W22 T (the same warp does both)
  if (warp < 6) {
    const int o = 16 * warp;
    double* Tb = T + warp * 8 * LDE;
    tileMma(Tb, LDE, S + (o + 8) * LDQ + o, LDQ, Wm + o * LDE + o, LDE, 8, 1.0, g, t);
    __syncwarp();
    tileMma(Wm + (o + 8) * LDE + o, LDE, Wm + (o + 8) * LDE + o + 8, LDE, Tb, LDE, 8, -1.0, g, t);
  }
  consumerBar();
  // 16 -> 32: three 32-blocks, 16 x 16 products (4 tiles each, 12 tiles per stage)
  for (int tile = warp; tile < 12; tile += 8) {
    const int o = 32 * (tile >> 2), ti = (tile >> 1) & 1, tj = tile & 1;
    tileMma(T + (tile >> 2) * 16 * LDE + ti * 8 * LDE + tj * 8, LDE, S + (o + 16 + 8 * ti) * LDQ + o, LDQ,
            Wm + o * LDE + o + 8 * tj, LDE, 16, 1.0, g, t);
  }
  consumerBar();
  for (int tile = warp; tile < 12; tile += 8) {
    const int o = 32 * (tile >> 2), ti = (tile >> 1) & 1, tj = tile & 1;
    tileMma(Wm + (o + 16 + 8 * ti) * LDE + o + 8 * tj, LDE, Wm + (o + 16 + 8 * ti) * LDE + o + 16, LDE,
            T + (tile >> 2) * 16 * LDE + tj * 8, LDE, 16, -1.0, g, t);
  }
  consumerBar();
  // 32 -> 96. T0 = L21 W11 (rows 0..31 of T), T1 = L32 W22 (rows 32..63): 32 tiles
  for (int tile = warp; tile < 32; tile += 8) {
    const int which = tile >> 4, ti = (tile >> 2) & 3, tj = tile & 3;
    const double* X = which ? S + (64 + 8 * ti) * LDQ + 32 : S + (32 + 8 * ti) * LDQ;
    const double* Y = which ? Wm + 32 * LDE + 32 + 8 * tj : Wm + 8 * tj;
    tileMma(T + (32 * which + 8 * ti) * LDE + 8 * tj, LDE, X, LDQ, Y, LDE, 32, 1.0, g, t);
  }
  consumerBar();
  // W21 = -W22 T0, W32 = -W33 T1
  for (int tile = warp; tile < 32; tile += 8) {
    const int which = tile >> 4, ti = (tile >> 2) & 3, tj = tile & 3;
    const int ro = which ? 64 : 32, co = which ? 32 : 0;
    tileMma(Wm + (ro + 8 * ti) * LDE + co + 8 * tj, LDE, Wm + (ro + 8 * ti) * LDE + ro, LDE,
            T + 32 * which * LDE + 8 * tj, LDE, 32, -1.0, g, t);
  }
  consumerBar();
  // T0 = [L31 L32] [W11; W21]  (K = 64: both operands are contiguous in k), then W31 = -W33 T0
  for (int tile = warp; tile < 16; tile += 8) {
    const int ti = tile >> 2, tj = tile & 3;
    tileMma(T + 8 * ti * LDE + 8 * tj, LDE, S + (64 + 8 * ti) * LDQ, LDQ, Wm + 8 * tj, LDE, 64, 1.0, g, t);
  }
  consumerBar();
  for (int tile = warp; tile < 16; tile += 8) {
    const int ti = tile >> 2, tj = tile & 3;
    tileMma(Wm + (64 + 8 * ti) * LDE + 8 * tj, LDE, Wm + (64 + 8 * ti) * LDE + 64, LDE, T + 8 * tj, LDE, 32, -1.0, g, t);
  }
  fenceProxyAsync();  // Wm leaves by a bulk copy (async proxy)
  consumerBar();
}

// ------------------------------------------------------------------------------------------------ the kernel
// Roles. The first CTA to arrive is the CHAIN CTA: it walks the diagonal blocks d = 0 .. nbc-1 and does the serial part
// of each (triangular product with the inverse of the previous block, which is still in its shared memory, D = P - L L^T,
// Cholesky of D, inverse), nothing else while there are any. Every other CTA takes JOBS by an arrival ticket from one
// list built on the host (buildJobs below): the tiles (i, c) below the sub-diagonal - rows i = c+2 .. nbr-1 of block
// column c, i = c+1 too where there is no diagonal block c+1 (rows below the lump's square) - and, per diagonal block, the
// two accumulate jobs whose results (M1 = A(d,d-1) - sum, P = A(d,d) - sum) are handed to the chain through global memory.
// The list is sorted by the chain step that makes a job runnable, hand-overs first: a job only waits for jobs with
// smaller tickets and for chain steps that in turn only wait for hand-overs with smaller tickets - no deadlock whatever
// the number of co-resident CTAs (tests/test_lumpchol_schedule.py replays this argument on the list).
// History (profiles/README.md): diagonal jobs as one unit on a set of chain CTAs (2.42 ms on the 5226-wide lump: a flag hop
// and a store / load of the inverse per block column on the chain) -> one chain CTA + dedicated accumulate CTAs (the
// accumulate jobs became the bound) -> one sorted list for everybody (2.28 ms).
struct Job {
  int i, c;     // tile (block row, block column); for P = (d, d) of the chain: i = d, c = d - 1
  int k0, k1;   // K blocks [k0, k1) of the tile's sum
  int seg;      // index of this segment among the tile's segments
  int diag;     // 0: regular tile, 1: M1 = (d, d-1), 2: P = (d, d)
  int last;     // the segment that finishes the tile (triangular product / hand-over to the chain)
};

// what the producer lane needs to request the operand tiles of a job
struct LoadCtx {
  const CUtensorMap* tmap;
  const unsigned* doneA;  // flags of block row i
  const unsigned* doneB;  // flags of block row c
  unsigned* abortFlag;
  unsigned epoch;
  int rowA0, rowB0, nBTiles;
  bool ok;  // false once the launch was aborted: stop waiting for flags, keep the pipeline protocol going
  long long waitCycles;  // diagnostics: cycles the producer lane spent waiting for source tiles
};
__device__ __forceinline__ void issueStage(LoadCtx& lc, int kt, unsigned q, uint32_t stagesBase, uint32_t fullBar0,
                                           uint32_t emptyBar0) {
  const unsigned stage = q % NST, parity = (q / NST) & 1;
  mbarWait(emptyBar0 + 8 * stage, parity ^ 1);  // every warp released the previous use of the slot
  if (kt % (TB / BK) == 0 && lc.ok) {           // first stage of a K block (segments start on one): its two source tiles must be published
    const int kb = kt / (TB / BK);
    const long long t0 = clock64();
    lc.ok = waitFlag(lc.doneA + kb, lc.epoch, lc.abortFlag) && waitFlag(lc.doneB + kb, lc.epoch, lc.abortFlag);
    lc.waitCycles += clock64() - t0;
    fenceProxyAsync();  // the tiles were written through the generic proxy, TMA reads through the async proxy
  }
  const uint32_t bar = fullBar0 + 8 * stage, dst = stagesBase + stage * kStageBytes;
  mbarArriveExpectTx(bar, kTileBytes * (1 + lc.nBTiles));
  tmaLoad2D(dst, lc.tmap, kt * BK, lc.rowA0, bar);
  tmaLoad2D(dst + kTileBytes, lc.tmap, kt * BK, lc.rowB0, bar);
  if (lc.nBTiles == 2) tmaLoad2D(dst + 2 * kTileBytes, lc.tmap, kt * BK, lc.rowB0 + TB, bar);
}

template <int TN>  // column tiles per warp: 6 (one 96-wide tile, warps 4 x 2 of 24 x 48) or 12 (two tiles, 24 x 96)
__device__ __forceinline__ void mainLoop(double (&acc)[3][12][2], int ktBegin, int ktEnd, uint32_t stagesBase,
                                         uint32_t fullBar0, uint32_t emptyBar0, unsigned& it, int rowA, int rowB, int g, int t,
                                         int lane, bool producer, LoadCtx& lc) {
  // swizzled position of this lane's fragment element inside a 128-byte row: 16-byte chunk (k >> 1) ^ (row & 7)
  uint32_t koff[BK / 4];
#pragma unroll
  for (int s = 0; s < BK / 4; s++) koff[s] = ((uint32_t)(((2 * s) | (t >> 1)) ^ g) << 4) | ((uint32_t)(t & 1) << 3);
  if (producer)
    for (int kt = ktBegin; kt < ktBegin + NST - 1 && kt < ktEnd; kt++)
      issueStage(lc, kt, it + (kt - ktBegin), stagesBase, fullBar0, emptyBar0);
  for (int kt = ktBegin; kt < ktEnd; kt++, it++) {
    const unsigned stage = it % NST, parity = (it / NST) & 1;
    mbarWait(fullBar0 + 8 * stage, parity);
    const uint32_t as = stagesBase + stage * kStageBytes + (uint32_t)(rowA + g) * 128;
    const uint32_t bs = stagesBase + stage * kStageBytes + kTileBytes + (uint32_t)(rowB + g) * 128;
#pragma unroll
    for (int s = 0; s < BK / 4; s++) {
      double af[3], bf[TN];
#pragma unroll
      for (int i = 0; i < 3; i++)
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(af[i]) : "r"(as + i * 8 * 128 + koff[s]));
#pragma unroll
      for (int j = 0; j < TN; j++)
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bf[j]) : "r"(bs + j * 8 * 128 + koff[s]));
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbarArrive(emptyBar0 + 8 * stage);
    // the slot of iteration kt - 1 is requested again for iteration kt + NST - 1
    if (producer && kt + NST - 1 < ktEnd) issueStage(lc, kt + NST - 1, it + NST - 1, stagesBase, fullBar0, emptyBar0);
    __syncwarp();
  }
}

// x = M[rbase .. rbase+24, :] W[cols, :]^T for the operands staged in E0 (M, [96][LDE]) and E1 (W): this warp's six 8-column
// tiles are the INTERLEAVED tiles 2 j + wn (j < 6) of the twelve. W is lower triangular, so column tile c only sees
// k < 8 c + 8: interleaving balances that triangular work between the two warps that share an SM sub-partition (with
// contiguous halves the warp of the left half finishes early and the other one issues DMMA alone, at half rate:
// 14.4 k cycles measured against 7.5 k of tensor-pipe time).
__device__ __forceinline__ void trsmProduct(double (&x)[3][6][2], const double* E0, const double* E1, int rbase, int wn,
                                            int g, int t) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) x[i][j][0] = x[i][j][1] = 0.0;
  const double* as = E0 + (rbase + g) * LDE + t;
  const double* bs = E1 + (8 * wn + g) * LDE + t;
  // Column tile j is live while kk < 16 j + 8 wn + 8. The dead tiles are skipped with a REAL branch (a switch that
  // falls through the live ones): an `if` around an mma is compiled to a predicated DMMA, and a predicated-off DMMA
  // still holds the tensor pipe for its 16 cycles (measured: no gain at all from the triangular structure).
  const int kkEnd = wn ? TB : TB - 8;
#pragma unroll 1
  for (int kk = 0; kk < kkEnd; kk += 4) {
    double af[3];
#pragma unroll
    for (int i = 0; i < 3; i++) af[i] = as[i * 8 * LDE + kk];
    const int jmin = kk < 8 * wn + 8 ? 0 : (kk - 8 * wn - 8) / 16 + 1;
#define LC_TRSM_TILE(J)                                                        \
  {                                                                            \
    const double b = bs[(J) * 16 * LDE + kk];                                  \
    dmma(x[0][J][0], x[0][J][1], af[0], b);                                    \
    dmma(x[1][J][0], x[1][J][1], af[1], b);                                    \
    dmma(x[2][J][0], x[2][J][1], af[2], b);                                    \
  }
    switch (jmin) {
      case 0: LC_TRSM_TILE(0)
      case 1: LC_TRSM_TILE(1)
      case 2: LC_TRSM_TILE(2)
      case 3: LC_TRSM_TILE(3)
      case 4: LC_TRSM_TILE(4)
      case 5: LC_TRSM_TILE(5)
      default: break;
    }
#undef LC_TRSM_TILE
  }
}

__global__ void __launch_bounds__(kThreads, 1) lump_chol_kernel(const __grid_constant__ CUtensorMap tmap, LcParams p) {
  extern __shared__ __align__(1024) unsigned char smemRawLc[];
  // NO integer round trip on this pointer (e.g. to align it by hand): the compiler would lose the address space and
  // turn every shared-memory access of the epilogues into a generic load / store (measured: the diagonal-block Cholesky
  // took 80 k cycles instead of 30 k). __align__(1024) places the dynamic segment on the 1024-byte boundary the
  // 128-byte swizzle needs; checked once below.
  unsigned char* smem = smemRawLc;
  double* E0 = reinterpret_cast<double*>(smem + kSmemE0);
  double* E1 = reinterpret_cast<double*>(smem + kSmemE1);
  double* colbuf = reinterpret_cast<double*>(smem + kSmemCol);
  double* ybuf = reinterpret_cast<double*>(smem + kSmemY);
  const uint32_t stagesBase = smemU32(smem);
  const uint32_t fullBar0 = smemU32(smem + kSmemBar), emptyBar0 = fullBar0 + 8 * NST, wBar = fullBar0 + 16 * NST;
  const uint32_t mBar = wBar + 8, pBar = wBar + 16;  // chain CTA: M1 / P of the next diagonal block have landed
  __shared__ int jobS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = tid == 0;
  const int g = lane >> 2, t = lane & 3;
  // consumer warps: 4 x 2. wn = warp >> 2, so that each column half (the T1 warps, the T2 warps of a diagonal job, which
  // work alone in parts of its epilogue) has one warp on every SM sub-partition (warp id mod 4) - with wn = warp & 1 the
  // four warps of a half share two sub-partitions and the DMMA rate of those phases halves
  const int wm = warp & 3, wn = warp >> 2;

  if (tid == 0) {
    if (stagesBase & 1023u) __trap();
    for (int s = 0; s < NST; s++) {
      mbarInit(fullBar0 + 8 * s, 1);
      mbarInit(emptyBar0 + 8 * s, kConsumers / 32);
    }
    mbarInit(wBar, 1);
    mbarInit(mBar, 1);
    mbarInit(pBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

#define LC_STAMP(i) \
  if (p.dbg && tid == 0 && job.diag == 1 && bi < 64) p.dbg[bi * 16 + (i)] = clock64();
  long long cycMain = 0, cycEpi = 0, cycWait = 0, nJobs = 0;
  unsigned wUses = 0;  // bulk loads of a block inverse so far (phase of wBar)
  unsigned it = 0;  // pipeline iteration counter, continues across the jobs of this CTA (producer and consumers agree)
  double* __restrict__ A = p.A;
  const int64_t ld = p.ld;

  __shared__ int chainS;
  if (tid == 0) chainS = (int)(atomicAdd(p.ctr + 2, 1u) - p.arriveBase);
  __syncthreads();
  const int arrival = chainS;
  // arrival 0: THE chain CTA (serial part of every diagonal block; once through, it joins the others); everybody else
  // works through the job list
  if (arrival == 0) {
    // ================================================================================ the chain of diagonal blocks
    // Step d:  L1 = L(d,d-1) = M1 W_{d-1}^T  ->  D = P - L1 L1^T  ->  potrf(D)  ->  W_d = L(d,d)^-1, with W_{d-1} still in
    // shared memory from the previous step (no flag hop, no store / load of the inverse on the chain) and M1, P arriving
    // from the accumulate jobs by bulk copies. The two 96 x LDE buffers swap roles every step: W_{d-1} | M1 -> X = L1,
    // then D is built over the dead W_{d-1} and W_d over the dead L1.
    double* Tk = reinterpret_cast<double*>(smem + kSmemT);  // P (packed tiles) until D is built, then inversion scratch
    unsigned mUses = 0, pUses = 0;
    bool wPending = false;  // thread 0: W_{d-1} is on its way to global memory, flag not yet published
    // thread 0: the operands of block dd, M1 into dstM; `block` = false: only if they are already published
    auto requestOperands = [&](int dd, double* dstM, bool block) -> bool {
      if (*(volatile unsigned*)p.abortFlag != p.epoch) {
        if (!block && ((dd > 0 && ldAcquire(p.mdone + dd) != p.epoch) || ldAcquire(p.pdone + dd) != p.epoch)) return false;
        const long long t0 = clock64();
        if (dd > 0) waitFlag(p.mdone + dd, p.epoch, p.abortFlag);
        waitFlag(p.pdone + dd, p.epoch, p.abortFlag);
        cycWait += clock64() - t0;
      }
      fenceProxyAsync();
      const double* src = p.mbuf + (int64_t)dd * kMbufDoubles;
      if (dd > 0) {
        mbarArriveExpectTx(mBar, TB * LDE * 8);
        bulkLoad(smemU32(dstM), src, TB * LDE * 8, mBar);
      }
      mbarArriveExpectTx(pBar, kPkDoubles * 8);
      bulkLoad(smemU32(Tk), src + TB * LDE, kPkDoubles * 8, pBar);
      return true;
    };
#define LC_CSTAMP(i) \
  if (p.dbg && tid == 0 && d < 64) p.dbg[d * 16 + (i)] = clock64();
    if (tid == 0) requestOperands(0, E1, true);
    const long long tChain = clock64();
    for (int d = 0; d < p.nbc; d++) {
      const int par = d & 1;
      double* Ea = E0 + par * (TB * LDE);        // W_{d-1}, then the diagonal block
      double* Eb = E0 + (par ^ 1) * (TB * LDE);  // M1, then L1, then W_d
      double* S = Ea;
      const int c = d - 1, rowA0 = d * TB;
      const int nd = min(TB, p.n - d * TB);
      const int rbase = 24 * wm;
      LC_CSTAMP(0)
      if (p.dbg && tid == 0 && d < 64) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.dbg[d * 16 + 2]));
      if (c >= 0) {
        mbarWait(mBar, mUses & 1);
        mUses++;
        LC_CSTAMP(3)
        double x[3][6][2];
        trsmProduct(x, Eb, Ea, rbase, wn, g, t);
        if (tid == 0 && wPending) {  // W_{d-1}: the copy out of Ea has had the whole product to complete
          bulkWaitAll();
          stRelease(p.wdone + c, p.epoch);
          if (p.dbg && c < 64) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.dbg[c * 16 + 12]));
          wPending = false;
        }
        consumerBar();  // every read of M1 and W_{d-1} is done
        LC_CSTAMP(4)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const int r = rbase + 8 * i + g, cc = 8 * (2 * j + wn) + 2 * t;
            const double2 v = make_double2(x[i][j][0], x[i][j][1]);
            *reinterpret_cast<double2*>(Eb + r * LDE + cc) = v;
            // L1 -> global straight from the registers (plain stores, nobody waits for them here: the flag is
            // published after the next product; bulk row copies out of Eb cost a proxy fence + 96 issues + a wait: 3 k cycles)
            if (rowA0 + r < p.rows) __stcg(reinterpret_cast<double2*>(A + ((int64_t)rowA0 + r) * ld + (int64_t)c * TB + cc), v);
          }
      }
      // a full interior block overwrites its whole lower triangle below (what is left of W_{d-1} above the diagonal of
      // the diagonal tiles is never read); a partial one must be zero outside its valid region
      if (nd < TB || rowA0 + TB > p.rows)
        for (int idx = tid; idx < TB * LDQ / 2; idx += kConsumers) reinterpret_cast<double2*>(S)[idx] = make_double2(0.0, 0.0);
      consumerBar();
      LC_CSTAMP(5)
      mbarWait(pBar, pUses & 1);  // P has landed
      pUses++;
      // D = P - L1 L1^T on the 78 lower tiles, straight into S. Warp (wm, wn): three tile rows chosen so that every wm
      // holds 18 - 21 lower tiles ({11,4,1}, {10,5,2}, {9,6,3}, {8,7,0}), column tiles 2 j + wn: 9 fragment loads per 18
      // DMMA, and the two warps of an SM sub-partition (same wm) carry the same load
      {
        double y[3][6][2];
        int rowT[3];
        rowT[0] = 11 - wm, rowT[1] = wm == 3 ? 7 : 4 + wm, rowT[2] = wm == 3 ? 0 : 1 + wm;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) y[i][j][0] = y[i][j][1] = 0.0;
        if (c >= 0) {
          const double* bs = Eb + (8 * wn + g) * LDE + t;
          // live column tiles of row i: 2 j + wn <= rowT[i]; dead ones skipped by a real branch (see trsmProduct)
          int nact[3];
#pragma unroll
          for (int i = 0; i < 3; i++) nact[i] = rowT[i] >= wn ? (rowT[i] - wn) / 2 + 1 : 0;
#pragma unroll 1
          for (int kk = 0; kk < TB; kk += 4) {
            double bf[6];
#pragma unroll
            for (int j = 0; j < 6; j++) bf[j] = bs[j * 16 * LDE + kk];
#pragma unroll
            for (int i = 0; i < 3; i++) {
              const double a = Eb[(8 * rowT[i] + g) * LDE + kk + t];
              switch (nact[i]) {
                case 6: dmma(y[i][5][0], y[i][5][1], a, bf[5]);
                case 5: dmma(y[i][4][0], y[i][4][1], a, bf[4]);
                case 4: dmma(y[i][3][0], y[i][3][1], a, bf[3]);
                case 3: dmma(y[i][2][0], y[i][2][1], a, bf[2]);
                case 2: dmma(y[i][1][0], y[i][1][1], a, bf[1]);
                case 1: dmma(y[i][0][0], y[i][0][1], a, bf[0]);
                default: break;
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const int ti = rowT[i], tj = 2 * j + wn;
            if (tj <= ti) {
              const double2 pv = *reinterpret_cast<const double2*>(Tk + (ti * (ti + 1) / 2 + tj) * 64 + g * 8 + 2 * t);
              const int r = 8 * ti + g, cc = 8 * tj + 2 * t;
              // rows >= nd of the block (rows below a partial last diagonal block, when the lump has rows below) ride
              // along the factorization as the extra rows of a trapezoid: they come out as M L^-T
              if (cc <= r && cc < nd && rowA0 + r < p.rows) S[r * LDQ + cc] = pv.x - y[i][j][0];
              if (cc + 1 <= r && cc + 1 < nd && rowA0 + r < p.rows) S[r * LDQ + cc + 1] = pv.y - y[i][j][1];
            }
          }
      }
      consumerBar();
      if (c >= 0 && tid == 0) stRelease(p.done + (int64_t)d * p.nbc + c, p.epoch);  // L(d,d-1): stored before the product
      LC_CSTAMP(7)
      potrfTile(S, nd, colbuf, tid, warp, lane, (p.dbg && d == 20) ? p.dbg + 63 * 16 : nullptr);
      LC_CSTAMP(8)
      // L(d,d) -> global (lower triangle, coalesced rows): plain stores, nobody inside the launch waits for them
      for (int r = warp; r < nd; r += 8)
#pragma unroll
        for (int u = 0; u < 3; u++) {
          const int cc = lane + 32 * u;
          if (cc <= r) A[((int64_t)rowA0 + r) * ld + (int64_t)d * TB + cc] = S[r * LDQ + cc];
        }
      if (nd < TB) {
        // ride-along rows -> global; then rows / columns beyond the block become identity for the inversion
        for (int r = nd + warp; r < TB; r += 8)
          if (rowA0 + r < p.rows) {
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const int cc = lane + 32 * u;
              if (cc < nd) A[((int64_t)rowA0 + r) * ld + (int64_t)d * TB + cc] = S[r * LDQ + cc];
            }
          }
        consumerBar();
        for (int idx = tid; idx < (TB - nd) * TB; idx += kConsumers) {
          const int r = nd + idx / TB, cc = idx % TB;
          S[r * LDQ + cc] = (r == cc) ? 1.0 : 0.0;
        }
        for (int idx = tid; idx < nd * (TB - nd); idx += kConsumers) S[(idx / (TB - nd)) * LDQ + nd + idx % (TB - nd)] = 0.0;
        consumerBar();
      }
      if (d + 1 < p.nbr) {  // somebody below needs W_d = L(d,d)^-1
        invertTile(S, Eb, Tk, tid, warp, lane);  // ends with a proxy fence + barrier: S and the scratch are dead after it
        LC_CSTAMP(9)
        if (tid == 0) {
          bulkStore(p.wbuf + (int64_t)d * TB * LDE, smemU32(Eb), TB * LDE * 8);
          bulkCommit();
          // the next block's operands: M1 over the dead diagonal block. When they are already there (the usual case) the
          // chain goes straight on and W_d is published from inside the next step, once its copy has completed (~5 k
          // cycles that thread 0 - and with it the next triangular product - would otherwise wait here); else W_d is
          // published first (regular tiles wait for it) and the request blocks
          wPending = d + 1 < p.nbc && requestOperands(d + 1, Ea, false);
          if (!wPending) {
            bulkWaitAll();
            stRelease(p.wdone + d, p.epoch);
            if (p.dbg && d < 64) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.dbg[d * 16 + 12]));
            if (d + 1 < p.nbc) requestOperands(d + 1, Ea, true);
          }
        }
        LC_CSTAMP(10)
      }
    }
#undef LC_CSTAMP
    fenceProxyAsync();
    __syncthreads();
    cycEpi += clock64() - tChain, nJobs += p.nbc;
  }

  for (;;) {
    if (tid == 0) jobS = (int)(atomicAdd(p.ctr, 1u) - p.ticketBase);
    __syncthreads();
    const int jt = jobS;
    if (jt >= p.numJobs) break;
    Job job;
    {
      const int4 q = __ldg(p.jobs + jt);
      job.i = q.x, job.c = q.y, job.k0 = q.z, job.k1 = q.w & 0xfffffff, job.diag = (q.w >> 28) & 3, job.last = (q.w >> 30) & 1;
      job.seg = (int)((unsigned)q.z >> 20);
      job.k0 &= 0xfffff;
    }
    const int c = job.c, bi = job.i;
    // block row of the B operand: c for a tile (i, c) - regular tiles and M1 = (d, d-1) -, d itself for P = (d, d)
    const int browB = job.diag == 2 ? bi : (c < 0 ? 0 : c);
    const int rowA0 = bi * TB, rowB0 = browB * TB;
    // this tile's own column block in A and its valid extent
    const int ownCol = job.diag == 2 ? bi : c;
    unsigned* segFlag = p.seg + (int64_t)bi * p.nbc + ownCol;

    const long long tJob = clock64();
    LC_STAMP(13)
    {
      // the tile's own entries are only needed after the K loop: pull them into L2 now, one 128-byte line per request
      for (int idx = tid; idx < TB * 6; idx += kConsumers) {
        const int r = idx / 6, l = idx % 6;
        const int64_t gr = (int64_t)rowA0 + r, gc = (int64_t)ownCol * TB + l * 16;
        if (gr < p.rows && gc < p.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + gr * ld + gc));
      }
    }
    double acc[3][12][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 12; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    LoadCtx lc;
    lc.tmap = &tmap, lc.doneA = p.done + (int64_t)bi * p.nbc, lc.doneB = p.done + (int64_t)browB * p.nbc;
    lc.abortFlag = p.abortFlag, lc.epoch = p.epoch, lc.rowA0 = rowA0, lc.rowB0 = rowB0, lc.nBTiles = 1;
    lc.ok = *(volatile unsigned*)p.abortFlag != p.epoch;
    lc.waitCycles = 0;
    mainLoop<6>(acc, job.k0 * (TB / BK), job.k1 * (TB / BK), stagesBase, fullBar0, emptyBar0, it, 24 * wm, 48 * wn, g, t, lane,
                producer, lc);
    // an earlier segment of this tile's sum has to have landed in the tile before it is read below
    if (tid == 0 && job.seg > 0 && lc.ok) {
      const long long t0 = clock64();
      waitFlag(segFlag, p.epoch * 256u + (unsigned)job.seg, p.abortFlag, p.epoch);
      lc.waitCycles += clock64() - t0;
    }
    __syncthreads();  // (A) every stage has been consumed: the stage memory becomes the epilogue workspace
    const long long tMain = clock64();
    cycWait += lc.waitCycles;
    LC_STAMP(14)

    const int nc = c >= 0 ? min(TB, p.n - c * TB) : 0;  // valid columns of block column c
    if (!job.last) {
      // ---- a segment of the tile's sum: A(tile) -= acc, in place (only the lower 8 x 8 tiles of a diagonal block: the
      // entries above its diagonal are never written)
      const int rbase = 24 * wm, cbase = 48 * wn;
      const int ncOwn = min(TB, p.n - ownCol * TB);
      double2 av[3][6];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int r = rbase + 8 * i + g, cc = cbase + 8 * j + 2 * t;
          const bool on = rowA0 + r < p.rows && cc < ncOwn && (job.diag != 2 || 6 * wn + j <= 3 * wm + i);
          av[i][j] = make_double2(0.0, 0.0);
          if (on) av[i][j] = __ldcg(reinterpret_cast<const double2*>(A + ((int64_t)rowA0 + r) * ld + (int64_t)ownCol * TB + cc));
        }
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int r = rbase + 8 * i + g, cc = cbase + 8 * j + 2 * t;
          const bool on = rowA0 + r < p.rows && cc < ncOwn && (job.diag != 2 || 6 * wn + j <= 3 * wm + i);
          if (on) {
            double* dst = A + ((int64_t)rowA0 + r) * ld + (int64_t)ownCol * TB + cc;
            if (job.diag == 2 && 6 * wn + j == 3 * wm + i) {  // a diagonal 8 x 8 tile: entries with column <= row only
              if (cc <= r) __stcg(dst, av[i][j].x - acc[i][j][0]);
              if (cc + 1 <= r) __stcg(dst + 1, av[i][j].y - acc[i][j][1]);
            } else {
              __stcg(reinterpret_cast<double2*>(dst), make_double2(av[i][j].x - acc[i][j][0], av[i][j].y - acc[i][j][1]));
            }
          }
        }
      __syncthreads();
      if (tid == 0) stRelease(segFlag, p.epoch * 256u + (unsigned)job.seg + 1u);
      __syncthreads();  // (B)
      cycMain += tMain - tJob, cycEpi += clock64() - tMain, nJobs++;
      continue;
    }
    // W_c -> E1 by one bulk copy (the buffer holds the padded operand layout). Requested right away when W_c is already
    // published - the copy then runs under the staging below - else after the staging
    bool wIssued = false;
    auto requestW = [&](bool block) {
      if (wIssued || c < 0) return;
      if (*(volatile unsigned*)p.abortFlag != p.epoch) {
        if (!block && ldAcquire(p.wdone + c) != p.epoch) return;
        const long long t0 = clock64();
        waitFlag(p.wdone + c, p.epoch, p.abortFlag);
        cycWait += clock64() - t0;
      }
      fenceProxyAsync();
      mbarArriveExpectTx(wBar, TB * LDE * 8);
      bulkLoad(smemU32(E1), p.wbuf + (int64_t)c * TB * LDE, TB * LDE * 8, wBar);
      wIssued = true;
    };
    if (tid == 0 && !job.diag) requestW(false);
    if (!job.diag) {
      // ---- regular tile: M = A(i,c) - acc -> E0 ; W_c -> E1 ; X = M W^T -> global
      const int rbase = 24 * wm, cbase = 48 * wn;
      // every load of the tile's own entries is issued before the first use (one L2 round trip instead of eighteen)
      {
        double2 av[3][6];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const int r = rbase + 8 * i + g, cc = cbase + 8 * j + 2 * t;
            const int64_t gr = (int64_t)rowA0 + r;
            av[i][j] = make_double2(0.0, 0.0);
            if (gr < p.rows && cc < nc) av[i][j] = __ldcg(reinterpret_cast<const double2*>(A + gr * ld + c * TB + cc));
          }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const int r = rbase + 8 * i + g, cc = cbase + 8 * j + 2 * t;
            *reinterpret_cast<double2*>(E0 + r * LDE + cc) =
                make_double2(av[i][j].x - acc[i][j][0], av[i][j].y - acc[i][j][1]);
          }
      }
      if (tid == 0) requestW(true);
      consumerBar();
      mbarWait(wBar, wUses & 1);  // W_c has landed in E1 (requested by thread 0 right after the main loop)
      wUses++;
      double x[3][6][2];
      trsmProduct(x, E0, E1, rbase, wn, g, t);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int r = rbase + 8 * i + g, cc = 8 * (2 * j + wn) + 2 * t;
          const int64_t gr = (int64_t)rowA0 + r;
          if (gr < p.rows && cc < nc)
            *reinterpret_cast<double2*>(A + gr * ld + c * TB + cc) = make_double2(x[i][j][0], x[i][j][1]);
        }
      fenceProxyAsync();
      consumerBar();
      // one fence by the releasing thread after the barrier (the pattern of a grid sync): the barrier orders the CTA's
      // stores before it, the release is cumulative at gpu scope
      if (tid == 0) stRelease(p.done + (int64_t)bi * p.nbc + c, p.epoch);  // release at gpu scope is cumulative over the barrier
      __syncthreads();  // (B)
      cycMain += tMain - tJob, cycEpi += clock64() - tMain, nJobs++;
      continue;
    }

    // ---- accumulate job of block d = bi: M1 = A(d,d-1) - sum (diag = 1) or P = A(d,d) - sum (diag = 2, lower tiles,
    // packed) goes to the block's slot of mbuf; the chain CTA picks them up with two bulk copies when it gets there
    const int d = bi;
    const int nd = min(TB, p.n - d * TB);  // valid rows / columns of the diagonal block
    const int rbase = 24 * wm, cbase = 48 * wn;
    double* __restrict__ mb = p.mbuf + (int64_t)d * kMbufDoubles;
    if (job.diag == 1) {
      double2 av[3][6];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int r = rbase + 8 * i + g, cc = cbase + 8 * j + 2 * t;
          av[i][j] = make_double2(0.0, 0.0);
          if (rowA0 + r < p.rows) av[i][j] = __ldcg(reinterpret_cast<const double2*>(A + ((int64_t)rowA0 + r) * ld + c * TB + cc));
        }
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int r = rbase + 8 * i + g, cc = cbase + 8 * j + 2 * t;
          __stcg(reinterpret_cast<double2*>(mb + r * LDE + cc), make_double2(av[i][j].x - acc[i][j][0], av[i][j].y - acc[i][j][1]));
        }
    } else {
      double* __restrict__ pk = mb + TB * LDE;
      double2 av[3][6];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {  // lower tiles only (warp uniform)
          const int ti = 3 * wm + i, tj = 6 * wn + j, r = rbase + 8 * i + g, cc = 8 * tj + 2 * t;
          av[i][j] = make_double2(0.0, 0.0);
          if (tj <= ti && cc < nd && rowA0 + r < p.rows)
            av[i][j] = __ldcg(reinterpret_cast<const double2*>(A + ((int64_t)rowA0 + r) * ld + (int64_t)d * TB + cc));
        }
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const int ti = 3 * wm + i, tj = 6 * wn + j;
          if (tj <= ti)
            __stcg(reinterpret_cast<double2*>(pk + (ti * (ti + 1) / 2 + tj) * 64 + g * 8 + 2 * t),
                   make_double2(av[i][j].x - acc[i][j][0], av[i][j].y - acc[i][j][1]));
        }
    }
    LC_STAMP(15)
    __syncthreads();
    if (tid == 0) {
      stRelease((job.diag == 1 ? p.mdone : p.pdone) + d, p.epoch);
      if (p.dbg && d < 64 && job.diag == 1) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.dbg[d * 16 + 1]));
      if (p.dbg && d < 64 && job.diag == 2) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.dbg[d * 16 + 6]));
    }
    __syncthreads();  // (B)
    cycMain += tMain - tJob, cycEpi += clock64() - tMain, nJobs++;
  }

#undef LC_STAMP
  if (p.dbg && tid == 0) {
    long long* q = p.dbg + 64 * 16 + (int64_t)blockIdx.x * 4;
    q[0] = cycMain, q[1] = cycEpi, q[2] = cycWait, q[3] = nJobs;
  }
  if (tid == 0 && *(volatile unsigned*)p.abortFlag == p.epoch) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    A[0] = nan;
  }
}

// ------------------------------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encodeTiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// flags, ticket, abort word and block inverses of the launches issued on one stream of one device. Launches on a stream
// run in order and every launch leaves the state consistent (fresh epoch, ticket base advanced), so a stream's launches
// share it; different streams (the lanes of concurrent lumps, other Solvers) get their own.
struct LcState {
  DevBuf<long long> dbg;
  DevBuf<unsigned> words;
  DevBuf<double> wbuf, mbuf;
  int nbcCap = 0;
  int64_t tileCap = 0;
  unsigned epoch = 0, ticketBase = 0, arriveBase = 0;
  // job lists by shape (block columns, block rows, segment length): built and uploaded once per stream and shape
  struct JobList {
    DevBuf<int4> dev;
    int64_t count = 0;
  };
  std::map<std::tuple<int, int, int>, std::unique_ptr<JobList>> jobLists;
};

// The job list of a shape. Every tile's sum over K blocks [0, K) is cut into segments of ~seglen blocks with staggered
// boundaries; a segment that does not finish the tile is applied to the tile in place. Ticket order: by the LAST K block
// a job needs (= the chain step that makes it runnable), inside that the hand-overs to the chain first, then the jobs
// that finish a tile, then the rest, nearest column first. A job only depends on jobs with a smaller key (the previous
// segment of its tile, the finishing jobs of the tiles it reads) and on the chain CTA, which in turn only waits for
// hand-over jobs with smaller keys: whatever the number of resident CTAs, the job with the smallest ticket still
// running can always finish. Without the segments a CTA held a tile from its first K block to its last, and in the
// first third of the run - when tiles can only advance one K block per chain step - nearly every CTA sat waiting with
// one tile while a thousand other tiles could have used the same K block (round 2b: 20 % of the CTA time in waits).
std::vector<int4> buildJobs(int nbc, int nbr, int seglen, int lag) {
  struct J {
    int key, rank, c, i, k0, k1, seg, type, last;
  };
  std::vector<J> js;
  auto addTile = [&](int i, int c, int type) {
    const int K = std::max(0, c);  // K blocks of the sum (c = d - 1 for the two tiles of diagonal block d)
    int b = K > 0 ? ((i * 7 + c * 3) % seglen) + 1 : 0, k0 = 0, seg = 0;
    for (;;) {
      int k1 = std::min(K, b);
      if (K - k1 < std::max((seglen + 1) / 2, lag + 1)) k1 = K;  // no short tail segment (and room for the lag below)
      const int last = k1 == K;
      const int rank = type != 0 ? (last ? 0 : 1) : (last ? 2 : 3);
      // a segment that does not finish its tile is not urgent: it takes its ticket `lag` chain steps after it became
      // runnable, behind the jobs the chain is waiting for (still before the tile's next segment / finishing job)
      js.push_back(J{last ? k1 : k1 + lag, rank, c, i, k0, k1, seg, type, last});
      if (last) break;
      k0 = k1, b = k1 + seglen, seg++;
    }
  };
  for (int d = 0; d < nbc; d++) {
    if (d > 0) addTile(d, d - 1, 1);
    addTile(d, d - 1, 2);
  }
  for (int c = 0; c < nbc; c++)
    for (int i = (c + 1 < nbc ? c + 2 : c + 1); i < nbr; i++) addTile(i, c, 0);
  std::stable_sort(js.begin(), js.end(), [](const J& a, const J& b) {
    return std::make_tuple(a.key, a.rank, a.c, a.i) < std::make_tuple(b.key, b.rank, b.c, b.i);
  });
  std::vector<int4> out;
  out.reserve(js.size());
  for (const J& j : js) {
    if (j.seg > 255 || j.k1 >= (1 << 20)) throw std::runtime_error("lump_chol: job list out of range");
    out.push_back(make_int4(j.i, j.c, j.k0 | (j.seg << 20), j.k1 | (j.type << 28) | (j.last << 30)));
  }
  return out;
}

long long*& lastDbg() {
  static long long* p = nullptr;
  return p;
}
std::mutex& lcMutex() {
  static std::mutex m;
  return m;
}
LcState& lcState(cudaStream_t st) {
  static std::map<std::pair<int, cudaStream_t>, LcState> states;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(lcMutex());
  return states[{dev, st}];
}

}  // namespace

// read at every call (a getenv per wide lump is noise): tests and probes flip the switches at run time
// diagnostics: the stamps of the last instrumented launch (synchronizes); returns the bytes copied
int64_t lumpCholDebugRead(void* out, int64_t bytes) {
  if (!lastDbg()) return 0;
  const int64_t n = std::min<int64_t>(bytes, (64 * 16 + 1024 * 4) * (int64_t)sizeof(long long));
  B200_CUDA(cudaDeviceSynchronize());
  B200_CUDA(cudaMemcpy(out, lastDbg(), n, cudaMemcpyDeviceToHost));
  return n;
}

// host only (no device needed): the job list lump_chol_kernel would run for a shape, as 5 numbers per job
// {block row, block column, first K block, end K block, type | last << 4} (type 0 regular tile, 1 = M1, 2 = P of a diagonal
// block); returns the number of jobs, copies at most `capJobs` of them. For the schedule-validity test.
int64_t lumpCholJobList(int nbc, int nbr, int seglen, int lag, int32_t* out, int64_t capJobs) {
  if (seglen <= 0) seglen = 1 << 20;
  const std::vector<int4> jobs = buildJobs(nbc, nbr, seglen, lag);
  for (int64_t i = 0; i < (int64_t)jobs.size() && i < capJobs; i++) {
    const int4 q = jobs[i];
    out[5 * i + 0] = q.x, out[5 * i + 1] = q.y, out[5 * i + 2] = q.z & 0xfffff, out[5 * i + 3] = q.w & 0xfffffff;
    out[5 * i + 4] = ((q.w >> 28) & 3) | (((q.w >> 30) & 1) << 4);
  }
  return (int64_t)jobs.size();
}

// hint of the caller's host thread: this many tile-DAG launches are about to run side by side (lumps of one tree level on
// their lanes); 0 / 1 = alone. The launch then asks for its share of the SMs instead of all it could use alone.
thread_local int tlsLumpConcurrency = 0;
void lumpCholSetConcurrency(int n) { tlsLumpConcurrency = n; }

int lumpCholMinWidth() {
  const char* e = getenv("BSPB200_LUMPCHOL_MIN");
  return e ? atoi(e) : 384;
}

// Cholesky of the (n + rowsBelow) x n trapezoid by the tile-DAG kernel; false when the shape / alignment is not eligible
// (the caller then runs the recursive schedule)
bool lumpCholesky(cudaStream_t st, int64_t n, int64_t rowsBelow, double* A, int64_t ld) {
  const char* sw = getenv("BSPB200_LUMPCHOL");
  if ((sw && atoi(sw) == 0) || n < lumpCholMinWidth() || n > (1 << 20) || n + rowsBelow > (1 << 24)) return false;
  if (ld % 2 != 0 || n % 2 != 0 || ((uintptr_t)A & 15) != 0 || !encodeTiled()) return false;  // TMA: 16-byte rows / base
  LcParams p;
  p.A = A, p.ld = ld, p.n = (int)n, p.rows = (int)(n + rowsBelow);
  p.nbc = ceilDiv(n, TB), p.nbr = ceilDiv(n + rowsBelow, TB);
  LcState& s = lcState(st);
  const int64_t tiles = (int64_t)p.nbr * p.nbc;
  if (p.nbc > s.nbcCap || tiles > s.tileCap) {
    B200_CUDA(cudaStreamSynchronize(st));
    s.nbcCap = std::max(p.nbc, s.nbcCap * 2);
    s.tileCap = std::max(tiles, s.tileCap * 2);
    s.words.resize((size_t)(4 + 3 * s.nbcCap + 2 * s.tileCap));
    s.wbuf.resize((size_t)s.nbcCap * TB * LDE);
    s.mbuf.resize((size_t)s.nbcCap * kMbufDoubles);
    B200_CUDA(cudaMemsetAsync(s.words.ptr(), 0, s.words.size() * sizeof(unsigned), st));
    s.epoch = 0, s.ticketBase = s.arriveBase = 0;
  }
  p.ctr = s.words.ptr(), p.abortFlag = s.words.ptr() + 1, p.wdone = s.words.ptr() + 4;
  p.mdone = s.words.ptr() + 4 + s.nbcCap, p.pdone = s.words.ptr() + 4 + 2 * s.nbcCap;
  p.done = s.words.ptr() + 4 + 3 * s.nbcCap, p.seg = p.done + s.tileCap;
  p.wbuf = s.wbuf.ptr(), p.mbuf = s.mbuf.ptr();
  p.epoch = ++s.epoch, p.ticketBase = s.ticketBase, p.arriveBase = s.arriveBase;
  // segment length: BSPB200_LUMPCHOL_SEG (K blocks of 96; 0 = whole sums, one job per tile: the default); at most 255
  // segments a tile. Measured on the 5226-wide lump (profiles/README.md): whole sums 2.28 ms; segments of 2 / 4 / 8 / 16 /
  // 24 K blocks 3.01 / 2.72 / 2.58 / 2.51 / 2.46 ms (every extra job costs ~20 k cycles of pipeline ramp, read-modify-
  // write of the tile and flag waits, and the flood of early segments holds back the jobs the chain waits for; giving
  // the segments a later ticket - BSPB200_LUMPCHOL_LAG - recovers little: 8/3 2.57, 8/6 2.54, 16/4 2.50 ms).
  int seglen = 0;
  if (const char* e = getenv("BSPB200_LUMPCHOL_SEG")) seglen = atoi(e);
  if (seglen <= 0) seglen = 1 << 20;
  seglen = std::max(seglen, p.nbc / 200 + 1);
  int lag = 0;
  if (const char* e = getenv("BSPB200_LUMPCHOL_LAG")) lag = std::max(0, atoi(e));
  auto& jl = s.jobLists[std::make_tuple(p.nbc, p.nbr, seglen * 64 + lag)];
  if (!jl) {
    jl = std::make_unique<LcState::JobList>();
    const std::vector<int4> host = buildJobs(p.nbc, p.nbr, seglen, lag);
    jl->count = (int64_t)host.size();
    jl->dev.resize(host.size());
    B200_CUDA(cudaMemcpyAsync(jl->dev.ptr(), host.data(), host.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaStreamSynchronize(st));  // `host` goes out of scope
  }
  p.jobs = jl->dev.ptr(), p.numJobs = (int)jl->count;
  const int64_t jobs = jl->count;
  p.dbg = nullptr;
  if (const char* e = getenv("BSPB200_LUMPCHOL_DBG")) {
    if (atoi(e) != 0) {
      if (s.dbg.size() == 0) s.dbg.resize(64 * 16 + 1024 * 4);
      B200_CUDA(cudaMemsetAsync(s.dbg.ptr(), 0, s.dbg.size() * sizeof(long long), st));
      p.dbg = s.dbg.ptr();
      lastDbg() = s.dbg.ptr();
    }
  }

  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)(n + rowsBelow)};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {BK, TB}, estr[2] = {1, 1};
  const CUresult r = encodeTiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;

  int dev = 0, sms = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // CTAs: the chain of diagonal blocks bounds the run time from below (~36.5 us per block column, measured); more CTAs
  // than it takes to finish the flops in that time only spin on flags - and occupy SMs that lumps factored concurrently
  // on other streams (independent lumps of a tree level) could use. Twice the break-even count at ~150 GF/s per SM
  // (232 GF/s is the DMMA peak of one SM; a CTA spends ~20 % of its time waiting for source tiles), plus the chain CTA
  // and the accumulate jobs running ahead of it. Measured (profiles/README.md): n = 2000: 24 CTAs 1.00 ms, 32: 0.80,
  // 48: 0.77, 96: 0.77; n = 1000 + 700 rows: 16 CTAs 0.60 ms, 24: 0.44, 32: 0.41, 48: 0.41.
  const double flops = (double)n * n * n / 3 + (double)rowsBelow * n * n;
  int64_t want = (int64_t)(2.0 * flops / (p.nbc * 36.5e-6 * 150e9)) + (int64_t)(0.4 * p.nbc) + 3;
  // concurrent lumps (GRID 120x120, leaf level on 8 lanes: factor 16.9 ms uncapped, 16.1 ms with 40 CTAs each)
  if (tlsLumpConcurrency >= 4) want = std::min<int64_t>(want, std::max<int64_t>(24, 2 * (int64_t)sms / tlsLumpConcurrency + 3));
  if (const char* e = getenv("BSPB200_LUMPCHOL_GRID")) want = atoi(e);
  // at least two CTAs: the chain CTA only consumes what the accumulate jobs of the others publish
  const int grid = (int)std::max<int64_t>(2, std::min<int64_t>({(int64_t)sms, jobs + 1, std::max<int64_t>(want, 16)}));
  ProfScope prof(st, KC_LUMP_CHOL, flops, 0);
  ensureDynSmem((const void*)lump_chol_kernel, kSmemBytes);
  lump_chol_kernel<<<grid, kThreads, kSmemBytes, st>>>(tmap, p);
  B200_LAUNCH_CHECK();
  // every CTA ends with one failing fetch of the ticket
  s.ticketBase += (unsigned)(jobs + grid);
  s.arriveBase += (unsigned)grid;
  return true;
}

}  // namespace b200
}  // namespace BaSpaCho
