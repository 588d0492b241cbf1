// Irregular kernels of the B200 backend (sm_100a): sparse elimination, assemble scatter, span-wise
// pseudo factor, elimination-range triangular solves and the vector gather/scatter of the dense solves.
// These are HBM-bound index-driven kernels: no tensor cores, coalesced row-wise accesses, fixed summation
// order (deterministic, no atomics - unlike reference MatOpsCuda.cu:235-331, 883-1012).
#include <type_traits>
#include "B200Sparse.h"

namespace BaSpaCho {
namespace b200 {
namespace {

// in-place lower Cholesky of an s x s row-major block (stride ld) by ONE thread
// (semantics of reference MathUtils.h:36-63 `cholesky`)
template <typename T>
__device__ __forceinline__ void choleskySerial(T* D, int s, int64_t ld) {
  for (int j = 0; j < s; j++) {
    T d = D[j * ld + j];
    for (int q = 0; q < j; q++) d -= D[j * ld + q] * D[j * ld + q];
    d = sqrt(d);
    D[j * ld + j] = d;
    T inv = T(1) / d;
    for (int i = j + 1; i < s; i++) {
      T v = D[i * ld + j];
      for (int q = 0; q < j; q++) v -= D[i * ld + q] * D[j * ld + q];
      D[i * ld + j] = v * inv;
    }
  }
}

// x <- x * tril(L)^-T for one row x of length s (reference MathUtils.h:66-79 `solveUpperT`)
template <typename T>
__device__ __forceinline__ void solveRowSerial(const T* L, int s, int64_t ldl, T* x) {
  for (int j = 0; j < s; j++) {
    T v = x[j];
    for (int q = 0; q < j; q++) v -= x[q] * L[j * ldl + q];
    x[j] = v / L[j * ldl + j];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Sparse elimination, step 1: one warp per lump. Lane 0 factors the (tiny) diagonal block, then the lanes take
// the below-diagonal rows (consecutive lanes -> consecutive rows -> coalesced).  S > 0: compile-time width.
template <typename T, int S>
__global__ void __launch_bounds__(128) elim_factor_lumps_kernel(DevSkel sk, Mats<T> mats, int64_t lumpsBegin,
                                                                int64_t lumpsEnd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t lump = lumpsBegin + (int64_t)blockIdx.x * 4 + warp;
  if (lump >= lumpsEnd) return;
  T* data = mats.at(blockIdx.z);
  const int s = S > 0 ? S : (int)(sk.lumpStart[lump + 1] - sk.lumpStart[lump]);
  const int64_t cb = sk.chainColPtr[lump], ce = sk.chainColPtr[lump + 1];
  T* D = data + sk.chainData[cb];
  const int64_t rowsBelow = sk.chainRowsTillEnd[ce - 1] - s;
  T* below = D + (int64_t)s * s;
  if constexpr (S > 0) {
    // the first row of every lane is requested BEFORE lane 0 factors the diagonal block: the (long) global latency of
    // the panel rows overlaps the serial sqrt/divide chain instead of following it
    T x0[S];
    const bool has0 = lane < rowsBelow;
    if (has0) {
#pragma unroll
      for (int j = 0; j < S; j++) x0[j] = below[(int64_t)lane * S + j];
    }
    // every lane factors the (tiny) diagonal block redundantly in registers: one broadcast load, no second trip to
    // memory, no intra-warp hand-off; lane 0 writes the factor back
    T Dr[S * S];
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
      for (int q = 0; q <= j; q++) Dr[j * S + q] = D[j * S + q];
    T L[S * (S + 1) / 2];
#pragma unroll
    for (int j = 0; j < S; j++) {
      T d = Dr[j * S + j];
#pragma unroll
      for (int q = 0; q < j; q++) d -= Dr[j * S + q] * Dr[j * S + q];
      d = sqrt(d);
      Dr[j * S + j] = d;
      const T inv = T(1) / d;
      L[j * (j + 1) / 2 + j] = inv;  // reciprocal of the diagonal entry
#pragma unroll
      for (int i = j + 1; i < S; i++) {
        T v = Dr[i * S + j];
#pragma unroll
        for (int q = 0; q < j; q++) v -= Dr[i * S + q] * Dr[j * S + q];
        Dr[i * S + j] = v * inv;
      }
    }
#pragma unroll
    for (int j = 0; j < S; j++)
#pragma unroll
      for (int q = 0; q < j; q++) L[j * (j + 1) / 2 + q] = Dr[j * S + q];
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < S; j++)
#pragma unroll
        for (int q = 0; q <= j; q++) D[j * S + q] = Dr[j * S + q];
    }
    for (int64_t r = lane; r < rowsBelow; r += 32) {
      T* xr = below + r * S;
      T x[S];
      if (r == lane) {
#pragma unroll
        for (int j = 0; j < S; j++) x[j] = x0[j];
      } else {
#pragma unroll
        for (int j = 0; j < S; j++) x[j] = xr[j];
      }
#pragma unroll
      for (int j = 0; j < S; j++) {
        T v = x[j];
#pragma unroll
        for (int q = 0; q < j; q++) v -= x[q] * L[j * (j + 1) / 2 + q];
        x[j] = v * L[j * (j + 1) / 2 + j];
      }
#pragma unroll
      for (int j = 0; j < S; j++) xr[j] = x[j];
    }
  } else {
    if (lane == 0) choleskySerial(D, s, (int64_t)s);
    __syncwarp();
    for (int64_t r = lane; r < rowsBelow; r += 32) solveRowSerial(D, s, (int64_t)s, below + r * s);
  }
}

// Sparse elimination, step 2 (destination-major gather): E consecutive threads own one destination block,
// thread e computes element (e / nc, e % nc) = sum over the block's pair tasks of B[r,:] . A[c,:], then
// subtracts it from the target once.
template <typename T>
__global__ void __launch_bounds__(256) elim_gather_kernel(DevElimPlan p, Mats<T> mats, int E) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t d = gid / E;
  if (d >= p.numDst) return;
  T* data = mats.at(blockIdx.z);
  const int nr = p.dstRows[d], nc = p.dstCols[d];
  const int tBegin = p.dstTaskPtr[d], tEnd = p.dstTaskPtr[d + 1];
  T* dst = data + p.dstOff[d];
  const int64_t stride = p.dstStride[d];
  for (int e = (int)(gid - d * E); e < nr * nc; e += E) {
    const int r = e / nc, c = e - r * nc;
    T acc = 0;
    for (int t = tBegin; t < tEnd; t++) {
      const int k = p.taskK[t];
      const T* __restrict__ a = data + p.taskA[t] + c * k;
      const T* __restrict__ b = data + p.taskB[t] + r * k;
      for (int q = 0; q < k; q++) acc += b[q] * a[q];
    }
    dst[r * stride + c] -= acc;
  }
}

// Fixed-shape variant (every destination NR x NC, every source lump K wide - e.g. 6x6 camera blocks fed by 3-wide
// points): LANES consecutive lanes own one destination and split its pair tasks; a lane keeps the whole NR x NC
// partial product in registers (each block is read once, 8-byte loads of consecutive addresses per lane), then a
// fixed-order butterfly over the LANES lanes and one read-modify-write of the target per entry.
template <typename T, int NR, int NC, int K, int LANES>
__global__ void __launch_bounds__(LANES > 32 ? LANES : 128)
    elim_gather_fixed_kernel(DevElimPlan p, Mats<T> mats, const int32_t* __restrict__ list, int64_t count) {
  // LANES <= 32: LANES adjacent lanes per destination; LANES > 32: the whole CTA (LANES threads) on one destination
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = gid / LANES;
  const int sub = (int)(gid % LANES);
  const bool live = slot < count;
  const int64_t d = live ? list[slot] : 0;
  T* data = mats.at(blockIdx.z);
  T acc[NR * NC];
#pragma unroll
  for (int e = 0; e < NR * NC; e++) acc[e] = T(0);
  if (live) {
    const int tEnd = p.dstTaskPtr[d + 1];
    int t = p.dstTaskPtr[d] + sub;
    uint32_t oa = 0, ob = 0;
    if (t < tEnd) oa = p.taskA[t], ob = p.taskB[t];
    while (t < tEnd) {
      const T* __restrict__ a = data + oa;
      const T* __restrict__ b = data + ob;
      const int tn = t + LANES;
      if (tn < tEnd) oa = p.taskA[tn], ob = p.taskB[tn];  // next task's offsets in flight with this task's blocks
      T av[NC * K], bv[NR * K];
#pragma unroll
      for (int i = 0; i < NC * K; i++) av[i] = a[i];
#pragma unroll
      for (int i = 0; i < NR * K; i++) bv[i] = b[i];
#pragma unroll
      for (int r = 0; r < NR; r++)
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int q = 0; q < K; q++) acc[r * NC + c] += bv[r * K + q] * av[c * K + q];
      t = tn;
    }
  }
  constexpr int WL = LANES > 32 ? 32 : LANES;
#pragma unroll
  for (int o = 1; o < WL; o <<= 1)
#pragma unroll
    for (int e = 0; e < NR * NC; e++) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  if constexpr (LANES > 32) {
    __shared__ T red[LANES / 32][NR * NC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < NR * NC; e++)
      if (lane == e % 32) red[warp][e] = acc[e];
    __syncthreads();
    if (live && threadIdx.x < NR * NC) {
      const int e = threadIdx.x;
      T tot = 0;
#pragma unroll
      for (int w = 0; w < LANES / 32; w++) tot += red[w][e];
      data[p.dstOff[d] + (e / NC) * (int64_t)p.dstStride[d] + (e % NC)] -= tot;
    }
  } else if (live) {
    T* dst = data + p.dstOff[d];
    const int64_t stride = p.dstStride[d];
#pragma unroll
    for (int e = 0; e < NR * NC; e++)
      if (e % LANES == sub) dst[(e / NC) * stride + (e % NC)] -= acc[e];
  }
}

// The per-lane gather with 16-byte operand loads (round 2). A block starts on an 8-byte boundary only: an even element
// offset is 16-byte aligned and is read as NE/2 double2, an odd one as one double + (NE-2)/2 double2 + one double. The
// two cases are separate straight-line paths (lanes of a warp may take either: both run, each with its lanes) that fill
// the same registers with static indices - the parity-SELECT formulation of round 1 (load 20, pick 18) cost 36 selects
// per block and 164 registers. Halves the L1 sector accesses per task (36 -> 18..20), which bound the 8-byte version
// (ncu r02: 261 M sectors, 0.96 per cycle and SM). Same task ownership and summation order as the plain kernel.
template <typename T, int NE>
__device__ __forceinline__ void loadBlockVec(T (&v)[NE], const T* __restrict__ src) {
  static_assert(sizeof(T) == 8 && NE % 2 == 0, "double blocks with an even element count");
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NE / 2; i++) {
      const double2 q = __ldg(reinterpret_cast<const double2*>(src) + i);
      v[2 * i] = q.x, v[2 * i + 1] = q.y;
    }
  } else {
    v[0] = __ldg(src);
#pragma unroll
    for (int i = 0; i < NE / 2 - 1; i++) {
      const double2 q = __ldg(reinterpret_cast<const double2*>(src + 1) + i);
      v[1 + 2 * i] = q.x, v[2 + 2 * i] = q.y;
    }
    v[NE - 1] = __ldg(src + NE - 1);
  }
}
template <int NR, int NC, int K, int LANES>
__global__ void __launch_bounds__(128)
    elim_gather_fixed_vec_kernel(DevElimPlan p, Mats<double> mats, const int32_t* __restrict__ list, int64_t count) {
  using T = double;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = gid / LANES;
  const int sub = (int)(gid % LANES);
  const bool live = slot < count;
  const int64_t d = live ? list[slot] : 0;
  T* data = mats.at(blockIdx.z);
  T acc[NR * NC];
#pragma unroll
  for (int e = 0; e < NR * NC; e++) acc[e] = T(0);
  if (live) {
    const int tEnd = p.dstTaskPtr[d + 1];
    int t = p.dstTaskPtr[d] + sub;
    uint32_t oa = 0, ob = 0;
    if (t < tEnd) oa = p.taskA[t], ob = p.taskB[t];
    while (t < tEnd) {
      const T* __restrict__ a = data + oa;
      const T* __restrict__ b = data + ob;
      const int tn = t + LANES;
      if (tn < tEnd) oa = p.taskA[tn], ob = p.taskB[tn];  // next task's offsets in flight with this task's blocks
      T av[NC * K], bv[NR * K];
      loadBlockVec<T, NC * K>(av, a);
      loadBlockVec<T, NR * K>(bv, b);
#pragma unroll
      for (int r = 0; r < NR; r++)
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int q = 0; q < K; q++) acc[r * NC + c] += bv[r * K + q] * av[c * K + q];
      t = tn;
    }
  }
#pragma unroll
  for (int o = 1; o < LANES; o <<= 1)
#pragma unroll
    for (int e = 0; e < NR * NC; e++) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  if (live) {
    T* dst = data + p.dstOff[d];
    const int64_t stride = p.dstStride[d];
#pragma unroll
    for (int e = 0; e < NR * NC; e++)
      if (e % LANES == sub) dst[(e / NC) * stride + (e % NC)] -= acc[e];
  }
}

// Gather on the fp64 tensor pipe (round 2): one WARP per destination block, one DMMA (mma.sync.m8n8k4.f64) per pair task.
// The NR x K block B is the A fragment (lane (g, t) holds B[g][t]), the NC x K block A - transposed - the B fragment
// (lane (g, t) holds A[g][t]); rows / columns beyond the block and k >= K are zero lanes that do not load. So a task costs
// each lane two 8-byte loads of CONSECUTIVE addresses across the warp (5 + 5 sectors per task instead of the 36
// one-sector accesses of the per-lane kernel), the products need neither shuffles (the warp-cooperative SIMT variant
// spent 11 per task) nor 36 accumulator registers per lane (two hold the lane's share of the 8 x 8 tile): ~30 registers,
// full occupancy, and the loads of UNROLL tasks are in flight together. The sum of a destination runs over its tasks in
// list order on UNROLL interleaved accumulators that are added in a fixed order at the end: deterministic.
__device__ __forceinline__ void dmmaAcc(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
template <int NR, int NC, int K, int UNROLL, bool HEAVY>
__device__ __forceinline__ void elimGatherDmmaBody(const DevElimPlan& p, double* data, int64_t block) {
  // blocks of up to 16 x 16 (e.g. 9-parameter cameras): MT x NT tiles of m8n8k4 per task, K <= 4 (the eliminated lump's width)
  static_assert(NR <= 16 && NC <= 16 && K <= 4, "at most 2 x 2 m8n8k4 tiles per task");
  constexpr int MT = (NR + 7) / 8, NT = (NC + 7) / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // light destinations: a warp each; heavy ones (long task lists): the eight warps of a CTA take consecutive slices of
  // the list and their partial tiles are added in warp order
  int64_t d;
  if (HEAVY) {
    d = p.heavyList[block];
  } else {
    const int64_t w = block * 8 + warp;
    if (w >= p.numLight) return;
    d = p.lightList[w];
  }
  const int idx = g * K + t;
  int tb = p.dstTaskPtr[d], te = p.dstTaskPtr[d + 1];
  // the destination's own entries are requested now, with the first tasks: one round trip less at the end
  double* dst = data + p.dstOff[d] + (int64_t)g * p.dstStride[d];
  const int64_t dstTile = 8 * (int64_t)p.dstStride[d];  // eight rows further down
  double old[MT][NT][2];
#pragma unroll
  for (int mi = 0; mi < MT; mi++)
#pragma unroll
    for (int ni = 0; ni < NT; ni++) {
      old[mi][ni][0] = old[mi][ni][1] = 0.0;
      if (!HEAVY && 8 * mi + g < NR) {
        if (8 * ni + 2 * t < NC) old[mi][ni][0] = dst[mi * dstTile + 8 * ni + 2 * t];
        if (8 * ni + 2 * t + 1 < NC) old[mi][ni][1] = dst[mi * dstTile + 8 * ni + 2 * t + 1];
      }
    }
  if (HEAVY) {
    const int chunk = (te - tb + 7) / 8;
    tb = min(te, tb + warp * chunk), te = min(te, tb + chunk);
  }
  double c[UNROLL][MT][NT][2];
#pragma unroll
  for (int u = 0; u < UNROLL; u++)
#pragma unroll
    for (int mi = 0; mi < MT; mi++)
#pragma unroll
      for (int ni = 0; ni < NT; ni++) c[u][mi][ni][0] = c[u][mi][ni][1] = 0.0;
  // the offsets of the next UNROLL tasks are fetched while the blocks of the current ones are in flight
  uint32_t oa[UNROLL], ob[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; u++) {
    oa[u] = ob[u] = 0;
    if (tb + u < te) oa[u] = __ldg(p.taskA + tb + u), ob[u] = __ldg(p.taskB + tb + u);
  }
  for (; tb < te; tb += UNROLL) {
    double a[UNROLL][MT], b[UNROLL][NT];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const bool live = tb + u < te && t < K;
#pragma unroll
      for (int mi = 0; mi < MT; mi++) a[u][mi] = live && 8 * mi + g < NR ? data[ob[u] + 8 * mi * K + idx] : 0.0;
#pragma unroll
      for (int ni = 0; ni < NT; ni++) b[u][ni] = live && 8 * ni + g < NC ? data[oa[u] + 8 * ni * K + idx] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
      if (tb + UNROLL + u < te) oa[u] = __ldg(p.taskA + tb + UNROLL + u), ob[u] = __ldg(p.taskB + tb + UNROLL + u);
    // a tile of zeros for a task beyond the end: adds nothing (the branch around an mma would only predicate it)
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
#pragma unroll
      for (int mi = 0; mi < MT; mi++)
#pragma unroll
        for (int ni = 0; ni < NT; ni++) dmmaAcc(c[u][mi][ni][0], c[u][mi][ni][1], a[u][mi], b[u][ni]);
  }
#pragma unroll
  for (int u = 1; u < UNROLL; u++)
#pragma unroll
    for (int mi = 0; mi < MT; mi++)
#pragma unroll
      for (int ni = 0; ni < NT; ni++) c[0][mi][ni][0] += c[u][mi][ni][0], c[0][mi][ni][1] += c[u][mi][ni][1];
  if (HEAVY) {
    __shared__ double red[8][MT * NT][32][2];
#pragma unroll
    for (int mi = 0; mi < MT; mi++)
#pragma unroll
      for (int ni = 0; ni < NT; ni++)
        red[warp][mi * NT + ni][lane][0] = c[0][mi][ni][0], red[warp][mi * NT + ni][lane][1] = c[0][mi][ni][1];
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int mi = 0; mi < MT; mi++)
#pragma unroll
      for (int ni = 0; ni < NT; ni++) {
        c[0][mi][ni][0] = red[0][mi * NT + ni][lane][0], c[0][mi][ni][1] = red[0][mi * NT + ni][lane][1];
#pragma unroll
        for (int w = 1; w < 8; w++)
          c[0][mi][ni][0] += red[w][mi * NT + ni][lane][0], c[0][mi][ni][1] += red[w][mi * NT + ni][lane][1];
      }
  }
#pragma unroll
  for (int mi = 0; mi < MT; mi++)
#pragma unroll
    for (int ni = 0; ni < NT; ni++)
      if (8 * mi + g < NR) {
        double* q = dst + mi * dstTile + 8 * ni + 2 * t;
        if (HEAVY) {
          if (8 * ni + 2 * t < NC) q[0] -= c[0][mi][ni][0];
          if (8 * ni + 2 * t + 1 < NC) q[1] -= c[0][mi][ni][1];
        } else {
          if (8 * ni + 2 * t < NC) q[0] = old[mi][ni][0] - c[0][mi][ni][0];
          if (8 * ni + 2 * t + 1 < NC) q[1] = old[mi][ni][1] - c[0][mi][ni][1];
        }
      }
}
// one launch: the first numHeavy CTAs take a heavy destination each (they are the long ones: first), the rest eight light
// destinations each - the two parts share the SMs (stress workload: 1.21 ms against 1.34 for two launches)
template <int NR, int NC, int K, int ULIGHT, int UHEAVY>
__global__ void __launch_bounds__(256) elim_gather_dmma_kernel(DevElimPlan p, Mats<double> mats) {
  double* data = mats.at(blockIdx.z);
  if ((int64_t)blockIdx.x < p.numHeavy) elimGatherDmmaBody<NR, NC, K, UHEAVY, true>(p, data, blockIdx.x);
  else elimGatherDmmaBody<NR, NC, K, ULIGHT, false>(p, data, (int64_t)blockIdx.x - p.numHeavy);
}
// one part per launch: with short light lists (one task in flight per warp) the light part alone needs 32 registers and
// runs with twice the warps of the merged kernel - it is bound by loads in flight (BAL: 0.65 ms for both launches
// against 0.90 merged)
template <int NR, int NC, int K, int UNROLL, bool HEAVY>
__global__ void __launch_bounds__(256) elim_gather_dmma_part_kernel(DevElimPlan p, Mats<double> mats) {
  elimGatherDmmaBody<NR, NC, K, UNROLL, HEAVY>(p, mats.at(blockIdx.z), blockIdx.x);
}

// Warp-cooperative variant of the fixed-shape gather (round 2). One WARP owns a destination (or, for the heavy ones, a
// contiguous slice of its tasks) and walks its pair tasks in order; for every task the two operand blocks are fetched
// COALESCED - lane i < NR*K reads word i of the B block, lane i < NC*K word i of the A block: two or three 32-byte sectors
// per block instead of one sector access per lane and word (the per-lane kernel above needs 36 sector accesses per
// task, ncu: 12.8 sectors per request, L1 sector throughput is its bound) - and the products are formed by exchanging
// the words with shuffles: lane (r, c0) computes the NCO = NR*NC / lanes outputs (r, c0 + j * CG) of its row, K shuffles
// for its B row + NCO * K for the A rows. The offsets of 32 tasks are read with one coalesced load per array and handed
// round by shuffle; four tasks are in flight per warp. Summation order per output: task order - deterministic, no
// atomics. Heavy destinations: the CTA's warps take consecutive slices of the task list, partial sums meet in shared
// memory and are added in warp order.
template <typename T, int NR, int NC, int K, int NCO, bool HEAVY>
__global__ void __launch_bounds__(HEAVY ? 256 : 128)
    elim_gather_warp_kernel(DevElimPlan p, Mats<T> mats, const int32_t* __restrict__ list, int64_t count) {
  static_assert(NC % NCO == 0, "outputs per lane must divide the columns");
  constexpr int CG = NC / NCO;      // column groups: lane (r, c0) owns columns c0 + j * CG
  constexpr int OL = NR * CG;       // lanes that own outputs
  static_assert(OL <= 32 && NR * K <= 32 && NC * K <= 32, "shape too large for one warp");
  constexpr int U = 4;              // tasks in flight
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int WPB = HEAVY ? 8 : 4;
  const int64_t slot = HEAVY ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * WPB + warp;
  if (slot >= count) return;
  const int64_t d = list[slot];
  T* data = mats.at(blockIdx.z);
  int tBegin = p.dstTaskPtr[d], tEnd = p.dstTaskPtr[d + 1];
  if (HEAVY) {  // this warp's slice of the tasks
    const int per = (tEnd - tBegin + WPB - 1) / WPB;
    tBegin = min(tEnd, tBegin + warp * per);
    tEnd = min(tEnd, tBegin + per);
  }
  const int r = lane / CG, c0 = lane % CG;  // meaningful for lane < OL
  T acc[NCO];
#pragma unroll
  for (int j = 0; j < NCO; j++) acc[j] = T(0);
  for (int base = tBegin; base < tEnd; base += 32) {
    const int n = min(32, tEnd - base);
    uint32_t offA = 0, offB = 0;
    if (lane < n) offA = p.taskA[base + lane], offB = p.taskB[base + lane];
    for (int i = 0; i < n; i += U) {
      T av[U], bv[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t oa = __shfl_sync(0xffffffffu, offA, (i + u) & 31), ob = __shfl_sync(0xffffffffu, offB, (i + u) & 31);
        const bool on = i + u < n;
        av[u] = (on && lane < NC * K) ? data[oa + lane] : T(0);
        bv[u] = (on && lane < NR * K) ? data[ob + lane] : T(0);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        T b[K];
#pragma unroll
        for (int q = 0; q < K; q++) b[q] = __shfl_sync(0xffffffffu, bv[u], (r * K + q) & 31);
#pragma unroll
        for (int j = 0; j < NCO; j++)
#pragma unroll
          for (int q = 0; q < K; q++) acc[j] += b[q] * __shfl_sync(0xffffffffu, av[u], ((c0 + j * CG) * K + q) & 31);
      }
    }
  }
  T* dst = data + p.dstOff[d];
  const int64_t stride = p.dstStride[d];
  if constexpr (HEAVY) {
    __shared__ T red[WPB][OL * NCO];
    if (lane < OL)
#pragma unroll
      for (int j = 0; j < NCO; j++) red[warp][lane * NCO + j] = acc[j];
    __syncthreads();
    if (warp == 0 && lane < OL) {
#pragma unroll
      for (int j = 0; j < NCO; j++) {
        T tot = T(0);
#pragma unroll
        for (int w = 0; w < WPB; w++) tot += red[w][lane * NCO + j];
        dst[r * stride + c0 + j * CG] -= tot;
      }
    }
  } else if (lane < OL) {
#pragma unroll
    for (int j = 0; j < NCO; j++) dst[r * stride + c0 + j * CG] -= acc[j];
  }
}

// Staged variant of the fixed-shape gather. Same task ownership and summation order as elim_gather_fixed_kernel (lane
// `sub` of a destination takes tasks sub, sub + LANES, ...; butterfly over the lanes; one read-modify-write of the
// target), but the operand blocks reach the lanes through shared memory: a lane reading its own two blocks with
// 8-byte loads touches 32 different cache lines per warp instruction and the L1 wavefront queue, not HBM, becomes the
// bound (ncu r01: 48 % L2 hits, IPC 0.1). Here the 32 tasks of a warp iteration are fetched COOPERATIVELY -
// consecutive lanes copy consecutive words of the same block with cp.async (2-3 lines per instruction) into a
// per-warp stage [task][word] with an odd stride (conflict-free both for the cooperative writes and for the
// per-lane reads), double buffered: the copies of iteration i + 1 are in flight while iteration i is multiplied.
template <int BYTES>
__device__ __forceinline__ void cpAsyncWord(void* smemDst, const void* src) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smemDst);
  if constexpr (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src));
}

constexpr int kStagedWarps = 4;
template <typename T, int NR, int NC, int K>
constexpr int stagedStride() {
  return ((NR + NC) * K) | 1;  // odd word stride of one task in the stage
}
template <typename T, int NR, int NC, int K>
constexpr size_t stagedSmemBytes() {
  return (size_t)kStagedWarps * 2 * 32 * stagedStride<T, NR, NC, K>() * sizeof(T);
}

template <typename T, int NR, int NC, int K, int LANES>
__global__ void __launch_bounds__(kStagedWarps * 32, 3)
    elim_gather_staged_kernel(DevElimPlan p, Mats<T> mats, const int32_t* __restrict__ list, int64_t count) {
  static_assert(LANES == 128 || (LANES <= 32 && 32 % LANES == 0), "LANES: a divisor of the warp, or the whole CTA");
  constexpr int WA = NC * K, WB = NR * K, WT = WA + WB;  // words of the A block, the B block, one task
  constexpr int STRIDE = ((NR + NC) * K) | 1;  // == stagedStride<T, NR, NC, K>()
  constexpr uint32_t kNone = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* stage = reinterpret_cast<T*>(smemRaw) + (size_t)warp * 2 * 32 * STRIDE;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = gid / LANES;
  const int sub = (int)(gid % LANES);
  const bool live = slot < count;
  const int64_t d = live ? list[slot] : 0;
  T* data = mats.at(blockIdx.z);
  T acc[NR * NC];
#pragma unroll
  for (int e = 0; e < NR * NC; e++) acc[e] = T(0);

  const int tEnd = live ? p.dstTaskPtr[d + 1] : 0;
  int t = live ? p.dstTaskPtr[d] + sub : 0;
  // offsets of the task whose copies are issued next (kNone: this lane has no more tasks)
  uint32_t oa = kNone, ob = kNone;
  if (t < tEnd) oa = p.taskA[t], ob = p.taskB[t];

  auto issue = [&](int buf) {  // cooperative copies of the 32 tasks (oa, ob) held by the lanes -> stage[buf]
    T* dstBase = stage + (size_t)buf * 32 * STRIDE;
    int task = lane / WT, word = lane % WT;  // flat word index f = r * 32 + lane = task * WT + word
#pragma unroll 2
    for (int r = 0; r < WT; r++) {  // WT rounds of 32 words
      const uint32_t sa = __shfl_sync(0xffffffffu, oa, task), sb = __shfl_sync(0xffffffffu, ob, task);
      if (sa != kNone) {
        const T* src = word < WA ? data + sa + word : data + sb + (word - WA);
        cpAsyncWord<sizeof(T)>(dstBase + task * STRIDE + word, src);
      }
      word += 32;
      task += word / WT;
      word %= WT;
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  bool cur = oa != kNone;  // this lane's task of the iteration being multiplied is valid
  issue(0);
  int buf = 0;
  while (__any_sync(0xffffffffu, cur)) {
    // next iteration: fetch its offsets, put its copies in flight
    t += LANES;
    oa = ob = kNone;
    if (t < tEnd) oa = p.taskA[t], ob = p.taskB[t];
    const bool nxt = oa != kNone;
    issue(buf ^ 1);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncwarp();
    if (cur) {
      const T* __restrict__ a = stage + (size_t)buf * 32 * STRIDE + lane * STRIDE;
      const T* __restrict__ b = a + WA;
      T av[WA], bv[WB];
#pragma unroll
      for (int i = 0; i < WA; i++) av[i] = a[i];
#pragma unroll
      for (int i = 0; i < WB; i++) bv[i] = b[i];
#pragma unroll
      for (int r = 0; r < NR; r++)
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int q = 0; q < K; q++) acc[r * NC + c] += bv[r * K + q] * av[c * K + q];
    }
    __syncwarp();  // the stage is rewritten two issues from now
    cur = nxt;
    buf ^= 1;
  }
  asm volatile("cp.async.wait_group 0;\n" ::);

  constexpr int WL = LANES > 32 ? 32 : LANES;
#pragma unroll
  for (int o = 1; o < WL; o <<= 1)
#pragma unroll
    for (int e = 0; e < NR * NC; e++) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  if constexpr (LANES > 32) {
    __shared__ T red[LANES / 32][NR * NC];
#pragma unroll
    for (int e = 0; e < NR * NC; e++)
      if (lane == e % 32) red[warp][e] = acc[e];
    __syncthreads();
    if (live && threadIdx.x < NR * NC) {
      const int e = threadIdx.x;
      T tot = 0;
#pragma unroll
      for (int w = 0; w < LANES / 32; w++) tot += red[w][e];
      data[p.dstOff[d] + (e / NC) * (int64_t)p.dstStride[d] + (e % NC)] -= tot;
    }
  } else if (live) {
    T* dst = data + p.dstOff[d];
    const int64_t stride = p.dstStride[d];
#pragma unroll
    for (int e = 0; e < NR * NC; e++)
      if (e % LANES == sub) dst[(e / NC) * stride + (e % NC)] -= acc[e];
  }
}

// one CTA (one warp) per span: diagonal block of the span + the rows below it, columns of this span only
// (reference factor_spans_kernel, MatOpsCuda.cu:188-233)
template <typename T>
__global__ void __launch_bounds__(32) pseudo_factor_spans_kernel(DevSkel sk, Mats<T> mats, int64_t spanBegin) {
  const int64_t span = spanBegin + blockIdx.x;
  T* data = mats.at(blockIdx.z);
  const int lane = threadIdx.x;
  const int64_t lump = sk.spanToLump[span];
  const int64_t w = sk.lumpStart[lump + 1] - sk.lumpStart[lump];
  const int s = (int)(sk.spanStart[span + 1] - sk.spanStart[span]);
  const int64_t ord = span - sk.lumpToSpan[lump];
  const int64_t cb = sk.chainColPtr[lump], ce = sk.chainColPtr[lump + 1];
  const int64_t colOff = sk.spanOffsetInLump[span];
  T* D = data + sk.chainData[cb + ord] + colOff;
  if (lane == 0) choleskySerial(D, s, w);
  __syncwarp();
  const int64_t rowsBelow = sk.chainRowsTillEnd[ce - 1] - sk.chainRowsTillEnd[cb + ord];
  T* below = D + (int64_t)s * w;  // chains of a lump are consecutive: the next row follows the block
  for (int64_t r = lane; r < rowsBelow; r += 32) solveRowSerial(D, s, w, below + r * w);
}

__global__ void prepare_assemble_kernel(DevSkel sk, int64_t* spanToChainOffset, int64_t chainBegin, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) spanToChainOffset[sk.chainRowSpan[chainBegin + i]] = sk.chainData[chainBegin + i];
}

// index of the chain (relative to `first`) containing row `row` of the column; rowsTillEnd is inclusive-cumulative
__device__ __forceinline__ int64_t chainOfRow(const int64_t* rowsTillEnd, int64_t count, int64_t row) {
  int64_t lo = 0, hi = count - 1;  // smallest x with rowsTillEnd[x] > row
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (rowsTillEnd[mid] > row) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// target[rowSpan r][colSpan c] -= temp block for block rows r, block cols c <= r, c < numBlockCols.
// One thread per scalar of the temp panel: consecutive threads -> consecutive columns (coalesced both sides).
// (reference assemble_kernel, MatOpsCuda.cu:370-406)
template <typename T>
__global__ void __launch_bounds__(256)
    assemble_kernel(DevSkel sk, const int64_t* __restrict__ spanToChainOffset, Mats<T> mats, Work<T> temp,
                    int64_t rectRowBegin, int64_t dstStride, int64_t srcColDataOffset, int64_t srcRectWidth,
                    int64_t numBlockRows, int64_t numBlockCols, int64_t numRows) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= numRows * srcRectWidth) return;
  const int64_t i = gid / srcRectWidth, j = gid - i * srcRectWidth;
  const int64_t* rowsTillEnd = sk.chainRowsTillEnd + srcColDataOffset;
  const int64_t* toSpan = sk.chainRowSpan + srcColDataOffset;
  const int64_t r = chainOfRow(rowsTillEnd, numBlockRows, i + rectRowBegin);
  const int64_t c = chainOfRow(rowsTillEnd, numBlockCols, j + rectRowBegin);
  if (c > r) return;
  const int64_t rBegin = rowsTillEnd[r - 1] - rectRowBegin, cBegin = rowsTillEnd[c - 1] - rectRowBegin;
  T* data = mats.at(blockIdx.z);
  const T* src = temp.at(blockIdx.z);
  T* dst = data + spanToChainOffset[toSpan[r]] + sk.spanOffsetInLump[toSpan[c]] + (i - rBegin) * dstStride + (j - cBegin);
  *dst -= src[i * srcRectWidth + j];
}

// ---------------------------------------------------------------------------------------------------------
// elimination-range solves. Vectors: column-major order x nRHS, leading dimension ldc.
// (1) per lump diagonal solve: thread per (lump, rhs)
template <typename T>
__global__ void __launch_bounds__(128) elim_diag_solve_kernel(DevSkel sk, Mats<T> mats, Mats<T> vecs, int64_t ldc,
                                                              int nRHS, int64_t lumpsBegin, int64_t lumpsEnd,
                                                              bool transposed) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t lump = lumpsBegin + gid / nRHS;
  if (lump >= lumpsEnd) return;
  const int rhs = (int)(gid % nRHS);
  const T* __restrict__ data = mats.at(blockIdx.z);
  T* x = vecs.at(blockIdx.z) + (int64_t)rhs * ldc + sk.lumpStart[lump];
  const int s = (int)(sk.lumpStart[lump + 1] - sk.lumpStart[lump]);
  const T* L = data + sk.chainData[sk.chainColPtr[lump]];
  if (!transposed) {
    for (int i = 0; i < s; i++) {
      T v = x[i];
      for (int q = 0; q < i; q++) v -= L[i * s + q] * x[q];
      x[i] = v / L[i * s + i];
    }
  } else {
    for (int i = s - 1; i >= 0; i--) {
      T v = x[i];
      for (int q = i + 1; q < s; q++) v -= L[q * s + i] * x[q];
      x[i] = v / L[i * s + i];
    }
  }
}

// (2) forward: one CTA per row span below the range gathers  v[span] -= sum_chains L(span, l) * x_l
// thread = (element e of the rows x nRHS result, chain group g); groups stride over the row's chains, then a
// shared-memory reduction over the groups in fixed order.
template <typename T>
__global__ void __launch_bounds__(256) elim_gather_solveL_kernel(DevSkel sk, DevElimPlan p, Mats<T> mats,
                                                                 Mats<T> vecs, int64_t ldc, int nRHS, int E) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* red = reinterpret_cast<T*>(smemRaw);
  const int64_t rel = blockIdx.x;
  const int cBegin = p.rowPtr[rel], cEnd = p.rowPtr[rel + 1];
  if (cBegin == cEnd) return;
  const int64_t span = rel + p.spanRowBegin;
  const int64_t r0 = sk.spanStart[span];
  const int rows = (int)(sk.spanStart[span + 1] - r0);
  const T* __restrict__ data = mats.at(blockIdx.z);
  T* C = vecs.at(blockIdx.z);
  const int G = blockDim.x / E;
  const int e = threadIdx.x % E, g = threadIdx.x / E;
  const int nElems = rows * nRHS;
  // elements beyond E are handled by looping (e, e + E, ...)
  for (int e0 = 0; e0 < nElems; e0 += E) {
    const int el = e0 + e;
    T acc = 0;
    if (el < nElems && g < G) {
      const int r = el % rows, rhs = el / rows;
      for (int c = cBegin + g; c < cEnd; c += G) {
        const int k = p.rowChainK[c];
        const T* __restrict__ blk = data + p.rowChainOff[c] + r * k;
        const T* __restrict__ x = C + (int64_t)rhs * ldc + p.rowChainCol[c];
        for (int q = 0; q < k; q++) acc += blk[q] * x[q];
      }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0 && el < nElems) {
      T tot = 0;
      for (int gg = 0; gg < G; gg++) tot += red[gg * E + e];
      const int r = el % rows, rhs = el / rows;
      C[(int64_t)rhs * ldc + r0 + r] -= tot;
    }
    __syncthreads();
  }
}

// (3) backward: thread per (lump, rhs): x_l -= sum_chains L(row, l)^T v[row]  (each lump owns its output)
// FUSE_DIAG (every lump at most MAXW wide): the transposed solve with the lump's own diagonal block follows in registers
// (the same operations in the same order as elim_diag_solve_kernel, one launch and one round trip of x less)
template <typename T, bool FUSE_DIAG>
__global__ void __launch_bounds__(128) elim_gather_solveLt_kernel(DevSkel sk, Mats<T> mats, Mats<T> vecs, int64_t ldc,
                                                                  int nRHS, int64_t lumpsBegin, int64_t lumpsEnd) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t lump = lumpsBegin + gid / nRHS;
  if (lump >= lumpsEnd) return;
  const int rhs = (int)(gid % nRHS);
  const T* __restrict__ data = mats.at(blockIdx.z);
  T* C = vecs.at(blockIdx.z) + (int64_t)rhs * ldc;
  const int s = (int)(sk.lumpStart[lump + 1] - sk.lumpStart[lump]);
  const int64_t c0 = sk.lumpStart[lump];
  const int64_t first = sk.chainColPtr[lump] + (sk.lumpToSpan[lump + 1] - sk.lumpToSpan[lump]);
  const int64_t end = sk.chainColPtr[lump + 1];
  constexpr int MAXW = 12;
  if (s <= MAXW) {
    // single pass over the chains, all s outputs accumulated in registers
    T acc[MAXW];
#pragma unroll
    for (int q = 0; q < MAXW; q++) acc[q] = q < s ? C[c0 + q] : T(0);
    for (int64_t ch = first; ch < end; ch++) {
      const int64_t span = sk.chainRowSpan[ch];
      const int64_t r0 = sk.spanStart[span];
      const int rows = (int)(sk.spanStart[span + 1] - r0);
      const T* __restrict__ blk = data + sk.chainData[ch];
      for (int r = 0; r < rows; r++) {
        const T v = C[r0 + r];
#pragma unroll
        for (int q = 0; q < MAXW; q++)
          if (q < s) acc[q] -= blk[r * s + q] * v;
      }
    }
    if (FUSE_DIAG) {
      const T* __restrict__ L = data + sk.chainData[sk.chainColPtr[lump]];
#pragma unroll
      for (int i = MAXW - 1; i >= 0; i--)
        if (i < s) {
          T v = acc[i];
#pragma unroll
          for (int q = i + 1; q < MAXW; q++)
            if (q < s) v -= L[q * s + i] * acc[q];
          acc[i] = v / L[i * s + i];
        }
    }
#pragma unroll
    for (int q = 0; q < MAXW; q++)
      if (q < s) C[c0 + q] = acc[q];
    return;
  }
  for (int q = 0; q < s; q++) {
    T acc = C[c0 + q];
    for (int64_t ch = first; ch < end; ch++) {
      const int64_t span = sk.chainRowSpan[ch];
      const int64_t r0 = sk.spanStart[span];
      const int rows = (int)(sk.spanStart[span + 1] - r0);
      const T* __restrict__ blk = data + sk.chainData[ch];
      for (int r = 0; r < rows; r++) acc -= blk[r * s + q] * C[r0 + r];
    }
    C[c0 + q] = acc;
  }
}

// dense-lump solves: C[rows of chain] += tmp  /  tmp = C[rows of chain]; thread per (row of tmp, rhs)
template <typename T, bool GATHER>
__global__ void __launch_bounds__(256) assemble_vec_kernel(DevSkel sk, Work<T> tmp, int64_t chainColPtr,
                                                           int64_t numColItems, int64_t numRows, Mats<T> vecs,
                                                           int64_t ldc, int nRHS) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= numRows * nRHS) return;
  const int64_t i = gid / nRHS;
  const int rhs = (int)(gid - i * nRHS);
  const int64_t* rowsTillEnd = sk.chainRowsTillEnd + chainColPtr;
  const int64_t startRow = rowsTillEnd[-1];
  const int64_t ch = chainOfRow(rowsTillEnd, numColItems, i + startRow);
  const int64_t rowOff = rowsTillEnd[ch - 1] - startRow;
  const int64_t span = sk.chainRowSpan[chainColPtr + ch];
  T* c = vecs.at(blockIdx.z) + (int64_t)rhs * ldc + sk.spanStart[span] + (i - rowOff);
  T* t = tmp.at(blockIdx.z) + i * nRHS + rhs;
  if (GATHER) *t = *c; else *c += *t;
}

template <typename T, int S>
void launchFactorLumps(cudaStream_t st, int batch, const DevSkel& sk, Mats<T> data, int64_t b, int64_t e) {
  elim_factor_lumps_kernel<T, S><<<dim3(ceilDiv(e - b, 4), 1, batch), 128, 0, st>>>(sk, data, b, e);
}

}  // namespace

template <typename T>
void elimFactorLumps(cudaStream_t st, int batch, const DevSkel& sk, Mats<T> data, int64_t lumpsBegin, int64_t lumpsEnd,
                     int uniformLumpSize, double profBytes) {
  if (lumpsEnd <= lumpsBegin) return;
  ProfScope prof(st, KC_ELIM_FACTOR, 0, profBytes * batch);
  switch (uniformLumpSize) {
    case 1: launchFactorLumps<T, 1>(st, batch, sk, data, lumpsBegin, lumpsEnd); break;
    case 2: launchFactorLumps<T, 2>(st, batch, sk, data, lumpsBegin, lumpsEnd); break;
    case 3: launchFactorLumps<T, 3>(st, batch, sk, data, lumpsBegin, lumpsEnd); break;
    case 4: launchFactorLumps<T, 4>(st, batch, sk, data, lumpsBegin, lumpsEnd); break;
    case 5: launchFactorLumps<T, 5>(st, batch, sk, data, lumpsBegin, lumpsEnd); break;
    case 6: launchFactorLumps<T, 6>(st, batch, sk, data, lumpsBegin, lumpsEnd); break;
    default: launchFactorLumps<T, 0>(st, batch, sk, data, lumpsBegin, lumpsEnd); break;
  }
  B200_LAUNCH_CHECK();
}

template <typename T>
void elimGather(cudaStream_t st, int batch, const DevElimPlan& plan, Mats<T> data) {
  if (plan.numDst == 0) return;
  ProfScope prof(st, KC_ELIM_GATHER, plan.gatherFlops * batch, plan.gatherBytes * sizeof(T) * batch);
  auto fixed = [&](auto light, auto heavy, int lanes) {
    if (plan.numLight > 0) {
      light<<<dim3(ceilDiv(plan.numLight * lanes, 128), 1, batch), 128, 0, st>>>(plan, data, plan.lightList, plan.numLight);
      B200_LAUNCH_CHECK();
    }
    if (plan.numHeavy > 0) {
      heavy<<<dim3((unsigned)plan.numHeavy, 1, batch), 256, 0, st>>>(plan, data, plan.heavyList, plan.numHeavy);
      B200_LAUNCH_CHECK();
    }
  };
  // Measured on B200 (profiles/README.md, round 1): the staged (cooperative cp.async) kernel wins when every lane of
  // the warp has a task in every iteration - the heavy destinations (stress workload, all heavy: 2.27 -> 1.78 ms) -
  // and loses on the light list, where most destinations hold one or two tasks and the fixed cost of a cooperative
  // iteration is paid for a few active lanes (BAL: 1.15 -> 1.94 ms). Default: direct loads for the light list, staged
  // for the heavy list. BSPB200_GATHER=0: direct everywhere, 2: staged everywhere. (A third variant - cooperative
  // coalesced ld.global into registers, then the stage - was measured slower on both workloads, BAL 1.92 ms and
  // stress 3.4 ms, and was removed: profiles/README.md.)
  const char* modeEnv = getenv("BSPB200_GATHER");  // read at every call: tests and probes compare the variants
  // default: 6 for fp64 (tensor-pipe gather: BAL 1.43 -> 0.95 ms, stress 2.31 -> 1.82 ms for the whole elimination), 1
  // otherwise (per-lane gather, staged heavy list). Measured negatives, kept opt-in: 3 (warp per destination, coalesced
  // loads + shuffle exchange: BAL 1.71, stress 3.80 ms - 7x the instructions per task), 5 (per-lane with 16-byte loads:
  // BAL 2.14 ms - 188 registers, half the warps in flight: the kernel is bound by loads in flight, not by L1 sectors);
  // a quad-transposed load variant measured 2.1 ms for the BAL gather and was dropped.
  const int mode = modeEnv ? atoi(modeEnv) : (std::is_same<T, double>::value ? 6 : 1);
  auto fixedStaged = [&](auto light, auto heavy, auto lightDirect, int lanes, size_t smem) {
    ensureDynSmem((const void*)light, smem);
    ensureDynSmem((const void*)heavy, smem);
    constexpr int NT = kStagedWarps * 32;
    if (plan.numLight > 0) {
      if (mode == 2)
        light<<<dim3(ceilDiv(plan.numLight * lanes, NT), 1, batch), NT, smem, st>>>(plan, data, plan.lightList, plan.numLight);
      else
        lightDirect<<<dim3(ceilDiv(plan.numLight * lanes, 128), 1, batch), 128, 0, st>>>(plan, data, plan.lightList, plan.numLight);
      B200_LAUNCH_CHECK();
    }
    if (plan.numHeavy > 0) {
      heavy<<<dim3((unsigned)plan.numHeavy, 1, batch), NT, smem, st>>>(plan, data, plan.heavyList, plan.numHeavy);
      B200_LAUNCH_CHECK();
    }
  };
  // BSPB200_GATHER=3 (opt-in): the warp-cooperative kernel (coalesced block loads + shuffle exchange)
  auto warpCoop = [&](auto light, auto heavy) {
    if (plan.numLight > 0) {
      light<<<dim3(ceilDiv(plan.numLight, 4), 1, batch), 128, 0, st>>>(plan, data, plan.lightList, plan.numLight);
      B200_LAUNCH_CHECK();
    }
    if (plan.numHeavy > 0) {
      heavy<<<dim3((unsigned)plan.numHeavy, 1, batch), 256, 0, st>>>(plan, data, plan.heavyList, plan.numHeavy);
      B200_LAUNCH_CHECK();
    }
  };
  if (mode == 3) {
    if (plan.uniRows == 6 && plan.uniCols == 6 && plan.uniK == 3)
      return warpCoop(elim_gather_warp_kernel<T, 6, 6, 3, 2, false>, elim_gather_warp_kernel<T, 6, 6, 3, 2, true>);
    if (plan.uniRows == 3 && plan.uniCols == 3 && plan.uniK == 3)
      return warpCoop(elim_gather_warp_kernel<T, 3, 3, 3, 1, false>, elim_gather_warp_kernel<T, 3, 3, 3, 1, true>);
  }
  // BSPB200_GATHER=6: the tensor-pipe gather (fp64, uniform shapes that fit one m8n8k4 tile)
  if constexpr (std::is_same<T, double>::value) {
    if (mode == 6 && plan.numLight + plan.numHeavy > 0) {
      auto launch = [&](auto kernel) {
        kernel<<<dim3((unsigned)(plan.numHeavy + ceilDiv(plan.numLight, 8)), 1, batch), 256, 0, st>>>(plan, data);
        B200_LAUNCH_CHECK();
      };
      // tasks in flight per warp: four for the heavy lists and for light lists of medium-sized destinations, ONE when the
      // typical light destination holds a handful of tasks (an unrolled round costs its DMMAs and predicated loads
      // whether or not the tasks exist). Measured, BAL (18 tasks per light destination on average): gather 0.65 ms
      // with 1, 0.69 with 2, 0.75 with 3, 0.81 with 4; stress (all destinations long): 1.39 / 1.39 / 1.36 / 1.32 ms.
      const bool shortLists = plan.numLight > 0 && plan.lightTasks < 32 * plan.numLight;
      auto launchParts = [&](auto light, auto heavy) {
        if (plan.numHeavy > 0) {
          heavy<<<dim3((unsigned)plan.numHeavy, 1, batch), 256, 0, st>>>(plan, data);
          B200_LAUNCH_CHECK();
        }
        if (plan.numLight > 0) {
          light<<<dim3((unsigned)ceilDiv(plan.numLight, 8), 1, batch), 256, 0, st>>>(plan, data);
          B200_LAUNCH_CHECK();
        }
      };
      // (two tasks in flight inside 40 registers - 6 CTAs per SM, 96 tasks in flight against 64 - measured 0.79 vs 0.70 ms)
      if (plan.uniRows == 6 && plan.uniCols == 6 && plan.uniK == 3)
        return shortLists ? launchParts(elim_gather_dmma_part_kernel<6, 6, 3, 1, false>, elim_gather_dmma_part_kernel<6, 6, 3, 4, true>)
                          : launch(elim_gather_dmma_kernel<6, 6, 3, 4, 4>);
      if (plan.uniRows == 3 && plan.uniCols == 3 && plan.uniK == 3)
        return shortLists ? launchParts(elim_gather_dmma_part_kernel<3, 3, 3, 1, false>, elim_gather_dmma_part_kernel<3, 3, 3, 4, true>)
                          : launch(elim_gather_dmma_kernel<3, 3, 3, 4, 4>);
      // 9-parameter cameras (pose + intrinsics: the camera model of the BAL data sets) with 3-d points: 2 x 2 tiles per task
      if (plan.uniRows == 9 && plan.uniCols == 9 && plan.uniK == 3)
        return shortLists ? launchParts(elim_gather_dmma_part_kernel<9, 9, 3, 1, false>, elim_gather_dmma_part_kernel<9, 9, 3, 2, true>)
                          : launch(elim_gather_dmma_kernel<9, 9, 3, 2, 2>);
    }
  }
  // BSPB200_GATHER=5: 16-byte operand loads for the light list (fp64, 6x6x3), staged kernel for the heavy list
  if constexpr (std::is_same<T, double>::value) {
    if (mode == 5 && plan.uniRows == 6 && plan.uniCols == 6 && plan.uniK == 3) {
      if (plan.numLight > 0) {
        elim_gather_fixed_vec_kernel<6, 6, 3, 8><<<dim3(ceilDiv(plan.numLight * 8, 128), 1, batch), 128, 0, st>>>(
            plan, data, plan.lightList, plan.numLight);
        B200_LAUNCH_CHECK();
      }
      if (plan.numHeavy > 0) {
        auto heavy = elim_gather_staged_kernel<T, 6, 6, 3, 128>;
        const size_t smem = stagedSmemBytes<T, 6, 6, 3>();
        ensureDynSmem((const void*)heavy, smem);
        heavy<<<dim3((unsigned)plan.numHeavy, 1, batch), kStagedWarps * 32, smem, st>>>(plan, data, plan.heavyList, plan.numHeavy);
        B200_LAUNCH_CHECK();
      }
      return;
    }
  }
  const bool staged = mode != 0;
  if (plan.uniRows == 6 && plan.uniCols == 6 && plan.uniK == 3) {
    if (staged)
      return fixedStaged(elim_gather_staged_kernel<T, 6, 6, 3, 8>, elim_gather_staged_kernel<T, 6, 6, 3, 128>,
                         elim_gather_fixed_kernel<T, 6, 6, 3, 8>, 8, stagedSmemBytes<T, 6, 6, 3>());
    return fixed(elim_gather_fixed_kernel<T, 6, 6, 3, 8>, elim_gather_fixed_kernel<T, 6, 6, 3, 256>, 8);
  }
  if (plan.uniRows == 3 && plan.uniCols == 3 && plan.uniK == 3) {
    if (staged)
      return fixedStaged(elim_gather_staged_kernel<T, 3, 3, 3, 4>, elim_gather_staged_kernel<T, 3, 3, 3, 128>,
                         elim_gather_fixed_kernel<T, 3, 3, 3, 4>, 4, stagedSmemBytes<T, 3, 3, 3>());
    return fixed(elim_gather_fixed_kernel<T, 3, 3, 3, 4>, elim_gather_fixed_kernel<T, 3, 3, 3, 256>, 4);
  }
  if (plan.uniRows == 6 && plan.uniCols == 6 && plan.uniK == 6) {
    if (staged)
      return fixedStaged(elim_gather_staged_kernel<T, 6, 6, 6, 8>, elim_gather_staged_kernel<T, 6, 6, 6, 128>,
                         elim_gather_fixed_kernel<T, 6, 6, 6, 8>, 8, stagedSmemBytes<T, 6, 6, 6>());
    return fixed(elim_gather_fixed_kernel<T, 6, 6, 6, 8>, elim_gather_fixed_kernel<T, 6, 6, 6, 256>, 8);
  }
  int E = std::min(plan.maxDstElems, 256);
  int64_t threads = plan.numDst * E;
  elim_gather_kernel<T><<<dim3(ceilDiv(threads, 256), 1, batch), 256, 0, st>>>(plan, data, E);
  B200_LAUNCH_CHECK();
}

template <typename T>
void pseudoFactorSpans(cudaStream_t st, int batch, const DevSkel& sk, Mats<T> data, int64_t spanBegin, int64_t spanEnd) {
  if (spanEnd <= spanBegin) return;
  pseudo_factor_spans_kernel<T><<<dim3((unsigned)(spanEnd - spanBegin), 1, batch), 32, 0, st>>>(sk, data, spanBegin);
  B200_LAUNCH_CHECK();
}

void prepareAssemble(cudaStream_t st, const DevSkel& sk, int64_t* spanToChainOffset, int64_t chainBegin,
                     int64_t numChains) {
  if (numChains <= 0) return;
  prepare_assemble_kernel<<<ceilDiv(numChains, 128), 128, 0, st>>>(sk, spanToChainOffset, chainBegin, numChains);
  B200_LAUNCH_CHECK();
}

template <typename T>
void assemble(cudaStream_t st, int batch, const DevSkel& sk, const int64_t* spanToChainOffset, Mats<T> data,
              Work<T> temp, int64_t rectRowBegin, int64_t dstStride, int64_t srcColDataOffset, int64_t srcRectWidth,
              int64_t numBlockRows, int64_t numBlockCols, int64_t numRows) {
  int64_t threads = numRows * srcRectWidth;
  if (threads <= 0) return;
  ProfScope prof(st, KC_ASSEMBLE, 0, 3.0 * threads * sizeof(T) * batch);
  assemble_kernel<T><<<dim3(ceilDiv(threads, 256), 1, batch), 256, 0, st>>>(
      sk, spanToChainOffset, data, temp, rectRowBegin, dstStride, srcColDataOffset, srcRectWidth, numBlockRows,
      numBlockCols, numRows);
  B200_LAUNCH_CHECK();
}

template <typename T>
void elimSolveL(cudaStream_t st, int batch, const DevSkel& sk, const DevElimPlan& plan, Mats<T> data, Mats<T> C,
                int64_t ldc, int nRHS) {
  int64_t n = (plan.lumpsEnd - plan.lumpsBegin) * nRHS;
  if (n <= 0) return;
  ProfScope prof(st, KC_SOLVE_ELIM, 0, plan.factorEntries * sizeof(T) * batch);
  elim_diag_solve_kernel<T><<<dim3(ceilDiv(n, 128), 1, batch), 128, 0, st>>>(sk, data, C, ldc, nRHS, plan.lumpsBegin,
                                                                             plan.lumpsEnd, false);
  B200_LAUNCH_CHECK();
  if (plan.numRowSpans <= 0) return;
  int elems = plan.maxRowSpanSize * nRHS;
  int E = 1;
  while (E < elems && E < 32) E *= 2;
  elim_gather_solveL_kernel<T><<<dim3((unsigned)plan.numRowSpans, 1, batch), 256, 256 * sizeof(T), st>>>(
      sk, plan, data, C, ldc, nRHS, E);
  B200_LAUNCH_CHECK();
}

template <typename T>
void elimSolveLt(cudaStream_t st, int batch, const DevSkel& sk, const DevElimPlan& plan, Mats<T> data, Mats<T> C,
                 int64_t ldc, int nRHS) {
  int64_t n = (plan.lumpsEnd - plan.lumpsBegin) * nRHS;
  if (n <= 0) return;
  ProfScope prof(st, KC_SOLVE_ELIM, 0, plan.factorEntries * sizeof(T) * batch);
  // (a warp-per-lump variant with coalesced row reads and the diagonal solve fused in measured slower on the BAL-shaped
  // problem - 0.52 vs 0.46 ms for both directions - and was removed)
  if (plan.uniformLumpSize >= 1 && plan.uniformLumpSize <= 12) {  // MAXW of the kernel: diagonal solve fused
    elim_gather_solveLt_kernel<T, true><<<dim3(ceilDiv(n, 128), 1, batch), 128, 0, st>>>(sk, data, C, ldc, nRHS,
                                                                                         plan.lumpsBegin, plan.lumpsEnd);
    B200_LAUNCH_CHECK();
    return;
  }
  elim_gather_solveLt_kernel<T, false><<<dim3(ceilDiv(n, 128), 1, batch), 128, 0, st>>>(sk, data, C, ldc, nRHS,
                                                                                        plan.lumpsBegin, plan.lumpsEnd);
  B200_LAUNCH_CHECK();
  elim_diag_solve_kernel<T><<<dim3(ceilDiv(n, 128), 1, batch), 128, 0, st>>>(sk, data, C, ldc, nRHS, plan.lumpsBegin,
                                                                             plan.lumpsEnd, true);
  B200_LAUNCH_CHECK();
}


// ---------------------------------------------------------------------------------------------------------
// out += alpha * A * in over the columns of a sparse-elimination range (addMvFrom; A symmetric, lower blocks stored).
// (1) column pass, thread per (lump, rhs): out_l += alpha (D_l in_l + sum over the lump's chains of B^T in[rows]) - each
//     lump owns its output; (2) row pass, CTA per row span below the range: out[span] += alpha sum over the span's chains
//     of B in_l - the row view of the elimination plan, fixed order, no atomics.
template <typename T>
__global__ void __launch_bounds__(128) elim_mv_cols_kernel(DevSkel sk, Mats<T> mats, Mats<T> inV, int64_t ldi, Mats<T> outV,
                                                           int64_t ldo, int nRHS, int64_t lumpsBegin, int64_t lumpsEnd,
                                                           T alpha) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t lump = lumpsBegin + gid / nRHS;
  if (lump >= lumpsEnd) return;
  const int rhs = (int)(gid % nRHS);
  const T* __restrict__ data = mats.at(blockIdx.z);
  const T* __restrict__ X = inV.at(blockIdx.z) + (int64_t)rhs * ldi;
  T* Y = outV.at(blockIdx.z) + (int64_t)rhs * ldo;
  const int s = (int)(sk.lumpStart[lump + 1] - sk.lumpStart[lump]);
  const int64_t c0 = sk.lumpStart[lump];
  const T* __restrict__ D = data + sk.chainData[sk.chainColPtr[lump]];
  const int64_t first = sk.chainColPtr[lump] + (sk.lumpToSpan[lump + 1] - sk.lumpToSpan[lump]);
  const int64_t end = sk.chainColPtr[lump + 1];
  for (int q = 0; q < s; q++) {
    T acc = 0;
    for (int j = 0; j < s; j++) acc += (j <= q ? D[q * s + j] : D[j * s + q]) * X[c0 + j];
    for (int64_t ch = first; ch < end; ch++) {
      const int64_t span = sk.chainRowSpan[ch];
      const int64_t r0 = sk.spanStart[span];
      const int rows = (int)(sk.spanStart[span + 1] - r0);
      const T* __restrict__ blk = data + sk.chainData[ch];
      for (int r = 0; r < rows; r++) acc += blk[r * s + q] * X[r0 + r];
    }
    Y[c0 + q] += alpha * acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) elim_mv_rows_kernel(DevSkel sk, DevElimPlan p, Mats<T> mats, Mats<T> inV, int64_t ldi,
                                                           Mats<T> outV, int64_t ldo, int nRHS, int E, T alpha) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T* red = reinterpret_cast<T*>(smemRaw);
  const int64_t rel = blockIdx.x;
  const int cBegin = p.rowPtr[rel], cEnd = p.rowPtr[rel + 1];
  if (cBegin == cEnd) return;
  const int64_t span = rel + p.spanRowBegin;
  const int64_t r0 = sk.spanStart[span];
  const int rows = (int)(sk.spanStart[span + 1] - r0);
  const T* __restrict__ data = mats.at(blockIdx.z);
  const T* __restrict__ X = inV.at(blockIdx.z);
  T* Y = outV.at(blockIdx.z);
  const int G = blockDim.x / E;
  const int e = threadIdx.x % E, g = threadIdx.x / E;
  const int nElems = rows * nRHS;
  for (int e0 = 0; e0 < nElems; e0 += E) {
    const int el = e0 + e;
    T acc = 0;
    if (el < nElems && g < G) {
      const int r = el % rows, rhs = el / rows;
      for (int c = cBegin + g; c < cEnd; c += G) {
        const int k = p.rowChainK[c];
        const T* __restrict__ blk = data + p.rowChainOff[c] + r * k;
        const T* __restrict__ x = X + (int64_t)rhs * ldi + p.rowChainCol[c];
        for (int q = 0; q < k; q++) acc += blk[q] * x[q];
      }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0 && el < nElems) {
      T tot = 0;
      for (int gg = 0; gg < G; gg++) tot += red[gg * E + e];
      const int r = el % rows, rhs = el / rows;
      Y[(int64_t)rhs * ldo + r0 + r] += alpha * tot;
    }
    __syncthreads();
  }
}

template <typename T>
void elimMV(cudaStream_t st, int batch, const DevSkel& sk, const DevElimPlan& plan, Mats<T> data, Mats<T> in, int64_t ldi,
            Mats<T> out, int64_t ldo, int nRHS, T alpha) {
  const int64_t n = (plan.lumpsEnd - plan.lumpsBegin) * nRHS;
  if (n <= 0) return;
  ProfScope prof(st, KC_SOLVE_ELIM, 0, plan.factorEntries * sizeof(T) * batch * 2);
  // the row pass first: it reads in[range] and writes out[below]; the column pass writes out[range] - with in == out
  // aliasing excluded by the interface (addMvFrom takes distinct vectors) the order does not matter
  if (plan.numRowSpans > 0) {
    int elems = plan.maxRowSpanSize * nRHS;
    int E = 1;
    while (E < elems && E < 32) E *= 2;
    elim_mv_rows_kernel<T><<<dim3((unsigned)plan.numRowSpans, 1, batch), 256, 256 * sizeof(T), st>>>(sk, plan, data, in, ldi, out,
                                                                                                     ldo, nRHS, E, alpha);
    B200_LAUNCH_CHECK();
  }
  elim_mv_cols_kernel<T><<<dim3(ceilDiv(n, 128), 1, batch), 128, 0, st>>>(sk, data, in, ldi, out, ldo, nRHS, plan.lumpsBegin,
                                                                          plan.lumpsEnd, alpha);
  B200_LAUNCH_CHECK();
}

template <typename T>
void assembleVec(cudaStream_t st, int batch, const DevSkel& sk, Work<T> tmp, int64_t chainColPtr, int64_t numColItems,
                 int64_t numRows, Mats<T> C, int64_t ldc, int nRHS) {
  int64_t n = numRows * nRHS;
  if (n <= 0) return;
  assemble_vec_kernel<T, false><<<dim3(ceilDiv(n, 256), 1, batch), 256, 0, st>>>(sk, tmp, chainColPtr, numColItems,
                                                                                 numRows, C, ldc, nRHS);
  B200_LAUNCH_CHECK();
}

template <typename T>
void assembleVecT(cudaStream_t st, int batch, const DevSkel& sk, Work<T> tmp, int64_t chainColPtr, int64_t numColItems,
                  int64_t numRows, Mats<T> C, int64_t ldc, int nRHS) {
  int64_t n = numRows * nRHS;
  if (n <= 0) return;
  assemble_vec_kernel<T, true><<<dim3(ceilDiv(n, 256), 1, batch), 256, 0, st>>>(sk, tmp, chainColPtr, numColItems,
                                                                                numRows, C, ldc, nRHS);
  B200_LAUNCH_CHECK();
}

#define B200_INSTANTIATE_SPARSE(T)                                                                                      \
  template void elimFactorLumps<T>(cudaStream_t, int, const DevSkel&, Mats<T>, int64_t, int64_t, int, double);                 \
  template void elimGather<T>(cudaStream_t, int, const DevElimPlan&, Mats<T>);                                          \
  template void pseudoFactorSpans<T>(cudaStream_t, int, const DevSkel&, Mats<T>, int64_t, int64_t);                    \
  template void assemble<T>(cudaStream_t, int, const DevSkel&, const int64_t*, Mats<T>, Work<T>, int64_t, int64_t,     \
                            int64_t, int64_t, int64_t, int64_t, int64_t);                                               \
  template void elimSolveL<T>(cudaStream_t, int, const DevSkel&, const DevElimPlan&, Mats<T>, Mats<T>, int64_t, int);  \
  template void elimSolveLt<T>(cudaStream_t, int, const DevSkel&, const DevElimPlan&, Mats<T>, Mats<T>, int64_t, int); \
  template void elimMV<T>(cudaStream_t, int, const DevSkel&, const DevElimPlan&, Mats<T>, Mats<T>, int64_t, Mats<T>,   \
                          int64_t, int, T);                                                                            \
  template void assembleVec<T>(cudaStream_t, int, const DevSkel&, Work<T>, int64_t, int64_t, int64_t, Mats<T>,         \
                               int64_t, int);                                                                           \
  template void assembleVecT<T>(cudaStream_t, int, const DevSkel&, Work<T>, int64_t, int64_t, int64_t, Mats<T>,        \
                                int64_t, int);
B200_INSTANTIATE_SPARSE(double)
B200_INSTANTIATE_SPARSE(float)


// ---------------------------------------------------------------------------------------------------------
// "Fragmented" whole-range ops (reference MatOps.h:168-183, CPU implementation MatOpsFast.cpp:613-1018; the reference's
// CUDA backend has none): every lump is a single span, nRHS = 1, so the matrix is a block-CSC of small blocks and the
// dense-lump machinery (one launch per lump) is pure overhead. Here: one warp per span, lane i owns entry i of the span
// (spans of at most 32 scalars), every block read once per use with the lanes walking down a block column, a span's
// row blocks gathered in a fixed order (no atomics: deterministic, unlike a scatter), triangular solves scheduled by
// LEVELS of the block dependency graph (all spans of a level are independent; one launch per level).
template <typename T>
__global__ void __launch_bounds__(128) frag_mv_kernel(FragDev f, DevSkel sk, Mats<T> data, Mats<T> X, Mats<T> Y,
                                                      int64_t spanBegin, int64_t spanEnd, T alpha) {
  const int lane = threadIdx.x & 31;
  const int64_t s = spanBegin + (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s >= sk.numSpans) return;
  const T* __restrict__ d = data.at(blockIdx.z);
  const T* __restrict__ x = X.at(blockIdx.z);
  T* y = Y.at(blockIdx.z);
  const int64_t s0 = sk.spanStart[s];
  const int sn = (int)(sk.spanStart[s + 1] - s0);
  T acc = T(0);
  if (s < spanEnd && lane < sn) {
    const int64_t cp = sk.chainColPtr[s], ce = sk.chainColPtr[s + 1];
    const T* D = d + sk.chainData[cp];  // sn x sn, lower triangle meaningful: symmetric product
    for (int j = 0; j <= lane; j++) acc += D[lane * sn + j] * x[s0 + j];
    for (int j = lane + 1; j < sn; j++) acc += D[j * sn + lane] * x[s0 + j];
    for (int64_t p = cp + 1; p < ce; p++) {  // blocks below in this column: y_s += B^T x_r
      const int64_t r = sk.chainRowSpan[p], r0 = sk.spanStart[r];
      const int rn = (int)(sk.spanStart[r + 1] - r0);
      const T* B = d + sk.chainData[p];  // rn x sn
      for (int i = 0; i < rn; i++) acc += B[i * sn + lane] * x[r0 + i];
    }
  }
  if (lane < sn) {
    for (int32_t q = f.rowPtr[s]; q < f.rowPtr[s + 1]; q++) {  // blocks in this row: y_s += B x_c
      const int32_t c = f.rowCol[q];
      if (c < spanBegin || c >= spanEnd) continue;
      const int64_t c0 = sk.spanStart[c];
      const int cn = (int)(sk.spanStart[c + 1] - c0);
      const T* B = d + f.rowOff[q];  // sn x cn
      for (int j = 0; j < cn; j++) acc += B[lane * cn + j] * x[c0 + j];
    }
    y[s0 + lane] += alpha * acc;
  }
}

// forward: spans list[0 .. count) (one level, or the rows behind the range when `diag` is false)
template <typename T>
__global__ void __launch_bounds__(128) frag_solve_l_kernel(FragDev f, DevSkel sk, Mats<T> data, Mats<T> Y,
                                                           const int32_t* __restrict__ list, int count, int64_t spanBegin,
                                                           int64_t spanEnd, bool diag) {
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (idx >= count) return;
  const int64_t s = list[idx];
  const T* __restrict__ d = data.at(blockIdx.z);
  T* y = Y.at(blockIdx.z);
  const int64_t s0 = sk.spanStart[s];
  const int sn = (int)(sk.spanStart[s + 1] - s0);
  T acc = lane < sn ? y[s0 + lane] : T(0);
  if (lane < sn)
    for (int32_t q = f.rowPtr[s]; q < f.rowPtr[s + 1]; q++) {
      const int32_t c = f.rowCol[q];
      if (c < spanBegin || c >= spanEnd) continue;
      const int64_t c0 = sk.spanStart[c];
      const int cn = (int)(sk.spanStart[c + 1] - c0);
      const T* B = d + f.rowOff[q];
      for (int j = 0; j < cn; j++) acc -= B[lane * cn + j] * y[c0 + j];
    }
  if (diag) {
    const T* D = d + sk.chainData[sk.chainColPtr[s]];
    for (int j = 0; j < sn; j++) {
      const T xj = __shfl_sync(0xffffffffu, acc, j) / D[j * sn + j];
      if (lane == j) acc = xj;
      else if (lane > j && lane < sn) acc -= D[lane * sn + j] * xj;
    }
  }
  if (lane < sn) y[s0 + lane] = acc;
}

// backward: one level of spans inside the range
template <typename T>
__global__ void __launch_bounds__(128) frag_solve_lt_kernel(DevSkel sk, Mats<T> data, Mats<T> Y,
                                                            const int32_t* __restrict__ list, int count) {
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (idx >= count) return;
  const int64_t s = list[idx];
  const T* __restrict__ d = data.at(blockIdx.z);
  T* y = Y.at(blockIdx.z);
  const int64_t s0 = sk.spanStart[s];
  const int sn = (int)(sk.spanStart[s + 1] - s0);
  const int64_t cp = sk.chainColPtr[s], ce = sk.chainColPtr[s + 1];
  T acc = lane < sn ? y[s0 + lane] : T(0);
  if (lane < sn)
    for (int64_t p = cp + 1; p < ce; p++) {
      const int64_t r = sk.chainRowSpan[p], r0 = sk.spanStart[r];
      const int rn = (int)(sk.spanStart[r + 1] - r0);
      const T* B = d + sk.chainData[p];
      for (int i = 0; i < rn; i++) acc -= B[i * sn + lane] * y[r0 + i];
    }
  const T* D = d + sk.chainData[cp];
  for (int j = sn - 1; j >= 0; j--) {
    const T xj = __shfl_sync(0xffffffffu, acc, j) / D[j * sn + j];
    if (lane == j) acc = xj;
    else if (lane < j) acc -= D[j * sn + lane] * xj;
  }
  if (lane < sn) y[s0 + lane] = acc;
}

template <typename T>
void fragMV(cudaStream_t st, int batch, const FragDev& f, const DevSkel& sk, Mats<T> data, Mats<T> x, Mats<T> y,
            int64_t spanBegin, int64_t spanEnd, T alpha, double bytes) {
  const int64_t n = sk.numSpans - spanBegin;
  if (n <= 0) return;
  ProfScope prof(st, KC_SOLVE_DENSE, 0, bytes * batch);
  frag_mv_kernel<T><<<dim3(ceilDiv(n, 4), 1, batch), 128, 0, st>>>(f, sk, data, x, y, spanBegin, spanEnd, alpha);
  B200_LAUNCH_CHECK();
}
template <typename T>
void fragSolveLLevel(cudaStream_t st, int batch, const FragDev& f, const DevSkel& sk, Mats<T> data, Mats<T> y,
                     const int32_t* list, int count, int64_t spanBegin, int64_t spanEnd, bool diag) {
  if (count <= 0) return;
  frag_solve_l_kernel<T><<<dim3(ceilDiv(count, 4), 1, batch), 128, 0, st>>>(f, sk, data, y, list, count, spanBegin,
                                                                            spanEnd, diag);
  B200_LAUNCH_CHECK();
}
template <typename T>
void fragSolveLtLevel(cudaStream_t st, int batch, const DevSkel& sk, Mats<T> data, Mats<T> y, const int32_t* list,
                      int count) {
  if (count <= 0) return;
  frag_solve_lt_kernel<T><<<dim3(ceilDiv(count, 4), 1, batch), 128, 0, st>>>(sk, data, y, list, count);
  B200_LAUNCH_CHECK();
}
#define B200_INST_FRAG(T)                                                                                              \
  template void fragMV<T>(cudaStream_t, int, const FragDev&, const DevSkel&, Mats<T>, Mats<T>, Mats<T>, int64_t,       \
                          int64_t, T, double);                                                                         \
  template void fragSolveLLevel<T>(cudaStream_t, int, const FragDev&, const DevSkel&, Mats<T>, Mats<T>,                \
                                   const int32_t*, int, int64_t, int64_t, bool);                                       \
  template void fragSolveLtLevel<T>(cudaStream_t, int, const DevSkel&, Mats<T>, Mats<T>, const int32_t*, int);
B200_INST_FRAG(double)
B200_INST_FRAG(float)

}  // namespace b200
}  // namespace BaSpaCho
