// Wavefront update kernel (sm_100a): for one tree level, every (small target lump, 64-row tile) work item pulls the
// contributions of ALL its source boards straight into the target:
//     target[rowMap(i), rowMap(j)] -= sum_q S[i][q] * S[j][q]     i over the source rows mapping into the tile,
//                                                                  j over the source rows of the board (-> target columns)
// The source rows are gathered into shared memory AT THEIR TARGET POSITIONS (rows the source does not touch stay
// zero), so the product is a plain dense 64 x 96 x k contraction: fp64 on DMMA m8n8k4 tiles with the accumulators in
// registers across all sources, one read-modify-write of the target tile at the end. Replaces
// saveSyrkGemm + prepareAssemble + assemble of the reference (MatOpsCuda.cu:471-498, 568-590), with a fixed summation
// order (sources ascending) -> deterministic, no temp buffer.
#include "B200Kernels.h"
#include "B200Wave.h"

namespace BaSpaCho {
namespace b200 {
namespace {

constexpr int kTR = WavePlan::kTileRows;      // 64 tile rows
constexpr int kTC = WavePlan::kMaxSmallWidth;  // 96 target columns (max)
constexpr int kBK = 16, kLD = kBK + 4;         // k chunk, padded smem stride (conflict-free fragment loads)
constexpr int kThreads = 256;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// first index in map[0..n) with map[idx] >= v
__device__ __forceinline__ int lowerBound(const int32_t* __restrict__ map, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (map[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    wave_update_kernel(Operand<T> dataOp, const WaveTile* __restrict__ tiles, const WaveTarget* __restrict__ targets,
                       const WaveSource* __restrict__ sources, const int32_t* __restrict__ rowMap) {
  __shared__ __align__(16) T Bs[kTR * kLD];
  __shared__ __align__(16) T As[kTC * kLD];
  __shared__ int srcOfRow[kTR];  // source row feeding tile row r (-1: none)
  __shared__ int srcOfCol[kTC];  // source row feeding target column c (-1: none)
  T* data = dataOp.at(blockIdx.z);
  const WaveTile tile = tiles[blockIdx.x];
  const WaveTarget tg = targets[tile.target];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int w = tg.width;
  const int tileRows = min(kTR, tg.totalRows - tile.row0);

  // warp tile: 4 (rows) x 2 (cols) warps -> 16 x 48 per warp = 2 x 6 DMMA tiles
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 48;
  const int g = lane >> 2, t = lane & 3;
  T acc[2][6][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 6; j++) acc[i][j][0] = acc[i][j][1] = T(0);

  for (int s = tg.srcBegin; s < tg.srcEnd; s++) {
    const WaveSource src = sources[s];
    const int32_t* map = rowMap + src.mapBegin;
    // which source rows land in this tile / in the target's columns (the row map is strictly increasing)
    int hit = 0;
    if (tid < kTR) {
      int idx = -1;
      if (tid < tileRows) {
        int p = lowerBound(map, src.rows, tile.row0 + tid);
        if (p < src.rows && map[p] == tile.row0 + tid) idx = p;
      }
      srcOfRow[tid] = idx;
      hit = idx >= 0;
    } else if (tid < kTR + kTC) {
      const int c = tid - kTR;
      int idx = -1;
      if (c < w) {
        int p = lowerBound(map, src.rows, c);
        if (p < src.rows && map[p] == c) idx = p;
      }
      srcOfCol[c] = idx;
    }
    if (!__syncthreads_or(hit)) continue;  // this source does not touch the tile
    const T* __restrict__ S = data + src.dataOff;
    const int k = src.k;
    for (int k0 = 0; k0 < k; k0 += kBK) {
      // gather the k-chunk of the source rows to their target positions (zero where there is no source row)
      for (int i = tid; i < kTR * kBK; i += kThreads) {
        const int r = i / kBK, kk = i % kBK;
        const int sr = srcOfRow[r];
        Bs[r * kLD + kk] = (sr >= 0 && k0 + kk < k) ? S[(int64_t)sr * k + k0 + kk] : T(0);
      }
      for (int i = tid; i < kTC * kBK; i += kThreads) {
        const int c = i / kBK, kk = i % kBK;
        const int sr = srcOfCol[c];
        As[c * kLD + kk] = (sr >= 0 && k0 + kk < k) ? S[(int64_t)sr * k + k0 + kk] : T(0);
      }
      __syncthreads();
      if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int kk = 0; kk < kBK; kk += 4) {
          double af[2], bf[6];
#pragma unroll
          for (int i = 0; i < 2; i++) af[i] = Bs[(wm + i * 8 + g) * kLD + kk + t];
#pragma unroll
          for (int j = 0; j < 6; j++) bf[j] = As[(wn + j * 8 + g) * kLD + kk + t];
#pragma unroll
          for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 6; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      } else {  // fp32: SIMT on the same fragment ownership (row wm+i*8+g, cols wn+j*8+2t+{0,1})
#pragma unroll
        for (int kk = 0; kk < kBK; kk++) {
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const T a = Bs[(wm + i * 8 + g) * kLD + kk];
#pragma unroll
            for (int j = 0; j < 6; j++) {
              acc[i][j][0] += a * As[(wn + j * 8 + 2 * t) * kLD + kk];
              acc[i][j][1] += a * As[(wn + j * 8 + 2 * t + 1) * kLD + kk];
            }
          }
        }
      }
      __syncthreads();
    }
  }
  // one read-modify-write of the tile
  T* C = data + tg.dataOff + (int64_t)tile.row0 * w;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int r = wm + i * 8 + g;
    if (r >= tileRows) continue;
#pragma unroll
    for (int j = 0; j < 6; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = wn + j * 8 + 2 * t + e;
        if (c < w) C[(int64_t)r * w + c] -= acc[i][j][e];
      }
  }
}

}  // namespace

template <typename T>
void waveUpdate(cudaStream_t st, int batch, Operand<T> data, const WaveTile* tiles, int64_t count,
                const WaveTarget* targets, const WaveSource* sources, const int32_t* rowMap) {
  if (count <= 0) return;
  ProfScope prof(st, KC_ASSEMBLE, 0, 0);
  wave_update_kernel<T><<<dim3((unsigned)count, 1, batch), kThreads, 0, st>>>(data, tiles, targets, sources, rowMap);
  B200_LAUNCH_CHECK();
}
template void waveUpdate<double>(cudaStream_t, int, Operand<double>, const WaveTile*, int64_t, const WaveTarget*,
                                 const WaveSource*, const int32_t*);
template void waveUpdate<float>(cudaStream_t, int, Operand<float>, const WaveTile*, int64_t, const WaveTarget*,
                                const WaveSource*, const int32_t*);

}  // namespace b200
}  // namespace BaSpaCho
