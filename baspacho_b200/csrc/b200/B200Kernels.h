// Launch wrappers of the hand-written sm_100a kernels (definitions in DenseKernels.cu, SparseKernels.cu,
// SolveKernels.cu). Everything is row-major; `Operand` locates a matrix inside the factor data (one
// pointer or a device array of batch pointers) or inside a per-batch-item workspace.
#pragma once

#include "B200Defs.h"

namespace BaSpaCho {
namespace b200 {

template <typename T>
struct Operand {
  T* base = nullptr;         // single matrix / workspace base
  T* const* many = nullptr;  // device array of batch pointers (overrides base)
  int64_t off = 0;           // element offset added to the selected pointer
  int64_t bstride = 0;       // workspace: distance between batch items
  __host__ __device__ T* at(int b) const { return (many ? many[b] : base + (int64_t)b * bstride) + off; }
};

template <typename T>
inline Operand<T> opnd(const Mats<T>& m, int64_t off) {
  Operand<T> o;
  o.base = m.one, o.many = m.many, o.off = off, o.bstride = 0;
  return o;
}
template <typename T>
inline Operand<T> opnd(const Work<T>& w, int64_t off) {
  Operand<T> o;
  o.base = w.base, o.many = nullptr, o.off = off, o.bstride = w.stride;
  return o;
}
template <typename T>
inline Operand<T> shifted(Operand<T> o, int64_t delta) {
  o.off += delta;
  return o;
}

// ---------------------------------------------------------------------------------- dense (DenseKernels.cu)
// C(m x n) = alpha * A(m x k) * B(n x k)^T + beta * C.  lowerOnly: only entries with col <= row are
// computed/stored (tiles strictly above the diagonal are skipped). double -> DMMA tensor tiles.
template <typename T>
void gemmNT(cudaStream_t st, int batch, int64_t m, int64_t n, int64_t k, T alpha, Operand<T> A, int64_t lda,
            Operand<T> B, int64_t ldb, T beta, Operand<T> C, int64_t ldc, bool lowerOnly);

// diagnostics: what == 0 -> phase time stamps (clock64) of the last instrumented panel_kernel launch
int64_t debugRead(int what, void* out, int64_t bytes);

// max diagonal block handled by one CTA (potrfBlock / trsmBlock / trsvBlock)
template <typename T>
int maxBlockDim();

// in-place Cholesky of the n x n (n <= maxBlockDim) block, lower triangle, one CTA per batch item
template <typename T>
void potrfBlock(cudaStream_t st, int batch, int n, Operand<T> A, int64_t lda);

// X * tril(L)^T = B in place on `rows` rows (L: n x n, n <= maxBlockDim)
template <typename T>
void trsmBlock(cudaStream_t st, int batch, int n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B, int64_t ldb);

// Cholesky of the trapezoid: A is (n + rowsBelow) x n row-major (ld), top n x n = diagonal block.
// Blocked recursively on top of potrfBlock / trsmBlock / gemmNT. rowsBelow = 0 -> plain potrf.
template <typename T>
void potrfTrapezoid(cudaStream_t st, int batch, int64_t n, int64_t rowsBelow, Operand<T> A, int64_t ld);

// the same by ONE persistent tile-DAG kernel (LumpCholKernel.cu: TMA-fed DMMA tiles, device-side dependency flags, the
// chain of diagonal blocks inside dedicated CTAs); fp64, single matrix, n >= lumpCholMinWidth(), even n / ld and a
// 16-byte aligned base (TMA). Returns false when not eligible - the caller then runs the recursive schedule.
bool lumpCholesky(cudaStream_t st, int64_t n, int64_t rowsBelow, double* A, int64_t ld);
int lumpCholMinWidth();
int64_t lumpCholDebugRead(void* out, int64_t bytes);
void lumpCholSetConcurrency(int n);
int64_t lumpCholJobList(int nbc, int nbr, int seglen, int lag, int32_t* out, int64_t capJobs);

// diagonal Cholesky + triangular solve of many small lump columns in one launch (work list on the device)
struct WavePanel;
template <typename T>
void potrfTrsmPanelBatch(cudaStream_t st, int batch, Operand<T> data, const WavePanel* work, int64_t count,
                         int numLumps, double flops);

// wavefront update of the small target lumps of one level (WaveKernels.cu)
struct WaveTile;
struct WaveTarget;
struct WaveSource;
template <typename T>
void waveUpdate(cudaStream_t st, int batch, Operand<T> data, const WaveTile* tiles, int64_t count,
                const WaveTarget* targets, const WaveSource* sources, const int32_t* rowMap);

// X * tril(L)^T = B for any n (blocked): L n x n (ldl), B rows x n (ldb)
template <typename T>
void trsmAny(cudaStream_t st, int batch, int64_t n, int64_t rows, Operand<T> L, int64_t ldl, Operand<T> B, int64_t ldb);

// ---------------------------------------------------------------------------------- vectors (SolveKernels.cu)
// Vectors are column-major (rows x nRHS, leading dimension ldc).

// one 96 x 96 diagonal block whose inverse the dense solves use: where it sits in the factor data and in the scratch
struct InvBlockDesc {
  int64_t dataOff;  // element offset of the block's (0,0) entry in the factor data
  int64_t wOff;     // element offset of its 2 x 96 x 96 slot (W, W^T) in the inverse scratch
  int32_t ld, jb;   // leading dimension (= lump width) and block size (<= 96)
};
// W = L_bb^-1 for every block of the list, one launch (the blocks of all wide lumps of a factor)
template <typename T>
void invertBlockList(cudaStream_t st, int batch, const InvBlockDesc* list, int64_t count, Operand<T> data,
                     Operand<T> invScratch);

// synchronisation state of the chained triangular solve (trsv_chain_kernel): device flags + arrival ticket, and the
// host-side running epoch / ticket values of the launches issued so far (one stream, launches in order)
struct ChainSync {
  unsigned* flags = nullptr;   // [batch][flagsPerItem], zeroed once (epoch 0 is never used)
  unsigned* ticket = nullptr;  // never reset
  unsigned epoch = 0, ticketBase = 0;
  int flagsPerItem = 0;
};

// tril(L) X = C (transposed=false) or tril(L)^T X = C (transposed=true), L n x n row-major (ldl), any n
// returns true when the update of the rows below (rowsBelow, rowMap, vec) was done as part of the solve
template <typename T>
bool trsvAny(cudaStream_t st, int batch, int64_t n, Operand<T> L, int64_t ldl, Operand<T> C, int64_t ldc, int nRHS,
             bool transposed, Operand<T> scratch /* n x nRHS per batch item, used when n > one block */,
             Operand<T> invScratch /* optional: 2 x 96 x 96 x ceil(n / 96) per batch item -> inverse-based block steps */,
             bool inversesReady = false /* the scratch already holds the inverses of this matrix */,
             ChainSync* chain = nullptr /* with invScratch: one chained launch instead of one launch per block */,
             int64_t rowsBelow = 0 /* forward + chain only: the lump's rows below the diagonal block (contiguous after */,
             const int64_t* rowMap = nullptr /* it, same ld) update vec[rowMap[r]] -= L21[r,:] x inside the same launch */,
             Operand<T> vec = Operand<T>() /* the whole vector (C is the lump's slice of it) */);

// lanes of the dense solves: v += the lanes' delta vectors (lane order), deltas zeroed
template <typename T>
void gatherLaneDeltas(cudaStream_t st, int batch, int64_t n, int nRHS, int64_t ldc, Operand<T> delta, int64_t laneStride,
                      int nLanes, Operand<T> v);

// out[i * outRowStride + c * outColStride] (+)= alpha * sum_q M[i][q] * X[c * ldx + q]   (M rows x cols, ldm)
// (tmp row-major rows x nRHS: strides (nRHS, 1); a column-major vector: strides (1, ld))
template <typename T>
void gemvRows(cudaStream_t st, int batch, int64_t rows, int64_t cols, T alpha, Operand<T> M, int64_t ldm, Operand<T> X,
              int64_t ldx, Operand<T> out, int64_t outRowStride, int64_t outColStride, int nRHS, bool accumulate,
              const int64_t* rowMap = nullptr /* device: output row of every M row (scatter = fused assembleVec) */);

// X[c * ldx + q] += alpha * sum_i M[i][q] * in[i * inRowStride + c * inColStride]
template <typename T>
void gemvColsT(cudaStream_t st, int batch, int64_t rows, int64_t cols, T alpha, Operand<T> M, int64_t ldm,
               Operand<T> in, int64_t inRowStride, int64_t inColStride, Operand<T> X, int64_t ldx, int nRHS,
               Operand<T> part = Operand<T>() /* optional scratch for row-chunk partial sums (tall panels) */,
               int64_t partCapacity = 0 /* elements per batch item */,
               const int64_t* rowMap = nullptr /* device: input row of every M row (gather = fused assembleVecT) */);

// y(col-major ldy)[0..n) += alpha * sym(M) * x(col-major ldx)[0..n), M n x n lower-stored row-major (ldm = n)
template <typename T>
void symmLower(cudaStream_t st, int batch, int64_t n, T alpha, Operand<T> M, Operand<T> X, int64_t ldx, Operand<T> Y,
               int64_t ldy, int nRHS);

}  // namespace b200
}  // namespace BaSpaCho
