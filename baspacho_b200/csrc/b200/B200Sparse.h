// Launch wrappers of the irregular (index-driven) kernels: sparse elimination, assemble scatter, and the
// vector gathers/scatters of the triangular solves. Definitions in SparseKernels.cu.
#pragma once

#include "B200Defs.h"
#include "B200Plan.h"

namespace BaSpaCho {
namespace b200 {

// device mirrors of the skeleton arrays (reference: CudaSymbolicCtx members, MatOpsCuda.cu:122-136)
struct DevSkel {
  const int64_t* spanStart;
  const int64_t* spanToLump;
  const int64_t* lumpStart;
  const int64_t* lumpToSpan;
  const int64_t* spanOffsetInLump;
  const int64_t* chainColPtr;
  const int64_t* chainRowSpan;
  const int64_t* chainData;
  const int64_t* chainRowsTillEnd;
  const int64_t* boardColPtr;
  const int64_t* boardRowLump;
  const int64_t* boardChainColOrd;
  int64_t numSpans, numLumps;
};

struct DevElimPlan {
  int64_t lumpsBegin, lumpsEnd, spanRowBegin;
  int uniformLumpSize;
  double factorEntries;  // entries of the eliminated columns (s^2 + s*r summed): algorithmic traffic = 2x (step 1), 1x (solve)
  double gatherFlops;    // 2 * sum over pair tasks of rowsA * rowsB * k
  double gatherBytes;    // ENTRIES moved algorithmically by step 2: every below-diagonal source block once + 2x targets
  int64_t numDst;
  int maxDstElems;
  int uniRows, uniCols, uniK;
  const int32_t* lightList;
  const int32_t* heavyList;
  int64_t numLight, numHeavy;
  int64_t lightTasks;  // pair tasks of the light destinations
  const int64_t* dstOff;
  const int32_t* dstStride;
  const int16_t* dstRows;
  const int16_t* dstCols;
  const int32_t* dstTaskPtr;
  const uint32_t* taskA;
  const uint32_t* taskB;
  const uint16_t* taskK;
  int64_t numRowSpans;
  int maxRowSpanSize;
  const int32_t* rowPtr;
  const int64_t* rowChainOff;
  const int32_t* rowChainCol;
  const int16_t* rowChainK;
};

// row view of a fragmented skeleton (every lump one span): per span the blocks found in its row, column ascending
struct FragDev {
  const int32_t* rowPtr;  // per span (+1)
  const int32_t* rowCol;  // column span of the block
  const int64_t* rowOff;  // data offset of the block (rows(span) x cols(column span), row-major)
};
template <typename T>
void fragMV(cudaStream_t st, int batch, const FragDev& f, const DevSkel& sk, Mats<T> data, Mats<T> x, Mats<T> y,
            int64_t spanBegin, int64_t spanEnd, T alpha, double bytes);
template <typename T>
void fragSolveLLevel(cudaStream_t st, int batch, const FragDev& f, const DevSkel& sk, Mats<T> data, Mats<T> y,
                     const int32_t* list, int count, int64_t spanBegin, int64_t spanEnd, bool diag);
template <typename T>
void fragSolveLtLevel(cudaStream_t st, int batch, const DevSkel& sk, Mats<T> data, Mats<T> y, const int32_t* list,
                      int count);

// step 1 of the sparse elimination: per lump, Cholesky of the diagonal block + X L^T = B on the rows below
template <typename T>
void elimFactorLumps(cudaStream_t st, int batch, const DevSkel& sk, Mats<T> data, int64_t lumpsBegin, int64_t lumpsEnd,
                     int uniformLumpSize, double profBytes = 0);
// step 2: destination-major gather of the block-pair products
template <typename T>
void elimGather(cudaStream_t st, int batch, const DevElimPlan& plan, Mats<T> data);

template <typename T>
void pseudoFactorSpans(cudaStream_t st, int batch, const DevSkel& sk, Mats<T> data, int64_t spanBegin, int64_t spanEnd);

void prepareAssemble(cudaStream_t st, const DevSkel& sk, int64_t* spanToChainOffset, int64_t targetLump,
                     int64_t numChains);

template <typename T>
void assemble(cudaStream_t st, int batch, const DevSkel& sk, const int64_t* spanToChainOffset, Mats<T> data,
              Work<T> temp, int64_t rectRowBegin, int64_t dstStride, int64_t srcColDataOffset, int64_t srcRectWidth,
              int64_t numBlockRows, int64_t numBlockCols, int64_t numRows);

template <typename T>
void elimSolveL(cudaStream_t st, int batch, const DevSkel& sk, const DevElimPlan& plan, Mats<T> data, Mats<T> C,
                int64_t ldc, int nRHS);
// out += alpha * A * in over the columns of the elimination range (both triangles of the symmetric matrix)
template <typename T>
void elimMV(cudaStream_t st, int batch, const DevSkel& sk, const DevElimPlan& plan, Mats<T> data, Mats<T> in, int64_t ldi,
            Mats<T> out, int64_t ldo, int nRHS, T alpha);

template <typename T>
void elimSolveLt(cudaStream_t st, int batch, const DevSkel& sk, const DevElimPlan& plan, Mats<T> data, Mats<T> C,
                 int64_t ldc, int nRHS);

// C[rows of the chains] += tmp (tmp: rows x nRHS row-major), and the gather mirror tmp = C[rows]
template <typename T>
void assembleVec(cudaStream_t st, int batch, const DevSkel& sk, Work<T> tmp, int64_t chainColPtr, int64_t numColItems,
                 int64_t numRows, Mats<T> C, int64_t ldc, int nRHS);
template <typename T>
void assembleVecT(cudaStream_t st, int batch, const DevSkel& sk, Work<T> tmp, int64_t chainColPtr, int64_t numColItems,
                  int64_t numRows, Mats<T> C, int64_t ldc, int nRHS);

}  // namespace b200
}  // namespace BaSpaCho
