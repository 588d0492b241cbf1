// TEST INFRASTRUCTURE: a "user kernel" in the sense of reference Accessor.h:19,111 ("no constructor, to be able to use
// it as argument of a Cuda kernel") - it receives Solver::deviceAccessor() BY VALUE and writes matrix blocks through
// blockOffset() / diagBlockOffset() on the device, the way a caller assembles its Hessian straight into the factor
// buffer (reference MatOpsCuda.cu:87-92, Solver.h:48). Built by tests/cpp/Makefile into libaccessor_kernel.so.
#include <cuda_runtime.h>
#include <cstdint>
#include "../../baspacho_b200/csrc/host/Accessor.h"

using BaSpaCho::PermutedCoalescedAccessor;

// value the test expects at entry (a, b) of user block (r, c)
__host__ __device__ inline double entryValue(int64_t r, int64_t c, int64_t a, int64_t b) {
  return 1000.0 * (double)r + 10.0 * (double)c + (double)a + 0.125 * (double)b;
}

__global__ void write_blocks_kernel(PermutedCoalescedAccessor acc, int64_t nBlocks, const int64_t* rowBlock,
                                    const int64_t* colBlock, double* data) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nBlocks) return;
  const int64_t r = rowBlock[i], c = colBlock[i];
  const int64_t nr = acc.paramSize(r), nc = acc.paramSize(c);
  if (r == c) {
    const auto os = acc.diagBlockOffset(r);
    for (int64_t a = 0; a < nr; a++)
      for (int64_t b = 0; b <= a; b++) data[os.first + a * os.second + b] = entryValue(r, c, a, b);
    return;
  }
  const auto loc = acc.blockOffset(r, c);
  const int64_t off = std::get<0>(loc), stride = std::get<1>(loc);
  const bool flipped = std::get<2>(loc);
  for (int64_t a = 0; a < nr; a++)
    for (int64_t b = 0; b < nc; b++) {
      const double v = entryValue(r, c, a, b);
      if (flipped) data[off + b * stride + a] = v;  // the stored block is the transpose of the requested one
      else data[off + a * stride + b] = v;
    }
}

// ptrs: the 8 device pointers of bspb200_device_accessor, in the member order of PermutedCoalescedAccessor
extern "C" int accessor_write_blocks(const int64_t* const* ptrs, int64_t nBlocks, const int64_t* devRowBlock,
                                     const int64_t* devColBlock, double* devData, void* stream) {
  PermutedCoalescedAccessor acc;
  acc.init(ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5], ptrs[6], ptrs[7]);
  const int threads = 128;
  write_blocks_kernel<<<(unsigned)((nBlocks + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      acc, nBlocks, devRowBlock, devColBlock, devData);
  return (int)cudaGetLastError();
}

extern "C" double accessor_entry_value(int64_t r, int64_t c, int64_t a, int64_t b) { return entryValue(r, c, a, b); }
