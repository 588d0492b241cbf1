// Micro-benchmarks of the fp64 paths of one B200 SM (run under gpurun): DFMA / DMMA.884 throughput and dependent
// latency, rsqrt(double) latency, shared-memory + barrier round trip. Prints cycles; 8 warps (2 per SMSP) like the
// panel kernel. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dfma_kernel(double* out, long long* cyc, double x, int iters) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = fma(acc[i], x, 1.0);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS>
__global__ void dmma_kernel(double* out, long long* cyc, double x, int iters) {
  double acc[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc[i][0] = acc[i][1] = threadIdx.x + i;
  double a = x + threadIdx.x, b = x - threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) dmma(acc[i][0], acc[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rsqrt_kernel(double* out, long long* cyc, double x, int iters) {
  double v = x + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) v = rsqrt(v) + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void sqrtdiv_kernel(double* out, long long* cyc, double x, int iters) {
  double v = x + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) v = 1.0 / sqrt(v) + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// publish -> barrier -> read back -> dependent fma, the communication pattern of one Cholesky column group
__global__ void smem_bar_kernel(double* out, long long* cyc, int iters) {
  __shared__ double buf[2][256];
  double v = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    buf[it & 1][threadIdx.x] = v;
    __syncthreads();
    v = fma(buf[it & 1][(threadIdx.x + 33) & 255], 0.5, 1.0);
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_kernel(double* out, long long* cyc, int iters) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) v = fma(__shfl_sync(0xffffffffu, v, (it + 1) & 31), 0.5, 1.0);
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out;
  long long *cyc, h;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 64);
  const int iters = 2000;
#define RUN(name, launch, perIter)                                                   \
  launch; launch;                                                                    \
  cudaDeviceSynchronize();                                                           \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                    \
  printf("%-44s %8.2f cycles per %s (err=%s)\n", name, (double)h / iters, perIter, cudaGetErrorString(cudaGetLastError()));
  RUN("DFMA latency (1 chain, 1 warp)", (dfma_kernel<1><<<1, 32>>>(out, cyc, 0.999, iters)), "dependent DFMA")
  RUN("DFMA 8 chains, 1 warp/SM", (dfma_kernel<8><<<1, 32>>>(out, cyc, 0.999, iters)), "8 DFMA of one warp")
  RUN("DFMA 8 chains, 4 warps (1/SMSP)", (dfma_kernel<8><<<1, 128>>>(out, cyc, 0.999, iters)), "8 DFMA per warp")
  RUN("DFMA 8 chains, 8 warps (2/SMSP)", (dfma_kernel<8><<<1, 256>>>(out, cyc, 0.999, iters)), "8 DFMA per warp")
  RUN("DFMA 8 chains, 16 warps (4/SMSP)", (dfma_kernel<8><<<1, 512>>>(out, cyc, 0.999, iters)), "8 DFMA per warp")
  RUN("DFMA 16 chains, 8 warps (2/SMSP)", (dfma_kernel<16><<<1, 256>>>(out, cyc, 0.999, iters)), "16 DFMA per warp")
  RUN("DMMA.884 latency (1 chain, 1 warp)", (dmma_kernel<1><<<1, 32>>>(out, cyc, 0.999, iters)), "dependent DMMA")
  RUN("DMMA.884 8 chains, 1 warp", (dmma_kernel<8><<<1, 32>>>(out, cyc, 0.999, iters)), "8 DMMA of one warp")
  RUN("DMMA.884 8 chains, 4 warps (1/SMSP)", (dmma_kernel<8><<<1, 128>>>(out, cyc, 0.999, iters)), "8 DMMA per warp")
  RUN("DMMA.884 8 chains, 8 warps (2/SMSP)", (dmma_kernel<8><<<1, 256>>>(out, cyc, 0.999, iters)), "8 DMMA per warp")
  RUN("DMMA.884 16 chains, 8 warps (2/SMSP)", (dmma_kernel<16><<<1, 256>>>(out, cyc, 0.999, iters)), "16 DMMA per warp")
  RUN("rsqrt(double)+add latency", (rsqrt_kernel<<<1, 32>>>(out, cyc, 3.0, iters)), "rsqrt+add")
  RUN("1/sqrt(double)+add latency", (sqrtdiv_kernel<<<1, 32>>>(out, cyc, 3.0, iters)), "div+sqrt+add")
  RUN("STS -> BAR(8 warps) -> LDS -> DFMA", (smem_bar_kernel<<<1, 256>>>(out, cyc, iters)), "round trip")
  RUN("SHFL.64 -> DFMA", (shfl_kernel<<<1, 32>>>(out, cyc, iters)), "round trip")
  return 0;
}
