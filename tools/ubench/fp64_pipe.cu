// Micro-benchmark: latency and issue rate of vector fp64 (DFMA), DMMA m8n8k4, rsqrt.approx.f64 and 64-bit shuffles on
// one SM. build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu ; run: ./fp64_pipe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE, int ILP>
__global__ void k(double* out, long long* cyc, int iters) {
  double v[ILP];
  for (int i = 0; i < ILP; i++) v[i] = 1.0 + threadIdx.x * 1e-3 + i;
  double a = 1.0000001, b = 1e-9, w[ILP];
  for (int i = 0; i < ILP; i++) w[i] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0) v[i] = fma(v[i], a, b);
      if (MODE == 1) dmma(v[i], w[i], a, b);
      if (MODE == 2) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(v[i]));
      if (MODE == 3) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1);
      if (MODE == 4) v[i] = v[i] * a;
      if (MODE == 5) v[i] = v[i] + a;
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += v[i] + w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE, ILP><<<1, threads>>>(out, cyc, iters);
  k<MODE, ILP><<<1, threads>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-8s ILP %d threads %4d : %.1f cycles per instruction per warp (%.2f per-SM warp-instr/cycle)\n", name, ILP, threads,
         (double)h / iters / ILP, (double)iters * ILP * (threads / 32) / h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 1>("dfma", 32); run<0, 8>("dfma", 32); run<0, 8>("dfma", 128); run<0, 8>("dfma", 256); run<0, 8>("dfma", 512);
  run<4, 1>("dmul", 32); run<5, 1>("dadd", 32);
  run<1, 1>("dmma", 32); run<1, 8>("dmma", 32); run<1, 8>("dmma", 128); run<1, 8>("dmma", 256);
  run<2, 1>("rsqrt", 32); run<2, 8>("rsqrt", 32);
  run<3, 1>("shfl64", 32); run<3, 8>("shfl64", 32);
  return 0;
}
