// Common definitions of the B200 (sm_100a) backend: error handling (exceptions, never abort: the
// reference's cuCHECK aborts the process, CudaDefs.h:27-65), RAII device buffers (role of the
// reference's DevMirror, CudaDefs.h:75-128), the batch-aware data reference passed to every kernel
// (role of the reference's Plain/Batched policies, MatOpsCuda.cu:345-368) and the launch counter.
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace BaSpaCho {
namespace b200 {

[[noreturn]] inline void cudaFail(cudaError_t e, const char* what, const char* file, int line) {
  std::stringstream ss;
  ss << "[" << file << ":" << line << "] CUDA error in " << what << ": " << cudaGetErrorString(e);
  throw std::runtime_error(ss.str());
}

#define B200_CUDA(call)                                                          \
  do {                                                                           \
    cudaError_t b200_err_ = (call);                                              \
    if (b200_err_ != cudaSuccess) ::BaSpaCho::b200::cudaFail(b200_err_, #call, __FILE__, __LINE__); \
  } while (0)

// number of kernels launched by this library (bench.py "gpu_launches")
std::atomic<int64_t>& launchCounter();
inline void countLaunch(int64_t n = 1) { launchCounter().fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-kernel-class profiling (bench.py roofline): CUDA events around every launch of a class on the
// launching stream, with the ALGORITHMIC flops / bytes of that launch accumulated beside the measured time.
enum KernelClass {
  KC_GEMM = 0,      // DMMA / SIMT GEMM-SYRK tiles (tensor/FMA bound)
  KC_POTRF_BLOCK,   // one-CTA diagonal block Cholesky (latency bound)
  KC_TRSM_BLOCK,    // panel triangular solve
  KC_ELIM_FACTOR,   // sparse elimination step 1 (HBM bound)
  KC_ELIM_GATHER,   // sparse elimination step 2 (HBM/L2 bound)
  KC_ASSEMBLE,      // scatter-subtract into target lumps
  KC_SOLVE_ELIM,    // elimination-range triangular solves (HBM bound)
  KC_SOLVE_DENSE,   // dense-lump triangular solves / gemv (HBM bound)
  KC_OTHER,
  KC_LUMP_CHOL,     // tile-DAG Cholesky of a wide lump column (LumpCholKernel.cu; tensor/FMA bound)
  KC_COUNT
};
void profileEnable(bool on);
bool profileEnabled();
void profileBegin(cudaStream_t st, int cls, double flops, double bytes);
void profileEnd(cudaStream_t st);
std::string profileReportJson();  // synchronizes, aggregates and clears the recorded launches

struct ProfScope {
  ProfScope(cudaStream_t st, int cls, double flops, double bytes) : st_(st), on_(profileEnabled()) {
    if (on_) profileBegin(st_, cls, flops, bytes);
  }
  ~ProfScope() {
    if (on_) profileEnd(st_);
  }
  cudaStream_t st_;
  bool on_;
};

#define B200_LAUNCH_CHECK()                 \
  do {                                      \
    ::BaSpaCho::b200::countLaunch();        \
    B200_CUDA(cudaGetLastError());          \
  } while (0)

// One matrix (`one`) or a batch of identically structured matrices (`many`: DEVICE array of `batch`
// device pointers). Kernels take the batch index from blockIdx.z.
template <typename T>
struct Mats {
  T* one = nullptr;
  T* const* many = nullptr;
  int batch = 1;
  __host__ __device__ T* at(int b) const { return many ? many[b] : one; }
};

// Per-batch-item workspace: item b lives at base + b * stride.
template <typename T>
struct Work {
  T* base = nullptr;
  int64_t stride = 0;
  __host__ __device__ T* at(int b) const { return base + (int64_t)b * stride; }
};

template <typename T>
class DevBuf {
 public:
  DevBuf() = default;
  explicit DevBuf(size_t n) { resize(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr, o.n_ = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p_ = o.p_, n_ = o.n_;
      o.p_ = nullptr, o.n_ = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }

  void resize(size_t n) {
    if (n == n_) return;
    release();
    if (n) B200_CUDA(cudaMalloc((void**)&p_, n * sizeof(T)));
    n_ = n;
  }
  void ensure(size_t n) {
    if (n > n_) resize(n);
  }
  void upload(const std::vector<T>& v) {
    resize(v.size());
    if (!v.empty()) B200_CUDA(cudaMemcpy(p_, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void release() {
    if (p_) cudaFree(p_);
    p_ = nullptr, n_ = 0;
  }
  T* ptr() const { return p_; }
  size_t size() const { return n_; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remembered per (kernel, device), so
// a process that drives several GPUs (or calls from several threads) sets it wherever the kernel is about to run
void ensureDynSmem(const void* kernel, size_t bytes);

inline int ceilDiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace b200
}  // namespace BaSpaCho
