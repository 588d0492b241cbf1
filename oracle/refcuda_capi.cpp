// ORACLE / MEASUREMENT INFRASTRUCTURE ONLY: C ABI of liboracle_refcuda.so - the restated reference CUDA backend
// (RefCudaOps.cu: cuSOLVER/cuBLAS + thread-per-item kernels with atomics) behind the same entry points as
// include/baspacho_b200.h with the prefix refcuda_ ; numeric buffers are DEVICE pointers. Used by bench.py
// (--impl ref_cuda) and tests/ to time and check the reference's GPU algorithm on the same B200.
#include <atomic>

#include "../include/baspacho_b200.h"

#define CAPI(name) refcuda_##name
#include "../baspacho_b200/csrc/capi_impl.h"

namespace BaSpaCho {
OpsPtr refCudaOps();
// this library never contains the product backend
OpsPtr b200Ops() { throw std::runtime_error("ref_cuda library: the B200 backend is not part of it"); }
}  // namespace BaSpaCho

namespace {
struct RegisterRefCuda {
  RegisterRefCuda() {
    BaSpaCho::registerBackend(BaSpaCho::BackendCuda, [](int) { return BaSpaCho::refCudaOps(); });
  }
} g_register;
}  // namespace

extern "C" {
const char* refcuda_version(void) { return "ref-cuda (restated reference cuBLAS/cuSOLVER backend) 0.1"; }
int refcuda_factor_solve_host(bspb200_solver*, int, const void*, void*, void*, int64_t, int) {
  capi_detail::lastError() = "refcuda_factor_solve_host: not provided (device-pointer baseline only)";
  return 1;
}
int64_t refcuda_launch_count(void) { return 0; }
}  // extern "C"
