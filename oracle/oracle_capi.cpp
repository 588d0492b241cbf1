// ORACLE / TEST INFRASTRUCTURE ONLY: C ABI of the CPU checker library (liboracle_cpu.so). Same entry points
// as include/baspacho_b200.h with the prefix oracle_ ; numeric buffers are HOST pointers and the
// backends are the CPU restatements in CpuOps.cpp (BackendRef = naive loops, BackendFast = BLAS + threads).
#include "../include/baspacho_b200.h"
#include "BlasLoader.h"

#define CAPI(name) oracle_##name
#include "../baspacho_b200/csrc/capi_impl.h"

namespace BaSpaCho {
OpsPtr oracleRefOps();
OpsPtr oracleFastOps(int numThreads);
// the oracle library never contains the device backend
OpsPtr b200Ops() { throw std::runtime_error("oracle library: no device backend here"); }
}  // namespace BaSpaCho

namespace {
struct RegisterCpuBackends {
  RegisterCpuBackends() {
    BaSpaCho::registerBackend(BaSpaCho::BackendRef, [](int) { return BaSpaCho::oracleRefOps(); });
    BaSpaCho::registerBackend(BaSpaCho::BackendFast, [](int n) { return BaSpaCho::oracleFastOps(n); });
  }
} g_register;
}  // namespace

extern "C" {

const char* oracle_version(void) { return "oracle-cpu (restated reference CPU backends) 0.1"; }

// bind the BLAS used by BackendFast; prefix/suffix decorate the Fortran symbol names
int oracle_load_blas(const char* path, const char* prefix, const char* suffix) {
  std::string err;
  if (!oracle_blas::load(path, prefix ? prefix : "", suffix ? suffix : "", &err)) {
    capi_detail::lastError() = err;
    return 1;
  }
  return 0;
}

int oracle_blas_threads(void) {
  return oracle_blas::api().get_num_threads ? oracle_blas::api().get_num_threads() : -1;
}

int oracle_factor_solve_host(bspb200_solver* s, int dtype, const void* host_data, void* host_factor_out, void* host_vec,
                             int64_t ld, int n_rhs) {
  return guarded([&] {
    const auto& sv = *reinterpret_cast<SolverBox*>(s)->solver;
    if (!host_factor_out) throw std::runtime_error("oracle_factor_solve_host needs an output factor buffer");
    size_t bytes = (size_t)sv.dataSize() * (dtype == 0 ? 8 : 4);
    if (host_factor_out != host_data) std::memcpy(host_factor_out, host_data, bytes);
    if (dtype == 0) {
      sv.factor((double*)host_factor_out);
      if (n_rhs > 0) sv.solve((const double*)host_factor_out, (double*)host_vec, ld, n_rhs);
    } else {
      sv.factor((float*)host_factor_out);
      if (n_rhs > 0) sv.solve((const float*)host_factor_out, (float*)host_vec, ld, n_rhs);
    }
  });
}

int64_t oracle_launch_count(void) { return 0; }

}  // extern "C"
