// ORACLE / TEST INFRASTRUCTURE ONLY - never linked into or called from the product library.
// Runtime binding (dlopen) of the Fortran BLAS/LAPACK entry points the reference's BackendFast uses
// (reference baspacho/baspacho/BlasDefs.h:20-55: potrf, trsm, syrk, gemm, symm). The image has no
// system BLAS; OpenBLAS copies bundled in python wheels are located by the python side and passed in.
#pragma once
#include <string>

namespace oracle_blas {

using blasint = int;  // LP64

struct Api {
  void (*dpotrf)(const char*, const blasint*, double*, const blasint*, blasint*) = nullptr;
  void (*spotrf)(const char*, const blasint*, float*, const blasint*, blasint*) = nullptr;
  void (*dtrsm)(const char*, const char*, const char*, const char*, const blasint*, const blasint*, const double*,
                const double*, const blasint*, double*, const blasint*) = nullptr;
  void (*strsm)(const char*, const char*, const char*, const char*, const blasint*, const blasint*, const float*,
                const float*, const blasint*, float*, const blasint*) = nullptr;
  void (*dsyrk)(const char*, const char*, const blasint*, const blasint*, const double*, const double*, const blasint*,
                const double*, double*, const blasint*) = nullptr;
  void (*ssyrk)(const char*, const char*, const blasint*, const blasint*, const float*, const float*, const blasint*,
                const float*, float*, const blasint*) = nullptr;
  void (*dgemm)(const char*, const char*, const blasint*, const blasint*, const blasint*, const double*, const double*,
                const blasint*, const double*, const blasint*, const double*, double*, const blasint*) = nullptr;
  void (*sgemm)(const char*, const char*, const blasint*, const blasint*, const blasint*, const float*, const float*,
                const blasint*, const float*, const blasint*, const float*, float*, const blasint*) = nullptr;
  void (*set_num_threads)(int) = nullptr;
  int (*get_num_threads)() = nullptr;
  std::string path;
  bool loaded = false;
};

// load from `path`, trying symbol name = prefix + name + "_" (+ suffix); returns false on failure
bool load(const std::string& path, const std::string& prefix, const std::string& suffix, std::string* err);
const Api& api();

}  // namespace oracle_blas
