// ORACLE / TEST INFRASTRUCTURE ONLY. See BlasLoader.h.
#include "BlasLoader.h"
#include <dlfcn.h>

namespace oracle_blas {

static Api g_api;
const Api& api() { return g_api; }

bool load(const std::string& path, const std::string& prefix, const std::string& suffix, std::string* err) {
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    if (err) *err = dlerror();
    return false;
  }
  Api a;
  auto sym = [&](const char* name) { return dlsym(h, (prefix + name + "_" + suffix).c_str()); };
  a.dpotrf = (decltype(a.dpotrf))sym("dpotrf");
  a.spotrf = (decltype(a.spotrf))sym("spotrf");
  a.dtrsm = (decltype(a.dtrsm))sym("dtrsm");
  a.strsm = (decltype(a.strsm))sym("strsm");
  a.dsyrk = (decltype(a.dsyrk))sym("dsyrk");
  a.ssyrk = (decltype(a.ssyrk))sym("ssyrk");
  a.dgemm = (decltype(a.dgemm))sym("dgemm");
  a.sgemm = (decltype(a.sgemm))sym("sgemm");
  a.set_num_threads = (decltype(a.set_num_threads))dlsym(h, (prefix + "openblas_set_num_threads" + suffix).c_str());
  a.get_num_threads = (decltype(a.get_num_threads))dlsym(h, (prefix + "openblas_get_num_threads" + suffix).c_str());
  if (!a.dpotrf || !a.dtrsm || !a.dsyrk || !a.dgemm || !a.spotrf || !a.strsm || !a.ssyrk || !a.sgemm) {
    if (err) *err = "missing BLAS/LAPACK symbols in " + path;
    return false;
  }
  a.path = path;
  a.loaded = true;
  g_api = a;
  return true;
}

}  // namespace oracle_blas
