// ORACLE / TEST INFRASTRUCTURE ONLY.
// C face of oracle/_ref/libref_host.so: the Eigen-free translation units of the REFERENCE compiled from where they lie
// under /root/reference (Utils.cpp, SparseStructure.cpp with the shim amd.h, testing/TestingUtils.cpp,
// testing/TestingMatGen.cpp, MathUtils.h) - no reference source is copied into this repo. Used by
// tests/test_ref_objects.py to hold this repo's restated generators, pattern algebra and small-block math to the
// reference's own object code, bit for bit.
#include "baspacho/baspacho/MathUtils.h"
#include "baspacho/baspacho/SparseStructure.h"
#include "baspacho/baspacho/Utils.h"
#include "baspacho/testing/TestingMatGen.h"
#include "baspacho/testing/TestingUtils.h"

#include "host_scenarios.h"

extern "C" int64_t ref_hostcheck(int id, const double* params, int n, int64_t* out, int64_t cap) {
  return hostcheck::entry(id, params, n, out, cap);
}
extern "C" const char* ref_hostcheck_origin(void) {
  return "reference object code: baspacho/baspacho/{Utils,SparseStructure}.cpp, baspacho/testing/{TestingUtils,TestingMatGen}.cpp, MathUtils.h";
}
