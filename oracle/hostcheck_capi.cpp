// ORACLE / TEST INFRASTRUCTURE ONLY.
// The scenarios of host_scenarios.h compiled against THIS repo's restated host layer (csrc/host, csrc/testing) and
// oracle/SmallBlockMath.h; its twin oracle/ref_capi.cpp compiles them against the reference's own sources.
#include "../baspacho_b200/csrc/host/SparseStructure.h"
#include "../baspacho_b200/csrc/host/Utils.h"
#include "../baspacho_b200/csrc/testing/TestingUtils.h"
#include "SmallBlockMath.h"

#include "host_scenarios.h"

extern "C" int64_t oracle_hostcheck(int id, const double* params, int n, int64_t* out, int64_t cap) {
  return hostcheck::entry(id, params, n, out, cap);
}
