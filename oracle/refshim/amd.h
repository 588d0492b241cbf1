/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * Shim for SuiteSparse's <amd.h>, which is absent from this image. It exists only so that the reference's
 * baspacho/baspacho/SparseStructure.cpp compiles unmodified with -DBASPACHO_USE_SUITESPARSE_AMD (its other
 * branch needs Eigen, also absent): every function of that file EXCEPT fillReducingPermutation
 * (SparseStructure.cpp:297-330) is then the reference's own object code. amd_l_order here returns the
 * identity ordering, so oracle/_ref says nothing about the fill-reducing ordering (pinned by the reference's
 * fill bound instead, tests/test_oracle_cpu.py::test_amd_fill_quality_bound). */
#ifndef ORACLE_REFSHIM_AMD_H_
#define ORACLE_REFSHIM_AMD_H_
#include <stdint.h>
#define AMD_CONTROL 5
#define AMD_INFO 20
#define AMD_OK 0
static inline void amd_l_defaults(double* control) { (void)control; }
static inline int amd_l_order(int64_t n, const int64_t* ap, const int64_t* ai, int64_t* p, double* control, double* info) {
  (void)ap, (void)ai, (void)control, (void)info;
  for (int64_t i = 0; i < n; i++) p[i] = i;
  return AMD_OK;
}
#endif
