// Runtime invariant checks -> std::runtime_error (host code only).
// Role of reference baspacho/baspacho/DebugMacros.h:17-50.
#pragma once
#include <sstream>
#include "Utils.h"

#define BASPACHO_CHECK(cond)                                                   \
  do {                                                                         \
    if (!(cond)) ::BaSpaCho::throwError(__FILE__, __LINE__, #cond);            \
  } while (0)

#define BASPACHO_CHECK_OP(a, b, op)                                            \
  do {                                                                         \
    auto bsp_a_ = (a);                                                         \
    auto bsp_b_ = (b);                                                         \
    if (!(bsp_a_ op bsp_b_)) {                                                 \
      std::stringstream bsp_ss_;                                               \
      bsp_ss_ << #a " " #op " " #b << " (" << bsp_a_ << ", " << bsp_b_ << ")"; \
      ::BaSpaCho::throwError(__FILE__, __LINE__, bsp_ss_.str());               \
    }                                                                          \
  } while (0)

#define BASPACHO_CHECK_EQ(a, b) BASPACHO_CHECK_OP(a, b, ==)
#define BASPACHO_CHECK_LE(a, b) BASPACHO_CHECK_OP(a, b, <=)
#define BASPACHO_CHECK_LT(a, b) BASPACHO_CHECK_OP(a, b, <)
#define BASPACHO_CHECK_GE(a, b) BASPACHO_CHECK_OP(a, b, >=)
#define BASPACHO_CHECK_GT(a, b) BASPACHO_CHECK_OP(a, b, >)
#define BASPACHO_CHECK_NOTNULL(a)                                              \
  do {                                                                         \
    if ((a) == nullptr)                                                        \
      ::BaSpaCho::throwError(__FILE__, __LINE__, "'" #a "' Must be non NULL"); \
  } while (0)
